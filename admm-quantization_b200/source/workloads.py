"""Layer enumeration and synthetic weights for the BASELINE.json configurations.

The reference factorizes one layer per process (`scripts/factorize.py --layer ...`) and names the
layers of a model in source/layer_map.py:1-98; the whole-model / sweep configurations of
BASELINE.json need the same lists as data.  There is no network for pretrained checkpoints, so weights
are synthetic with the initialisation the architecture itself uses (Kaiming-normal, fan_out, as
torchvision's ResNet does) - SURVEY 8(d).
"""
import math

import torch

from .solver import layer_weight_as_tensor, rank_from_reduction_rate

# (name, Cout, Cin, kh, kw) of the 3x3 convolutions factorized for ResNet-18
# (`get_layer_list('resnet18', downsample=False, conv1=False)`, source/layer_map.py:10-13)
_RESNET18_WIDTHS = {"layer1": (64, 64), "layer2": (128, 64), "layer3": (256, 128), "layer4": (512, 256)}


def resnet18_conv_layers():
    layers = []
    for stage, (width, prev) in _RESNET18_WIDTHS.items():
        for block in (0, 1):
            for conv in (1, 2):
                cin = prev if (block == 0 and conv == 1) else width
                layers.append((f"{stage}.{block}.conv{conv}", width, cin, 3, 3))
    return layers


def llama7b_linear_layers():
    """Config 5 shapes: the (out, in) matrices are used as they are."""
    return [("q_proj", 4096, 4096, 1, 1), ("gate_proj", 11008, 4096, 1, 1)]


def resnet50_layer4_layers():
    """Config 4 shapes."""
    return [("layer4.conv2", 512, 512, 3, 3), ("layer4.conv3", 2048, 512, 1, 1)]


def synthetic_weight(cout, cin, kh, kw, seed, name=""):
    """Kaiming-normal (fan_out, relu gain) conv weight, CPU generator so every rank / host sees the
    same numbers for the same (seed, name)."""
    g = torch.Generator().manual_seed((int(seed) * 1000003 + sum(ord(c) * (i + 1) for i, c in enumerate(name))) % (2 ** 31))
    std = math.sqrt(2.0 / (cout * kh * kw))
    return torch.randn(cout, cin, kh, kw, generator=g) * std


def random_init(shape, rank, seed):
    """`init_factors(..., init='random')` on a CPU generator (source/admm.py:22-28)."""
    g = torch.Generator().manual_seed(int(seed))
    return [torch.randn(int(d), rank, generator=g) for d in shape]


def build_problems(layers, reduction_rate=2.0, weight_seed=42, init_seed=42):
    """[(name, W (CPU, 2-D/3-D), rank, [init factors (CPU)])] for a list of layer specs."""
    out = []
    for name, cout, cin, kh, kw in layers:
        W = layer_weight_as_tensor(synthetic_weight(cout, cin, kh, kw, weight_seed, name)).contiguous()
        rank = rank_from_reduction_rate(W, reduction_rate)
        out.append((name, W, rank, random_init(W.shape, rank, init_seed)))
    return out


def solve_cost(shape, rank, num_attempts=200, c_eval=8.0):
    """Relative cost of one outer sweep: per inner iteration 2*I*R^2 (ridge product) +
    num_attempts * c_eval * I * R (clip search), summed over modes; used for LPT sharding."""
    return sum(2.0 * d * rank * rank + num_attempts * c_eval * d * rank for d in shape)


def lpt_assign(costs, n_bins):
    """Longest-processing-time-first assignment; returns bin index per item."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * n_bins
    where = [0] * len(costs)
    for i in order:
        b = min(range(n_bins), key=lambda k: load[k])
        where[i] = b
        load[b] += costs[i]
    return where
