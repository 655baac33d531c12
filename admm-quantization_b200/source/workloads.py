"""Layer enumeration and synthetic weights for the BASELINE.json configurations.

The reference factorizes one layer per process (`scripts/factorize.py --layer ...`) and names the
layers of a model in source/layer_map.py:1-98; the whole-model / sweep configurations of
BASELINE.json need the same lists as data.  There is no network for pretrained checkpoints, so weights
are synthetic with the initialisation the architecture itself uses (Kaiming-normal, fan_out, as
torchvision's ResNet does) - SURVEY 8(d).
"""
import math

import torch

from .shapes import layer_weight_as_tensor, rank_from_reduction_rate

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


def resnet50_conv_layers():
    """`get_layer_list('resnet50')` (source/layer_map.py:24-31): conv1..conv3 of every bottleneck of torchvision's
    ResNet-50 (1x1 reduce, 3x3, 1x1 expand; the first block of a stage reads the previous stage's width)."""
    blocks = {1: 3, 2: 4, 3: 6, 4: 3}
    layers, inplanes = [], 64
    for i in range(1, 5):
        planes = 64 * 2 ** (i - 1)
        for j in range(blocks[i]):
            layers.append((f"layer{i}.{j}.conv1", planes, inplanes, 1, 1))
            layers.append((f"layer{i}.{j}.conv2", planes, planes, 3, 3))
            layers.append((f"layer{i}.{j}.conv3", planes * 4, planes, 1, 1))
            inplanes = planes * 4
    return layers


def llama7b_block_layers(n_blocks=1):
    """The seven linear maps of a Llama-7B decoder block, (out, in) as stored (notebooks/LlamaADMMQuant.ipynb)."""
    layers = []
    for b in range(n_blocks):
        for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
            layers.append((f"layers.{b}.self_attn.{name}", 4096, 4096, 1, 1))
        layers.append((f"layers.{b}.mlp.gate_proj", 11008, 4096, 1, 1))
        layers.append((f"layers.{b}.mlp.up_proj", 11008, 4096, 1, 1))
        layers.append((f"layers.{b}.mlp.down_proj", 4096, 11008, 1, 1))
    return layers


def model_layers(model_name):
    """Layer specs (name, Cout, Cin, kh, kw) of the models the whole-model driver knows."""
    if model_name == "resnet18":
        return resnet18_conv_layers()
    if model_name == "resnet50":
        return resnet50_conv_layers()
    if model_name == "llama7b":
        return llama7b_block_layers(1)
    raise ValueError(f"Unknown model {model_name}")


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


def unit_cost(shape, rank, bits=4, num_attempts=200):
    """Cost of one outer sweep of a unit for sharding and SM budgets: solve_cost with the clip search weighted by its
    threshold count (2^bits - 1 thresholds per candidate: an 8-bit grid makes the search several times dearer than a
    4-bit one).  Calibrated on the 256-unit sweep: per-rank times follow the summed cost to ~6 %."""
    return solve_cost(shape, rank, num_attempts) * (1.0 + ((1 << int(bits)) - 1) / 60.0)


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


# ---------------------------------------------------------------------------------------------------------------
# SM budgets for concurrently running solves.  A layer's persistent kernels run on `g` CTAs (max_ctas); the tensor-core
# ridge product works in whole waves of 128 x bn tiles, so its time is a STEP function of g (72 tiles of width 64 take 3
# waves on 24 .. 35 CTAs and 2 on 36), while the clip search / dual update scale like 1/g above a latency floor.  The
# budgets are chosen from that model, fitted to the phase times the kernels report (admmq_loop_report.phase_ns).
_TILE_FIXED = 16           # csrc/admm_loop.cu kTileFixed: tile cost that does not depend on its width
_PHASE_FLOOR_US = 14.0     # P2 + P3 of a factor that is too small to matter: barriers and L2 round trips


_STAGE_CAP = 23296         # csrc/search.cuh kStageCap
_STAGE_FIXED_US = 7.0      # scan + threshold pass of one more stage (measured: 20.7 k elements as two stages 27 us, as one 20 us)


def search_stages(rows, rank, g):
    """Stages of the clip search per CTA when rows x rank elements are split over g CTAs (chunks are multiples of 64)."""
    chunk = -(-(rows * rank) // g)
    chunk = -(-chunk // 64) * 64
    return max(1, -(-chunk // _STAGE_CAP))


def product_cost(rows, rank, g, solve_precision=1):
    """Relative time of one ridge product H_ls = RHS . Minv on g CTAs (mirrors launch_loop in csrc/admm_loop.cu)."""
    if rows <= 16 or solve_precision != 1 or rows < 64 or rank < 32:
        return 1.0 / g                                   # column strips / small FFMA tiles: scales with the grid
    tiles_m = (rows + 127) // 128
    best = None
    for bn in range(160, 15, -16):
        tiles = tiles_m * ((rank + bn - 1) // bn)
        cost = ((tiles + g - 1) // g) * (_TILE_FIXED + bn) * (1.05 if bn > 128 else 1.0)   # mirrors launch_loop
        best = cost if best is None or cost < best else best
    return float(best)


def predict_sweep_ms(measured_ms, g0, g, modes, solve_precision=1):
    """Sweep time of one layer on g CTAs from a measurement on g0 CTAs.
    modes: [(rows, rank, p1_us, p23_us, iterations)] per factor - P1 (ridge product) and P2 + P3 per inner iteration."""
    loop0 = sum(it * (p1 + p23) for _, _, p1, p23, it in modes) / 1e3
    rest = max(measured_ms - loop0, 0.0)                 # Gram, MTTKRP, inverse, projection, errors: taken as fixed
    loop = 0.0
    for rows, rank, p1, p23, it in modes:
        p1g = p1 * product_cost(rows, rank, g, solve_precision) / product_cost(rows, rank, g0, solve_precision)
        floor = min(_PHASE_FLOOR_US, p23)
        p23g = floor + (p23 - floor) * g0 / g
        # the clip search works in stages of _STAGE_CAP elements per CTA; every further stage repeats its fixed costs
        p23g += _STAGE_FIXED_US * (search_stages(rows, rank, g) - search_stages(rows, rank, g0))
        loop += it * (p1g + max(p23g, floor)) / 1e3
    return rest + loop


def allocate_ctas_modelled(layers, sm_count, solve_precision=1):
    """layers: [(measured_ms, g0, modes)] as for predict_sweep_ms.  Minimax over the model: the smallest common
    deadline T for which every layer has a budget g_l with predicted time <= T and sum g_l <= sm_count (bisection on T;
    per layer the smallest such g_l, so budgets sit on wave boundaries), then the SMs that are left go, step by step,
    to whichever layer is predicted to finish last and still profits.  Returns the budgets (sum <= sm_count)."""
    n = len(layers)
    if n >= sm_count:
        return [1] * n
    table = [[predict_sweep_ms(m, g0, g, modes, solve_precision) for g in range(1, sm_count + 1)] for m, g0, modes in layers]

    def budgets_for(T):
        out = []
        for row in table:
            g = next((k + 1 for k, t in enumerate(row) if t <= T), None)
            if g is None:
                return None
            out.append(g)
        return out if sum(out) <= sm_count else None

    lo = max(min(row) for row in table)      # nobody can beat its own best time
    hi = max(row[0] for row in table)        # one CTA each always fits (n < sm_count)
    best = budgets_for(hi)
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        b = budgets_for(mid)
        if b is None:
            lo = mid
        else:
            hi, best = mid, b
    g = list(best)
    free = sm_count - sum(g)
    while free > 0:
        order = sorted(range(n), key=lambda i: -table[i][g[i] - 1])
        moved = False
        for k in order:
            cur = table[k][g[k] - 1]
            step = next((e for e in range(1, free + 1) if table[k][g[k] - 1 + e] < cur * (1.0 - 1e-3)), None)
            if step is not None:
                g[k] += step
                free -= step
                moved = True
                break
        if not moved:
            break
    return g
