"""CP-format layers built from factor matrices - the consumer side of the factor files the solver writes
(reference source/models.py:24-74, used by scripts/calibrate.py:151-189).

A rank-R CP factorization W[o, i, k] = sum_r A[o, r] B[i, r] C[k, r] of a (Cout, Cin, kh*kw) convolution weight is
the composition of three convolutions: a 1x1 Cin -> R with weight B^T, a depthwise kh x kw on the R channels with
the r-th filter C[:, r] (which carries the padding and stride of the original layer), and a 1x1 R -> Cout with
weight A (which carries the bias).  A 1x1 convolution (2-D factorization W = A B^T) needs only the two pointwise
layers.  Module and parameter names (conv1 / conv2 / conv3, fc1 / fc2) follow the reference so that state dicts are
interchangeable.
"""
from collections import OrderedDict

import torch
from torch import nn


def _set(conv, weight, bias=None):
    assert tuple(conv.weight.shape) == tuple(weight.shape), \
        f"Expected shape: {tuple(conv.weight.shape)}, but got {tuple(weight.shape)}"
    with torch.no_grad():
        conv.weight = nn.Parameter(weight.detach().clone().contiguous(), requires_grad=True)
        if bias is not None:
            assert tuple(conv.bias.shape) == tuple(bias.shape), \
                f"Expected shape: {tuple(conv.bias.shape)}, but got {tuple(bias.shape)}"
            conv.bias = nn.Parameter(bias.detach().clone(), requires_grad=True)


def build_cp_layer(rank, factors, bias, cin, cout, kernel_size, padding, stride, groups):
    """reference source/models.py:24-50.  factors = [A (cout, R), B (cin, R), C (kh*kw, R)] or None/[] for an
    uninitialised skeleton."""
    has_bias = bias is not None
    seq = nn.Sequential(OrderedDict([
        ("conv1", nn.Conv2d(cin, rank, kernel_size=(1, 1), groups=groups, bias=False)),
        ("conv2", nn.Conv2d(rank, rank, kernel_size=kernel_size, groups=rank, padding=padding, stride=stride, bias=False)),
        ("conv3", nn.Conv2d(rank, cout, kernel_size=(1, 1), bias=has_bias)),
    ]))
    if factors:
        A, B, C = factors
        kh, kw = kernel_size
        _set(seq.conv1, B.t()[:, :, None, None])                                   # (R, cin, 1, 1)
        _set(seq.conv2, C.reshape(kh, kw, rank).permute(2, 0, 1)[:, None, :, :])   # (R, 1, kh, kw)
        _set(seq.conv3, A[:, :, None, None], bias if has_bias else None)          # (cout, R, 1, 1)
    return seq


def build_cp2conv_layer(rank, factors, bias, cin, cout, padding, stride):
    """reference source/models.py:53-74: a factorized 1x1 convolution, factors = [A (cout, R), B (cin, R)]."""
    has_bias = bias is not None
    seq = nn.Sequential(OrderedDict([
        ("conv1", nn.Conv2d(cin, rank, kernel_size=(1, 1), padding=padding, stride=stride, bias=False)),
        ("conv2", nn.Conv2d(rank, cout, kernel_size=(1, 1), bias=has_bias)),
    ]))
    if factors:
        A, B = factors
        _set(seq.conv1, B.t()[:, :, None, None])
        _set(seq.conv2, A[:, :, None, None], bias if has_bias else None)
    return seq


def build_cpfc_layer(rank, factors, bias, fin, fout):
    """reference source/models.py:77-95: a factorized linear layer, factors = [A (fout, R), B (fin, R)]."""
    has_bias = bias is not None
    seq = nn.Sequential(OrderedDict([
        ("fc1", nn.Linear(fin, rank, bias=False)),
        ("fc2", nn.Linear(rank, fout, bias=has_bias)),
    ]))
    if factors:
        A, B = factors
        _set(seq.fc1, B.t())
        _set(seq.fc2, A, bias if has_bias else None)
    return seq


def get_submodule(model, path):
    mod = model
    for attr in path.split("."):
        mod = getattr(mod, attr)
    return mod


def replace_with_cp(model, layer_path, factors, rank=None):
    """Swap the convolution at `layer_path` for its CP form built from `factors` (the loop body of
    reference scripts/calibrate.py:161-189).  Returns the new module."""
    layer = get_submodule(model, layer_path)
    rank = factors[0].shape[1] if rank is None else rank
    bias = layer.bias.detach() if layer.bias is not None else None
    dev = layer.weight.device
    factors = [f.to(device=dev, dtype=torch.float32) for f in factors]
    if tuple(layer.kernel_size) != (1, 1):
        new = build_cp_layer(rank, factors, bias, layer.in_channels, layer.out_channels, tuple(layer.kernel_size),
                             layer.padding, layer.stride, layer.groups)
    else:
        new = build_cp2conv_layer(rank, factors, bias, layer.in_channels, layer.out_channels, layer.padding, layer.stride)
    new = new.to(dev)
    parent_path, _, leaf = layer_path.rpartition(".")
    setattr(get_submodule(model, parent_path) if parent_path else model, leaf, new)
    return new
