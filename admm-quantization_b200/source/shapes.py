"""Shape rules of the reference CLI - pure Python, no native library (importable by the CPU-only legs of bench.py)."""


def rank_from_reduction_rate(weight, reduction_rate):
    """scripts/factorize.py:157-158."""
    return int(weight.numel() / sum(list(weight.shape)) / reduction_rate)


def rank_for_shape(shape, reduction_rate):
    """The same rule from a shape tuple."""
    numel = 1
    for d in shape:
        numel *= int(d)
    return int(numel / sum(int(d) for d in shape) / reduction_rate)


def layer_weight_as_tensor(weight):
    """Intended reshape of scripts/factorize.py:138-145 / scripts/calibrate.py:178-184:
    conv (Cout,Cin,kh,kw) -> (Cout,Cin,kh*kw); 1x1 conv -> (Cout,Cin)."""
    if weight.ndim == 4:
        if tuple(weight.shape[2:]) == (1, 1):
            return weight.reshape(weight.shape[0], weight.shape[1])
        return weight.reshape(weight.shape[0], weight.shape[1], -1)
    return weight
