"""ADMM inner loop, factor initialisation and error metric - drop-in for the reference's
source/admm.py:14-67.  The arithmetic runs in libadmmq.so (csrc/admm_loop.cu)."""
import numpy as np
import torch

from . import _native
from .quantization import quantize_tensor  # noqa: F401  (re-exported like the reference module)
from .utils import unfold


def squared_relative_diff(X, Y):
    """sqrt(sum((X-Y)^2) / sum(X^2)) as a Python float (reference source/admm.py:14-15).
    The solver's own per-sweep errors never materialise Y (admmq_recon_error); this helper keeps
    the reference signature for callers that already hold a reconstruction."""
    return torch.sqrt(torch.sum((X - Y) ** 2) / torch.sum(X ** 2)).item()


def init_factors(tensor, rank, init='random', device=None, seed=None, rng=None):
    """reference source/admm.py:21-48.  'random' and 'svd' draw from a torch.Generator on `device`
    exactly like the reference; 'parafac' / 'parafac-epc' run the restated ALS / ALS+EPC
    (source/parafac_epc.py) because tensorly/musco are third-party packages.  `rng` (optional numpy RandomState, an
    addition to the reference signature) replaces numpy's global stream for the 'parafac-epc' start."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    factors = []
    if init == 'random':
        for mode in range(tensor.ndim):
            factors.append(torch.randn(tensor.shape[mode], rank, generator=gen, device=device))
    elif init == 'svd':
        for mode in range(tensor.ndim):
            Umat, _, _ = torch.linalg.svd(unfold(tensor, mode), full_matrices=False)
            if tensor.shape[mode] < rank:
                pad = torch.randn(Umat.shape[0], rank - tensor.shape[mode], generator=gen, device=device)
                Umat = torch.cat((Umat, pad.to(Umat.device)), axis=1)
            factors.append(Umat[:, :rank])
    elif init == 'parafac':
        from .parafac_epc import parafac_als
        _, factors = parafac_als(tensor, rank, n_iter_max=100, tol=1e-5, random_state=seed,
                                 normalize_factors=False, dtype=torch.float32)
    elif init == 'parafac-epc':
        from .parafac_epc import parafac_epc
        _, factors = parafac_epc(tensor, rank=rank, init='random', als_maxiter=50, epc_maxiter=50, rng=rng)
        factors = [f.to(dtype=torch.float) for f in factors]
    else:
        raise NotImplementedError(init)
    return factors


last_report = None  # LoopReport of the most recent admm_iteration call (diagnostics)


def admm_iteration(H, U, F, G, max_iter, eps, bits, qscheme, num_attempts=200, return_codes=False, precision=0):
    """reference source/admm.py:51-67: `max_iter - 1` iterations of
    { H_ls = (G + rho I)^-1 (F + rho (H + U)); H = Q(H_ls - U); U += H - H_ls }, rho = trace(G)/R.
    Returns a NEW tensor H and the caller's U, which is updated in place (:60, :67)."""
    global last_report
    _native.require_cuda(H, U, F, G)
    Hc = _native.f32c(H).clone()
    Uc = U if (U.dtype == torch.float32 and U.is_contiguous()) else _native.f32c(U).clone()
    codes = torch.empty(Hc.shape, dtype=torch.int8, device=Hc.device) if return_codes else None
    report = _native.admm_iteration_inplace(Hc, Uc, _native.f32c(F), _native.f32c(G), max_iter, eps, bits, qscheme,
                                            num_attempts, codes, precision)
    last_report = _native.read_report(report)  # raises LinAlgError if G + rho I is not positive definite
    if Uc is not U:
        U.copy_(Uc)
    if return_codes:
        return Hc, U, codes
    return Hc, U
