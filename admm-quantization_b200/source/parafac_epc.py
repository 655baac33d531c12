"""ALS + EPC initialisation - drop-in for the reference's source/parafac_epc.py:12-82.

The reference delegates the arithmetic to two third-party packages that are neither vendored nor
installed here: `tensorly.decomposition.parafac` (0.4.5) and musco's `cp_anc` (1.0.6).  Both are
restated from their published algorithms (PARITY UNPINNED - there is nothing to
compare bit for bit): ALS with normal-equation solves and column normalisation, and error-preserving
correction (Phan, Tichavsky, Cichocki, IEEE TSP 2019): minimise the sum of squared component norms
subject to ||Y - Yhat|| <= delta.

Everything runs in float64 ON THE GPU (there is no CPU path: a CPU tensor is moved to the current CUDA
device and the results come back on the tensor's device, like a host-buffer call).  The contractions
that dominate a pass are libadmmq kernels (csrc/contract.cu): the MTTKRP forms the Khatri-Rao operand on
the fly (tensorly materialises it: (I J) x R doubles, 2.4 GB for mode 2 of a 512 x 512 x 9 layer, every
pass), the Gram-Hadamard product and the column normalisation are one kernel each.  The R x R dense
factorizations of a mode update - `solve` in ALS, the symmetric eigen-decomposition in EPC - stay with
torch (cuSOLVER), see SURVEY K9 / K10.  The wrapper logic (float64 copy, ascending mode permutation,
rounds, stop rules, original mode order) follows source/parafac_epc.py line by line.
"""
import math

import numpy as np
import torch

from . import _native
from .utils import unfold


def _device_for(tensor):
    if tensor.is_cuda:
        return tensor.device
    if not torch.cuda.is_available():
        raise RuntimeError("parafac_epc / parafac_als need a CUDA device: the contractions are libadmmq kernels "
                           "(there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class _Tensor:
    """A float64 tensor on the GPU with its mode unfoldings made once (the operands of the MTTKRP kernel)."""

    def __init__(self, Y):
        assert Y.is_cuda and Y.dtype == torch.float64
        if Y.ndim not in (2, 3):
            raise ValueError("Incorrect number of dimentions in weight tensor")   # scripts/factorize.py:154
        self.Y, self.N = Y, Y.ndim
        self.unf = [unfold(Y, m).contiguous() for m in range(Y.ndim)]
        self.norm2 = float(torch.sum(Y * Y))
        self.ws = None

    def others(self, factors, mode):
        o = [f for k, f in enumerate(factors) if k != mode]
        return o[0], (o[1] if len(o) > 1 else None)

    def mttkrp(self, factors, mode):
        X, Yf = self.others(factors, mode)
        need = int(_native.lib.admmq_mttkrp_f64_workspace_bytes(self.unf[mode].shape[0], X.shape[0],
                                                                1 if Yf is None else Yf.shape[0], X.shape[1]))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.Y.device)
        return _native.mttkrp_f64(self.unf[mode], X.contiguous(), None if Yf is None else Yf.contiguous(), ws=self.ws)

    def gram(self, factors, mode):
        X, Yf = self.others(factors, mode)
        return _native.gram_hadamard_f64(X.contiguous(), None if Yf is None else Yf.contiguous())


def _reconstruct(weights, factors):
    N = len(factors)
    letters = "".join(chr(105 + k) for k in range(N))
    spec = "r," + ",".join(f"{c}r" for c in letters) + "->" + letters
    return torch.einsum(spec, weights, *factors)


def _als(T, rank, n_iter_max, tol, rng, normalize_factors):
    """tensorly 0.4.5 `parafac(init='random')` on a `_Tensor` (float64).  Returns (weights, factors)."""
    Y, N, dev = T.Y, T.N, T.Y.device
    factors = [torch.from_numpy(rng.random_sample((Y.shape[m], rank))).to(device=dev, dtype=torch.float64) for m in range(N)]
    if normalize_factors:
        factors = [f / (torch.linalg.norm(f, dim=0) + 1e-12) for f in factors]
    weights = torch.ones(rank, dtype=torch.float64, device=dev)
    norm_y = math.sqrt(T.norm2)
    errs = []
    for it in range(n_iter_max):
        mt = None
        for m in range(N):
            gram = T.gram(factors, m)
            mt = T.mttkrp(factors, m)
            f = torch.linalg.solve(gram.T, mt.T).T.contiguous()
            if normalize_factors:
                weights = torch.linalg.norm(f, dim=0)
                weights = torch.where(weights <= torch.finfo(torch.float64).eps, torch.ones_like(weights), weights)
                f = f / weights
            factors[m] = f
        if tol:
            gram_all = T.gram(factors, N - 1) * _native.gram_hadamard_f64(factors[N - 1].contiguous())
            norm_rec2 = (weights[:, None] * weights[None, :] * gram_all).sum()
            inner = (weights * (mt * factors[N - 1]).sum(dim=0)).sum()
            err = math.sqrt(abs(T.norm2 + float(norm_rec2) - 2 * float(inner))) / norm_y
            errs.append(err)
            if it >= 1 and abs(errs[-2] - errs[-1]) < tol:
                break
    return weights, factors


def parafac_als(tensor, rank, n_iter_max=100, tol=1e-8, random_state=None, normalize_factors=False,
                dtype=None):
    """`tensorly.decomposition.parafac(tensor, rank, init='random', ...)` as the reference calls it
    (source/admm.py:38-39, source/parafac_epc.py:42-43, scripts/factorize.py:324-325).
    Returns (weights, factors) in `dtype` (default: the tensor's) on the tensor's device; the iteration itself runs in
    float64.  `random_state=None` draws from numpy's GLOBAL stream, like tensorly."""
    if isinstance(random_state, np.random.RandomState):
        rng = random_state
    else:
        rng = np.random.mtrand._rand if random_state is None else np.random.RandomState(random_state)
    dtype = tensor.dtype if dtype is None else dtype
    dev = _device_for(tensor)
    weights, factors = _als(_Tensor(tensor.detach().to(device=dev, dtype=torch.float64).contiguous()), rank, n_iter_max,
                            tol, rng, normalize_factors)
    return weights.to(device=tensor.device, dtype=dtype), [f.to(device=tensor.device, dtype=dtype) for f in factors]


def _multiplier(sig_c, s_c, norm_y2, target):
    """mu >= 0 with residual(mu) = ||Y||^2 - sum_i s_i (sig_i + 2 mu) / (sig_i + mu)^2 = target (increasing in mu;
    mu = 0 when the least-squares residual already reaches the target): 64-way subdivision on the host - 9 vectorised
    evaluations reach 1e-15 relative width (numpy on R-vectors: cheaper than any device round trip)."""
    def resid(mu):
        d = sig_c + mu
        return norm_y2 - float(np.sum(s_c * (d + mu) / (d * d)))

    floor = float(sig_c.max()) * 1e-14
    mu = 0.0
    if resid(floor) < target:
        lo, hi = floor, max(float(sig_c.max()), 1e-300)
        while resid(hi) < target and hi < 1e300:
            hi *= 2.0
        for _ in range(12):
            grid = lo + (hi - lo) * (np.arange(1, 64) / 64.0)
            d = sig_c[None, :] + grid[:, None]
            vals = norm_y2 - np.sum(s_c[None, :] * (d + grid[:, None]) / (d * d), axis=1)
            k = int(np.searchsorted(vals >= target, True))      # first grid point with residual >= target
            lo, hi = (grid[k - 1] if k > 0 else lo), (grid[k] if k < 63 else hi)
            if hi - lo <= 1e-15 * hi:
                break
        mu = 0.5 * (lo + hi)
    return max(mu, floor)


def _epc_sweep(T, factors, delta):
    """One error-preserving-correction pass over all modes (the body of musco's `cp_anc`); `factors` are float64 CUDA
    tensors owned by the caller and updated in place / replaced."""
    rank = factors[0].shape[1]
    target = delta * delta
    for m in range(T.N):
        for k in range(T.N):                        # unit columns in the other factors (their scale moves into mode m,
            if k != m:                              # whose factor is recomputed below)
                if not factors[k].is_contiguous():
                    factors[k] = factors[k].contiguous()
                _native.normalize_columns_f64(factors[k])
        gamma = T.gram(factors, m)
        Tm = T.mttkrp(factors, m)
        sig, V = torch.linalg.eigh(gamma)
        sig = torch.clamp(sig, min=0.0)
        Tt = Tm @ V
        s = (Tt * Tt).sum(dim=0)
        both = torch.stack([sig, s]).cpu().numpy()  # one device -> host copy per mode update
        mu = _multiplier(both[0], both[1], T.norm2, target)
        factors[m] = ((Tt / (sig + mu)) @ V.T).contiguous()
    return factors


def epc_sweep(Y, factors, delta):
    """Public form of one EPC pass: Y and factors anywhere; returns new float64 factors on Y's device."""
    dev = _device_for(Y)
    T = _Tensor(Y.detach().to(device=dev, dtype=torch.float64).contiguous())
    fac = _epc_sweep(T, [f.detach().to(device=dev, dtype=torch.float64).contiguous().clone() for f in factors], float(delta))
    return [f.to(Y.device) for f in fac]


def _intensities(factors):
    lam = torch.ones(factors[0].shape[1], dtype=factors[0].dtype, device=factors[0].device)
    for f in factors:
        lam = lam * torch.linalg.norm(f, dim=0)
    return lam


def parafac_epc(tensor, rank, als_maxiter=5000, als_tol=1e-5, num_threads=4, init="random",
                epc_maxiter=5000, epc_rounds=50, epc_tol=1e-5, stop_tol=1e-4, ratio_tol=1e-3,
                ratio_max_iters=10, info=None, rng=None):
    """reference source/parafac_epc.py:12-82.  Returns (lmbda, Us) with Us in the ORIGINAL mode order,
    float64, on the device of `tensor`.  `num_threads` is accepted for signature compatibility; the
    reference's `torch.set_num_threads` (:33) is a CPU-side global side effect that has no GPU meaning.
    `info` (a dict, optional) receives delta, the pass counts and the seconds spent in ALS and EPC.  `rng` (optional
    numpy RandomState) replaces numpy's GLOBAL stream, which tensorly's random ALS start draws from (:42-43 pass no
    random_state): a driver that initialises several layers concurrently gives every layer RandomState(seed), the stream
    a fresh process gets from `np.random.seed(seed)` (scripts/factorize.py:21-24)."""
    import time
    if init != "random":
        raise NotImplementedError(init)
    dev = _device_for(tensor)
    t0 = time.perf_counter()
    Y = tensor.detach().to(device=dev, dtype=torch.float64)                 # :36
    order = np.argsort(Y.shape)                                             # :38
    T = _Tensor(Y.permute(tuple(int(o) for o in order)).contiguous())       # :40
    weights, factors = _als(T, rank, als_maxiter, als_tol, np.random.mtrand._rand if rng is None else rng, True)   # :42-43
    delta = float(torch.linalg.norm(T.Y - _reconstruct(weights, factors)))  # :51
    lam_prev_norm = float(torch.linalg.norm(weights))                       # :52
    factors[-1] = (factors[-1] * weights).contiguous()                      # :53
    alpha_prev = float(weights.max() / weights.min())                       # :57
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    stopflag, passes, rounds = 0, 0, 0
    lam = _intensities(factors)
    for _ in range(epc_rounds):                                             # :61
        prev = None
        rounds += 1
        for _it in range(epc_maxiter):                                      # cp_anc(maxiter, tol)  :63
            factors = _epc_sweep(T, factors, delta)
            passes += 1
            cur = float((_intensities(factors) ** 2).sum())
            if prev is not None and abs(prev - cur) < epc_tol * prev:
                break
            prev = cur
        lam = _intensities(factors)
        lam_norm = float(torch.linalg.norm(lam))
        alpha = float(lam.max() / lam.min())
        if abs(lam_prev_norm - lam_norm) < stop_tol * lam_prev_norm:        # :67
            break
        stopflag = stopflag + 1 if abs(alpha_prev - alpha) < ratio_tol else 0   # :69
        lam_prev_norm, alpha_prev = lam_norm, alpha
        if stopflag >= ratio_max_iters:                                     # :74
            break
    torch.cuda.synchronize(dev)
    if info is not None:
        info.update(delta=delta, norm=math.sqrt(T.norm2), als_s=t1 - t0, epc_s=time.perf_counter() - t1,
                    epc_passes=passes, epc_rounds=rounds)
    inv = np.argsort(order)
    return lam.to(tensor.device), [factors[int(i)].to(tensor.device) for i in inv]   # :77-82
