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
factorizations of a mode update stay with torch (cuSOLVER / cuBLAS), see SURVEY K9 / K10: `solve` in ALS; in EPC
the symmetric eigen-decomposition only for the first update of a run - every later update expands the residual
and the factor in power series around a warm-started multiplier from ONE Cholesky factorization
(`_ridge_factor_chol`: 2 ms against 15 ms at R = 1141, same multiplier and factor to 1e-12).  The wrapper logic
(float64 copy, ascending mode permutation, rounds, stop rules, original mode order) follows
source/parafac_epc.py line by line.
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


_EYE = {}
_EYE_LOCK = __import__("threading").Lock()     # layers are initialised concurrently on host threads (bench.py --full)


def _eye(n, like):
    """Identity cached per (n, device).  The cache is shared by the host threads that initialise several layers
    concurrently, each on its own stream: the creating stream is synchronised once so that no other stream can read
    the constant before it is written."""
    key = (n, like.device)
    e = _EYE.get(key)
    if e is None:
        e = torch.eye(n, dtype=like.dtype, device=like.device)
        if e.is_cuda:
            torch.cuda.current_stream(e.device).synchronize()
        with _EYE_LOCK:
            while sum(1 for k in _EYE if k[0] != "mom") >= 16:  # a sweep over many ranks must not pin one n x n matrix each
                _EYE.pop(next(k for k in _EYE if k[0] != "mom"), None)
            _EYE[key] = e
    return e


def _moment_index(terms, device):
    """Flat positions in the (terms + 1) x (terms + 1) Gram matrix of [Tm, F_1 .. F_terms] that hold a_1 .. a_{2 terms}:
    a_1 = <Tm, F_1>, a_k = <F_{k // 2}, F_{k - k // 2}>.  Cached per (terms, device): no index upload per call."""
    key = ("mom", terms, device)
    if key not in _EYE:
        ii = [0] + [k // 2 for k in range(2, 2 * terms + 1)]
        jj = [1] + [k - k // 2 for k in range(2, 2 * terms + 1)]
        idx = torch.tensor([i * (terms + 1) + j for i, j in zip(ii, jj)], dtype=torch.long, device=device)
        if idx.is_cuda:
            torch.cuda.current_stream(idx.device).synchronize()   # shared across streams, see _eye
        with _EYE_LOCK:
            _EYE[key] = idx
    return _EYE[key]


def _ridge_eval(gamma, Tm, mu, terms):
    """Everything the multiplier search needs around one value of mu from ONE Cholesky factorization of
    M = Gamma + mu I: the explicit inverse M^-1 = L^-T L^-1 (potrf, one triangular solve against the identity, one
    GEMM: half the time of cuSOLVER's potri-style `cholesky_inverse`), the sequence F_k = Tm M^-k, k = 1..terms (one
    float64 GEMM each, written into one buffer) and the moments a_k = tr(Tm M^-k Tm^T), k = 1..2 terms, read off the
    Gram matrix of the flattened [Tm, F_1 .. F_terms] (ONE GEMM): a_1 = <F_1, Tm>, a_{i+j} = <F_i, F_j>.
    Returns (buffer [Tm, F_1 .. F_terms] of shape (terms + 1) x I x R, stats) with
    stats = [a_1 .. a_{2 terms}, ||F_1 M - Tm|| / ||Tm||, info] still on the device."""
    R = gamma.shape[0]
    eye = _eye(R, gamma)
    M = torch.add(gamma, eye, alpha=mu)
    L, info = torch.linalg.cholesky_ex(M)
    Li = torch.linalg.solve_triangular(L, eye, upper=False)
    Minv = Li.T @ Li
    buf = torch.empty((terms + 1,) + tuple(Tm.shape), dtype=Tm.dtype, device=Tm.device)
    buf[0].copy_(Tm)
    for k in range(terms):
        torch.matmul(buf[k], Minv, out=buf[k + 1])
    flat = buf.view(terms + 1, -1)
    G = flat @ flat.T                                        # G[i, j] = <F_i, F_j>, F_0 = Tm
    stats = torch.empty(2 * terms + 2, dtype=Tm.dtype, device=Tm.device)
    torch.index_select(G.view(-1), 0, _moment_index(terms, G.device), out=stats[:2 * terms])
    stats[2 * terms] = torch.linalg.norm(torch.addmm(Tm, buf[1], M, beta=-1.0)) / torch.sqrt(G[0, 0])
    stats[2 * terms + 1] = info
    return buf, stats


def _series_residual(a, mu, d, norm_y2):
    """residual(mu + d) from the moments a[k - 1] = a_k at mu (|d| < mu + sigma_min):
    a_1(mu + d) = sum_j (-d)^j a_{1+j},  a_2(mu + d) = sum_j (j + 1) (-d)^j a_{2+j},  residual = ||Y||^2 - a_1 - (mu + d) a_2."""
    n = len(a)
    s1 = s2 = 0.0
    p = 1.0
    for j in range(n):
        s1 += p * a[j]
        if j + 1 < n:
            s2 += (j + 1) * p * a[j + 1]
        p *= -d
    return norm_y2 - s1 - (mu + d) * s2


def _ridge_factor_chol(gamma, Tm, norm_y2, target, mu0, floor, max_evals=5, counters=None, terms=12):
    """The EPC mode update without an eigen-decomposition: the same mu and F = Tm (Gamma + mu I)^-1 as `_multiplier`
    gives.  residual(mu) = ||Y||^2 - a_1 - mu a_2 with a_k = tr(Tm M^-k Tm^T), M = Gamma + mu I, is analytic in mu:
    around a warm start mu0 (the multiplier of the previous mode update - it moves by a few per cent per update) both
    the residual and F(mu0 + d) = sum_j (-d)^j Tm M0^-(j+1) are power series in d that converge for |d| < mu0.  ONE
    Cholesky factorization + explicit inverse at mu0 and `terms` float64 GEMMs give the moments and the series terms
    (`_ridge_eval`); the root d of residual(mu0 + d) = target is then found on the host from the moments (bisection on
    scalars), and F is summed from the terms - against cuSOLVER's symmetric eigen-decomposition of the eigen form
    (syevd: 15 ms at R = 1141, the whole cost of an update).  |d| / mu0 <= 10^(-12 / terms) keeps the truncation of the F
    series at 1e-12 (the caller picks 8, 12 or 16 terms from the size of the previous step); a larger step moves mu0
    to the series' root and factorizes again (rare once the passes settle).
    residual is increasing in mu, which gives the direction when the root lies outside the series' trust interval.
    Returns (mu, F, |mu - mu0| / mu0) or None when the factorization fails (Gamma + mu I numerically singular), the explicit
    inverse is not accurate enough (||F_1 M - Tm|| > 1e-9 ||Tm||) or the budget is spent - the caller then takes the
    eigen form, which copes with a singular Gamma."""
    mu = mu_start = max(float(mu0), floor)
    for _ in range(max_evals):
        Fs, stats = _ridge_eval(gamma, Tm, mu, terms)
        st = stats.tolist()                                  # one device -> host copy per factorization
        a, chk, info = st[:-2], st[-2], st[-1]
        if counters is not None:
            counters["chol_evals"] = counters.get("chol_evals", 0) + 1
        if info != 0 or not all(math.isfinite(v) for v in st) or chk > 1e-9:
            return None
        accept = 10.0 ** (-12.0 / terms)
        g = lambda d: _series_residual(a, mu, d, norm_y2) - target
        dlo, dhi = max(-0.5 * mu, floor - mu), 0.5 * mu
        glo, ghi = g(dlo), g(dhi)
        slope = 2.0 * mu * a[2]                              # d residual / d mu at mu
        if ghi < 0.0:                                        # the root is above the trust interval
            step = -g(0.0) / slope if slope > 0.0 else 3.0 * mu
            mu += min(max(step, 0.5 * mu), 3.0 * mu)
            terms = max(terms, 12)
            continue
        if glo > 0.0:                                        # ... or below it
            if mu + dlo <= floor:                            # even the least-squares fit leaves more than the target
                if mu <= floor:
                    return floor, Fs[1], 0.0
                mu = floor
                continue
            step = -g(0.0) / slope if slope > 0.0 else -0.75 * mu
            mu = max(mu + min(max(step, -0.75 * mu), -0.5 * mu), floor)
            terms = max(terms, 12)
            continue
        for _b in range(200):                                # bisection on the series (scalars on the host)
            d = 0.5 * (dlo + dhi)
            if g(d) < 0.0:
                dlo = d
            else:
                dhi = d
            if dhi - dlo <= 1e-16 * mu:
                break
        d = 0.5 * (dlo + dhi)
        r = abs(d) / mu
        if r <= accept:
            if d == 0.0:
                return mu, Fs[1], abs(mu - mu_start) / mu_start
            coef = [0.0] + [(-d) ** k for k in range(terms)]                   # F(mu + d) = sum_k (-d)^k F_{k+1}: one GEMV
            F = torch.mv(Fs.view(terms + 1, -1).T, torch.tensor(coef, dtype=Fs.dtype, device=Fs.device)).view(Fs.shape[1:])
            return mu + d, F, abs(mu + d - mu_start) / mu_start
        mu += d                                              # the series' root is good to ~r^(2 terms): next time |d| is tiny
        terms = 8 if r <= 0.25 else 12                       # the remaining step is ~r^(2 terms) of mu
    return None


def _epc_sweep(T, factors, delta, state=None):
    """One error-preserving-correction pass over all modes (the body of musco's `cp_anc`); `factors` are float64 CUDA
    tensors owned by the caller and updated in place / replaced.  `state` (a dict the caller keeps across passes)
    carries the last multiplier: with it a mode update is the Cholesky form (`_ridge_factor_chol`), without it - the
    first update of a run, a stand-alone pass, or whenever the Cholesky form gives up - the eigen form."""
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
        got = None
        if state is not None and state.get("mu") is not None:
            # warm start: the last multiplier times the ratio the same two consecutive updates had one pass ago (the
            # multipliers of the modes differ systematically; the prediction is good to ~1e-3 once the passes settle);
            # series length from the size of the last step: 10^(-12 / terms) stays above it with a margin
            h = state["hist"]
            mu0 = h[-1] * (h[-T.N] / h[-T.N - 1]) if len(h) > T.N else h[-1]
            r_last = state.get("r", 1.0)
            terms = 4 if r_last <= 4e-4 else 6 if r_last <= 4e-3 else 8 if r_last <= 0.015 else 12 if r_last <= 0.05 else 16
            got = _ridge_factor_chol(gamma, Tm, T.norm2, target, mu0, state["sig_max"] * 1e-14, counters=state, terms=terms)
        if got is None:
            sig, V = torch.linalg.eigh(gamma)
            sig = torch.clamp(sig, min=0.0)
            Tt = Tm @ V
            s = (Tt * Tt).sum(dim=0)
            both = torch.stack([sig, s]).cpu().numpy()  # one device -> host copy per mode update
            mu = _multiplier(both[0], both[1], T.norm2, target)
            factors[m] = ((Tt / (sig + mu)) @ V.T).contiguous()
            if state is not None:
                state["sig_max"] = float(both[0].max())
                state["eigh_updates"] = state.get("eigh_updates", 0) + 1
        else:
            mu, F, state["r"] = got
            factors[m] = F.contiguous()
        if state is not None:
            state["mu"] = mu
            state.setdefault("hist", []).append(mu)
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
    state = {}                                                              # last multiplier (warm start of the next update)
    lam = _intensities(factors)
    for _ in range(epc_rounds):                                             # :61
        prev = None
        rounds += 1
        for _it in range(epc_maxiter):                                      # cp_anc(maxiter, tol)  :63
            factors = _epc_sweep(T, factors, delta, state)
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
                    epc_passes=passes, epc_rounds=rounds, epc_chol_evals=state.get("chol_evals", 0),
                    epc_eigh_updates=state.get("eigh_updates", 0))
    inv = np.argsort(order)
    return lam.to(tensor.device), [factors[int(i)].to(tensor.device) for i in inv]   # :77-82
