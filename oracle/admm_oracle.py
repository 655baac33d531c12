"""CPU oracle for the quantization-aware CP factorization hot path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module,
and there only as the checker (or as the timed CPU baseline), never as part of the
shipped path.  The product (`admm-quantization_b200/`) never imports `oracle/`.

Parity status
-------------
* ADMM inner loop, projections, Gram/MTTKRP/error, outer loop: **pinned** against the
  unmodified reference run in the dev container (`oracle/make_golden.py` ->
  `tests/golden/*.npz`, `tests/test_oracle_pinned.py`).
* ALS (`tensorly.parafac` 0.4.5) and EPC (`musco ... cp_anc` 1.0.6) initialisation:
  **parity unpinned** - both packages are absent from /root/reference and from this
  image, so `als_fp64` / `epc_fp64` restate the published algorithms
  and are covered by property tests only.

Each function cites the reference file:line (relative to the reference repo root) it
restates.  Arithmetic is float32 torch-CPU, single-threaded, because that is what
the reference executes; reductions therefore follow ATen's order on the host that
runs the oracle.  `sum_mode="exact"` replaces the one order-dependent reduction of
the projection (the per-candidate MSE) by a correctly rounded sum - that is the
semantics the CUDA kernel implements (see DESIGN.md "projection").
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

QSCHEMES = ("tensor_mseminmax_symmetric", "tensor_minmax", "tensor_symmetric", "tensor_affine")


# --------------------------------------------------------------------------- rank rule
def rank_for(shape: Sequence[int], reduction_rate: float) -> int:
    """scripts/factorize.py:157-158 - `int(numel / sum(shape) / reduction_rate)`."""
    numel = 1
    for d in shape:
        numel *= int(d)
    return int(numel / sum(int(d) for d in shape) / reduction_rate)


def conv_weight_to_tensor(w: torch.Tensor) -> torch.Tensor:
    """Intended reshape of scripts/factorize.py:138-145 (commented block) and
    scripts/calibrate.py:178-184: conv (Cout,Cin,kh,kw) -> (Cout,Cin,kh*kw); 1x1 -> (Cout,Cin)."""
    if w.ndim == 4:
        if w.shape[2] == 1 and w.shape[3] == 1:
            return w.reshape(w.shape[0], w.shape[1])
        return w.reshape(w.shape[0], w.shape[1], -1)
    return w


# --------------------------------------------------------------------------- projection
def candidate_grid(mx: float, num_attempts: int) -> np.ndarray:
    """source/quantization.py:129-131: `torch.linspace(0.2*mx.item(), 1.2*mx.item(), n)`
    (float32, built on the CPU).  ATen evaluates it as a fused multiply-add from `start`
    for the first half and from `end` for the second half; the float64 expression below is
    exact for these magnitudes (34-bit product + 24-bit addend) so one rounding remains."""
    mx32 = np.float32(mx)
    start = np.float32(0.2 * float(mx32))
    end = np.float32(1.2 * float(mx32))
    n = int(num_attempts)
    if n == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((end - start) / np.float32(n - 1))
    idx = np.arange(n, dtype=np.float64)
    lo = np.float64(step) * idx + np.float64(start)
    hi = np.float64(end) - np.float64(step) * (np.float64(n - 1) - idx)
    out = np.where(idx < n // 2, lo, hi).astype(np.float32)
    return out


def _levels(bits: int) -> Tuple[int, int]:
    q = 2 ** (bits - 1)
    return q, 2 * q - 1


def project_mse(x: torch.Tensor, bits: int, num_attempts: int = 200, sum_mode: str = "aten"):
    """source/quantization.py:118-144 `quantize_tensor_mse`.

    Returns (xq, codes int8, scale float32, best_index, mses float32[num_attempts]).
    sum_mode: "aten"  - per-candidate `mean` exactly as the reference (ATen float32 order);
              "exact" - float64 sum of the float32 squares, rounded to float32, then / N
                        in float32 (the CUDA kernel's definition)."""
    assert x.dtype == torch.float32
    q, denom = _levels(bits)
    mx = torch.max(torch.abs(x.min()), torch.abs(x.max()))          # :129
    grid = torch.from_numpy(candidate_grid(mx.item(), num_attempts))  # :130-131
    mses = torch.empty(num_attempts, dtype=torch.float32)
    n32 = np.float32(x.numel())
    for i in range(num_attempts):                                    # :136-139
        scale = 2 * grid[i] / denom                                   # :125
        xq = torch.clamp(torch.round(x / scale), -q, q - 1) * scale   # :127
        sq = (x - xq) ** 2
        if sum_mode == "aten":
            mses[i] = sq.mean()
        else:
            tot = np.float32(sq.double().sum().item())
            mses[i] = float(np.float32(tot / n32))
    best = int(torch.argmin(mses).item())                            # :141 (first minimum)
    scale = 2 * grid[best] / denom
    codes_f = torch.clamp(torch.round(x / scale), -q, q - 1)         # :144
    return codes_f * scale, codes_f.to(torch.int8), float(scale), best, mses


def project_minmax(x: torch.Tensor, bits: int) -> torch.Tensor:
    """source/quantization.py:48-66 `min_max_quantize`."""
    assert bits >= 1, bits
    if bits == 1:
        return torch.sign(x) - 1
    lo, hi = x.min(), x.max()
    unit = (x - lo) / (hi - lo)
    n = math.pow(2.0, bits) - 1
    k = torch.floor(unit * n + 0.5)
    return k * (hi - lo) / n + lo


def project_symmetric(x: torch.Tensor, bits: int) -> torch.Tensor:
    """source/quantization.py:91-95 (`tensor_symmetric`)."""
    q, denom = _levels(bits)
    tmax, tmin = x.max(), x.min()
    scale = 2 * torch.where(tmin.abs() > tmax, tmin.abs(), tmax) / denom
    return torch.clamp(torch.round(x / scale), -q, q - 1).to(int) * scale


def project_affine(x: torch.Tensor, bits: int, tmin=None, tmax=None) -> torch.Tensor:
    """source/quantization.py:97-106 (`tensor_affine`)."""
    q, denom = _levels(bits)
    if tmin is None or tmax is None:
        tmax, tmin = x.max(), x.min()
    scale = (tmax - tmin) / denom
    zp = (-q - (tmin / scale).int()).int()
    zp = torch.clamp(zp, -q, q - 1)
    return (torch.clamp(torch.round(x / scale) + zp, -q, q - 1).to(int) - zp) * scale


def project(x: torch.Tensor, bits: int, qscheme: str, sum_mode: str = "aten", **kw) -> torch.Tensor:
    """source/quantization.py:69-115 `quantize_tensor` dispatch (tensor_* schemes only:
    the channel_* schemes cannot be reached from the solver)."""
    if qscheme == "tensor_mseminmax_symmetric":
        return project_mse(x, bits, kw.get("num_attempts", 200), sum_mode)[0]
    if qscheme == "tensor_minmax":
        return project_minmax(x, bits)
    if qscheme == "tensor_symmetric":
        return project_symmetric(x, bits)
    if qscheme == "tensor_affine":
        return project_affine(x, bits, kw.get("tmin"), kw.get("tmax"))
    raise NotImplementedError(qscheme)


# --------------------------------------------------------------------------- ADMM inner loop
def ridge_rho(G: torch.Tensor) -> torch.Tensor:
    """source/admm.py:52-53: rho = trace(G)/R (ATen's CPU trace accumulates in double)."""
    return torch.trace(G) / G.shape[0]


def admm_iteration(H, U, F, G, max_iter, eps, bits, qscheme, sum_mode="aten", trace=None,
                   num_attempts=200):
    """source/admm.py:51-67.  Returns (H, U, iterations_done); mutates U in place like the
    reference.  `trace`, if a list, receives per-iteration dicts (V, H, scale index...)."""
    R = H.shape[1]
    rho = ridge_rho(G)
    L = torch.linalg.cholesky(G + rho * torch.eye(R), upper=False)          # :54
    done = 0
    for _ in range(1, max_iter):                                             # :55
        rhs = F + rho * (H + U)
        Hls = torch.cholesky_solve(rhs.T, L, upper=False).T                  # :56-57
        V = Hls - U
        H_prev = H
        if qscheme == "tensor_mseminmax_symmetric":
            Hq, codes, scale, best, _ = project_mse(V, bits, num_attempts, sum_mode)
        else:
            Hq, codes, scale, best = project(V, bits, qscheme), None, None, None
        H = Hq                                                               # :59
        U += H - Hls                                                         # :60
        r = torch.sum((H - Hls) ** 2) / torch.sum(H ** 2)                    # :62
        s = torch.sum((H - H_prev) ** 2) / torch.sum(U ** 2)                 # :63
        done += 1
        if trace is not None:
            trace.append(dict(V=V.clone(), Hls=Hls.clone(), H=H.clone(), U=U.clone(), codes=codes,
                              scale=scale, best=best, r=float(r), s=float(s)))
        if r < eps and s < eps:                                              # :64-65
            break
    return H, U, done


# --------------------------------------------------------------------------- per-sweep contractions
def gram_hadamard(others: Sequence[torch.Tensor]) -> torch.Tensor:
    """scripts/factorize.py:215,226,236 (3-D) and :276,286 (2-D)."""
    G = others[0].T @ others[0]
    for M in others[1:]:
        G = G * (M.T @ M)
    return G


def mttkrp(W: torch.Tensor, factors: Sequence[torch.Tensor], mode: int) -> torch.Tensor:
    """scripts/factorize.py:217,227,237 (einsum MTTKRP) and :277,287 (matrix case)."""
    if W.ndim == 2:
        return W @ factors[1] if mode == 0 else W.T @ factors[0]
    A, B, C = factors
    if mode == 0:
        return torch.einsum("abc,cr,br->ar", W, C, B)
    if mode == 1:
        return torch.einsum("abc,cr,ar->br", W, C, A)
    return torch.einsum("abc,br,ar->cr", W, B, A)


def reconstruct(factors: Sequence[torch.Tensor]) -> torch.Tensor:
    if len(factors) == 2:
        return factors[0] @ factors[1].T
    return torch.einsum("ir,jr,kr->ijk", *factors)


def rel_error(X: torch.Tensor, Y: torch.Tensor) -> float:
    """source/admm.py:14-15 `squared_relative_diff` (it is the *root* of the ratio)."""
    return torch.sqrt(torch.sum((X - Y) ** 2) / torch.sum(X ** 2)).item()


def init_random(shape: Sequence[int], rank: int, seed: int) -> List[torch.Tensor]:
    """source/admm.py:22-28 with device=None: one CPU generator, modes drawn in order."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    return [torch.randn(int(d), rank, generator=gen) for d in shape]


def factorize(W: torch.Tensor, factors: Sequence[torch.Tensor], bits: int, qscheme: str,
              max_iter_als: int, max_iter_admm: int, eps: float = 1e-8, tol: float = 1e-5,
              init_is_random: bool = True, sum_mode: str = "aten", stop_rules: bool = True,
              on_mode=None):
    """Outer AO-ADMM loop, scripts/factorize.py:192-310 restated for N = 2 or 3 modes.

    Returns (factors, factors_requantized, loss_hist, loss_quant_hist, duals)."""
    N = W.ndim
    fac = [f.clone() for f in factors]
    duals = [torch.zeros_like(f) for f in fac]                              # :209-212 / :272-273
    facq = [None] * N
    loss, lossq = [], []
    if not init_is_random:                                                   # :192-201
        fq0 = [project(f, bits, qscheme, sum_mode) for f in fac]
        loss.append(rel_error(W, reconstruct(fac)))
        lossq.append(rel_error(W, reconstruct(fq0)))
    back = 5 if N == 3 else 10                                               # :262 vs :306
    for sweep in range(max_iter_als):
        for m in range(N):
            others = [fac[k] for k in range(N) if k != m]
            G = gram_hadamard(others)
            Fm = mttkrp(W, fac, m)
            fac[m], duals[m], _ = admm_iteration(fac[m], duals[m], Fm, G, max_iter_admm, eps, bits,
                                                 qscheme, sum_mode)
            facq[m] = project(fac[m], bits, qscheme, sum_mode)               # :222,232,242
            if on_mode is not None:
                on_mode(sweep, m, fac, duals, Fm, G)
        loss.append(rel_error(W, reconstruct(fac)))
        lossq.append(rel_error(W, reconstruct(facq)))
        if stop_rules:
            if len(loss) > 1 and abs(loss[-2] - loss[-1]) < tol:             # :259
                break
            if len(loss) > 10 and loss[-1] - loss[-back] > 1e-3:             # :262 / :306
                break
    return fac, facq, loss, lossq, duals


# --------------------------------------------------------------------------- ALS + EPC (parity unpinned)
def _unfold(T: torch.Tensor, mode: int) -> torch.Tensor:
    """source/utils.py:60-74."""
    return torch.reshape(torch.moveaxis(T, mode, 0), (T.shape[mode], -1))


def _khatri_rao(mats: Sequence[torch.Tensor]) -> torch.Tensor:
    out = mats[0]
    for M in mats[1:]:
        out = (out[:, None, :] * M[None, :, :]).reshape(-1, out.shape[1])
    return out


def _mttkrp_nd(T: torch.Tensor, factors: Sequence[torch.Tensor], mode: int) -> torch.Tensor:
    return _unfold(T, mode) @ _khatri_rao([f for k, f in enumerate(factors) if k != mode])


def als_fp64(Y: torch.Tensor, rank: int, n_iter_max: int, tol: float, rng: np.random.RandomState,
             normalize: bool = True):
    """tensorly 0.4.5 `parafac` as called at source/parafac_epc.py:42-43 (restated from the
    published algorithm, SURVEY App. B.1 - PARITY UNPINNED): uniform-random init, per mode
    normal-equation solve (no ridge), optional column normalisation into `weights`,
    stop on |delta rec_error| < tol."""
    N = Y.ndim
    factors = [torch.from_numpy(rng.random_sample((Y.shape[m], rank))).to(Y.dtype) for m in range(N)]
    if normalize:
        factors = [f / (torch.linalg.norm(f, dim=0) + 1e-12) for f in factors]
    weights = torch.ones(rank, dtype=Y.dtype)
    normY = torch.linalg.norm(Y)
    errs: List[float] = []
    for it in range(n_iter_max):
        for m in range(N):
            gram = torch.ones(rank, rank, dtype=Y.dtype)
            for k in range(N):
                if k != m:
                    gram = gram * (factors[k].T @ factors[k])
            mt = _mttkrp_nd(Y, factors, m)
            f = torch.linalg.solve(gram.T, mt.T).T
            if normalize:
                weights = torch.linalg.norm(f, dim=0)
                weights = torch.where(weights <= torch.finfo(Y.dtype).eps, torch.ones_like(weights), weights)
                f = f / weights
            factors[m] = f
        if tol:
            gram_all = torch.ones(rank, rank, dtype=Y.dtype)
            for k in range(N):
                gram_all = gram_all * (factors[k].T @ factors[k])
            norm_rec2 = (weights[:, None] * weights[None, :] * gram_all).sum()
            inner = (weights * (mt * factors[N - 1]).sum(dim=0)).sum()
            err = math.sqrt(abs(float(normY ** 2 + norm_rec2 - 2 * inner))) / float(normY)
            errs.append(err)
            if it >= 1 and abs(errs[-2] - errs[-1]) < tol:
                break
    return weights, factors, errs


def epc_sweep_fp64(Y: torch.Tensor, factors: List[torch.Tensor], delta: float) -> List[torch.Tensor]:
    """One EPC pass over all modes (Phan, Tichavsky, Cichocki 2019; SURVEY App. B.2 -
    PARITY UNPINNED): minimise sum_r prod_n ||u_r^(n)||^2 subject to ||Y - Yhat|| <= delta."""
    N = Y.ndim
    rank = factors[0].shape[1]
    normY2 = float(torch.sum(Y * Y))
    for m in range(N):
        scale = torch.ones(rank, dtype=Y.dtype)
        for k in range(N):
            if k != m:
                nk = torch.linalg.norm(factors[k], dim=0)
                nk = torch.where(nk == 0, torch.ones_like(nk), nk)
                factors[k] = factors[k] / nk
                scale = scale * nk
        factors[m] = factors[m] * scale
        gamma = torch.ones(rank, rank, dtype=Y.dtype)
        for k in range(N):
            if k != m:
                gamma = gamma * (factors[k].T @ factors[k])
        T = _mttkrp_nd(Y, factors, m)
        sig, V = torch.linalg.eigh(gamma)
        sig = torch.clamp(sig, min=0.0)
        Tt = T @ V
        s = (Tt * Tt).sum(dim=0)

        def resid(mu: float) -> float:
            return normY2 - float((s * (sig + 2 * mu) / (sig + mu) ** 2).sum())

        target = delta * delta
        floor = float(sig.max()) * 1e-14
        mu = 0.0
        if resid(floor) < target:
            lo, hi = floor, max(float(sig.max()), 1e-300)
            while resid(hi) < target:
                hi *= 2.0
                if hi > 1e300:
                    break
            for _ in range(200):
                mid = 0.5 * (lo + hi)
                if resid(mid) < target:
                    lo = mid
                else:
                    hi = mid
                if hi - lo <= 1e-15 * hi:
                    break
            mu = 0.5 * (lo + hi)
        factors[m] = (Tt / (sig + max(mu, floor))) @ V.T
    return factors


def parafac_epc_fp64(tensor: torch.Tensor, rank: int, als_maxiter=5000, als_tol=1e-5, epc_maxiter=5000,
                     epc_rounds=50, epc_tol=1e-5, stop_tol=1e-4, ratio_tol=1e-3, ratio_max_iters=10,
                     rng: Optional[np.random.RandomState] = None, info: Optional[dict] = None):
    """source/parafac_epc.py:12-82 wrapper logic around the two restated routines."""
    rng = rng if rng is not None else np.random.mtrand._rand
    Y = tensor.to(torch.float64)
    order = np.argsort(Y.shape)                                             # :38
    Yp = Y.permute(tuple(int(o) for o in order))                            # :40
    weights, factors, _ = als_fp64(Yp, rank, als_maxiter, als_tol, rng)     # :42-43
    rec = torch.einsum("r," + ",".join(f"{chr(105 + k)}r" for k in range(Y.ndim)) + "->" +
                       "".join(chr(105 + k) for k in range(Y.ndim)), weights, *factors)
    delta = float(torch.linalg.norm(Yp - rec))                              # :51
    factors[-1] = factors[-1] * weights                                     # :53
    if info is not None:
        info.update(delta=delta, norm=float(torch.linalg.norm(Yp)), als_intensity2=float((weights ** 2).sum()))

    def intensities(fs):
        lam = torch.ones(rank, dtype=torch.float64)
        for f in fs:
            lam = lam * torch.linalg.norm(f, dim=0)
        return lam

    lam_prev_norm = float(torch.linalg.norm(weights))                       # :52
    alpha_prev = float(weights.max() / weights.min())                       # :57
    stopflag = 0
    for _ in range(epc_rounds):                                             # :61
        prev = None
        for _it in range(epc_maxiter):                                      # cp_anc(maxiter, tol)
            factors = epc_sweep_fp64(Yp, factors, delta)
            cur = float((intensities(factors) ** 2).sum())
            if prev is not None and abs(prev - cur) < epc_tol * prev:
                break
            prev = cur
        lam = intensities(factors)
        lam_norm = float(torch.linalg.norm(lam))
        alpha = float(lam.max() / lam.min())
        if abs(lam_prev_norm - lam_norm) < stop_tol * lam_prev_norm:        # :67
            break
        stopflag = stopflag + 1 if abs(alpha_prev - alpha) < ratio_tol else 0   # :69
        lam_prev_norm, alpha_prev = lam_norm, alpha
        if stopflag >= ratio_max_iters:                                      # :74
            break
    inv = np.argsort(order)
    return intensities(factors), [factors[int(i)] for i in inv]             # :77-82 (original mode order)
