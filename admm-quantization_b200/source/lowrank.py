"""Two-block ADMM splitting W ~ W_q + W_r (quantized + low-rank), the solver of the reference's
scripts/factorize_lowrank.py:80-170 (the notebook behind BASELINE config 5).

Each outer iteration runs two inner ADMM loops (:152-154): the quantized block (projection = quantize_tensor) in one
persistent kernel of libadmmq (admmq_split_loop: the least-squares step is elementwise), and the low-rank block
(projection = SVD truncation, :80-82), whose 49 singular value decompositions per call stay with torch.linalg.svd.
"""
import torch

from . import _native


def project_rank(H, rank):
    """reference scripts/factorize_lowrank.py:80-82."""
    U, S, Vt = torch.linalg.svd(H)
    return U[:, :rank] @ torch.diag(S[:rank]) @ Vt[:rank]


def admm_iteration_quantized(H, U, W, H2, bits, qscheme, rho=1.0, max_iter=50, eps=1e-8, num_attempts=200, max_ctas=0):
    """`admm_iteration(H, U, W, H2, quantize_func, rho, max_iter, eps)` of the reference (:84-101) on the GPU kernel.
    Returns a NEW H; U is updated in place and returned, like the reference."""
    _native.require_cuda(H, U, W, H2)
    Hc = _native.f32c(H).clone()
    Uc = U if (U.dtype == torch.float32 and U.is_contiguous()) else _native.f32c(U).clone()
    rep = _native.split_loop_inplace(Hc, Uc, _native.f32c(W), _native.f32c(H2), rho, max_iter, eps, bits, qscheme,
                                     num_attempts, max_ctas)
    if Uc is not U:
        U.copy_(Uc)
    return Hc, U, rep


def admm_iteration_projected(H, U, W, H2, proj_func, rho=1.0, max_iter=50, eps=1e-8):
    """The same inner loop for an arbitrary projection (used for the low-rank block): torch elementwise ops in the
    reference's order (:84-101)."""
    for _ in range(1, max_iter):
        H_ = (rho * (H + U) + W - H2) / (1 + rho)
        H_prev = H.clone()
        H = proj_func(H_ - U)
        U += H - H_
        r = torch.sum((H - H_) ** 2) / torch.sum(H ** 2)
        s = torch.sum((H - H_prev) ** 2) / torch.sum(U ** 2)
        if r < eps and s < eps:
            break
    return H, U


def factorize_lowrank(W, bits, rank, qscheme="tensor_minmax", max_iter=100, seed=42, rho=1.0, inner_max_iter=50,
                      eps=1e-8, log=None):
    """Outer loop of scripts/factorize_lowrank.py:130-170: random initial W_q, random rank-projected W_r, alternate the
    two blocks, stop when the relative error jumps (:168).  Returns (W_q, W_r, history of rel_admm_diff)."""
    _native.require_cuda(W)
    W = _native.f32c(W)
    dev = W.device
    torch.manual_seed(seed)                                               # set_seed (:15-18)
    W_q = torch.randn(*W.shape, device=dev)                               # :138
    U_q = torch.zeros_like(W_q)
    W_r = project_rank(torch.randn(*W.shape, device=dev), rank)           # :144-145
    U_r = torch.zeros_like(W_r)
    proj = lambda X: project_rank(X, rank)
    hist, prev = [], None
    norm_w = torch.linalg.norm(W)
    for i in range(max_iter):
        W_q, U_q, _ = admm_iteration_quantized(W_q, U_q, W, W_r, bits, qscheme, rho=rho, max_iter=inner_max_iter, eps=eps)
        W_r, U_r = admm_iteration_projected(W_r, U_r, W, W_q, proj, rho=rho, max_iter=inner_max_iter, eps=eps)
        rel = float(torch.linalg.norm(W - W_r - W_q) / norm_w)            # :160-161
        if log is not None and i % 10 == 0:
            log(i, rel)
        if prev is not None and prev and prev < rel - 1:                  # :168
            hist.append(rel)
            break
        hist.append(rel)
        prev = rel
    return W_q, W_r, hist
