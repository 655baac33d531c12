"""TEST INFRASTRUCTURE - CPU restatement of the two-block splitting of the reference's scripts/factorize_lowrank.py.

Only tests/ may import this module.  `admm_iteration` restates scripts/factorize_lowrank.py:84-101 operation by
operation (float32 torch-CPU, single thread), `project_rank` :80-82.  The reference script itself cannot be imported
offline (it imports transformers' model loader and bitsandbytes at module level and needs a Hugging Face checkpoint),
so this restatement is pinned only through the projection it calls (oracle.admm_oracle.project, pinned against the
reference's quantize_tensor) - the loop around it is eight lines of elementwise torch.
"""
import torch

from . import admm_oracle as orc


def project_rank(H, rank):
    U, S, Vt = torch.linalg.svd(H)
    return U[:, :rank] @ torch.diag(S[:rank]) @ Vt[:rank]


def admm_iteration(H, U, W, H2, proj_func, rho=1.0, max_iter=50, eps=1e-8, trace=None):
    for j in range(1, max_iter):
        H_ = (rho * (H + U) + W - H2) / (1 + rho)
        H_prev = H.clone()
        H = proj_func(H_ - U)
        U += H - H_
        r = torch.sum((H - H_) ** 2) / torch.sum(H ** 2)
        s = torch.sum((H - H_prev) ** 2) / torch.sum(U ** 2)
        if trace is not None:
            trace.append((H.clone(), U.clone(), float(r), float(s)))
        if r < eps and s < eps:
            break
    return H, U


def quantize_func(bits, qscheme):
    return lambda x: orc.project(x, bits, qscheme)
