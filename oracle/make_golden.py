"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden.py [--out tests/golden]

/root/reference does not exist on the GPU box, so its outputs are committed here as small
fixtures together with this script.  Everything is single-threaded float32 torch-CPU with
fixed seeds; inputs are stored next to outputs so no RNG has to be reproduced elsewhere.
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_import import import_reference, REFERENCE_ROOT  # noqa: E402

MSE = "tensor_mseminmax_symmetric"


def run_mse_with_internals(ref, x, bits, num_attempts):
    """Call the reference projection, then recover (index, scale, codes) by re-evaluating the
    reference's own expressions for the chosen candidate (source/quantization.py:125-144)."""
    xq = ref.quantize_tensor(x, bits, MSE, num_attempts=num_attempts)
    q = 2 ** (bits - 1)
    denom = 2 * q - 1
    mx = torch.max(torch.abs(x.min()), torch.abs(x.max()))
    grid = torch.linspace(0.2 * mx.item(), 1.2 * mx.item(), num_attempts)
    found = None
    for i in range(num_attempts):
        scale = 2 * grid[i] / denom
        cand = torch.clamp(torch.round(x / scale), -q, q - 1) * scale
        if torch.equal(cand, xq) or (torch.isnan(xq).all() and torch.isnan(cand).all()):
            found = i
            break
    assert found is not None
    scale = 2 * grid[found] / denom
    codes = torch.clamp(torch.round(x / scale), -q, q - 1)
    return xq, found, float(scale), codes.to(torch.int8)


def projection_cases(ref):
    g = torch.Generator().manual_seed(1234)
    cases = []

    def add(name, x, bits, n=200):
        cases.append((name, x.contiguous().float(), bits, n))

    for bits in (2, 3, 4, 6, 8):
        add(f"gauss64x134_b{bits}", torch.randn(64, 134, generator=g), bits)
    add("gauss9x134_b4", torch.randn(9, 134, generator=g), 4)
    add("gauss128x278_b4", torch.randn(128, 278, generator=g) * 0.05, 4)
    add("gauss256x566_b4", torch.randn(256, 566, generator=g) * 3.0, 4)
    add("gauss64x134_b4_n1000", torch.randn(64, 134, generator=g), 4, 1000)
    add("gauss64x134_b4_n7", torch.randn(64, 134, generator=g), 4, 7)
    add("uniform_b4", torch.rand(40, 77, generator=g) - 0.5, 4)
    add("positive_only_b4", torch.rand(33, 65, generator=g) + 0.1, 4)
    add("negative_only_b3", -torch.rand(33, 65, generator=g) - 0.1, 3)
    x = torch.randn(64, 134, generator=g)
    x[7, 11] = 25.0
    add("one_outlier_b4", x, 4)
    add("constant_0p3_b4", torch.full((16, 16), 0.3), 4)
    add("ties_b4", torch.tensor([[0.5, 1.5, 2.5, -0.5, -7.5, 7.0, 3.5, -2.5]]), 4)
    add("single_element_b4", torch.tensor([[1.7]]), 4)
    add("ragged_1x513_b4", torch.randn(1, 513, generator=g), 4)
    add("ragged_257x3_b6", torch.randn(257, 3, generator=g), 6)
    add("tiny_values_b4", torch.randn(32, 32, generator=g) * 1e-20, 4)
    add("huge_values_b4", torch.randn(32, 32, generator=g) * 1e18, 4)
    add("all_zero_b4", torch.zeros(8, 8), 4)
    # grid-valued input (re-projection is not idempotent)
    y = ref.quantize_tensor(torch.randn(64, 134, generator=g), 4, MSE)
    add("regrid_b4", y, 4)
    out = {}
    meta = []
    for name, x, bits, n in cases:
        xq, idx, scale, codes = run_mse_with_internals(ref, x, bits, n)
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/xq"] = xq.numpy()
        out[f"{name}/codes"] = codes.numpy()
        out[f"{name}/idx_scale"] = np.array([idx, scale], dtype=np.float64)
        meta.append(dict(name=name, bits=bits, num_attempts=n, shape=list(x.shape)))
    # the other tensor_* schemes (source/quantization.py:48-66, 91-106)
    for scheme in ("tensor_minmax", "tensor_symmetric", "tensor_affine"):
        for bits in (1, 2, 4, 8) if scheme == "tensor_minmax" else (2, 4, 8):
            x = torch.randn(48, 100, generator=g) * 0.7 + 0.1
            y = ref.quantize_tensor(x, bits, scheme)
            name = f"{scheme}_b{bits}"
            out[f"{name}/x"] = x.numpy()
            out[f"{name}/xq"] = y.float().numpy()
            meta.append(dict(name=name, bits=bits, scheme=scheme, shape=list(x.shape)))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def contraction_cases(ref):
    g = torch.Generator().manual_seed(77)
    out = {}
    meta = []
    for name, (I, J, K, R) in dict(c3_small=(12, 10, 9, 17), c3_l1=(64, 64, 9, 134), c3_odd=(33, 20, 49, 40)).items():
        W = torch.randn(I, J, K, generator=g) * 0.06
        A, B, C = (torch.randn(d, R, generator=g) for d in (I, J, K))
        out[f"{name}/W"], out[f"{name}/A"], out[f"{name}/B"], out[f"{name}/C"] = (t.numpy() for t in (W, A, B, C))
        out[f"{name}/G0"] = (B.T @ B * (C.T @ C)).numpy()
        out[f"{name}/G1"] = (A.T @ A * (C.T @ C)).numpy()
        out[f"{name}/G2"] = (A.T @ A * (B.T @ B)).numpy()
        out[f"{name}/F0"] = torch.einsum("abc,cr,br->ar", W, C, B).numpy()
        out[f"{name}/F1"] = torch.einsum("abc,cr,ar->br", W, C, A).numpy()
        out[f"{name}/F2"] = torch.einsum("abc,br,ar->cr", W, B, A).numpy()
        out[f"{name}/err"] = np.array([ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", A, B, C))])
        meta.append(dict(name=name, ndim=3, dims=[I, J, K], rank=R))
    for name, (I, J, R) in dict(c2_small=(40, 24, 15), c2_rect=(96, 200, 48)).items():
        W = torch.randn(I, J, generator=g) * 0.02
        A, B = torch.randn(I, R, generator=g), torch.randn(J, R, generator=g)
        out[f"{name}/W"], out[f"{name}/A"], out[f"{name}/B"] = W.numpy(), A.numpy(), B.numpy()
        out[f"{name}/G0"] = (B.T @ B).numpy()
        out[f"{name}/G1"] = (A.T @ A).numpy()
        out[f"{name}/F0"] = (W @ B).numpy()
        out[f"{name}/F1"] = (W.T @ A).numpy()
        out[f"{name}/err"] = np.array([ref.squared_relative_diff(W, A @ B.T)])
        meta.append(dict(name=name, ndim=2, dims=[I, J], rank=R))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def config1_weight():
    """BASELINE config 1 input: resnet18(weights=None) under seed 42,
    layer1.0.conv1 reshaped to (64, 64, 9)."""
    import torchvision
    torch.manual_seed(42)
    m = torchvision.models.resnet18(weights=None)
    return m.layer1[0].conv1.weight.detach().reshape(64, 64, 9).contiguous()


def admm_iteration_cases(ref):
    """One reference `admm_iteration` call per case, replayed one inner iteration at a time
    (`max_iter=2` runs exactly one iteration, source/admm.py:55) so every iterate is recorded."""
    out = {}
    meta = []
    W = config1_weight()
    R = 134
    specs = [
        dict(name="l1_mode0_b4", mode=0, bits=4, iters=120, qscheme=MSE),
        dict(name="l1_mode2_b4", mode=2, bits=4, iters=60, qscheme=MSE),
        dict(name="l1_mode1_b3", mode=1, bits=3, iters=40, qscheme=MSE),
        dict(name="l1_mode0_b8_minmax", mode=0, bits=8, iters=30, qscheme="tensor_minmax"),
    ]
    for sp in specs:
        fac = ref.init_factors(W, R, init="random", device=None, seed=42)
        A, B, C = fac
        m = sp["mode"]
        if m == 0:
            G = B.T @ B * (C.T @ C); F = torch.einsum("abc,cr,br->ar", W, C, B)
        elif m == 1:
            G = A.T @ A * (C.T @ C); F = torch.einsum("abc,cr,ar->br", W, C, A)
        else:
            G = A.T @ A * (B.T @ B); F = torch.einsum("abc,br,ar->cr", W, B, A)
        H = fac[m].clone()
        U = torch.zeros_like(H)
        name = sp["name"]
        out[f"{name}/H0"], out[f"{name}/F"], out[f"{name}/G"] = H.numpy().copy(), F.numpy(), G.numpy()
        Hs, Us = [], []
        for it in range(sp["iters"]):
            H, U = ref.admm_iteration(H, U, F, G, max_iter=2, eps=1e-8, bits=sp["bits"], qscheme=sp["qscheme"])
            Hs.append(H.numpy().copy())
            Us.append(U.numpy().copy())
        keep = sorted(set(list(range(0, 12)) + list(range(12, sp["iters"], 9)) + [sp["iters"] - 1]))
        out[f"{name}/keep"] = np.array(keep)
        out[f"{name}/H"] = np.stack([Hs[k] for k in keep])
        out[f"{name}/U"] = np.stack([Us[k] for k in keep])
        # every iterate's H as grid values is large; store all of them as float16-safe codes instead:
        # the grid value / min positive spacing is recovered in the tests from H itself.
        out[f"{name}/H_all_sum"] = np.array([float(np.abs(h).astype(np.float64).sum()) for h in Hs])
        # and one genuine multi-iteration call to pin the loop/in-place semantics
        fac2 = ref.init_factors(W, R, init="random", device=None, seed=42)
        H2 = fac2[m].clone(); U2 = torch.zeros_like(H2)
        Hn, Un = ref.admm_iteration(H2, U2, F, G, max_iter=sp["iters"] + 1, eps=1e-8, bits=sp["bits"],
                                    qscheme=sp["qscheme"])
        assert Un is U2
        assert np.array_equal(Hn.numpy(), Hs[-1]) and np.array_equal(Un.numpy(), Us[-1])
        meta.append(dict(name=name, bits=sp["bits"], qscheme=sp["qscheme"], iters=sp["iters"], mode=m, rank=R))
    # 2-D case, larger ridge system than rows
    g = torch.Generator().manual_seed(5)
    Wm = torch.randn(96, 40, generator=g) * 0.02
    Rm = 30
    A, B = torch.randn(96, Rm, generator=g), torch.randn(40, Rm, generator=g)
    G = B.T @ B; F = Wm @ B
    H = A.clone(); U = torch.zeros_like(H)
    out["mat_mode0_b4/H0"], out["mat_mode0_b4/F"], out["mat_mode0_b4/G"] = H.numpy().copy(), F.numpy(), G.numpy()
    Hs, Us = [], []
    for it in range(25):
        H, U = ref.admm_iteration(H, U, F, G, max_iter=2, eps=1e-8, bits=4, qscheme=MSE)
        Hs.append(H.numpy().copy()); Us.append(U.numpy().copy())
    out["mat_mode0_b4/keep"] = np.arange(25)
    out["mat_mode0_b4/H"] = np.stack(Hs); out["mat_mode0_b4/U"] = np.stack(Us)
    out["mat_mode0_b4/H_all_sum"] = np.array([float(np.abs(h).astype(np.float64).sum()) for h in Hs])
    meta.append(dict(name="mat_mode0_b4", bits=4, qscheme=MSE, iters=25, mode=0, rank=Rm))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def outer_loop_cases(ref):
    """scripts/factorize.py:207-310 replayed verbatim with the reference's functions."""
    out = {}
    meta = []
    W = config1_weight()
    out["config1/W"] = W.numpy()

    def run3(W, R, bits, qscheme, sweeps, max_iter_admm, seed):
        A, B, C = ref.init_factors(W, R, init="random", device=None, seed=seed)
        init = [A.numpy().copy(), B.numpy().copy(), C.numpy().copy()]
        U_A, U_B, U_C = torch.zeros_like(A), torch.zeros_like(B), torch.zeros_like(C)
        loss, lossq = [], []
        for _ in range(sweeps):
            G = B.T @ B * (C.T @ C)
            F = torch.einsum("abc,cr,br->ar", W, C, B)
            A, U_A = ref.admm_iteration(A, U_A, F, G, max_iter=max_iter_admm, eps=1e-8, bits=bits, qscheme=qscheme)
            Aq = ref.quantize_tensor(A, qscheme=qscheme, bits=bits)
            G = A.T @ A * (C.T @ C)
            F = torch.einsum("abc,cr,ar->br", W, C, A)
            B, U_B = ref.admm_iteration(B, U_B, F, G, max_iter=max_iter_admm, eps=1e-8, bits=bits, qscheme=qscheme)
            Bq = ref.quantize_tensor(B, qscheme=qscheme, bits=bits)
            G = A.T @ A * (B.T @ B)
            F = torch.einsum("abc,br,ar->cr", W, B, A)
            C, U_C = ref.admm_iteration(C, U_C, F, G, max_iter=max_iter_admm, eps=1e-8, bits=bits, qscheme=qscheme)
            Cq = ref.quantize_tensor(C, qscheme=qscheme, bits=bits)
            loss.append(ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", A, B, C)))
            lossq.append(ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", Aq, Bq, Cq)))
        return init, [A, B, C], [U_A, U_B, U_C], loss, lossq

    t0 = time.time()
    for name, kw in dict(config1_full=dict(sweeps=2, max_iter_admm=1000),
                         config1_short=dict(sweeps=6, max_iter_admm=60)).items():
        init, fac, duals, loss, lossq = run3(W, 134, 4, MSE, seed=42, **kw)
        for m in range(3):
            out[f"{name}/init{m}"] = init[m]
            out[f"{name}/fac{m}"] = fac[m].numpy()
            out[f"{name}/dual{m}"] = duals[m].numpy()
        out[f"{name}/loss"] = np.array(loss)
        out[f"{name}/lossq"] = np.array(lossq)
        meta.append(dict(name=name, W="config1", rank=134, bits=4, qscheme=MSE, seed=42, **kw))
        print(name, loss, f"{time.time() - t0:.0f}s", flush=True)
    # matrix branch (scripts/factorize.py:269-310)
    g = torch.Generator().manual_seed(9)
    Wm = torch.randn(128, 48, generator=g) * 0.02
    out["mat/W"] = Wm.numpy()
    R = 17
    A, B = ref.init_factors(Wm, R, init="random", device=None, seed=3)
    out["mat/init0"], out["mat/init1"] = A.numpy().copy(), B.numpy().copy()
    U_A, U_B = torch.zeros_like(A), torch.zeros_like(B)
    loss, lossq = [], []
    for _ in range(5):
        G = B.T @ B; F = Wm @ B
        A, U_A = ref.admm_iteration(A, U_A, F, G, max_iter=80, eps=1e-8, bits=4, qscheme=MSE)
        Aq = ref.quantize_tensor(A, qscheme=MSE, bits=4)
        G = A.T @ A; F = Wm.T @ A
        B, U_B = ref.admm_iteration(B, U_B, F, G, max_iter=80, eps=1e-8, bits=4, qscheme=MSE)
        Bq = ref.quantize_tensor(B, qscheme=MSE, bits=4)
        loss.append(ref.squared_relative_diff(Wm, A @ B.T))
        lossq.append(ref.squared_relative_diff(Wm, Aq @ Bq.T))
    out["mat/fac0"], out["mat/fac1"] = A.numpy(), B.numpy()
    out["mat/loss"], out["mat/lossq"] = np.array(loss), np.array(lossq)
    meta.append(dict(name="mat", rank=R, bits=4, qscheme=MSE, seed=3, sweeps=5, max_iter_admm=80))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def self_divergence_cases(ref):
    """How far the UNMODIFIED reference drifts from itself when every MTTKRP output is multiplied
    by (1 +- 6e-8) (half a float32 ulp): BASELINE config 1, full inner budget, 2 sweeps.  This is the
    yardstick for free-running error-history parity: an implementation
    whose float32 ridge solve is not bit-identical to LAPACK's cannot track closer than this."""
    W = config1_weight()
    out, meta = {}, []
    for trial, noise_seed in enumerate((101, 202)):
        gn = torch.Generator().manual_seed(noise_seed)
        A, B, C = ref.init_factors(W, 134, init="random", device=None, seed=42)
        U_A, U_B, U_C = torch.zeros_like(A), torch.zeros_like(B), torch.zeros_like(C)

        def jitter(F):
            sign = torch.randint(0, 2, F.shape, generator=gn).float() * 2 - 1
            return F * (1 + 6e-8 * sign)

        loss, lossq = [], []
        for _ in range(2):
            G = B.T @ B * (C.T @ C)
            F = jitter(torch.einsum("abc,cr,br->ar", W, C, B))
            A, U_A = ref.admm_iteration(A, U_A, F, G, max_iter=1000, eps=1e-8, bits=4, qscheme=MSE)
            Aq = ref.quantize_tensor(A, qscheme=MSE, bits=4)
            G = A.T @ A * (C.T @ C)
            F = jitter(torch.einsum("abc,cr,ar->br", W, C, A))
            B, U_B = ref.admm_iteration(B, U_B, F, G, max_iter=1000, eps=1e-8, bits=4, qscheme=MSE)
            Bq = ref.quantize_tensor(B, qscheme=MSE, bits=4)
            G = A.T @ A * (B.T @ B)
            F = jitter(torch.einsum("abc,br,ar->cr", W, B, A))
            C, U_C = ref.admm_iteration(C, U_C, F, G, max_iter=1000, eps=1e-8, bits=4, qscheme=MSE)
            Cq = ref.quantize_tensor(C, qscheme=MSE, bits=4)
            loss.append(ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", A, B, C)))
            lossq.append(ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", Aq, Bq, Cq)))
        out[f"trial{trial}/loss"], out[f"trial{trial}/lossq"] = np.array(loss), np.array(lossq)
        meta.append(dict(name=f"trial{trial}", noise_seed=noise_seed, eps=6e-8, sweeps=2, max_iter_admm=1000))
        print("self-divergence trial", trial, loss, flush=True)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


# ------------------------------------------------------------------ round 2: solve-level self divergence, long runs
def _codes_of(xq):
    """int8 codes and the scale of a grid-valued tensor produced by the reference's clip search (xq = codes * scale
    exactly in float32): the scale is the smallest positive |value| that reproduces every value, checked exactly."""
    v = xq.abs()
    pos = torch.unique(v[v > 0])
    for s in pos[:4]:
        codes = torch.round(xq / s)
        if torch.equal(codes * s, xq) and float(codes.abs().max()) <= 128:
            return codes.to(torch.int8), float(s)
    raise AssertionError("could not recover the codes of a grid-valued tensor")


class _Recorder:
    """Wraps `quantize_tensor` AS SEEN BY the reference's source/admm.py (the function itself stays unmodified): every
    call inside admm_iteration is one inner iteration, its argument is V = H_T - U (source/admm.py:59) and, at call time,
    the caller's U tensor still holds the scaled dual ENTERING the iteration (it is updated in place at :60)."""

    def __init__(self, ref):
        self.glob = ref.admm_iteration.__globals__
        self.orig = self.glob["quantize_tensor"]
        self.on_call = None

    def __enter__(self):
        def wrapped(x, *a, **k):
            y = self.orig(x, *a, **k)
            if self.on_call is not None:
                self.on_call(x, y)
            return y
        self.glob["quantize_tensor"] = wrapped
        return self

    def __exit__(self, *exc):
        self.glob["quantize_tensor"] = self.orig


def _sweep3(ref, W, fac, duals, bits, qscheme, max_iter_admm, on_mode=None, jitter=None):
    """One outer sweep of scripts/factorize.py:214-255 with the reference's functions; returns (error, quantized error)."""
    A, B, C = fac
    for m in range(3):
        A, B, C = fac
        if m == 0:
            G = B.T @ B * (C.T @ C); F = torch.einsum("abc,cr,br->ar", W, C, B)
        elif m == 1:
            G = A.T @ A * (C.T @ C); F = torch.einsum("abc,cr,ar->br", W, C, A)
        else:
            G = A.T @ A * (B.T @ B); F = torch.einsum("abc,br,ar->cr", W, B, A)
        if jitter is not None:
            F = jitter(F)
        if on_mode is not None:
            on_mode(m, fac[m], duals[m], F, G)
        fac[m], duals[m] = ref.admm_iteration(fac[m], duals[m], F, G, max_iter=max_iter_admm, eps=1e-8, bits=bits,
                                              qscheme=qscheme)
    q = [ref.quantize_tensor(f, qscheme=qscheme, bits=bits) for f in fac]
    return (ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", *fac)),
            ref.squared_relative_diff(W, torch.einsum("ir,jr,kr->ijk", *q)))


def solve_divergence_cases(ref):
    """How far the UNMODIFIED reference drifts from itself when the OUTPUT OF ITS RIDGE SOLVE (H_ of source/admm.py:56)
    is multiplied element-wise by (1 +- eps) in every inner iteration: eps = 3e-7 is the measured distance of LAPACK's
    own float32 potrs from the exact solution on this system, eps = 6e-8 is half a float32 ulp - the
    distance between any two correctly rounded solves.  BASELINE config 1, full inner budget, 2 sweeps.  The spread of
    rec_error over the trials is the tolerance an implementation that is not bit-identical to LAPACK can be held to."""
    W = config1_weight()
    out, meta = {}, []
    real_solve = torch.cholesky_solve
    trials = [(3e-7, s) for s in (11, 12, 13, 14, 15)] + [(6e-8, s) for s in (21, 22, 23)]
    try:
        for trial, (eps, noise_seed) in enumerate(trials):
            gn = torch.Generator().manual_seed(noise_seed)

            def jittered(b, L, upper=False):
                x = real_solve(b, L, upper=upper)
                sign = torch.randint(0, 2, x.shape, generator=gn).float() * 2 - 1
                return x * (1 + eps * sign)

            torch.cholesky_solve = jittered      # the reference looks the function up on the torch module at call time
            fac = list(ref.init_factors(W, 134, init="random", device=None, seed=42))
            duals = [torch.zeros_like(f) for f in fac]
            loss, lossq = [], []
            for _ in range(2):
                e, eq = _sweep3(ref, W, fac, duals, 4, MSE, 1000)
                loss.append(e); lossq.append(eq)
            out[f"trial{trial}/loss"], out[f"trial{trial}/lossq"] = np.array(loss), np.array(lossq)
            meta.append(dict(name=f"trial{trial}", noise_seed=noise_seed, eps=eps, sweeps=2, max_iter_admm=1000))
            print("solve-level self-divergence trial", trial, eps, loss, flush=True)
    finally:
        torch.cholesky_solve = real_solve
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def short_divergence_cases(ref):
    """The same yardstick for the SHORT-budget history the outer-loop test follows over six sweeps (config 1,
    max_iter_admm = 60, the `config1_short` case of outer_loop.npz): the unmodified reference with the output of its own
    ridge solve jittered by +-3e-7 (five trials) / +-6e-8 (three trials) per inner iteration.  The per-sweep spread of
    rec_error over the trials is what `test_outer_loop_against_reference_history` holds the CUDA path to from sweep 2 on
    (sweeps 0 and 1 are asserted at north_star's 1e-3)."""
    W = config1_weight()
    out, meta = {}, []
    real_solve = torch.cholesky_solve
    trials = [(3e-7, s) for s in (31, 32, 33, 34, 35)] + [(6e-8, s) for s in (41, 42, 43)]
    try:
        for trial, (eps, noise_seed) in enumerate(trials):
            gn = torch.Generator().manual_seed(noise_seed)

            def jittered(b, L, upper=False):
                x = real_solve(b, L, upper=upper)
                sign = torch.randint(0, 2, x.shape, generator=gn).float() * 2 - 1
                return x * (1 + eps * sign)

            torch.cholesky_solve = jittered
            fac = list(ref.init_factors(W, 134, init="random", device=None, seed=42))
            duals = [torch.zeros_like(f) for f in fac]
            loss, lossq = [], []
            for _ in range(6):
                e, eq = _sweep3(ref, W, fac, duals, 4, MSE, 60)
                loss.append(e); lossq.append(eq)
            out[f"trial{trial}/loss"], out[f"trial{trial}/lossq"] = np.array(loss), np.array(lossq)
            meta.append(dict(name=f"trial{trial}", noise_seed=noise_seed, eps=eps, sweeps=6, max_iter_admm=60))
            print("short-budget self-divergence trial", trial, eps, loss, flush=True)
        # ... and for the 2-D branch (scripts/factorize.py:269-310): the `mat` case of outer_loop.npz, same jitter
        g = torch.Generator().manual_seed(9)
        Wm = torch.randn(128, 48, generator=g) * 0.02
        for trial, (eps, noise_seed) in enumerate(trials):
            gn = torch.Generator().manual_seed(noise_seed + 100)

            def jittered2(b, L, upper=False):
                x = real_solve(b, L, upper=upper)
                sign = torch.randint(0, 2, x.shape, generator=gn).float() * 2 - 1
                return x * (1 + eps * sign)

            torch.cholesky_solve = jittered2
            A, B = ref.init_factors(Wm, 17, init="random", device=None, seed=3)
            U_A, U_B = torch.zeros_like(A), torch.zeros_like(B)
            loss = []
            for _ in range(5):
                A, U_A = ref.admm_iteration(A, U_A, Wm @ B, B.T @ B, max_iter=80, eps=1e-8, bits=4, qscheme=MSE)
                B, U_B = ref.admm_iteration(B, U_B, Wm.T @ A, A.T @ A, max_iter=80, eps=1e-8, bits=4, qscheme=MSE)
                loss.append(ref.squared_relative_diff(Wm, A @ B.T))
            out[f"mat_trial{trial}/loss"] = np.array(loss)
            print("2-D self-divergence trial", trial, eps, loss, flush=True)
    finally:
        torch.cholesky_solve = real_solve
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def long_run_cases(ref, sweeps_cap=200):
    """BASELINE config 1 run to the reference's own stop rule (scripts/factorize.py:259-263) with the full inner budget:
      * sweep 0, each mode: the int8 codes of EVERY inner iteration as a CRC32 (first-N check over all 999 iterations)
        plus the codes themselves at every 10th iteration;
      * teacher-forcing states (H, U entering inner iteration k, F, G -> codes after it) at k in {1, 500, 999} of every
        mode in sweep 1, in a middle sweep (30) and in the LAST sweep;
      * the complete error histories and the number of inner iterations every call ran (the early-exit record)."""
    import zlib
    W = config1_weight()
    out = {}
    fac = list(ref.init_factors(W, 134, init="random", device=None, seed=42))
    for m in range(3):
        out[f"init{m}"] = fac[m].numpy().copy()
    duals = [torch.zeros_like(f) for f in fac]
    loss, lossq, iters_run = [], [], []
    keep_k = (1, 500, 999)
    st = dict(m=0, k=0, U=None, Hprev=None, tag=None, sink=out, per_mode=[])
    crcs, scales, dense = {}, {}, {}

    def on_mode(m, H, U, F, G):
        if m > 0:
            st["per_mode"].append(st["k"])
        st.update(m=m, k=0, U=U, Hprev=H)
        st["sink"][f"{st['tag']}/m{m}/F"] = F.numpy().copy()
        st["sink"][f"{st['tag']}/m{m}/G"] = G.numpy().copy()

    def on_call(V, Hq):
        st["k"] += 1
        k, m, tag, sink = st["k"], st["m"], st["tag"], st["sink"]
        if tag == "sweep0":
            codes, scale = _codes_of(Hq)
            crcs.setdefault(m, []).append(zlib.crc32(codes.numpy().tobytes()))
            scales.setdefault(m, []).append(scale)
            if k % 10 == 0 or k == 999:
                dense.setdefault(m, []).append(codes.numpy().copy())
        if k in keep_k:
            codes, scale = _codes_of(Hq)
            key = f"{tag}/m{m}/k{k}"
            hp = st["Hprev"]
            if k == 1 and tag == "sweep0":
                sink[key + "/Hin"] = hp.numpy().copy()           # the random init, not grid-valued
            else:
                cin, sin = _codes_of(hp)
                sink[key + "/Hin_codes"], sink[key + "/Hin_scale"] = cin.numpy(), np.array([sin], np.float32)
            sink[key + "/Uin"] = st["U"].numpy().copy()
            sink[key + "/codes"], sink[key + "/scale"] = codes.numpy(), np.array([scale], np.float32)
        st["Hprev"] = Hq

    t0 = time.time()
    last = {}
    with _Recorder(ref) as rec:
        rec.on_call = on_call
        for sweep in range(sweeps_cap):
            named = {0: "sweep0", 1: "sweep1", 30: "sweep30"}.get(sweep)
            # the last sweep is not known in advance: an unnamed sweep records into a scratch dict that is kept only
            # if the stop rule fires after it
            scratch = {}
            st.update(tag=named or "last", sink=out if named else scratch, per_mode=[])
            e, eq = _sweep3(ref, W, fac, duals, 4, MSE, 1000, on_mode=on_mode)
            st["per_mode"].append(st["k"])
            iters_run.append(list(st["per_mode"]))
            if not named:
                last = scratch
            loss.append(e); lossq.append(eq)
            print(f"long run sweep {sweep}: {e:.6f} {eq:.6f} iters {st['per_mode']} {time.time() - t0:.0f}s", flush=True)
            if len(loss) > 1 and abs(loss[-2] - loss[-1]) < 1e-5:
                break
            if len(loss) > 10 and loss[-1] - loss[-5] > 1e-3:
                break
    out.update(last)
    for m in range(3):
        out[f"sweep0/m{m}/crc"] = np.array(crcs[m], dtype=np.uint32)
        out[f"sweep0/m{m}/scales"] = np.array(scales[m], dtype=np.float32)
        out[f"sweep0/m{m}/codes_every10"] = np.stack(dense[m])
        out[f"final{m}"] = fac[m].numpy().copy()
    out["loss"], out["lossq"] = np.array(loss), np.array(lossq)
    out["iters_run"] = np.array(iters_run, dtype=np.int32)
    meta = dict(sweeps=len(loss), keep_k=list(keep_k), rank=134, bits=4, qscheme=MSE, seed=42, max_iter_admm=1000,
                last_sweep=len(loss) - 1, tags=["sweep0", "sweep1", "sweep30", "last"])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def early_exit_cases(ref, src=None):
    """Inner loops that LEAVE EARLY (r < eps and s < eps, source/admm.py:62-65).  The states entering such calls were
    captured from the CUDA solver on a B200 (tools/early_exit_probe.py --dump: ResNet-18-shaped layers, tap factor
    9 x R of later sweeps) and are replayed here through the UNMODIFIED reference: the fixture keeps the state, the
    iteration at which the reference's own exit test fires and the codes it returns."""
    src = src or os.path.join(os.path.dirname(HERE), "gpurun_out", "r2a", "early")
    picks = ["p0_layer2.0.conv2_m2_s9", "p0_layer3.0.conv1_m2_s10", "p1_layer2.0.conv2_m2_s15", "p1_layer2.1.conv1_m2_s23"]
    out, meta = {}, []
    for name in picks:
        z = np.load(os.path.join(src, name + ".npz"))
        probe = json.loads(bytes(z["meta"]).decode())
        H, U, F, G = (torch.from_numpy(z[k].copy()) for k in ("H", "U", "F", "G"))
        calls = [0]
        with _Recorder(ref) as rec:
            rec.on_call = lambda V, Hq: calls.__setitem__(0, calls[0] + 1)
            Hn, Un = ref.admm_iteration(H.clone(), U.clone(), F, G, max_iter=1000, eps=1e-8, bits=4, qscheme=MSE)
        codes, scale = _codes_of(Hn)
        for k in ("H", "U", "F", "G"):
            out[f"{name}/{k}"] = z[k]
        out[f"{name}/codes"], out[f"{name}/scale"] = codes.numpy(), np.array([scale], np.float32)
        out[f"{name}/Uout"] = Un.numpy()
        meta.append(dict(name=name, reference_iterations=calls[0], cuda_iterations=probe["iterations"],
                         cuda_precision=probe["precision"], layer=probe["layer"], sweep=probe["sweep"], shape=probe["shape"]))
        print("early exit", name, "reference stops after", calls[0], "iterations; CUDA solver reported", probe["iterations"], flush=True)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def rank_table():
    """source/rank_map.py is pure data; it pins the rank rule (scripts/factorize.py:157-158)."""
    spec = importlib.util.spec_from_file_location("ref_rank_map", os.path.join(REFERENCE_ROOT, "source", "rank_map.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    table = {}
    for rate in (1.5, 2, 3, 4):
        table[str(rate)] = {k: v for k, v in mod.get_rank_map("resnet18", rate).items() if k.startswith("layer") or k == "conv1"}
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.set_num_threads(1)
    os.makedirs(args.out, exist_ok=True)
    ref = import_reference()
    jobs = dict(projection=projection_cases, contractions=contraction_cases,
                admm_iteration=admm_iteration_cases, outer_loop=outer_loop_cases,
                self_divergence=self_divergence_cases, solve_divergence=solve_divergence_cases,
                short_divergence=short_divergence_cases,
                long_run=long_run_cases, early_exit=early_exit_cases)
    for name, fn in jobs.items():
        if args.only and name not in args.only.split(","):
            continue
        if not args.only and name in ("long_run", "early_exit"):
            continue   # ~45 CPU-minutes: only on request (--only long_run)
        t0 = time.time()
        data = fn(ref)
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **data)
        print(f"{name}: {len(data)} arrays, {time.time() - t0:.1f}s", flush=True)
    if not args.only or "rank" in args.only:
        with open(os.path.join(args.out, "rank_table.json"), "w") as f:
            json.dump(dict(source="source/rank_map.py get_rank_map('resnet18', rate)", table=rank_table()), f, indent=1, sort_keys=True)
    with open(os.path.join(args.out, "PROVENANCE.json"), "w") as f:
        json.dump(dict(generator="oracle/make_golden.py", reference=REFERENCE_ROOT, torch=torch.__version__,
                       numpy=np.__version__, threads=1, cpu_capability=torch.backends.cpu.get_cpu_capability()), f, indent=1)


if __name__ == "__main__":
    main()
