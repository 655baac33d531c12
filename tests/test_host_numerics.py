"""CPU-only: the kernels' float32 recipe (csrc/numerics.cuh, compiled for the host by
tests/hostcheck) against the oracle and the reference golden vectors; ABI surface checks."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from oracle import admm_oracle as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "admm-quantization_b200")
torch.set_num_threads(1)


@pytest.fixture(scope="session")
def hc():
    src = os.path.join(REPO, "tests", "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(REPO, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    deps = [src, os.path.join(PKG, "csrc", "numerics.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hc_mse.restype = ctypes.c_longlong
    lib.hc_fastpath_violations.restype = ctypes.c_longlong
    return lib


def fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_candidate_grid_and_scales_match_oracle(hc):
    g = torch.Generator().manual_seed(1)
    for _ in range(300):
        mx = float((torch.rand(1, generator=g) * 10 ** float(torch.randint(-8, 8, (1,), generator=g))).float())
        for n in (1, 2, 7, 200, 1000):
            for bits in (2, 4, 8):
                clip = np.empty(n, np.float32)
                scale = np.empty(n, np.float32)
                hc.hc_candidates(ctypes.c_float(mx), n, bits, fp(clip), fp(scale))
                ref = orc.candidate_grid(mx, n)
                assert np.array_equal(clip, ref)
                q = 2 ** (bits - 1)
                ref_scale = (2 * torch.from_numpy(ref) / (2 * q - 1)).numpy()
                assert np.array_equal(scale, ref_scale)


def test_kernel_mse_recipe_selects_reference_candidate(hc, golden_projection):
    gp = golden_projection
    for m in gp.meta:
        if "scheme" in m or m["name"] == "all_zero_b4":
            continue
        x = np.ascontiguousarray(gp[m["name"] + "/x"]).reshape(-1)
        mx = np.float32(max(abs(x.min()), abs(x.max())))
        nc = m["num_attempts"]
        mse = np.empty(nc, np.float32)
        best = ctypes.c_int()
        hc.hc_mse(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), m["bits"], nc, fp(mse), ctypes.byref(best), 0)
        idx, scale = gp[m["name"] + "/idx_scale"]
        assert best.value == int(idx), m["name"]
        xq = np.empty_like(x)
        codes = np.empty(x.size, np.int8)
        hc.hc_quantize(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(scale), m["bits"], fp(xq), fp(codes))
        assert np.array_equal(xq.view(np.uint32), gp[m["name"] + "/xq"].reshape(-1).view(np.uint32)), m["name"]
        assert np.array_equal(codes, gp[m["name"] + "/codes"].reshape(-1))


def test_threshold_form_selects_reference_candidate(hc, golden_projection):
    """The threshold form of the clip search (what the kernels run) picks the reference's candidate on every golden
    vector, and its sums agree with the direct float32 evaluation to float32 summation noise."""
    gp = golden_projection
    checked = 0
    for m in gp.meta:
        if "scheme" in m or m["name"] == "all_zero_b4":
            continue
        x = np.ascontiguousarray(gp[m["name"] + "/x"]).reshape(-1)
        mx = np.float32(max(abs(x.min()), abs(x.max())))
        if not (2.0 ** -40 <= mx <= 2.0 ** 40):
            continue  # outside the threshold form's range the kernels use the direct form
        nc = m["num_attempts"]
        mse, sums, best = np.empty(nc, np.float32), np.empty(nc, np.float64), ctypes.c_int()
        hc.hc_mse_thresholds(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), m["bits"], nc, fp(mse), fp(sums), ctypes.byref(best))
        idx, scale = gp[m["name"] + "/idx_scale"]
        assert best.value == int(idx), m["name"]
        direct = np.empty(nc, np.float32)
        hc.hc_mse(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), m["bits"], nc, fp(direct), ctypes.byref(best), 0)
        ok = direct > 0
        assert np.all(np.abs(mse[ok] - direct[ok]) <= 3e-6 * direct[ok]), m["name"]
        checked += 1
    assert checked >= 10


def test_code_thresholds_are_exact(hc):
    g = torch.Generator().manual_seed(11)
    for bits in (1, 2, 3, 4, 6, 8):
        for trial in range(200):
            scale = np.float32(float(torch.rand(1, generator=g) + 0.01) * 10 ** float(torch.randint(-9, 9, (1,), generator=g)))
            assert hc.hc_threshold_violations(ctypes.c_float(scale), bits) == 0, (bits, scale)
        # scales where the quotient is exact or the rounding boundary itself is a float: powers of two, mantissas of all
        # ones / one bit, and the ends of the range the threshold form accepts (abs-max in [2^-40, 2^40])
        for e in (-46, -40, -23, -1, 0, 1, 7, 24, 37):
            for mant in (1.0, 1.5, 1.0 + 2.0 ** -23, 2.0 - 2.0 ** -23, 1.25, 1.0 + 2.0 ** -12):
                scale = np.float32(mant * 2.0 ** e)
                assert hc.hc_threshold_violations(ctypes.c_float(scale), bits) == 0, (bits, scale)


def test_threshold_form_matches_oracle_argmin_on_random_tensors(hc):
    g = torch.Generator().manual_seed(77)
    for shape, bits, nc in [((64, 134), 4, 200), ((9, 134), 4, 200), ((33, 57), 3, 200), ((128, 97), 8, 200),
                            ((7, 5), 2, 33), ((40, 40), 6, 1000), ((1, 1), 4, 200), ((2, 3), 1, 50)]:
        for trial in range(3):
            xt = torch.randn(*shape, generator=g) * float(10 ** torch.randint(-3, 3, (1,), generator=g).item())
            _, _, _, ref_best, mses = orc.project_mse(xt, bits, nc, "aten")
            x = np.ascontiguousarray(xt.numpy()).reshape(-1)
            mx = np.float32(np.abs(x).max())
            mse, sums, best = np.empty(nc, np.float32), np.empty(nc, np.float64), ctypes.c_int()
            hc.hc_mse_thresholds(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), bits, nc, fp(mse), fp(sums), ctypes.byref(best))
            if best.value != ref_best:
                gap = abs(float(mses[best.value]) - float(mses[ref_best])) / float(mses[ref_best])
                assert gap < 3e-7, (shape, bits, nc, best.value, ref_best, gap)
            # the sums themselves against a float64 evaluation of the reference's expression
            clip, scale = np.empty(nc, np.float32), np.empty(nc, np.float32)
            hc.hc_candidates(ctypes.c_float(mx), nc, bits, fp(clip), fp(scale))
            q = 2 ** (bits - 1)
            for c in (0, nc // 3, nc - 1):
                codes = np.clip(np.rint((x / scale[c]).astype(np.float32)), -q, q - 1).astype(np.float32)
                y = (codes * scale[c]).astype(np.float32)
                exact = float(((x.astype(np.float64) - y.astype(np.float64)) ** 2).sum())
                assert abs(sums[c] - exact) <= 1e-9 * exact + 1e-300, (shape, bits, c, sums[c], exact)


def test_fast_path_equals_exact_division(hc):
    """x*(1/s) + magic-number rounding agrees with rint(x/s) whenever the fast path accepts, incl.
    values placed right at the rounding boundaries (k + 0.5) * s."""
    g = torch.Generator().manual_seed(9)
    total_slow = 0
    for bits in (2, 3, 4, 6, 8):
        q = 2 ** (bits - 1)
        for trial in range(40):
            scale = np.float32(float(torch.rand(1, generator=g)) * 10 ** float(torch.randint(-6, 6, (1,), generator=g)) + 1e-30)
            ks = np.arange(-q - 2, q + 2, dtype=np.float64) + 0.5
            edge = (ks * float(scale)).astype(np.float32)
            near = np.concatenate([np.nextafter(edge, np.float32(np.inf)), np.nextafter(edge, np.float32(-np.inf)), edge])
            rnd = (torch.randn(4000, generator=g).numpy() * float(scale) * q * 0.7).astype(np.float32)
            x = np.ascontiguousarray(np.concatenate([near, rnd, np.zeros(3, np.float32)]))
            bad = hc.hc_fastpath_violations(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(scale), bits)
            assert bad == 0, (bits, scale)
            # and the whole per-candidate recipe with/without forcing the exact path gives the same MSEs
            mx = np.float32(np.abs(x).max())
            a, b = np.empty(50, np.float32), np.empty(50, np.float32)
            best = ctypes.c_int()
            slow = hc.hc_mse(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), bits, 50, fp(a), ctypes.byref(best), 0)
            hc.hc_mse(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(mx), bits, 50, fp(b), ctypes.byref(best), 1)
            assert np.array_equal(a, b)
            total_slow += slow
    assert total_slow > 0  # the boundary inputs do exercise the fallback


def test_other_scheme_recipes_match_reference(hc, golden_projection):
    gp = golden_projection
    for m in gp.meta:
        if m.get("scheme") not in ("tensor_minmax", "tensor_affine"):
            continue
        x = np.ascontiguousarray(gp[m["name"] + "/x"]).reshape(-1)
        out = np.empty_like(x)
        fn = hc.hc_minmax if m["scheme"] == "tensor_minmax" else hc.hc_affine
        fn(fp(x), ctypes.c_longlong(x.size), ctypes.c_float(x.min()), ctypes.c_float(x.max()), m["bits"], fp(out))
        assert np.array_equal(out.view(np.uint32), gp[m["name"] + "/xq"].reshape(-1).view(np.uint32)), m["name"]


# ------------------------------------------------------------------ ABI surface (no compute without a GPU)
def _declared_symbols():
    text = open(os.path.join(REPO, "include", "admmq.h")).read()
    return sorted(set(re.findall(r"ADMMQ_API [\w\s\*]*?\b(admmq_\w+)\(", text)))


def test_library_exports_every_declared_symbol():
    import sys
    sys.path.insert(0, PKG)
    from source import _native
    names = _declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(_native.lib, n), n
    assert sorted(_native.EXPORTS) == names
    assert _native.lib.admmq_version() == 200
    assert _native.lib.admmq_padded_ld(134) == 136
    assert _native.lib.admmq_admm_iteration_workspace_bytes(64, 134, 200) > 0


def test_binding_argument_counts_match_the_header():
    """Every prototype of include/admmq.h against the ctypes signature of source/_native.py: same number of
    parameters, pointer parameters bound as pointers (an ABI drift between header and binding corrupts the stack)."""
    import ctypes
    import sys
    sys.path.insert(0, PKG)
    from source import _native
    text = open(os.path.join(REPO, "include", "admmq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"ADMMQ_API\s+[\w\s\*]*?\b(admmq_\w+)\(([^;]*?)\);", text, flags=re.S)
    assert len(protos) >= 25
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        fn = getattr(_native.lib, name)
        assert fn.argtypes is not None and len(fn.argtypes) == len(plist), (name, len(fn.argtypes or []), plist)
        for decl, ct in zip(plist, fn.argtypes):
            is_ptr = "*" in decl
            bound_ptr = ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "contents") or issubclass(ct, ctypes._Pointer)
            assert is_ptr == bound_ptr, (name, decl, ct)


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    from source import _native
    from source.quantization import quantize_tensor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        quantize_tensor(torch.randn(4, 4), 4, "tensor_mseminmax_symmetric")
    a = ctypes.c_int()
    assert _native.lib.admmq_device_info(ctypes.byref(a), ctypes.byref(a), ctypes.byref(a)) == _native.E_CUDA
    assert "no CPU fallback" in _native.last_error()


def test_product_never_touches_the_oracle():
    bad = []
    for root, _, files in os.walk(PKG):
        if os.path.basename(root) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text or "admm_oracle" in text:
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_epc_cholesky_form_equals_the_eigen_form():
    """source/parafac_epc.py::_ridge_factor_chol (power series of the residual and of the factor around a warm-started
    multiplier, one Cholesky factorization + explicit inverse per expansion point) against the eigen form of the same mode update (`_multiplier`, the restatement the oracle
    pins): same multiplier and same factor to rounding, from good and from poor starts, when the constraint is
    inactive (mu = floor), and a clean give-up (None) on a singular Gram matrix.  The dense algebra is torch.linalg, so
    this host check runs the product's own function on CPU tensors (no libadmmq call is involved)."""
    import math
    import sys
    sys.path.insert(0, PKG)
    from source.parafac_epc import _multiplier, _ridge_factor_chol
    g = torch.Generator().manual_seed(11)
    worst_mu, worst_f, evals = 0.0, 0.0, []
    for (I, R, spread) in ((40, 24, 1.0), (9, 60, 3.0), (128, 96, 2.0), (64, 134, 4.0)):
        A = torch.randn(3 * R, R, generator=g, dtype=torch.float64) * torch.logspace(0, -spread, R, dtype=torch.float64)
        gamma = A.T @ A
        gamma = gamma / gamma.diagonal().max()
        Tm = torch.randn(I, R, generator=g, dtype=torch.float64)
        sig, V = torch.linalg.eigh(gamma)
        sig = sig.clamp(min=0.0)
        Tt = Tm @ V
        s = (Tt * Tt).sum(0)
        ls = float((s / sig).sum())                       # what the unconstrained least-squares fit explains
        norm_y2 = ls * 1.5
        for frac in (0.9, 0.5, 0.1):
            target = norm_y2 - frac * ls                  # residual the multiplier has to reach (> the LS residual)
            mu_e = _multiplier(sig.numpy(), s.numpy(), norm_y2, target)
            F_e = (Tt / (sig + mu_e)) @ V.T
            floor = float(sig.max()) * 1e-14
            for start in (1.02, 0.7, 30.0):
                cnt = {}
                got = _ridge_factor_chol(gamma, Tm, norm_y2, target, mu_e * start, floor, max_evals=12, counters=cnt)
                assert got is not None, (I, R, frac, start)
                mu, F = got[0], got[1]
                worst_mu = max(worst_mu, abs(mu - mu_e) / mu_e)
                worst_f = max(worst_f, float((F - F_e).abs().max() / F_e.abs().max()))
                if start == 1.02:
                    evals.append(cnt["chol_evals"])
        # inactive constraint: even the least-squares fit leaves more than the target -> mu = floor, F = the LS solution
        target = (norm_y2 - ls) * 0.8
        mu_e = _multiplier(sig.numpy(), s.numpy(), norm_y2, target)
        got = _ridge_factor_chol(gamma, Tm, norm_y2, target, 0.05, float(sig.max()) * 1e-14, max_evals=40)
        if got is not None:                               # (None = Cholesky of Gamma + 1e-14 I failed: eigen form takes over)
            assert got[0] == mu_e == float(sig.max()) * 1e-14
    assert worst_mu <= 1e-9 and worst_f <= 1e-9, (worst_mu, worst_f)
    assert max(evals) == 1, evals                         # a 2 % warm start: ONE factorization
    # singular Gram matrix: no exception, the caller falls back to the eigen form
    B = torch.randn(10, 4, generator=g, dtype=torch.float64)
    sing = (B @ B.T)
    assert _ridge_factor_chol(sing, torch.randn(5, 10, generator=g, dtype=torch.float64), 10.0, 1.0, 0.0, 0.0) is None
