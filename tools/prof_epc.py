"""Dev tool: where does one ALS / EPC pass spend its time?   python tools/prof_epc.py I J K [passes]

Prints the time of an ALS run (50 iterations budget), of `passes` EPC passes and of the pieces of one mode update:
the native float64 MTTKRP / Gram-Hadamard / column normalisation (csrc/contract.cu) against the materialised
Khatri-Rao form, and torch's (cuSOLVER's) symmetric eigen-decomposition, which is what remains."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
import numpy as np, torch
from source import _native, parafac_epc as pe
shp = tuple(int(a) for a in sys.argv[1:4])
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 10
g = torch.Generator().manual_seed(42)
W = (torch.randn(*shp, generator=g) * 0.05).cuda().double()
order = np.argsort(W.shape)
T = pe._Tensor(W.permute(tuple(int(o) for o in order)).contiguous())
R = int(W.numel() / sum(W.shape) / 2.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
w, fac = pe._als(T, R, 50, 1e-5, np.random.RandomState(42), True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"shape {shp} rank {R}: ALS {t1 - t0:.3f} s")
delta = float(torch.linalg.norm(T.Y - pe._reconstruct(w, fac))); fac[-1] = (fac[-1] * w).contiguous()
torch.cuda.synchronize(); t0 = time.perf_counter()
fac_e = [f.clone() for f in fac]
for i in range(passes):
    fac_e = pe._epc_sweep(T, fac_e, delta)                 # eigen form (no state)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"EPC, eigen form: {(t1 - t0) / passes * 1e3:.2f} ms per pass -> {2500 * (t1 - t0) / passes:.1f} s for the reference's budget of 50 rounds x 50 passes")
state = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(passes):
    fac = pe._epc_sweep(T, fac, delta, state)              # Cholesky form with a warm-started multiplier
torch.cuda.synchronize(); t1 = time.perf_counter()
diff = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(fac, fac_e))
print(f"EPC, Cholesky / series form: {(t1 - t0) / passes * 1e3:.2f} ms per pass ({state.get('chol_evals', 0)} factorizations and "
      f"{state.get('eigh_updates', 0)} eigen-form updates in {3 * passes} mode updates); factors after {passes} passes differ by {diff:.1e}")
mu = state["mu"]
Tm_ = T.mttkrp(fac, T.N - 1); gm_ = T.gram(fac, T.N - 1)
def tm(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
m = T.N - 1   # the largest mode
gamma = T.gram(fac, m)
def kr_form():
    o = [f for k, f in enumerate(fac) if k != m]
    kr = (o[0][:, None, :] * o[1][None, :, :]).reshape(-1, R)
    return T.unf[m] @ kr
print(f"mode {m}: native MTTKRP {tm(lambda: T.mttkrp(fac, m)):.3f} ms (materialised Khatri-Rao + matmul {tm(kr_form):.3f} ms), "
      f"native Gram-Hadamard {tm(lambda: T.gram(fac, m)):.3f} ms, column normalisation {tm(lambda: _native.normalize_columns_f64(fac[0].clone())):.3f} ms, "
      f"eigh {tm(lambda: torch.linalg.eigh(gamma)):.3f} ms, T @ V {tm(lambda: T.mttkrp(fac, m) @ gamma):.3f} ms, "
      f"one series expansion (potrf + trsm + inverse GEMM + 8 GEMMs + moment GEMM) {tm(lambda: pe._ridge_eval(gm_, Tm_, mu, 8)[1].tolist()):.3f} ms, "
      f"with 4 / 16 terms {tm(lambda: pe._ridge_eval(gm_, Tm_, mu, 4)[1].tolist()):.3f} / {tm(lambda: pe._ridge_eval(gm_, Tm_, mu, 16)[1].tolist()):.3f} ms")
Mm = gm_ + mu * torch.eye(R, dtype=torch.float64, device='cuda')
Lm = torch.linalg.cholesky(Mm)
eye = torch.eye(R, dtype=torch.float64, device='cuda')
def via_trtri():
    Li = torch.linalg.solve_triangular(Lm, eye, upper=False)
    return Li.T @ Li
print(f"pieces: potrf {tm(lambda: torch.linalg.cholesky_ex(Mm)):.3f} ms, cholesky_inverse {tm(lambda: torch.cholesky_inverse(Lm)):.3f} ms, "
      f"inverse as trsm(L, I) + GEMM {tm(via_trtri):.3f} ms (difference {float((via_trtri() - torch.cholesky_inverse(Lm)).abs().max() / torch.cholesky_inverse(Lm).abs().max()):.1e}), "
      f"GEMM Tm @ Minv {tm(lambda: Tm_ @ Mm):.3f} ms")
