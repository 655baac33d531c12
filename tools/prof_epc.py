import sys, time
sys.path[:0]=['/root/repo','/root/repo/admm-quantization_b200']
import numpy as np, torch
from source import parafac_epc as pe
shp = tuple(int(a) for a in sys.argv[1:4])
g = torch.Generator().manual_seed(42)
W = (torch.randn(*shp, generator=g) * 0.05).cuda().double()
order = np.argsort(W.shape); Yp = W.permute(tuple(int(o) for o in order)).contiguous()
R = int(W.numel() / sum(W.shape) / 2.0)
np.random.seed(42)
torch.cuda.synchronize(); t0=time.perf_counter()
w, fac = pe.parafac_als(Yp, R, n_iter_max=50, tol=1e-5, normalize_factors=True)
torch.cuda.synchronize(); t1=time.perf_counter()
print("ALS", t1-t0, "s")
delta = float(torch.linalg.norm(Yp - pe._reconstruct(w, fac))); fac[-1] = fac[-1]*w
n2 = float((Yp*Yp).sum())
for name, cache in (("eigh", None), ("chol", [None]*3)):
    f = [x.clone() for x in fac]
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(10): f = pe.epc_sweep(Yp, f, delta, n2, cache)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print(name, "per sweep", (t1-t0)/10*1e3, "ms")
# parts
m=0
gamma = (f[1].T@f[1])*(f[2].T@f[2]); T = pe._mttkrp(Yp, f, 0)
def tm(fn,n=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("eigh ms", tm(lambda: torch.linalg.eigh(gamma)))
L = torch.linalg.cholesky(gamma + 0.1*torch.eye(R, device='cuda', dtype=torch.float64))
S = T.T@T
print("cholesky ms", tm(lambda: torch.linalg.cholesky_ex(gamma + 0.1*torch.eye(R, device='cuda', dtype=torch.float64))))
print("cholesky_solve RxR ms", tm(lambda: torch.cholesky_solve(S, L)))
print("mttkrp ms", tm(lambda: pe._mttkrp(Yp, f, 0)))
print("inverse via solve_triangular ms", tm(lambda: torch.linalg.solve_triangular(L, torch.eye(R, device='cuda', dtype=torch.float64), upper=False)))
print("matmul RxR ms", tm(lambda: S@S))
