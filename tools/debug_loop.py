"""Debug: one ADMM iteration on the GPU vs the oracle with intermediates (dev tool, not product)."""
import os, sys, json
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from oracle import admm_oracle as orc
from source import _native as nat
from source import admm as A
torch.set_num_threads(1)
z = np.load(os.path.join(REPO, "tests/golden/admm_iteration.npz"))
n = "l1_mode0_b4"
H0, F, G = (torch.from_numpy(z[f"{n}/{k}"]) for k in ("H0", "F", "G"))
U0 = torch.zeros_like(H0)
tr = []
Ho, Uo, _ = orc.admm_iteration(H0.clone(), U0.clone(), F, G, 2, 1e-8, 4, "tensor_mseminmax_symmetric", trace=tr)
t = tr[0]
print("oracle: rho", float(orc.ridge_rho(G)), "scale", t["scale"], "best", t["best"], "absmax", float(t["V"].abs().max()))
Ud = U0.clone().cuda()
Hd, _ = A.admm_iteration(H0.cuda(), Ud, F.cuda(), G.cuda(), 2, 1e-8, 4, "tensor_mseminmax_symmetric")
r = A.last_report
print("gpu   : rho", r.rho, "scale", r.scale, "best", r.best_index, "absmax", r.absmax, "iters", r.iterations, "status", r.status, "r", r.r, "s", r.s)
print("oracle r,s", t["r"], t["s"])
# Minv check
Minv, rho, st = nat.spd_inverse(G.cuda())
R = G.shape[0]
M = Minv[:, :R].cpu()
rhs = F + float(rho.item()) * (H0 + U0)
Hls_m = rhs @ M
print("Hls via Minv vs oracle Hls maxdiff", float((Hls_m - t["Hls"]).abs().max()), "scale", float(t["Hls"].abs().max()))
# U after = U + H - Hls => Hls_gpu = H - U_new
Hls_gpu = (Hd - Ud).cpu()
print("Hls gpu(kernel) vs oracle maxdiff", float((Hls_gpu - t["Hls"]).abs().max()))
print("H agreement", float((Hd.cpu() == t["H"]).float().mean()))
d = (Hls_gpu - t["Hls"]).abs()
idx = torch.nonzero(d > 1e-3)
print("bad Hls elements", idx.shape[0], idx[:10].tolist())
