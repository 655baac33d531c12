"""Dev tool: per-kernel time of one mode update of a layer at a given cooperative-grid budget."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat, workloads as wl
from source.solver import LayerSolver
which = sys.argv[1] if len(sys.argv) > 1 else "layer4.1.conv1"
g = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
(name, W, rank, init), = wl.build_problems([l for l in wl.resnet18_conv_layers() if l[0] == which])
s = LayerSolver(W.cuda(), [f.cuda() for f in init], 4, "tensor_mseminmax_symmetric", max_iter_admm=iters + 1, solve_precision=1, max_ctas=g)
def T(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for mode in range(3):
    o = s._others[mode]; X = s.factors[o[0]]; Y = s.factors[o[1]]
    for rep in range(2):
        t = {}
        t["gram"] = T(lambda: nat.gram_hadamard(X, Y, out=s.G))
        t["mttkrp"] = T(lambda: nat.mttkrp(s.unfoldings[mode], X, Y, 0, out=s.F[mode], ws=s.ws_mttkrp))
        t["inverse"] = T(lambda: nat.spd_inverse(s.G, out=(s.Minv, s.rho, s.inv_status), ws=s.ws_inv, max_ctas=g))
        t["loop"] = T(lambda: nat.admm_loop_inplace(s.factors[mode], s.duals[mode], s.F[mode], s.Minv, s.rho, s.inv_status, s.max_iter_admm, s.eps, 4, s.qscheme, 200, None, report=s.reports_dev[mode], ws=s.ws_loop, precision=1, max_ctas=g))
        t["project"] = T(lambda: nat.project(s.factors[mode], 4, s.qscheme, 200, out=s.factors_q[mode], ws=s.ws_proj))
    r = nat.read_report(s.reports_dev[mode])
    print(f"{name} mode {mode} grid {g}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in t.items()) + f" | loop per iter {t['loop'] / iters * 1e3:.1f} us phases {[round(x / 1e3 / iters, 1) for x in r.phase_ns[:3]]}")
print("recon", T(lambda: s._error_sums(s.factors, out=s.err_sums[0])), "ms")
