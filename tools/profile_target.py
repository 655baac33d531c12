"""Dev tool: a short run of the dominant kernel for `ncu --set full` (layer4-sized factor update)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import workloads as wl
from source.solver import LayerSolver
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "layer4.1.conv1"
g = int(sys.argv[3]) if len(sys.argv) > 3 else 0
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
layers = [l for l in wl.resnet18_conv_layers() if l[0] == which]
(name, W, rank, init), = wl.build_problems(layers)
s = LayerSolver(W.cuda(), [f.cuda() for f in init], 4, "tensor_mseminmax_symmetric", max_iter_admm=iters + 1,
                solve_precision=prec, mttkrp_precision=prec, max_ctas=g)
for _ in range(2):
    s.update_mode(mode)
torch.cuda.synchronize()
rep = s.reports_dev[mode]
from source import _native
r = _native.read_report(rep)
print(name, "mode", mode, "ctas", g, "rank", rank, "iters", r.iterations, "phase us/iter:", [round(x / 1e3 / max(r.iterations, 1), 2) for x in r.phase_ns])
