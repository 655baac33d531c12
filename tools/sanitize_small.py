"""Dev tool: a tiny end-to-end exercise for compute-sanitizer (memcheck / racecheck): projection, general loop on a
few CTAs (tensor-core and FFMA products), resident single-CTA loop, two-block splitting, tensor-core MTTKRP."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
MSE = "tensor_mseminmax_symmetric"
g = torch.Generator().manual_seed(0)
x = torch.randn(70, 90, generator=g).cuda()
nat.project(x, 4, MSE, 50, want_codes=True)
for (I, R, ctas, prec) in [(64, 96, 1, 0), (9, 96, 1, 0), (130, 96, 3, 1), (130, 96, 3, 0), (9, 200, 2, 0)]:
    B = torch.randn(40, R, generator=g).cuda()
    G = nat.gram_hadamard(B, None)
    F = (torch.randn(I, R, generator=g) * 10).cuda()
    H = torch.randn(I, R, generator=g).cuda()
    U = torch.zeros_like(H)
    rep = nat.admm_iteration_inplace(H, U, F, G, 4, 1e-8, 4, MSE, num_attempts=50, precision=prec, max_ctas=ctas)
    print(I, R, ctas, prec, nat.read_report(rep).iterations, float(H.abs().max()))
W = torch.randn(60, 40, generator=g).cuda()
nat.split_loop_inplace(torch.randn(60, 40, generator=g).cuda(), torch.zeros(60, 40).cuda(), W, torch.zeros(60, 40).cuda(), 1.0, 4, 1e-8, 4, "tensor_minmax", max_ctas=2)
Wt = torch.randn(20, 12, 9, generator=g).cuda()
V = nat.permute_myx(Wt.reshape(20, 108), 12, 9)
nat.mttkrp_tc(V, 20, torch.randn(12, 33, generator=g).cuda(), torch.randn(9, 33, generator=g).cuda())
torch.cuda.synchronize()
print("done")
