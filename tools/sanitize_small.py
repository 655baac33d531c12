"""Dev tool: a tiny end-to-end exercise for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): projection,
ridge inverse, the general loop on a few CTAs with every form of the ridge product (float64 parity tiles, tcgen05
3xTF32 with its TMA ring / TMEM / mbarrier pipeline, float32 FFMA tiles, column strips), the shared-memory-resident
single-CTA loop, the two-block splitting, the tensor-core GEMM and MTTKRP (epilogue fold), the float64 kernels of the
ALS / EPC initialisation.      compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
MSE = "tensor_mseminmax_symmetric"
g = torch.Generator().manual_seed(0)
x = torch.randn(70, 90, generator=g).cuda()
nat.project(x, 4, MSE, 50, want_codes=True)
# (rows, rank, CTA budget, precision): resident kernel (1 CTA, precision 2), parity tiles, tensor cores, FFMA tiles, strips
for (I, R, ctas, prec) in [(64, 96, 1, 2), (9, 96, 1, 2), (130, 96, 3, 0), (130, 96, 3, 1), (130, 96, 3, 2), (9, 200, 2, 1), (9, 200, 2, 0)]:
    B = torch.randn(40, R, generator=g).cuda()
    G = nat.gram_hadamard(B, None)
    F = (torch.randn(I, R, generator=g) * 10).cuda()
    H = torch.randn(I, R, generator=g).cuda()
    U = torch.zeros_like(H)
    rep = nat.admm_iteration_inplace(H, U, F, G, 4, 1e-8, 4, MSE, num_attempts=50, precision=prec, max_ctas=ctas)
    print(I, R, ctas, prec, nat.read_report(rep).iterations, float(H.abs().max()))
W = torch.randn(60, 40, generator=g).cuda()
nat.split_loop_inplace(torch.randn(60, 40, generator=g).cuda(), torch.zeros(60, 40).cuda(), W, torch.zeros(60, 40).cuda(), 1.0, 4, 1e-8, 4, "tensor_minmax", max_ctas=2)
A = torch.randn(140, 72, generator=g).cuda()
Bm = torch.randn(50, 72, generator=g).cuda()
C = nat.gemm_nt(A, Bm)
print("gemm_nt max err", float((C - A @ Bm.T).abs().max()))
Wt = torch.randn(20, 12, 9, generator=g).cuda()
V = nat.permute_myx(Wt.reshape(20, 108), 12, 9)
nat.mttkrp_tc(V, 20, torch.randn(12, 33, generator=g).cuda(), torch.randn(9, 33, generator=g).cuda())
nat.mttkrp(Wt.reshape(20, 108), torch.randn(12, 33, generator=g).cuda(), torch.randn(9, 33, generator=g).cuda(), 0)
Yd = torch.randn(10, 8, 6, generator=g, dtype=torch.float64).cuda()
fx, fy = torch.randn(8, 7, generator=g, dtype=torch.float64).cuda(), torch.randn(6, 7, generator=g, dtype=torch.float64).cuda()
nat.mttkrp_f64(Yd.reshape(10, 48).contiguous(), fx, fy)
nat.gram_hadamard_f64(fx, fy)
nat.normalize_columns_f64(fx.clone(), carry=torch.ones(7, dtype=torch.float64).cuda())
torch.cuda.synchronize()
print("done")
