"""Dev tool: us per inner ADMM iteration of the persistent kernel for the ResNet-18 shapes."""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
MSE = "tensor_mseminmax_symmetric"
shapes = [(64, 134, 64, 9), (9, 134, 64, 64), (128, 278, 128, 9), (9, 278, 128, 128), (256, 566, 256, 9), (9, 566, 256, 256),
          (512, 759, 256, 9), (256, 759, 512, 9), (512, 1141, 512, 9), (9, 1141, 512, 512), (2048, 204, 512, 1), (4096, 1024, 4096, 1)]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator().manual_seed(0)
for (I, R, n1, n2) in shapes:
    B = torch.randn(n1, R, generator=g).cuda(); C = torch.randn(n2, R, generator=g).cuda()
    G = nat.gram_hadamard(B, C if n2 > 1 else None)
    F = (torch.randn(I, R, generator=g) * (n1 * n2) ** 0.5).cuda()
    H = torch.randn(I, R, generator=g).cuda(); U = torch.zeros_like(H)
    for rep in range(2):
        Hc, Uc = H.clone(), U.clone()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        Minv, rho, st = nat.spd_inverse(G)
        e1.record()
        r = nat.admm_iteration_inplace(Hc, Uc, F, G, iters + 1, 1e-8, 4, MSE, precision=prec)
        e2.record()
        torch.cuda.synchronize()
    rp = nat.read_report(r)
    t_inv = e0.elapsed_time(e1); t_all = e1.elapsed_time(e2)
    per = (t_all - t_inv) / rp.iterations * 1e3
    evals = 200.0 * I * R / (per * 1e-6) / 1e12
    gflops = 2.0 * I * R * R / (per * 1e-6) / 1e12
    print(f"I={I:5d} R={R:5d}: inverse {t_inv:8.3f} ms; loop {per:8.2f} us/iter ({rp.iterations} its)  "
          f"{evals:6.3f} T cand-evals/s  gemm-equiv {gflops:6.2f} TFLOP/s  phases us/iter "
          f"{[round(x / 1e3 / max(rp.iterations, 1), 1) for x in rp.phase_ns[:3]]}", flush=True)
