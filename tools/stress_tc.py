"""Dev tool: random-shape stress of the tensor-core tile product (pre-split path via mttkrp_tc, on-the-fly path via
gemm_nt) and of the persistent loop with the tensor-core ridge product at random SM budgets; checks against float64."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
g = torch.Generator().manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 60):
    M, R, nx = ri(1, 900), ri(1, 1500), ri(1, 1300)
    if it % 7 == 0:
        nx = ri(1, 130)            # one or two K-blocks
    if it % 11 == 0:
        M = ri(3000, 6000)         # many tiles per CTA
    W = torch.randn(M, nx, generator=g).cuda()
    X = torch.randn(nx, R, generator=g).cuda()
    ldv = (nx + 3) // 4 * 4
    V = torch.zeros(M, ldv, device="cuda"); V[:, :nx] = W
    F = nat.mttkrp_tc(V, M, X, None)
    ref = W.double() @ X.double()
    e1 = float((F.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    Xt = torch.zeros(R, ldv, device="cuda"); Xt[:, :nx] = X.t()
    C = nat.gemm_nt(V, Xt)
    e2 = float((C.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    worst = max(worst, e1, e2)
    assert e1 < 2e-5 and e2 < 2e-5, (M, R, nx, e1, e2)
print("tile product ok, worst rel err", worst, flush=True)
MSE = "tensor_mseminmax_symmetric"
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    I, R, ctas = ri(64, 700), ri(32, 900), ri(1, 148)
    B = torch.randn(ri(8, 300), R, generator=g).cuda()
    G = nat.gram_hadamard(B, None)
    F = (torch.randn(I, R, generator=g) * 10).cuda()
    H0 = torch.randn(I, R, generator=g).cuda()
    outs = []
    for c in (0, ctas):
        H, U = H0.clone(), torch.zeros_like(H0)
        rep = nat.admm_iteration_inplace(H, U, F, G, 6, 1e-8, 4, MSE, precision=1, max_ctas=c)
        nat.read_report(rep)
        outs.append((H, U))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (I, R, ctas)
print("loop ok (bit-identical across SM budgets)", flush=True)
