"""Dev tool: MTTKRP timings (float64 CUDA-core vs 3xTF32 tensor-core) on the largest ResNet-18 layer."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
I, J, K, R = 512, 512, 9, 1141
g = torch.Generator().manual_seed(0)
W = (torch.randn(I, J, K, generator=g) * 0.02).cuda()
fac = [torch.randn(d, R, generator=g).cuda() for d in (I, J, K)]
unf = [W.reshape(I, J * K), nat.unfold3(W, 1), nat.unfold3(W, 2)]
dims = [I, J, K]
def T(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
def TG(fn, n=20):
    """GPU time per call without host launch overhead: n calls captured into one CUDA graph, replayed three times."""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_, stream=s):
            for _ in range(n): fn()
        g_.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3): g_.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) / (3 * n)
flop = 2.0 * I * J * K * R
for mode in range(3):
    o = [k for k in range(3) if k != mode]
    X, Y = fac[o[0]], fac[o[1]]
    V = nat.permute_myx(unf[mode], X.shape[0], Y.shape[0])
    ws = torch.empty(nat.mttkrp_tc_workspace_bytes(dims[mode], X.shape[0], Y.shape[0], R), dtype=torch.uint8, device="cuda")
    F1 = torch.empty(dims[mode], R, device="cuda")
    t1 = T(lambda: nat.mttkrp_tc(V, dims[mode], X, Y, out=F1, ws=ws))
    try:
        tg = TG(lambda: nat.mttkrp_tc(V, dims[mode], X, Y, out=F1, ws=ws))
    except Exception as e:  # noqa: BLE001
        tg = float("nan"); print("graph timing failed:", str(e)[:100])
    t0 = T(lambda: nat.mttkrp(unf[mode], X, Y, 0))
    F0 = nat.mttkrp(unf[mode], X, Y, 0)
    err = ((F1 - F0).abs().max() / F0.abs().max()).item()
    print(f"mode {mode}: f64 CUDA-core {t0 * 1e3:8.1f} us ({flop / t0 / 1e9:6.1f} TFLOP/s)   3xTF32 tcgen05 {t1 * 1e3:8.1f} us per call from Python, {tg * 1e3:6.1f} us in a CUDA graph ({flop / tg / 1e9:6.1f} TFLOP/s fp32-equivalent, transpose/split kernel included)   max rel diff {err:.2e}")

# 2-D (matrix) MTTKRP of BASELINE config 5: F = W . B (scripts/factorize.py:277), W (out x in), B (in x R)
for (M, nx, R2) in [(4096, 4096, 1024), (11008, 4096, 1492), (4096, 11008, 1492)]:
    Wm = (torch.randn(M, nx, generator=g) * 0.02).cuda()
    X = torch.randn(nx, R2, generator=g).cuda()
    ws = torch.empty(nat.mttkrp_tc_workspace_bytes(M, nx, 1, R2), dtype=torch.uint8, device="cuda")
    F1 = torch.empty(M, R2, device="cuda")
    t1 = T(lambda: nat.mttkrp_tc(Wm, M, X, None, out=F1, ws=ws))
    ref = (Wm.double() @ X.double())
    err = ((F1.double() - ref).abs().max() / ref.abs().max()).item()
    fl = 2.0 * M * nx * R2
    print(f"matrix {M} x {nx}, R = {R2}: 3xTF32 tcgen05 {t1 * 1e3:8.1f} us ({fl / t1 / 1e9:6.1f} TFLOP/s fp32-equivalent)   max rel err vs float64 {err:.2e}")
