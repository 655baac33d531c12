"""Dev tool: admmq_gemm_nt (3xTF32 tcgen05) vs float64 on the GPU, plus timing."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
def pad4(x):
    K = x.shape[1]; Kp = (K + 3) // 4 * 4
    out = torch.zeros(x.shape[0], Kp); out[:, :K] = x
    return out.cuda()[:, :K] if False else out.cuda()
for (M, N, K) in [(128, 32, 32), (128, 64, 64), (128, 16, 8), (64, 134, 134), (100, 50, 70), (512, 1141, 1141), (256, 566, 566), (9, 1141, 1141), (2048, 204, 204), (300, 300, 1000)]:
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    Ad, Bd = pad4(A), pad4(B)
    Kp = Ad.shape[1]
    Ac = Ad[:, :K] if Kp == K else Ad   # pass padded (zero columns) with logical K
    C = torch.empty(M, N, device="cuda")
    nat.check(nat.lib.admmq_gemm_nt(nat.ptr(Ad), Kp, M, nat.ptr(Bd), Kp, N, K, nat.ptr(C), N, nat.stream_ptr(C.device)))
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T)
    c32 = (A.cuda() @ B.cuda().T).cpu().double()
    err = (C.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    err32 = (c32 - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M:5d} N={N:5d} K={K:5d}: max rel err 3xTF32 {err:.2e}   cuBLAS fp32 {err32:.2e}", flush=True)
M, N, K = 512, 1141, 1144
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nat.lib.admmq_gemm_nt(nat.ptr(A), K, M, nat.ptr(B), K, N, K, nat.ptr(C), N, nat.stream_ptr(C.device))
    e1.record(); torch.cuda.synchronize()
print(f"512x1141x1144: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call = {2.0 * M * N * K / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e12:.1f} TFLOP/s fp32-equivalent")
M, N, K = 4096, 4096, 4096
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        nat.lib.admmq_gemm_nt(nat.ptr(A), K, M, nat.ptr(B), K, N, K, nat.ptr(C), N, nat.stream_ptr(C.device))
    e1.record(); torch.cuda.synchronize()
print(f"4096^3: {e0.elapsed_time(e1) / 5:.3f} ms per call = {2.0 * M * N * K / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12:.1f} TFLOP/s fp32-equivalent")
