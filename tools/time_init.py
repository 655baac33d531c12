"""Dev tool: wall time of the ALS+EPC initialisation (init_factors(..., 'parafac-epc')) per layer shape."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source.admm import init_factors
from source.solver import rank_from_reduction_rate
shapes = [(64, 64, 9), (128, 128, 9), (256, 256, 9), (512, 512, 9)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for shp in shapes:
    g = torch.Generator().manual_seed(42)
    W = (torch.randn(*shp, generator=g) * (2.0 / (shp[1] * 9)) ** 0.5).cuda()
    R = rank_from_reduction_rate(W, 2.0)
    np.random.seed(42)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fac = init_factors(W, R, init="parafac-epc", device="cuda", seed=42)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec = torch.einsum("ir,jr,kr->ijk", *fac)
    err = float(torch.linalg.norm(W - rec) / torch.linalg.norm(W))
    print(f"{shp} R={R}: parafac-epc init {t1 - t0:7.2f} s, rel. error of the init {err:.4f}", flush=True)
