"""Dev tool: ALS / EPC seconds, passes and factorization counts of parafac_epc per layer shape.   python tools/time_epc_info.py [nshapes]"""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source.parafac_epc import parafac_epc
from source.solver import rank_from_reduction_rate
shapes = [(64, 64, 9), (128, 128, 9), (256, 256, 9), (512, 512, 9)]
if len(sys.argv) > 1:
    shapes = [shapes[int(a)] for a in sys.argv[1:]]
for shp in shapes:
    g = torch.Generator().manual_seed(42)
    W = (torch.randn(*shp, generator=g) * (2.0 / (shp[1] * 9)) ** 0.5).cuda()
    R = rank_from_reduction_rate(W, 2.0)
    info = {}
    np.random.seed(42)
    lam, Us = parafac_epc(W, R, als_maxiter=50, epc_maxiter=50, epc_rounds=50, info=info)   # budgets of source/admm.py:42-43
    n_up = 3 * info["epc_passes"]
    print(f"{shp} R={R}: ALS {info['als_s']:.2f} s, EPC {info['epc_s']:.2f} s = {info['epc_passes']} passes in {info['epc_rounds']} rounds "
          f"({info['epc_s'] / info['epc_passes'] * 1e3:.2f} ms per pass), {info['epc_chol_evals']} factorizations + {info['epc_eigh_updates']} eigen-form "
          f"updates in {n_up} mode updates", flush=True)
