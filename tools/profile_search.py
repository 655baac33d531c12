"""Dev tool: a few launches of the clip-search kernel for ncu (k_mse_sums)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
I, R, bits, nc, ctas, method = (int(a) for a in sys.argv[1:7])
x = (torch.randn(I, R, generator=torch.Generator().manual_seed(0)) * 0.05).cuda()
for _ in range(4):
    nat.clip_search_sums(x, bits, nc, method, ctas)
torch.cuda.synchronize()
