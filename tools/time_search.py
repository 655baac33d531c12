"""Dev tool: time the two forms of the clip search (admmq_clip_search_sums) on a few shapes."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import _native as nat
g = torch.Generator().manual_seed(0)
cases = [((512, 1141), 4, 200, 0), ((512, 1141), 4, 200, 31), ((512, 1141), 8, 200, 31), ((64, 134), 4, 200, 1),
         ((256, 566), 4, 200, 8), ((4096, 1024), 4, 200, 0), ((512, 1141), 4, 1000, 0), ((512, 1141), 8, 1000, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "crossover":   # elements per CTA x bits: where does the direct form win?
    cases = [((rows, 1141), bits, 200, 36) for bits in (4, 6, 8) for rows in (9, 32, 64, 128, 256, 512)]
for shape, bits, nc, ctas in cases:
    x = (torch.randn(*shape, generator=g) * 0.05).cuda()
    for method in (0, 1):
        for _ in range(3):
            nat.clip_search_sums(x, bits, nc, method, ctas)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            nat.clip_search_sums(x, bits, nc, method, ctas)
        e1.record()
        torch.cuda.synchronize()
        print(f"{shape} bits={bits} nc={nc} ctas={ctas or 148} method={method}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us (3 launches incl. minmax)", flush=True)
