"""Dev tool: n copies of one layer's sweep on n streams with g CTAs each vs one copy alone."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "admm-quantization_b200")]
from source import workloads as wl
from source.solver import LayerSolver
which, n, g, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
(name, W, rank, init), = wl.build_problems([l for l in wl.resnet18_conv_layers() if l[0] == which])
solvers = [LayerSolver(W.cuda(), [f.cuda() for f in init], 4, "tensor_mseminmax_symmetric", max_iter_admm=iters + 1,
                       solve_precision=1, max_ctas=g, time_loops=True) for _ in range(n)]
streams = [torch.cuda.Stream() for _ in range(n)]
def run(k):
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    joins = []
    for s, st in list(zip(solvers, streams))[:k]:
        st.wait_event(e0)
        with torch.cuda.stream(st):
            s.enqueue_sweep()
            j = torch.cuda.Event(enable_timing=True); j.record(st); joins.append(j)
        main.wait_event(j)
    e1.record(main); torch.cuda.synchronize()
    for s in solvers[:k]: s.collect()
    loops = [sum(a.elapsed_time(b) for _, a, b in s.loop_events[-3:]) for s in solvers[:k]]
    return e0.elapsed_time(e1), [e0.elapsed_time(j) for j in joins], loops
for k in (1, n, 1, n):
    tot, per, loops = run(k)
    print(f"{name} x{k} @ {g} CTAs, {iters} inner its: total {tot:.1f} ms; per stream {[round(x,1) for x in per]}; loop kernels {[round(x,1) for x in loops]}", flush=True)
