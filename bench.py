#!/usr/bin/env python
"""Benchmark of the quantization-aware CP factorization hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): every 3x3 conv layer of ResNet-18 (16 layers, synthetic
Kaiming-normal weights, rank from reduction-rate 2.0), ADMM with 4-bit tensor_mseminmax_symmetric
projection, max_iter_admm = 1000.  One STEP = one outer sweep over all 16 layers = 16 x 3 factor updates
x 999 inner ADMM iterations (+ Gram, MTTKRP, ridge inverse, re-projection, two reconstruction errors).
Metric: inner ADMM iterations per second (the unit of work is one inner iteration of one
factor).  With N GPUs every rank factorizes the whole layer set for its own seed (the reference's
`--seed` axis: independent solves, no data-path collective) -> weak scaling; the factors are gathered
once with NCCL after the timed region.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# One hardware work queue per layer stream: with the default of 8 connections, streams that share a queue falsely
# serialise behind each other's long-running persistent kernels (measured: the last-launched layers started 0.4 s late).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "admm-quantization_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

QSCHEME = "tensor_mseminmax_symmetric"
METRIC = "admm_inner_iter_per_s"
UNIT = "inner ADMM iterations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="resnet18", choices=["resnet18", "layer1", "resnet50-l4", "llama7b"])
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--reduction-rate", type=float, default=2.0)
    ap.add_argument("--max-iter-admm", type=int, default=1000)
    ap.add_argument("--cpu-budget-s", type=float, default=15.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mttkrp-precision", type=int, default=1,
                    help="0 = float64-accumulating CUDA-core MTTKRP, 1 = 3xTF32 tcgen05 MTTKRP (default)")
    ap.add_argument("--concurrency", default="prop", choices=["off", "prop"],
                    help="off: layers run one after the other on every SM; prop: every layer gets a share of the SMs "
                         "proportional to its cost and all layers run concurrently on their own streams")
    ap.add_argument("--min-ctas", type=int, default=2)
    ap.add_argument("--balance", default="model", choices=["model", "prop"],
                    help="re-balancing of the SM budgets during warm-up: 'model' = wave model of the tensor-core product "
                         "fitted to the reported phase times (source/workloads.py), 'prop' = proportional to time x CTAs")
    ap.add_argument("--budgets", default=None,
                    help="experiment knob: comma-separated SM budgets, one per layer in workload order; fixes the budgets "
                         "(no re-balancing during warm-up)")
    ap.add_argument("--trace-layer", default=None, help="print the per-kernel-group times of this layer's last sweep")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="SMs kept out of the cooperative-grid budgets so that the ordinary kernels between the loops "
                         "(Gram, MTTKRP, projection, errors) never wait for a persistent kernel to finish")
    ap.add_argument("--solve-precision", type=int, default=1,
                    help="ridge product inside the ADMM loop: 0 = float32 FFMA, 1 = 3xTF32 on tcgen05 (default)")
    return ap.parse_args()


def workload_layers(name):
    from source import workloads as wl
    if name == "resnet18":
        return wl.resnet18_conv_layers(), "ResNet-18 all 16 3x3 conv layers"
    if name == "layer1":
        return wl.resnet18_conv_layers()[:1], "ResNet-18 layer1.0.conv1 (64x64x3x3)"
    if name == "resnet50-l4":
        return wl.resnet50_layer4_layers(), "ResNet-50-shaped layer4 (512x512x3x3, 2048x512 1x1)"
    return wl.llama7b_linear_layers(), "Llama-7B-shaped linear (4096x4096, 11008x4096)"


def config_dict(args, desc, extra=None):
    cfg = {"workload": f"{desc}, ADMM {args.bits}-bit {QSCHEME}, reduction-rate {args.reduction_rate}, init=random, "
                       f"max_iter_admm={args.max_iter_admm}; one step = one outer sweep over every layer",
           "bits": args.bits, "qscheme": QSCHEME, "reduction_rate": args.reduction_rate,
           "max_iter_admm": args.max_iter_admm, "num_attempts": 200,
           "ridge_product": "3xTF32 tcgen05" if getattr(args, "solve_precision", 0) == 1 else "float32 FFMA"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 9 and r[5 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(args, layers, budget_s, threads=None):
    """Times the reference algorithm's inner loop (oracle/admm_oracle.py::admm_iteration = the op-for-op
    restatement of source/admm.py:51-67 + source/quantization.py:118-144 in torch-CPU float32) on every
    distinct factor shape of the workload for a few inner iterations, and converts to the workload's
    inner-iterations/s.  Returns (value, description, threads, seconds spent)."""
    import torch
    from oracle import admm_oracle as orc
    from source import workloads as wl
    nproc = os.cpu_count() or 1
    shapes = {}
    for name, cout, cin, kh, kw in layers:
        dims = (cout, cin, kh * kw) if kh * kw > 1 else (cout, cin)
        numel = 1
        for d in dims:
            numel *= d
        rank = int(numel / sum(dims) / args.reduction_rate)
        for m, d in enumerate(dims):
            others = [x for k, x in enumerate(dims) if k != m]
            key = (d, rank)
            shapes.setdefault(key, [0, others])[0] += 1
    g = torch.Generator().manual_seed(7)

    def one(I, R, others, iters):
        mats = [torch.randn(o, R, generator=g) for o in others]
        G = orc.gram_hadamard(mats)
        scale = 1.0
        for o in others:
            scale *= o
        F = torch.randn(I, R, generator=g) * scale ** 0.5
        H = torch.randn(I, R, generator=g)
        U = torch.zeros(I, R)
        t0 = time.perf_counter()
        orc.admm_iteration(H, U, F, G, iters + 1, 1e-8, args.bits, QSCHEME)
        return (time.perf_counter() - t0) / iters

    # thread count: the 200-pass projection is a chain of small elementwise ops; pick what is fastest here
    if threads is None:
        big = max(shapes, key=lambda k: k[0] * k[1])
        best = None
        for t in sorted({1, max(1, nproc // 2), nproc}):
            torch.set_num_threads(t)
            one(big[0], big[1], shapes[big][1], 1)
            dt = one(big[0], big[1], shapes[big][1], 1)
            if best is None or dt < best[1]:
                best = (t, dt)
        threads = best[0]
    torch.set_num_threads(threads)
    t_start = time.perf_counter()
    total_s_per_sweep = 0.0
    sampled = 0
    share = budget_s / len(shapes)          # equal CPU time per distinct shape
    for (I, R), (count, others) in sorted(shapes.items()):
        probe = one(I, R, others, 2)
        iters = int(min(60, max(2, round((share - 2 * probe) / probe))))
        per_iter = one(I, R, others, iters)
        sampled += iters + 2
        total_s_per_sweep += count * (args.max_iter_admm - 1) * per_iter
    inner_per_sweep = sum(c for c, _ in shapes.values()) * (args.max_iter_admm - 1)
    spent = time.perf_counter() - t_start
    desc = (f"oracle admm_iteration (torch-CPU float32, {threads} threads): {sampled} inner iterations spread over the "
            f"{len(shapes)} distinct factor shapes of the workload (Cholesky set-up included), extrapolated to one sweep "
            f"= {inner_per_sweep} inner iterations; Gram/MTTKRP/error terms (<1 % of CPU time) not included")
    return inner_per_sweep / total_s_per_sweep, desc, threads, spent


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    layers, desc = workload_layers(args.workload)
    vals, threads = [], None
    t0 = time.perf_counter()
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample, threads, _ = cpu_reference_sample(args, layers, 6.0, threads)
        if i >= args.warmup:
            vals.append(v)
    value = len(vals) / sum(1.0 / v for v in vals)
    inner = sum(3 if l[3] * l[4] > 1 else 2 for l in layers) * (args.max_iter_admm - 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": inner / value * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, desc),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
            "note": "ms_per_step is the extrapolated CPU time of one full sweep; each timed step ran a bounded sample"}
    print(json.dumps(line), flush=True)
    return 0


def allocate_ctas(costs, sm_count, min_ctas):
    """Share of the SMs for every independent solve, proportional to its estimated cost (largest-remainder rounding,
    at least `min_ctas` each, sum == sm_count when there are enough SMs)."""
    n = len(costs)
    total = float(sum(costs))
    if n * min_ctas >= sm_count:
        return [max(1, sm_count // n)] * n
    spare = sm_count - n * min_ctas
    raw = [c / total * spare for c in costs]
    out = [min_ctas + int(r) for r in raw]
    rest = sm_count - sum(out)
    for i in sorted(range(n), key=lambda k: raw[k] - int(raw[k]), reverse=True)[:rest]:
        out[i] += 1
    return out


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from source import _native, workloads as wl
    from source.solver import LayerSolver
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the solver has no CPU path (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    layers, desc = workload_layers(args.workload)
    # weak scaling: every rank factorizes the whole layer set for its own seed (reference: one run per --seed)
    problems = wl.build_problems(layers, args.reduction_rate, weight_seed=42, init_seed=42 + rank)
    host = []   # pinned host state for the end-to-end leg
    solvers = []
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    budgets = allocate_ctas([wl.solve_cost(W.shape, rnk) for _, W, rnk, _ in problems], sm_count - args.reserve_sms,
                            args.min_ctas) \
        if args.concurrency == "prop" else [0] * len(problems)
    if args.budgets and args.concurrency == "prop":
        budgets = [int(x) for x in args.budgets.split(",")]
        assert len(budgets) == len(problems) and sum(budgets) <= sm_count, (len(problems), sum(budgets), sm_count)
    streams = [torch.cuda.Stream(device=dev) for _ in problems] if args.concurrency == "prop" else None
    for (name, W, rnk, init), g in zip(problems, budgets):
        solvers.append(LayerSolver(W.to(dev), [f.to(dev) for f in init], args.bits, QSCHEME,
                                   max_iter_admm=args.max_iter_admm, mttkrp_precision=args.mttkrp_precision,
                                   solve_precision=args.solve_precision, time_loops=True, max_ctas=g))
        host.append({"W": W.pin_memory(), "factors": [f.clone().pin_memory() for f in init],
                     "duals": [torch.zeros_like(f).pin_memory() for f in init],
                     "factors_q": [torch.zeros_like(f).pin_memory() for f in init],
                     "err": torch.zeros(2, 2, dtype=torch.float64).pin_memory()})
    for (name, _, _, _), s in zip(problems, solvers):
        if args.trace_layer and name == args.trace_layer:
            s.part_events = []
    inner_per_step = sum(s.inner_iterations_per_sweep() for s in solvers)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sweep_events = []

    def enqueue_all(prepare=None):
        """One sweep of every layer.  Concurrent mode: every layer on its own stream (fork from / join into the
        current stream with events, so CUDA events recorded on the current stream bracket all of the work)."""
        if streams is None:
            for k, s in enumerate(solvers):
                if prepare is not None:
                    prepare(k)
                s.enqueue_sweep()
            return
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event(enable_timing=True)
        fork.record(main)
        sweep_events.clear()
        for k, (s, st) in enumerate(zip(solvers, streams)):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                if prepare is not None:
                    prepare(k)
                s.enqueue_sweep()
                join = torch.cuda.Event(enable_timing=True)
                join.record(st)
            main.wait_event(join)
            sweep_events.append((fork, join))

    def step():
        enqueue_all()
        return [s.collect() for s in solvers]

    def step_e2e():
        """The same sweep through the public solver API starting from HOST buffers: weights, factors and
        duals go host->device (pinned), results (factors, duals, re-projected factors, error sums) come back."""
        moved = [0, 0]

        def h2d_copy(k):
            moved[0] += solvers[k].load_from_host(host[k]["W"], host[k]["factors"], host[k]["duals"])

        enqueue_all(prepare=h2d_copy)
        for s, h in zip(solvers, host):   # results come back on the (joined) current stream
            moved[1] += s.store_to_host(h["factors"], h["duals"], h["factors_q"], h["err"])
        torch.cuda.synchronize()
        return moved[0], moved[1]

    tried = []   # (step ms, budgets) of every warm-up step: the timed steps run with the best MEASURED budgets
    if streams is not None:
        step()   # one-time costs (module load, kernel attributes, allocator) stay out of the comparison below
    for w in range(args.warmup):
        flush.fill_(1)
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        step()
        w1.record()
        torch.cuda.synchronize()
        tried.append((w0.elapsed_time(w1), [s.max_ctas for s in solvers]))
        if streams is not None and w < args.warmup - 1 and not args.budgets:
            # re-balance the SM budgets from what was just measured: per layer the sweep time and, per factor, the
            # phase times its persistent kernel reported, through the wave model of source/workloads.py
            fitted = []
            for (a, b), s in zip(sweep_events, solvers):
                modes = []
                for f, r in zip(s.factors, s.last_reports):
                    it = max(r.iterations, 1)
                    modes.append((f.shape[0], s.R, r.phase_ns[0] / 1e3 / it, (r.phase_ns[1] + r.phase_ns[2]) / 1e3 / it, it))
                fitted.append((a.elapsed_time(b), max(s.max_ctas, 1), modes))
            new = wl.allocate_ctas_modelled(fitted, sm_count - args.reserve_sms, args.solve_precision) \
                if args.balance == "model" else \
                allocate_ctas([m * g0 for m, g0, _ in fitted], sm_count - args.reserve_sms, 1)
            for s, g in zip(solvers, new):
                s.max_ctas = g
    if streams is not None and len(tried) > 1:
        best_ms, best = min((ms, b) for ms, b in tried)
        for s, g in zip(solvers, best):
            s.max_ctas = g
    budgets[:] = [s.max_ctas for s in solvers]
    for s in solvers:
        s.loop_events.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    step_ms = []
    # inner iterations actually executed in the timed steps, per layer and factor, from the kernels' device reports (the
    # loop may leave early: r < eps and s < eps, source/admm.py:64-65) - the metric counts these, not the nominal 999
    done = [[0] * s.N for s in solvers]
    early = nonfinite = 0
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)                      # L2 flush between timed steps (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        errs = step()
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        for k, sv in enumerate(solvers):
            for m, r in enumerate(sv.last_reports):
                done[k][m] += int(r.iterations)
                early += int(r.iterations) < sv.max_iter_admm - 1
                nonfinite += bool(int(r.status) & _native.ST_NONFINITE)
    barrier()
    done_total = sum(sum(d) for d in done)
    launches = _native.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    done_all = torch.tensor([done_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(done_all, op=dist.ReduceOp.SUM)
    value = float(done_all.item()) / (total_ms / 1e3)

    # ---- the dominant kernel (persistent ADMM loop), timed with CUDA events on its own stream.  In concurrent mode a
    # launch holds only its share g of the SMs, so its duration is weighted by g / SMs: the sum is the time the whole
    # GPU would have been busy with these launches (equals the plain sum when every launch uses every SM).
    loop_ms, loop_gpu_ms, per_layer = 0.0, 0.0, {}
    for k, ((name, _, _, _), s) in enumerate(zip(problems, solvers)):
        ms = sum(a.elapsed_time(b) for _, a, b in s.loop_events)
        loop_ms += ms
        loop_gpu_ms += ms * ((s.max_ctas or sm_count) / sm_count)
        per_layer[name] = round(sum(done[k]) / (ms / 1e3), 1)
    n_loop = sum(len(s.loop_events) for s in solvers)
    # algorithmic work of the iterations that ran
    alg_bytes = sum(it * (16 * f.shape[0] * s.R + 4 * s.R * s.R) for s, d in zip(solvers, done) for f, it in zip(s.factors, d))
    evals = sum(it * (s.num_attempts if s.qscheme == QSCHEME else 0) * f.shape[0] * s.R
                for s, d in zip(solvers, done) for f, it in zip(s.factors, d))
    flops = sum(it * 2 * f.shape[0] * s.R * s.R for s, d in zip(solvers, done) for f, it in zip(s.factors, d))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_achieved = alg_bytes / (loop_gpu_ms / 1e3) / 1e9
    # The kernel's dominant phase is the ridge product H_ls = RHS . Minv (2 I R^2 flop per inner iteration, float32
    # accuracy through three TF32 tensor-core products per term), so it is reported against the tensor roofline:
    # dense TF32 = bf16 / 2, 3xTF32 = TF32 / 3, from the MEASURED sustained bf16 rate (the kernel runs inside a long step).
    bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0 * 0.62)))
    tensor_peak = bf16 / 2.0 / 3.0
    achieved = flops / (loop_gpu_ms / 1e3) / 1e12
    roofline = {"kernel": "k_admm_loop (persistent ADMM inner loop: ridge product on tcgen05 + clip search + dual update)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32) / 3 (3xTF32 split), fp32-equivalent" if peaks
                                else "fallback: 0.62 x 2250 bf16 TFLOP/s / 6"),
                "traffic": (traffic or {}).get("bytes") if isinstance(traffic, dict) else traffic,
                "traffic_note": (traffic or {}).get("note") if isinstance(traffic, dict) else None, "launches": n_loop, "avg_launch_ms": loop_ms / max(n_loop, 1),
                "gpu_share_weighted_ms": loop_gpu_ms, "share_of_step": loop_gpu_ms / sum(step_ms),
                "algorithmic_flops_per_launch": flops / max(n_loop, 1),
                "algorithmic_bytes_per_launch": alg_bytes / max(n_loop, 1),
                "note": "achieved = 2 I R^2 flop per inner iteration (fp32-equivalent; the hardware executes 3x that in "
                        "TF32) / (launch duration x the launch's share of the SMs); the launch also contains the clip "
                        "search, the dual update and three device-wide barriers per iteration, so frac is a lower bound "
                        "for the product itself (tools/microbench/tcprof.cu times the product alone)",
                "hbm": {"achieved_gbs": hbm_achieved, "peak_gbs": hbm_peak, "frac": hbm_achieved / hbm_peak,
                        "note": "state is L2-resident by design (ncu: see traffic), HBM is not the binding resource"},
                "clip_search": {"candidates_x_elements_per_s": evals / (loop_gpu_ms / 1e3),
                                "note": "threshold form: O(1) work per element + (2^bits - 1) x candidates thresholds per CTA"}}

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        t_e2e = []
        h2d = d2h = 0
        done_e2e = 0
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h2d, d2h = step_e2e()
            e1.record()
            torch.cuda.synchronize()
            t_e2e.append(e0.elapsed_time(e1))
            done_e2e += sum(int(_native.read_report(r).iterations) for sv in solvers for r in sv.reports_dev)
        barrier()
        tot = torch.tensor([sum(t_e2e)], dtype=torch.float64, device=dev)
        cnt = torch.tensor([done_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        e2e = {"value": float(cnt.item()) / (float(tot.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(tot.item()) / args.steps}

    # ---- final factor gather over NCCL (once per job, outside the timed region)
    gather_ms = None
    if world > 1:
        packed = torch.cat([f.reshape(-1) for s in solvers for f in s.factors])
        out = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.gather(packed, out, dst=0)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t0) * 1e3

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sample, threads, spent = cpu_reference_sample(args, layers, args.cpu_budget_s)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "host_cores": os.cpu_count(), "seconds": round(spent, 1)}
    if rank == 0 and args.trace_layer:
        for (name, _, _, _), s in zip(problems, solvers):
            if s.part_events:
                sys.stderr.write(f"[trace {name}] " + ", ".join(f"{k} {v:.1f}" for k, v in s.part_times_ms()) + "\n")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, desc, {"l2": "flushed between timed steps (256 MiB write)",
                                                   "concurrency": args.concurrency,
                                                   "ctas_per_layer": dict(zip([p[0] for p in problems], budgets)),
                                                   "parallelism": f"{world} independent seeds, one per GPU; no data-path collective",
                                                   "inner_iterations_per_step_per_gpu": done_total / args.steps,
                                                   "inner_iterations_nominal_per_step": inner_per_step,
                                                   "loops_left_early": f"{early} of {args.steps * sum(sv.N for sv in solvers)} "
                                                                       "(exit test r < eps and s < eps, eps = 1e-8)",
                                                   "loops_nonfinite": nonfinite}),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "per_layer_inner_iter_per_s": per_layer,
                "per_layer_sweep_ms_last_step": {p[0]: round(a.elapsed_time(b), 1) for p, (a, b) in zip(problems, sweep_events)},
                "rec_error_last_step": {n[0]: round(e[0], 6) for n, e in zip(problems, errs)},
                "factor_gather_ms": gather_ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.steps < 1 or args.warmup < 0:
        raise SystemExit("--steps >= 1 and --warmup >= 0")
    sys.exit(run_reference(args) if args.impl == "reference" else run_native(args))


if __name__ == "__main__":
    main()
