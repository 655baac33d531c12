#!/usr/bin/env python
"""Which inner loops leave early (r < eps and s < eps, source/admm.py:62-65), and with what state?

    python tools/early_exit_probe.py --sweeps 26 --precision 1 [--dump gpurun_out/early_exit]

Runs the bench workload layer by layer (every layer on all SMs, same kernels as bench.py) for `--sweeps` outer sweeps,
prints one line per loop that stopped before max_iter_admm - 1 iterations - (layer, mode, sweep, iterations, r, s) -
and, with --dump, saves the state ENTERING the few smallest such calls (H, U, F, G and what the kernel reported) as
.npz, so that the same call can be replayed through the unmodified reference on the CPU (oracle/make_golden.py
--only early_exit) and pinned as a fixture.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "admm-quantization_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from source import workloads as wl  # noqa: E402
from source.solver import LayerSolver  # noqa: E402

QS = "tensor_mseminmax_symmetric"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=26)
    ap.add_argument("--precision", type=int, default=1, help="solve precision; the MTTKRP follows (1 -> tensor cores)")
    ap.add_argument("--layers", nargs="*", default=None)
    ap.add_argument("--dump", default=None)
    ap.add_argument("--max-dump", type=int, default=6)
    ap.add_argument("--max-dump-elems", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    layers = wl.resnet18_conv_layers()
    if args.layers:
        layers = [l for l in layers if l[0] in args.layers]
    problems = wl.build_problems(layers, 2.0, weight_seed=42, init_seed=args.seed)
    found, dumped = [], 0
    if args.dump:
        os.makedirs(args.dump, exist_ok=True)
    for name, W, rank, init in problems:
        s = LayerSolver(W.to(dev), [f.to(dev) for f in init], 4, QS, max_iter_admm=1000,
                        mttkrp_precision=1 if args.precision == 1 else 0, solve_precision=args.precision)
        s.snapshot_inputs = True
        for sweep in range(args.sweeps):
            err, errq = s.sweep()
            for m, r in enumerate(s.last_reports):
                if r.iterations < 999:
                    rec = dict(layer=name, mode=m, sweep=sweep, iterations=int(r.iterations), r=float(r.r), s=float(r.s),
                               shape=list(s.factors[m].shape), status=int(r.status), precision=args.precision)
                    found.append(rec)
                    print("EARLY", json.dumps(rec), flush=True)
                    H, U, F, G = s.snapshots[m]
                    if args.dump and dumped < args.max_dump and H.numel() <= args.max_dump_elems:
                        np.savez_compressed(os.path.join(args.dump, f"p{args.precision}_{name}_m{m}_s{sweep}.npz"),
                                            H=H.cpu().numpy(), U=U.cpu().numpy(), F=F.cpu().numpy(), G=G.cpu().numpy(),
                                            Hout=s.factors[m].cpu().numpy(), meta=np.frombuffer(json.dumps(rec).encode(), dtype=np.uint8))
                        dumped += 1
        print(f"{name}: rec_error {s.loss_hist[-1]:.6f} after {args.sweeps} sweeps", flush=True)
    print(f"precision {args.precision}: {len(found)} early exits in {len(problems) * args.sweeps * 3} loops")
    if args.dump:
        with open(os.path.join(args.dump, f"summary_p{args.precision}.json"), "w") as f:
            json.dump(found, f, indent=1)


if __name__ == "__main__":
    main()
