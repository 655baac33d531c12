#!/usr/bin/env python
"""Whole-model / sweep driver (new: the reference has only the per-layer CLI and external shell loops).

    torchrun --nproc-per-node 8 scripts/factorize_model.py --model-name resnet18 --bits 4 --reduction-rate 2 \
        --qscheme tensor_mseminmax_symmetric --seed 42 --max_iter_als 1000

Enumerates layers x reduction rates x bit-widths (BASELINE configs 2 and 3), shards the independent solves over the
ranks (LPT), runs them with `LayerSolver`, gathers the factors once, and writes the reference's file layout
(`{bits}bit_{qscheme}/factors_admm_seed{seed}/{layer}_admm_{init}_rank_{rank}_mode_{m}.pt`) on rank 0.
"""
import os
import sys
import time
from argparse import ArgumentParser

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from source import workloads as wl  # noqa: E402
from source.admm import init_factors  # noqa: E402
from source.distributed import gather_results, shard_units  # noqa: E402
from source.solver import LayerSolver, layer_weight_as_tensor, rank_from_reduction_rate  # noqa: E402


def parse_args(argv=None):
    ap = ArgumentParser()
    ap.add_argument("--model-name", default="resnet18", choices=["resnet18"])
    ap.add_argument("--bits", type=int, nargs="+", default=[4])
    ap.add_argument("--reduction-rate", type=float, nargs="+", default=[2.0])
    ap.add_argument("--qscheme", default="tensor_mseminmax_symmetric")
    ap.add_argument("--init", default="random")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--max_iter_als", type=int, default=5000)
    ap.add_argument("--max_iter_admm", type=int, default=1000)
    ap.add_argument("--solve-precision", type=int, default=1)
    ap.add_argument("--outroot", default=".")
    ap.add_argument("--layers", nargs="*", default=None, help="subset of layer names")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("factorize_model.py needs CUDA devices: the solver has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    layers = wl.resnet18_conv_layers()
    if args.layers:
        layers = [l for l in layers if l[0] in args.layers]
    units = []
    for name, cout, cin, kh, kw in layers:
        W = layer_weight_as_tensor(wl.synthetic_weight(cout, cin, kh, kw, args.seed, name)).contiguous()
        for rr in args.reduction_rate:
            r = rank_from_reduction_rate(W, rr)
            for bits in args.bits:
                units.append({"key": (name, rr, bits), "W": W, "shape": tuple(W.shape), "rank": r, "bits": bits})
    owner = shard_units(units, world)
    t0 = time.time()
    results = {}
    for u, o in zip(units, owner):
        if o != rank:
            continue
        Wd = u["W"].to(dev)
        factors = init_factors(Wd, rank=u["rank"], init=args.init, device=dev, seed=args.seed)
        s = LayerSolver(Wd, factors, u["bits"], args.qscheme, max_iter_admm=args.max_iter_admm,
                        init_is_random=(args.init == "random"), solve_precision=args.solve_precision)
        sweeps = s.run(args.max_iter_als)
        results[u["key"]] = {"factors": [f.clone() for f in s.factors], "loss": s.loss_hist, "loss_quant": s.loss_quant_hist}
        print(f"[rank {rank}] {u['key']} rank {u['rank']}: {sweeps} sweeps, rec_error {s.loss_hist[-1]:.6f}, "
              f"quant {s.loss_quant_hist[-1]:.6f}", flush=True)
    merged = gather_results(results, device=dev)
    if rank == 0:
        for (name, rr, bits), res in merged.items():
            r = res["factors"][0].shape[1]
            outdir = os.path.join(args.outroot, f"{bits}bit_{args.qscheme}", f"factors_admm_seed{args.seed}")
            os.makedirs(outdir, exist_ok=True)
            prefix = f"{name}_admm_{args.init}_rank_{r}"
            for m, f in enumerate(res["factors"]):
                torch.save(f.clone(), os.path.join(outdir, prefix + f"_mode_{m}.pt"))
            torch.save(res["loss"], os.path.join(outdir, prefix + "_losshist.pt"))
            torch.save(res["loss_quant"], os.path.join(outdir, prefix + "_lossquanthist.pt"))
        print(f"{len(merged)} solves on {world} GPU(s) in {time.time() - t0:.1f} s")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
