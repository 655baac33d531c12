"""W ~ W_q + W_r for one weight matrix - the CLI of the reference's scripts/factorize_lowrank.py (same flags and output
file names).  There is no network for Hugging Face checkpoints, so without a local `--weight-file` (a torch-saved
2-D tensor) the matrix is synthetic with the Llama-like scale of BASELINE config 5 (`--synthetic OUT IN`).

    python admm-quantization_b200/scripts/factorize_lowrank.py --layer model.layers.0.self_attn.q_proj \
        --max-iter 20 --bits 4 --rank 16 --output-dir out --synthetic 4096 4096
"""
import argparse
import logging
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
logging.basicConfig(format="%(asctime)s - %(message)s", level=logging.INFO)

from source.lowrank import factorize_lowrank  # noqa: E402
from source.quantization import quantize_tensor  # noqa: E402


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-id", type=str, help="[huggyllama/llama-7b, ] (needs a local checkpoint; see --weight-file)")
    ap.add_argument("--cache-dir", type=str)
    ap.add_argument("--output-dir", type=str, default=".")
    ap.add_argument("--with-wandb", action="store_true")
    ap.add_argument("--layer", type=str, default="layer")
    ap.add_argument("--max-iter", required=True, type=int)
    ap.add_argument("--bits", required=True, type=int)
    ap.add_argument("--rank", required=True, type=int)
    ap.add_argument("--qscheme", default="tensor_minmax", type=str)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--weight-file", type=str, default=None, help="torch.save'd 2-D float tensor to split")
    ap.add_argument("--synthetic", type=int, nargs=2, default=None, metavar=("OUT", "IN"))
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("factorize_lowrank.py needs a CUDA device (libadmmq has no CPU fallback)")
    device = torch.device("cuda:0")
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    random.seed(args.seed)
    if args.weight_file:
        W = torch.load(args.weight_file).to(device=device, dtype=torch.float32)
    elif args.synthetic:
        g = torch.Generator().manual_seed(args.seed)
        W = (torch.randn(*args.synthetic, generator=g) * 0.02).to(device)
    else:
        raise SystemExit("no network for Hugging Face checkpoints here: pass --weight-file or --synthetic OUT IN")
    if W.ndim != 2:
        raise ValueError("Incorrect number of dimentions in weight tensor")
    run = None
    if args.with_wandb:
        import wandb
        run = wandb.init(config=args, name=f"b={args.bits}_r={args.rank}_s={args.seed}")
    logging.info(f"Bits: {args.bits}, Rank: {args.rank}")
    abs_q = torch.linalg.norm(W - quantize_tensor(W, qscheme=args.qscheme, bits=args.bits))
    logging.info(f"Diff between W and quantized W abs: {abs_q:.4f}, rel: {abs_q / torch.linalg.norm(W):.4f}")

    def log(i, rel):
        logging.info(f"Diff between W and (W_q + W_r) rel: {rel:.4f}")
        if run:
            run.log({"rel_admm_diff": rel})

    W_q, W_r, hist = factorize_lowrank(W, args.bits, args.rank, args.qscheme, args.max_iter, args.seed, log=log)
    os.makedirs(args.output_dir, exist_ok=True)
    rel = hist[-1]
    torch.save(W_q.cpu(), os.path.join(args.output_dir, f"{args.layer}_{args.bits}_{args.rank}_{rel:.3f}_Q.pt"))
    torch.save(W_r.cpu(), os.path.join(args.output_dir, f"{args.layer}_{args.bits}_{args.rank}_{rel:.3f}_R.pt"))
    if run:
        run.finish()
    return rel


if __name__ == "__main__":
    main()
