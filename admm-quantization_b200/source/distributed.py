"""Sharding of independent solves over the GPUs of one box and the final gather of their results.

The reference factorizes one (layer, rank, bits, seed) per process invocation and exchanges results through files
(`scripts/factorize.py:315-318, 345-347`); sweeps over layers / reduction rates / bit-widths are external shell loops.
Here one process per GPU (torchrun) takes its share of the units by longest-processing-time-first, runs them with no
data-path collective, and the packed factors are gathered ONCE at the end (NCCL over NVLink on GPUs; gloo in the CPU
tests).  Because no arithmetic crosses a shard boundary, N-GPU results are bitwise identical to 1-GPU results.
"""
import torch
import torch.distributed as dist

from .workloads import lpt_assign, unit_cost


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(units, world_size):
    """units: list of dicts with 'shape' and 'rank' (and optionally 'bits', 'num_attempts').  Returns the owner rank of each."""
    costs = [unit_cost(u["shape"], u["rank"], u.get("bits", 4), u.get("num_attempts", 200)) for u in units]
    return lpt_assign(costs, world_size)


def pack_factors(factors):
    """Concatenate float32 factor matrices into one flat buffer + their shapes."""
    shapes = [tuple(f.shape) for f in factors]
    flat = torch.cat([f.reshape(-1).to(torch.float32) for f in factors]) if factors else torch.empty(0)
    return flat, shapes


def unpack_factors(flat, shapes):
    out, off = [], 0
    for shp in shapes:
        n = 1
        for d in shp:
            n *= d
        out.append(flat[off:off + n].reshape(shp))
        off += n
    return out


def gather_results(local, device=None, dst=0):
    """local: {unit_key: {'factors': [tensors], 'loss': [...], 'loss_quant': [...]}} of this rank.
    Returns the merged dict on `dst` (None elsewhere).  Metadata travels as Python objects (tiny), the factor payload
    as ONE padded float32 buffer per rank through `dist.gather` / `all_gather` (NCCL needs equal sizes)."""
    rank, ws = world()
    if ws == 1:
        return dict(local)
    keys = sorted(local)
    flats, meta = [], []
    for k in keys:
        flat, shapes = pack_factors(local[k]["factors"])
        flats.append(flat)
        meta.append((k, shapes, list(local[k].get("loss", [])), list(local[k].get("loss_quant", []))))
    dev = device if device is not None else (flats[0].device if flats else torch.device("cpu"))
    payload = torch.cat(flats).to(dev) if flats else torch.empty(0, device=dev)
    all_meta = [None] * ws
    dist.all_gather_object(all_meta, (meta, payload.numel()))
    width = max(m[1] for m in all_meta)
    padded = torch.zeros(max(width, 1), dtype=torch.float32, device=dev)
    padded[:payload.numel()] = payload
    bufs = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(bufs, padded)           # one collective for the whole job
    if rank != dst:
        return None
    merged = {}
    for r in range(ws):
        off = 0
        for k, shapes, loss, lossq in all_meta[r][0]:
            n = sum(int(torch.tensor(s).prod()) for s in shapes)
            merged[k] = {"factors": unpack_factors(bufs[r][off:off + n].cpu(), shapes), "loss": loss, "loss_quant": lossq}
            off += n
    return merged
