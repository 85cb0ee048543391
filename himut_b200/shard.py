"""Sharding across ranks and the final reduction; used by genome.py (chunk runs -> ranks) and bench.py.

The path has no data-path collective: a run of consecutive chunks of a contig is owned by exactly one rank (the
som_seen carry inside it is the worker's; between runs it is replayed at the merge, genome.py).  What crosses ranks at the end is tiny: the 15 / 14
log counters and 2 x 33 trinucleotide bins (all-reduce, sum) and the site records (gathered on
rank 0 for the reference's natsort + writers).  Backend agnostic: NCCL on GPUs, gloo in tests.
"""
import numpy as np


def lpt_assign(weights, world):
    """longest-processing-time assignment: {name: weight} -> {name: rank}, heaviest first onto the
    least loaded rank (ties: lower rank), deterministic"""
    load = [0] * world
    out = {}
    for name, w in sorted(weights.items(), key=lambda kv: (-kv[1], str(kv[0]))):
        r = min(range(world), key=lambda i: (load[i], i))
        out[name] = r
        load[r] += w
    return out


def my_contigs(weights, rank, world):
    a = lpt_assign(weights, world)
    return [c for c in weights if a[c] == rank]


def all_reduce_sum(vec, device=None):
    """sum an int64 vector over ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(vec, np.int64)
    t = torch.as_tensor(np.asarray(vec, np.int64), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def all_reduce_max(vec, device=None):
    """element-wise maximum of an int64 vector over ranks (identity without a process group)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(vec, np.int64)
    t = torch.as_tensor(np.asarray(vec, np.int64), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def gather_dicts(local, dst=0):
    """merge per-contig dicts {chrom: value} onto rank `dst` (None elsewhere)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if bucket is None:
        return None
    merged = {}
    for part in bucket:
        merged.update(part)
    return merged
