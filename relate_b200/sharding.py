"""Host-side sharding of the Paint path over ranks/GPUs (no data-path collective).

Targets are independent given the replicated bit matrix and chunks are independent of each other
(SURVEY.md section 8e), so a job is split either by contiguous target ranges with roughly equal visited-site
counts (one chunk over several GPUs) or by whole chunks (several chunks).  The only communication is
bookkeeping: a barrier and a max/sum over ranks of a handful of scalars, done with torch.distributed
(NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def balanced_target_ranges(site_counts, world: int):
    """Cut targets 0..N-1 into `world` contiguous ranges with ~equal sum of D_k.  -> list of (k_begin, k_end)."""
    c = np.asarray(site_counts, dtype=np.float64)
    N = len(c)
    if world < 1:
        raise ValueError("world must be >= 1")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side="left"))
        k = min(max(k, cuts[-1]), N)
        cuts.append(k)
    cuts.append(N)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def chunks_for_rank(n_chunks: int, rank: int, world: int, sizes=None):
    """Whole chunks per rank: largest first, each to the currently lightest rank (LPT).  -> sorted chunk ids."""
    sizes = np.ones(n_chunks) if sizes is None else np.asarray(sizes, dtype=np.float64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world)
    mine = []
    for c in order:
        r = int(np.argmin(load))
        load[r] += sizes[c]
        if r == rank:
            mine.append(int(c))
    return sorted(mine)


def allreduce_scalars(values, op: str, device=None):
    """max / sum of a list of python floats over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()
