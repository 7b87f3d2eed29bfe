"""Multi-rank plumbing of bench.py (one process per GPU): no data-path collective.

Targets are independent given the replicated bit matrix and chunks are independent of each other.
Inside one process the C library shards a chunk's targets over devices (rp_paint_chunk) or whole chunks over devices
(rp_paint_chunks) by itself; across ranks every rank paints its own chunk.  The only communication is bookkeeping: a
barrier and a max / sum over ranks of a handful of scalars through torch.distributed (NCCL on the GPU box, gloo in the
CPU test), plus a CPU-side (gloo) group on which ranks can wait without parking a spinning kernel on their GPU while
rank 0 drives all devices for the strong-scaling leg.
"""
from __future__ import annotations


def allreduce_scalars(values, op: str, device=None):
    """max / sum of a list of python floats over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()


def cpu_barrier_group():
    """A gloo process group over all ranks (None when not distributed): `dist.barrier(group=...)` on it blocks on the host."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.new_group(backend="gloo")
