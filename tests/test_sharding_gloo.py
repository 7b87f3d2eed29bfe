"""world_size-2 gloo test of the multi-rank host logic bench.py uses (run on CPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relate_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what bench.py does with N ranks: every rank paints its own chunk (here: a stand-in count of visited sites and a
    # device time), then max / sum over ranks, and a host-side group to wait on while rank 0 runs the strong-scaling leg
    counts = np.random.default_rng(7 + rank).integers(50, 500, size=333)
    my_sites = float(counts.sum())
    my_ms = 10.0 + rank  # stand-in for this rank's device time
    dist.barrier()
    (t_max,) = sharding.allreduce_scalars([my_ms], "max")
    (sites,) = sharding.allreduce_scalars([my_sites], "sum")
    g = sharding.cpu_barrier_group()
    assert g is not None
    dist.barrier(group=g)
    q.put((rank, t_max, sites, my_sites))
    dist.barrier(group=g)
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(mine for _, _, _, mine in out)
    for rank, t_max, sites, mine in out:
        assert t_max == 11.0            # max over ranks
        assert sites == total           # sum over ranks
