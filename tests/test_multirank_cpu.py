"""Multi-rank host logic without a cluster: world_size-2 gloo processes exercise the partition
rule, the rank-local seeding and the max-over-ranks timing reduction that bench.py uses under
torchrun. (The data path itself has no collective: problems are independent.)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dto_b200 import sharding
    from util import rng_for
    sl = sharding.rank_slice(total, rank, world)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[sl] = 1
    dist.all_reduce(owned)  # every problem owned exactly once
    seed_draw = float(rng_for(2, rank).uniform())
    draws = [None] * world
    dist.all_gather_object(draws, seed_draw)
    tmax = sharding.reduce_max(10.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, sl.start, sl.stop, bool((owned == 1).all()), draws, tmax))


def test_two_rank_partition_and_reduction():
    world, total = 2, 4097
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 2049), (2049, 4097)]
    assert all(r[3] for r in res)
    assert res[0][4] == res[1][4] and len(set(res[0][4])) == world  # distinct per-rank input streams
    assert all(r[5] == 11.0 for r in res)  # max over ranks
