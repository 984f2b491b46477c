"""world_size-2 test of the multi-GPU host logic on CPU (gloo): segment sharding is a disjoint cover and the
cross-rank aggregation is SUM(rows) / MAX(time). The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as graft


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    zkl = graft.import_zkl()
    mine = zkl.sharding.segments_for_rank(16, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    rows, ms = zkl.sharding.aggregate(len(mine) * 4096, 100.0 + 50.0 * rank, dist)
    weak = zkl.sharding.weak_scaling_segments(16, rank)
    if rank == 0:
        out.put((gathered, rows, ms, weak))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_aggregation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    gathered, rows, ms, weak = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(gathered[0] + gathered[1]) == list(range(16))
    assert not set(gathered[0]) & set(gathered[1])
    assert gathered[0] == [0, 2, 4, 6, 8, 10, 12, 14]
    assert rows == 16 * 4096 and ms == 150.0
    assert weak == list(range(16))


def test_sharding_uneven():
    zkl = graft.import_zkl()
    parts = [zkl.sharding.segments_for_rank(16, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(16))
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert zkl.sharding.aggregate(10, 5.0) == (10, 5.0)
