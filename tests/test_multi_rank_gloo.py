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


def _shard_worker(rank, world, port, out):
    """Column-sharded commitment, index logic only: every rank fabricates the 'digests' of the partitions it owns
    (value = 1000 * partition + row), the ranks all-gather them over gloo, assemble() must return global partition order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    zkl = graft.import_zkl()
    ok = True
    for width, psize in ((219, 16), (40, 10), (33, 16), (204, 13)):
        p = zkl.sharded.plan(width, psize, world)
        N = 8
        mine = torch.zeros(p.parts_per_rank, N, dtype=torch.int64)
        for i in range(p.num_parts[rank]):
            mine[i] = 1000 * (p.first_part[rank] + i) + torch.arange(N)
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        rows = zkl.sharded.assemble(torch.cat(got), p)
        ok &= len(rows) == p.num_partitions and all(bool((rows[q] == 1000 * q + torch.arange(N)).all()) for q in range(p.num_partitions))
        lo, hi = p.columns(rank)
        cover = [None] * world
        dist.all_gather_object(cover, (lo, hi))
        ok &= cover[0][0] == 0 and all(cover[r][1] == cover[r + 1][0] or cover[r + 1][0] == cover[r + 1][1] for r in range(world - 1)) and max(c[1] for c in cover) == width
        rr = [zkl.sharded.row_range(64, r, world) for r in range(world)]
        ok &= rr[0][0] == 0 and rr[-1][1] == 64 and all(rr[r][1] == rr[r + 1][0] for r in range(world - 1))
    if rank == 0:
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_column_shard_plan_world2_and_3():
    for world in (2, 3):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
        [p.start() for p in procs]
        assert q.get(timeout=120) is True
        [p.join(timeout=60) for p in procs]
        assert all(p.exitcode == 0 for p in procs)


def test_column_shard_plan_rejects_single_partition():
    import pytest
    zkl = graft.import_zkl()
    with pytest.raises(ValueError):
        zkl.sharded.plan(7, 16, 2)
    p = zkl.sharded.plan(219, 16, 8)
    assert p.num_partitions == 14 and p.parts_per_rank == 2 and sum(p.num_parts) == 14 and p.columns(7) == (219, 219)
    assert [p.columns(r) for r in range(8)][:7] == [(32 * r, min(219, 32 * (r + 1))) for r in range(7)]


def test_sharding_uneven():
    zkl = graft.import_zkl()
    parts = [zkl.sharding.segments_for_rank(16, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(16))
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert zkl.sharding.aggregate(10, 5.0) == (10, 5.0)
