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


def _comm_worker(rank, world, port, out):
    """The `zkl_comm` callbacks of TorchComm (zk-lisp_b200/sharded.py) on HOST buffers over gloo: the semantics the column-sharded prover
    relies on — all_gather: recv[s] = send of rank s; all_to_all: recv[s] = block `rank` of rank s's send; all_reduce_sum_u64 in place,
    exact modulo 2^64."""
    import ctypes as C
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    zkl = graft.import_zkl()
    comm = zkl.sharded.TorchComm(dist, torch.device("cpu"))
    ok = comm.c.rank == rank and comm.c.world == world
    nb = 48
    send = (np.arange(nb) + 50 * rank).astype(np.uint8)
    recv = np.zeros(world * nb, dtype=np.uint8)
    ok &= comm.c.all_gather(None, send.ctypes.data, recv.ctypes.data, nb) == 0
    ok &= all((recv[s * nb:(s + 1) * nb] == (np.arange(nb) + 50 * s).astype(np.uint8)).all() for s in range(world))
    send2 = np.concatenate([np.full(nb, 10 * rank + d, dtype=np.uint8) for d in range(world)])
    recv2 = np.zeros(world * nb, dtype=np.uint8)
    ok &= comm.c.all_to_all(None, send2.ctypes.data, recv2.ctypes.data, nb) == 0
    ok &= all((recv2[s * nb:(s + 1) * nb] == 10 * s + rank).all() for s in range(world))
    buf = np.array([2**63 + rank, 7, 2**64 - 1 if rank == 0 else 1], dtype=np.uint64)
    ok &= comm.c.all_reduce_sum_u64(None, buf.ctypes.data, buf.shape[0]) == 0
    want = [(sum(2**63 + r for r in range(world))) % 2**64, 7 * world, (2**64 - 1 + (world - 1)) % 2**64]
    ok &= [int(x) for x in buf] == want
    ok &= comm.calls == 3 and comm.error is None
    # column ownership as the library computes it (pure host code of the CUDA library): a disjoint cover by whole partitions
    lib = zkl.capi.load_library()
    for width, parts in ((219, (16, 16)), (204, (16, 16)), (204, (32, 8))):
        oc = zkl.air.ProofOptions(num_partitions=parts[0], hash_rate=parts[1]).to_c()
        ps = int(lib.zkl_partition_size(C.byref(oc), width))
        f, w = C.c_size_t(), C.c_size_t()
        ok &= lib.zkl_shard_columns(C.byref(oc), width, rank, world, C.byref(f), C.byref(w)) == 0
        cover = [None] * world
        dist.all_gather_object(cover, (f.value, w.value))
        ok &= cover[0][0] == 0 and all(cover[r][0] + cover[r][1] == cover[r + 1][0] for r in range(world - 1)) and sum(c[1] for c in cover) == width
        ok &= all(c[0] % ps == 0 and c[1] > 0 for c in cover)
        ok &= max(-(-c[1] // ps) for c in cover) - min(-(-c[1] // ps) for c in cover) <= 1   # balanced: X's 14 partitions over 8 ranks
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        out.put(all(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_torch_comm_callbacks_world2_and_4():
    for world in (2, 4):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q)) for r in range(world)]
        [p.start() for p in procs]
        assert q.get(timeout=60) is True
        [p.join(timeout=60) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
