"""Column-sharded trace commitment of ONE segment across the GPUs of a box ("single huge segment").

The reference commits to a trace row as `merge_many(hash_elements(partition_0), ..., hash_elements(partition_{P-1}))`
when `PartitionOptions` asks for several partitions (restated at zk-lisp-proof-winterfell/src/agg/child.rs:1025-1045;
partition selection src/utils.rs:394-409). Partitions are column ranges, so a rank that owns WHOLE partitions can run
the column transforms (iNTT, blowup LDE) and the partition hashes without seeing any other column. The only data
that has to cross NVLink is one 16-byte digest per (LDE row, partition):

    rank g:  columns of its partitions -> iNTT -> LDE -> partition digests [p_g][N]          (local, C ABI)
    all ranks: all_gather(partition digests)                                                  (NCCL, 16 B per row/partition
                                                                                               instead of 16*W B per row)
    rank g:  leaves[rows_g] = merge_many(digests[:, rows_g])                                  (row-sharded, local)
    all ranks: all_gather(leaves) -> MerkleTree::new -> root                                   (NCCL + local)

`plan()` and `assemble()` are pure index logic (tested on CPU with gloo, tests/test_multi_rank_gloo.py); the device
work goes through include/zkl_b200.h (`zkl_lde_dev`, `zkl_hash_partitions_dev`, `zkl_merge_partition_digests_dev`,
`zkl_merkle_dev`). The root is bit-identical to the single-GPU `zkl_trace_lde_commit` and to the CPU restatement
(tests/test_gpu_sharded.py, tools/sharded_commit_check.py under torchrun).
"""
from dataclasses import dataclass
from typing import List


@dataclass
class ShardPlan:
    width: int
    partition_size: int
    num_partitions: int          # P
    world: int
    parts_per_rank: int          # ceil(P / world): slots every rank contributes to the all-gather (padded)
    first_part: List[int]        # per rank: first owned partition
    num_parts: List[int]         # per rank: owned partitions (0 for ranks beyond P)

    def columns(self, rank: int):
        """[first, last) trace columns owned by `rank`."""
        lo = min(self.width, self.first_part[rank] * self.partition_size)
        hi = min(self.width, (self.first_part[rank] + self.num_parts[rank]) * self.partition_size)
        return lo, max(lo, hi)

    def slot_of_partition(self, q: int) -> int:
        """Index of global partition q in the gathered [world * parts_per_rank] digest array."""
        r = min(q // self.parts_per_rank, self.world - 1)
        return r * self.parts_per_rank + (q - self.first_part[r])


def plan(width: int, partition_size: int, world: int) -> ShardPlan:
    """Contiguous blocks of whole partitions per rank."""
    if partition_size <= 0 or width <= 0 or world <= 0:
        raise ValueError("bad shard plan arguments")
    P = (width + partition_size - 1) // partition_size
    if P < 2:
        raise ValueError("column sharding needs at least two partitions (PartitionOptions (1, 0) hashes whole rows)")
    per = (P + world - 1) // world
    first = [min(r * per, P) for r in range(world)]
    cnt = [max(0, min(P, (r + 1) * per) - first[r]) for r in range(world)]
    return ShardPlan(width, partition_size, P, world, per, first, cnt)


def row_range(N: int, rank: int, world: int):
    """Rows whose leaves `rank` merges (N is a power of two >= world)."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


def assemble(gathered, p: ShardPlan):
    """gathered: array-like [world * parts_per_rank][N] (rank-major, padded slots at the end of each rank block).
    Returns the list of row arrays in global partition order (views, no copy)."""
    return [gathered[p.slot_of_partition(q)] for q in range(p.num_partitions)]


class ShardedTraceCommit:
    """Runs the flow above on the calling rank. `dist` is torch.distributed (nccl backend), `ctx` a zk_lisp_b200 Context
    pinned to this rank's GPU."""

    def __init__(self, ctx, dist, opts):
        self.ctx, self.dist, self.opts = ctx, dist, opts
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def commit(self, local_cols, width: int, n: int):
        """local_cols: numpy (ncols_local, n, 2) uint64 — this rank's columns per plan(). Returns (root bytes, plan)."""
        import ctypes as C
        import numpy as np
        import torch
        lib, h = self.ctx.lib, self.ctx.h
        oc = self.opts.to_c()
        ps = int(lib.zkl_partition_size(C.byref(oc), width))
        p = plan(width, ps, self.world)
        B = int(self.opts.blowup)
        N = n * B
        lo, hi = p.columns(self.rank)
        ncols = hi - lo
        assert local_cols.shape[0] == ncols, (local_cols.shape, lo, hi)
        dev = torch.device("cuda", self.ctx.device)
        u8 = torch.uint8

        def buf(elems):
            return torch.empty(max(elems, 1) * 16, dtype=u8, device=dev)

        # local: upload, iNTT, LDE, partition digests
        mine = buf(p.parts_per_rank * N)
        mine.zero_()
        torch.cuda.synchronize(dev)   # torch's stream and the context's stream are independent
        if ncols:
            cols_d, polys_d, lde_d = buf(ncols * n), buf(ncols * n), buf(ncols * N)
            host = np.ascontiguousarray(local_cols, dtype=np.uint64)
            self.ctx._check(lib.zkl_dev_upload(h, C.c_void_p(cols_d.data_ptr()), host.ctypes.data, host.nbytes))
            self.ctx._check(lib.zkl_lde_dev(h, C.c_void_p(cols_d.data_ptr()), ncols, n, B, C.c_void_p(polys_d.data_ptr()), C.c_void_p(lde_d.data_ptr())))
            self.ctx._check(lib.zkl_hash_partitions_dev(h, C.c_void_p(lde_d.data_ptr()), ncols, n, B, ps, C.c_void_p(mine.data_ptr())))
        self.ctx._check(lib.zkl_sync(h))
        # exchange 1: partition digests (16 B per row and partition)
        allp = buf(self.world * p.parts_per_rank * N)
        self.dist.all_gather_into_tensor(allp, mine)
        torch.cuda.synchronize(dev)
        # compact to global partition order if the last rank blocks are padded
        if p.num_partitions != self.world * p.parts_per_rank:
            view = allp.view(self.world * p.parts_per_rank, N * 16)
            allp = torch.stack(assemble(view, p)).contiguous().view(-1)
        # local: leaves of this rank's rows
        r0, r1 = row_range(N, self.rank, self.world)
        per = (N + self.world - 1) // self.world
        leaves_mine = buf(per)
        leaves_mine.zero_()
        torch.cuda.synchronize(dev)
        if r1 > r0:
            self.ctx._check(lib.zkl_merge_partition_digests_dev(h, C.c_void_p(allp.data_ptr()), p.num_partitions, N, r0, r1 - r0,
                                                                C.c_void_p(leaves_mine.data_ptr())))
        self.ctx._check(lib.zkl_sync(h))
        # exchange 2: leaves, then the tree on every rank
        leaves = buf(self.world * per)
        self.dist.all_gather_into_tensor(leaves, leaves_mine)
        torch.cuda.synchronize(dev)
        nodes = buf(N)
        self.ctx._check(lib.zkl_merkle_dev(h, C.c_void_p(leaves.data_ptr()), N, C.c_void_p(nodes.data_ptr())))
        root = np.zeros(16, dtype=np.uint8)
        self.ctx._check(lib.zkl_dev_download(h, root.ctypes.data, C.c_void_p(nodes.data_ptr() + 16), 16))
        self.bytes_exchanged = (self.world - 1) * (p.parts_per_rank * N * 16 + per * 16)
        self.bytes_unsharded_rows = (self.world - 1) * ncols * N * 16
        return root.tobytes() + bytes(16), p


# ---------------------------------------------------------------------------------------------------------------------
# The whole proof on column-sharded data (include/zkl_b200.h: zkl_comm, zkl_prove_segment_sharded_*): the flow lives in the library
# (csrc/api_stark.inl: prove_segment_impl with a shard_view); this side provides the collectives and the column plan.
# ---------------------------------------------------------------------------------------------------------------------
class _DevMem:
    """A device pointer as a torch tensor (no copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(torch, device, ptr: int, nbytes: int):
    """`nbytes` at `ptr` as a uint8 tensor without a copy: device memory on a CUDA device, host memory otherwise (the gloo tests)."""
    if getattr(device, "type", "cuda") == "cuda":
        return torch.as_tensor(_DevMem(ptr, nbytes), device=device)
    import ctypes as C
    import numpy as np
    return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)))


class TorchComm:
    """`zkl_comm` over torch.distributed: all-gather, all-to-all and a u64 sum on the calling rank's pointers. NCCL with device
    pointers is the product path (one process per GPU); with `device` = cpu and a gloo group the same callbacks move host memory,
    which is how the CPU suite checks their semantics against the in-process communicator's (tests/test_multi_rank_gloo.py)."""

    def __init__(self, dist, device):
        import torch
        from . import capi
        self.dist, self.device, self.torch = dist, device, torch
        self.bytes_moved, self.calls, self.error = 0, 0, None
        self.seconds = {"all_gather": 0.0, "all_to_all": 0.0, "all_reduce_sum_u64": 0.0}   # host wall time inside the callbacks
        cuda = getattr(device, "type", "cuda") == "cuda"
        world = dist.get_world_size()

        def view(ptr, nbytes):
            return _view(torch, device, ptr, nbytes)

        def guard(fn, name):
            import time

            def run(*a):
                try:
                    t0 = time.perf_counter()
                    fn(*a)
                    if cuda:
                        torch.cuda.synchronize(device)
                    self.seconds[name] += time.perf_counter() - t0
                    self.calls += 1
                    return 0
                except Exception as e:      # a Python exception must not unwind through the C caller
                    self.error = repr(e)
                    return 1
            return run

        def all_gather(user, send, recv, nbytes):
            self.bytes_moved += (world - 1) * nbytes
            dist.all_gather_into_tensor(view(recv, nbytes * world), view(send, nbytes))

        def all_to_all(user, send, recv, nbytes):
            self.bytes_moved += (world - 1) * nbytes
            dist.all_to_all_single(view(recv, nbytes * world), view(send, nbytes * world))

        def all_reduce(user, buf, count):
            self.bytes_moved += 8 * count
            dist.all_reduce(view(buf, 8 * count).view(torch.int64), op=dist.ReduceOp.SUM)   # two's-complement sum = unsigned sum mod 2^64

        self._keep = (capi.COMM_MOVE(guard(all_gather, "all_gather")), capi.COMM_MOVE(guard(all_to_all, "all_to_all")),
                      capi.COMM_SUM(guard(all_reduce, "all_reduce_sum_u64")))
        self.c = capi.Comm(dist.get_rank(), world, None, *self._keep)


def shard_columns(lib, opts, width: int, rank: int, world: int):
    """`zkl_shard_columns`: [first, first + count) of the columns `rank` owns (whole partitions, balanced contiguous blocks)."""
    import ctypes as C
    oc = opts.to_c()
    f, w = C.c_size_t(), C.c_size_t()
    if lib.zkl_shard_columns(C.byref(oc), width, rank, world, C.byref(f), C.byref(w)):
        raise ValueError("column sharding needs PartitionOptions with at least one whole partition per rank")
    return f.value, w.value


def prove_segment_sharded(ctx, comm_c, local_trace_dev: int, first_col: int, local_width: int, width: int, n: int, opts, air_pi,
                          summary=None) -> bytes:
    """This rank's part of the column-sharded proof of one segment; every rank returns the same bytes — the single-GPU proof.
    `comm_c`: a capi.Comm (TorchComm(...).c, `lib.zkl_comm_local_rank(group, r)` for the threads of one process, or
    `lib.zkl_comm_loopback()` for one rank). `local_trace_dev`: device pointer to the rank's columns (local_width x n, column-major)."""
    import ctypes as C
    import numpy as np
    from . import capi
    from .prover import program_constants, public_inputs_to_c
    core = air_pi.core
    prog, pi, o = program_constants(core.program_id, core.program_commitment), public_inputs_to_c(air_pi), opts.to_c()
    buf = np.empty(4 << 20, dtype=np.uint8)
    cap = C.c_size_t(buf.shape[0])
    summ = summary if summary is not None else capi.ProofSummary()
    ctx._check(ctx.lib.zkl_prove_segment_sharded_pi_dev(ctx.h, C.c_void_p(local_trace_dev), first_col, local_width, width, n, C.byref(prog), C.byref(pi),
                                                        C.byref(o), comm_c, buf.ctypes.data, C.byref(cap), C.byref(summ)))
    return buf[:cap.value].tobytes()


class DeviceColumns:
    """The columns [first, first + count) of a host trace (W, n, 2) uint64 in device memory of `ctx` (freed on close())."""

    def __init__(self, ctx, trace, first: int, count: int):
        import ctypes as C
        import numpy as np
        self.ctx = ctx
        host = np.ascontiguousarray(trace[first:first + count], dtype=np.uint64)
        p = C.c_void_p()
        ctx._check(ctx.lib.zkl_dev_alloc(ctx.h, host.nbytes, C.byref(p)))
        self.ptr = p.value
        ctx._check(ctx.lib.zkl_dev_upload(ctx.h, C.c_void_p(self.ptr), host.ctypes.data, host.nbytes))

    def close(self):
        import ctypes as C
        if self.ptr:
            self.ctx.lib.zkl_dev_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = None


def prove_segment_thread_ranks(contexts, trace, opts, air_pi, min_leaves=None):
    """The column-sharded proof with one THREAD per rank inside this process (`zkl_comm_local_*`): contexts[r] is rank r's context —
    all on one device (several shards of a segment on one GPU) or one per device of a box. Returns the list of per-rank proofs (they
    are all the same bytes)."""
    import ctypes as C
    import threading
    world = len(contexts)
    lib = contexts[0].lib
    W, n = trace.shape[0], trace.shape[1]
    group = C.c_void_p()
    if lib.zkl_comm_local_create(world, C.byref(group)):
        raise RuntimeError("zkl_comm_local_create failed")
    out, errs = [None] * world, [None] * world

    def run(r):
        cols = None
        try:
            ctx = contexts[r]
            if min_leaves is not None:
                ctx._check(lib.zkl_set_shard_min_leaves(ctx.h, min_leaves))
            first, cnt = shard_columns(lib, opts, W, r, world)
            cols = DeviceColumns(ctx, trace, first, cnt)
            out[r] = prove_segment_sharded(ctx, lib.zkl_comm_local_rank(group, r), cols.ptr, first, cnt, W, n, opts, air_pi)
        except Exception as e:   # noqa: BLE001 - reported to the caller below
            errs[r] = e
        finally:
            if cols is not None:
                cols.close()

    try:
        # a rank that fails before its first collective leaves the others waiting: the group's barrier times out after 600 s; columns
        # that cannot be sharded fail on every rank alike, before any collective
        ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    finally:
        lib.zkl_comm_local_destroy(group)
    for e in errs:
        if e is not None:
            raise e
    return out
