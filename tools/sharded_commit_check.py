#!/usr/bin/env python
"""Column-sharded trace commitment across the GPUs of one box (run under torchrun, one rank per GPU):
every rank extends and hashes the column partitions it owns, the ranks all-gather the 16-byte partition digests over
NCCL, merge them row-sharded, all-gather the leaves and build the tree. Rank 0 checks the root against the single-GPU
hook and (for sizes the CPU finishes quickly) the oracle, and reports timing and the bytes that crossed NVLink."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4096, dest="n")
ap.add_argument("--width", type=int, default=219)
ap.add_argument("--partitions", type=int, default=16)
ap.add_argument("--hash-rate", type=int, default=16)
ap.add_argument("--oracle", action="store_true")
a = ap.parse_args()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from util import fast_random_trace  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
zkl = graft.import_zkl()
ctx = zkl.Context(local)
opts = zkl.air.ProofOptions(num_queries=8, blowup=16, grinding=0, num_partitions=a.partitions, hash_rate=a.hash_rate)
trace = fast_random_trace(a.width, a.n, seed=1234)           # same on every rank (seeded); each rank keeps its columns
import ctypes as C  # noqa: E402
oc = opts.to_c()
ps = int(ctx.lib.zkl_partition_size(C.byref(oc), a.width))
plan = zkl.sharded.plan(a.width, ps, world)
lo, hi = plan.columns(rank)
sc = zkl.sharded.ShardedTraceCommit(ctx, dist, opts)
sc.commit(trace[lo:hi], a.width, a.n)                         # warm-up (tables, NCCL channels)
dist.barrier(); torch.cuda.synchronize()
t0 = time.time()
root, _ = sc.commit(trace[lo:hi], a.width, a.n)
torch.cuda.synchronize(); dist.barrier()
dt = time.time() - t0
roots = [None] * world
dist.all_gather_object(roots, root)
ok = all(r == roots[0] for r in roots)
if rank == 0:
    prover = zkl.ZkProver(ctx, opts, None)
    t1 = time.time()
    lde, _ = prover.new_trace_lde(trace)
    single = lde.get_main_trace_commitment(); lde.free()
    dt1 = time.time() - t1
    ok &= single == root
    msg = f"world={world} n={a.n} W={a.width} partitions={plan.num_partitions} x {ps} cols: sharded {dt * 1e3:.1f} ms, single-GPU hook (incl. H2D of all columns) {dt1 * 1e3:.1f} ms; " \
          f"NVLink bytes per rank {sc.bytes_exchanged / 1e6:.2f} MB (raw rows would be {sc.bytes_unsharded_rows / 1e6:.1f} MB); roots equal: {ok}"
    if a.oracle:
        import oracle_lib
        oracle_lib.build(); oracle_lib.load()
        oroot, _, _ = oracle_lib.trace_lde_commit(trace, opts.to_c(), want_lde=False)
        ok &= oroot == root
        msg += f"; oracle root equal: {oroot == root}"
    print(msg)
    print("SHARDED COMMIT OK" if ok else "SHARDED COMMIT MISMATCH")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
