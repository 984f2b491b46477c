#!/usr/bin/env python
"""Column-sharded PROOF of one segment across the GPUs of one box, one process per GPU over NCCL (run under torchrun):
every rank uploads the columns it owns, proves through `zkl_prove_segment_sharded_pi_dev` with the NCCL collectives of
zk-lisp_b200/sharded.py: TorchComm, and rank 0 compares every rank's bytes with the single-GPU proof of the whole trace (which the
GPU suite pins against the CPU checker) and, with --verify, hands them to the restated verifier.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/sharded_prove_check.py --logn 10 12 --verify
"""
import argparse
import dataclasses
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logn", type=int, nargs="+", default=[8, 12])
ap.add_argument("--queries", type=int, default=32)
ap.add_argument("--grind", type=int, default=8)
ap.add_argument("--rich", action="store_true", help="RAM + Merkle + every ALU gadget live (tests/synth.py: rich_segment) instead of the plain VM trace")
ap.add_argument("--min-leaves", type=int, default=None, help="zkl_set_shard_min_leaves (0: exchange every tree and FRI layer)")
ap.add_argument("--verify", action="store_true")
a = ap.parse_args()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
zkl = graft.import_zkl()
import synth  # noqa: E402  (needs the package registered by import_zkl)
ctx = zkl.Context(local)
if a.min_leaves is not None:
    ctx._check(ctx.lib.zkl_set_shard_min_leaves(ctx.h, a.min_leaves))
comm = zkl.sharded.TorchComm(dist, dev)
ok = True
for logn in a.logn:
    n = 1 << logn
    # the generators are seeded: every rank builds the same trace and keeps its columns
    if a.rich:
        trace, air_pi, opts = synth.rich_segment(n=n, seed=31, merkle=True, queries=a.queries, grind=a.grind)
    else:
        trace, air_pi, opts = synth.valid_segment(n=n, queries=a.queries, blowup=16, grind=a.grind, seed=31, full_layout=True)
    opts = dataclasses.replace(opts, num_partitions=16, hash_rate=16)
    W = trace.shape[0]
    first, cnt = zkl.sharded.shard_columns(ctx.lib, opts, W, rank, world)
    cols = zkl.sharded.DeviceColumns(ctx, trace, first, cnt)
    summ = zkl.capi.ProofSummary()
    proof = zkl.sharded.prove_segment_sharded(ctx, comm.c, cols.ptr, first, cnt, W, n, opts, air_pi, summary=summ)   # warm-up: tables, NCCL channels
    calls0, bytes0 = comm.calls, comm.bytes_moved
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    proof = zkl.sharded.prove_segment_sharded(ctx, comm.c, cols.ptr, first, cnt, W, n, opts, air_pi, summary=summ)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.time() - t0
    cols.close()
    every = [None] * world
    dist.all_gather_object(every, proof)
    if rank == 0:
        prover = zkl.ZkProver(ctx, opts, air_pi)
        single = prover.prove(trace)
        t1 = time.time(); single = prover.prove(trace); dt1 = time.time() - t1
        same = all(p == single for p in every)
        ok &= same
        stages = {ctx.lib.zkl_stage_name(k).decode(): round(summ.stage_ms[k], 2) for k in range(9)}
        msg = (f"world={world} n=2^{logn} W={W} cols/rank={cnt}: sharded {dt * 1e3:.1f} ms, single GPU (incl. H2D of all columns) {dt1 * 1e3:.1f} ms; "
               f"{comm.calls - calls0} collectives, {(comm.bytes_moved - bytes0) / 1e6:.2f} MB sent per rank; proof {len(proof)} B; "
               f"every rank's bytes == single-GPU proof: {same}; rank-0 stages {stages}")
        if a.verify:
            import oracle_lib
            oracle_lib.build(); oracle_lib.load()
            ai = zkl.air.Air(W, n, air_pi, opts)
            rc, vmsg = oracle_lib.verify_segment(every[-1], W, n, ai.to_c(), zkl.prover.ints_to_fe(ai.seed_elements()), opts.to_c())
            ok &= rc == 0
            msg += f"; restated verifier on rank {world - 1}'s bytes: {'ACCEPT' if rc == 0 else 'REJECT ' + vmsg}"
        print(msg, flush=True)
    dist.barrier()
if rank == 0:
    print("SHARDED PROOF OK" if ok else "SHARDED PROOF MISMATCH", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag[0]) else 1)
