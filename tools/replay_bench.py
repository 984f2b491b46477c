#!/usr/bin/env python
"""Child-proof replay timing on the GPU: 16 rollup-bench-size children (n = 4096, q = 64).
  per-child path : replay.verify_child_openings  — one C-ABI call per tree level (~250 small batches per child)
  batched path   : replay.verify_children_openings — leaves in three batches, every tree in ONE zkl_merge_plan
Usage: python tools/replay_bench.py [children]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

zkl = graft.import_zkl()
import synth  # noqa: E402
from util import single_segment_step  # noqa: E402

children = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = zkl.Context(0)
steps = []
for i in range(min(children, 4)):
    trace, air_pi, opts = synth.valid_segment(n=4096, queries=64, blowup=16, grind=16, seed=100 + i)
    proof = zkl.ZkProver(ctx, opts, air_pi).prove(trace)
    steps.append(single_segment_step(zkl, air_pi, opts, 4096, proof))
steps = [steps[i % len(steps)] for i in range(children)]
t0 = time.time()
fss = [zkl.replay.replay_fs_from_step(ctx, s) for s in steps]
t0b = time.time()
assert zkl.replay.replay_fs_from_steps(ctx, steps) == fss
t1a = time.time()
assert zkl.replay.replay_fs_from_steps(ctx, steps) == fss
t1b = time.time()
t1 = time.time()
l0 = ctx.launches
for s, fs in zip(steps, fss):
    zkl.replay.verify_child_openings(ctx, s, fs)
t2 = time.time()
l1 = ctx.launches
zkl.replay.verify_children_openings(ctx, steps, fss)
t3 = time.time()
zkl.replay.verify_children_openings(ctx, steps, fss)
t4 = time.time()
l2 = ctx.launches
assert zkl.replay.replay_children_native(ctx, steps) == fss
t5 = time.time()
assert zkl.replay.replay_children_native(ctx, steps) == fss
t6 = time.time()
l3 = ctx.launches
hashed = sum(len(fs.query_positions) for fs in fss)
print(f"children {children}, opened positions {hashed}, proof bytes {sum(len(s.proof.inner_bytes) for s in steps)}")
print(f"fs replay            {1e3 * (t0b - t0) / children:8.2f} ms/child  (per child, one C-ABI call per transcript step)")
print(f"fs replay, lockstep  {1e3 * (t1b - t1a) / children:8.2f} ms/child  (all children per transcript step)")
print(f"openings, per child  {1e3 * (t2 - t1) / children:8.2f} ms/child  ({(l1 - l0) // children} launches/child)")
print(f"openings, one plan   {1e3 * (t4 - t3) / children:8.2f} ms/child  ({(l2 - l1) // 2} launches for all children; first call {1e3 * (t3 - t2):.1f} ms)")
print(f"native, everything   {1e3 * (t6 - t5) / children:8.2f} ms/child  (zkl_replay_children: transcripts + openings, {(l3 - l2) // 2} launches for all children)")
cap = {}
zkl.replay.replay_fs_from_step(ctx, steps[0], cap)
for bits in (16, 20):
    ctx.grind(cap["grinding_seed"], bits)
    t7 = time.time()
    nonce = ctx.grind(cap["grinding_seed"], bits)
    print(f"zkl_grind bits={bits}: nonce {nonce} in {1e3 * (time.time() - t7):.2f} ms")
