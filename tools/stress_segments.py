"""Stress: a prover pool of two contexts (different kernel paths) proving the segments of one program concurrently, each
step proof compared with a proof of the host-sliced segment on the first context. Any mismatch is a race."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft
zkl = graft.import_zkl()
import synth
A, S = zkl.air, zkl.segments
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
extra = [zkl.Context(0) for _ in range(2)]
ctx = zkl.Context(0)
ctx._check(ctx.lib.zkl_set_bulk_threshold(ctx.h, 1)); ctx._check(ctx.lib.zkl_set_hash_path(ctx.h, 1)); ctx._check(ctx.lib.zkl_set_coop_path(ctx.h, 1))
n_full, seg_rows = 512, 128
full, air_pi_full, opts = synth.valid_segment(n=n_full, queries=8, blowup=16, grind=4, seed=21, full_layout=True)
segs = S.plan_fixed_segments(n_full, seg_rows)
feats = [S.SegmentFeatures(vm=True, sponge=True)] * len(segs)
full_cols = A.Columns.baseline()
lay = S.SegmentLayout(full_cols, A.LayoutConfig(ram=False, merkle=False))
bad = 0
ref = None
for it in range(iters):
    ctx2 = zkl.Context(0)
    steps = S.prove_program([ctx, ctx2], full, segs, air_pi_full.core, feats, opts)
    proofs = [s.proof for s in steps]
    if ref is None:
        ref = []
        for s, seg in zip(steps, segs):
            host_seg = np.ascontiguousarray(full[lay.seg_to_full.astype(np.int64), seg.r_start:seg.r_end])
            ref.append(zkl.ZkProver(ctx, s.options, s.air_pi).prove(host_seg))
    for i, (s, seg) in enumerate(zip(steps, segs)):
        host_seg = np.ascontiguousarray(full[lay.seg_to_full.astype(np.int64), seg.r_start:seg.r_end])
        again = zkl.ZkProver(ctx, s.options, s.air_pi).prove(host_seg)
        if again != ref[i] or s.proof != ref[i]:
            bad += 1
            d1 = [k for k in range(min(len(again), len(ref[i]))) if again[k] != ref[i][k]][:4]
            d2 = [k for k in range(min(len(s.proof), len(ref[i]))) if s.proof[k] != ref[i][k]][:4]
            print(f"iter {it} segment {i}: again!=ref {again != ref[i]} at {d1}; pool!=ref {s.proof != ref[i]} at {d2}", flush=True)
    ctx2.close()
print("iterations", iters, "mismatches", bad)
