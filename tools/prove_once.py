#!/usr/bin/env python
"""Prove a few rollup-bench-sized segments on one stream (profiling target for ncu launch lists)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--count", type=int, default=2)
ap.add_argument("--grind", type=int, default=16)
ap.add_argument("--bulk-threshold", type=int, default=0)
a = ap.parse_args()
zkl = graft.import_zkl()
import synth  # noqa: E402

ctx = zkl.Context(0)
if a.bulk_threshold:
    ctx.lib.zkl_set_bulk_threshold(ctx.h, a.bulk_threshold)
trace, air_pi, opts = synth.valid_segment(n=a.n, queries=64, blowup=16, grind=a.grind, seed=7)
prover = zkl.ZkProver(ctx, opts, air_pi)
for i in range(a.count):
    proof = prover.prove(trace)
    s = prover.last_summary
    print(f"proof {i}: {len(proof)} bytes; stage ms:", {zkl.capi.load_library().zkl_stage_name(k).decode(): round(s.stage_ms[k], 3) for k in range(11)},
          "total", round(sum(s.stage_ms[:9]), 2), "launches", ctx.launches)
