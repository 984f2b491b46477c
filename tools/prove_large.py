#!/usr/bin/env python
"""Prove ONE large AIR-valid segment on one GPU and check the proof with the restated verifier (oracle, CPU; the verifier's
cost does not grow with n beyond the query paths). Sizes far above the reference's 4096-row segment cap: the reference's
O(n^2) boundary-divisor table makes them unprovable there; here they are the single-GPU slice of the
X configuration (BASELINE.json configs[4])."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logn", type=int, nargs="+", default=[16, 18])
ap.add_argument("--queries", type=int, default=64)
ap.add_argument("--grind", type=int, default=16)
ap.add_argument("--no-verify", action="store_true")
a = ap.parse_args()
zkl = graft.import_zkl()
import synth  # noqa: E402
import oracle_lib  # noqa: E402

if not a.no_verify:
    oracle_lib.build(); oracle_lib.load()
ctx = zkl.Context(0)
for logn in a.logn:
    n = 1 << logn
    t0 = time.time()
    trace, air_pi, opts = synth.valid_segment(n=n, queries=a.queries, blowup=16, grind=a.grind, seed=11)
    t_gen = time.time() - t0
    prover = zkl.ZkProver(ctx, opts, air_pi)
    best = None
    for rep in range(2):   # first pass builds the per-size tables
        t0 = time.time(); proof = prover.prove(trace); wall = time.time() - t0
        s = prover.last_summary
        gpu_ms = sum(s.stage_ms[:9])
        best = gpu_ms if best is None else min(best, gpu_ms)
    stages = {zkl.capi.load_library().zkl_stage_name(k).decode(): round(s.stage_ms[k], 2) for k in range(9)}
    line = f"n=2^{logn} W={trace.shape[0]} LDE {trace.shape[0] * n * 16 * 16 / 2**30:.1f} GiB: proof {len(proof)} B, device {best:.1f} ms " \
           f"({n / best * 1e3 / 1e3:.0f} k rows/s), wall {wall * 1e3:.0f} ms incl. H2D, trace generation {t_gen:.0f} s; stages {stages}"
    if not a.no_verify:
        ai = zkl.air.Air(trace.shape[0], n, air_pi, opts)
        t0 = time.time()
        rc, msg = oracle_lib.verify_segment(proof, trace.shape[0], n, ai.to_c(), zkl.prover.ints_to_fe(ai.seed_elements()), opts.to_c())
        line += f"; restated verifier: {'ACCEPT' if rc == 0 else 'REJECT ' + msg} ({time.time() - t0:.1f} s)"
    print(line, flush=True)
    del trace, proof
