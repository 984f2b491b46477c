#!/usr/bin/env python
"""Verify ZKLSTP1 step-proof files on the GPU hasher (zk-lisp_b200/verify.py: transcript, OOD consistency, every opening, DEEP, FRI),
then the chain and the children root of the batch — the host counterpart of `zk-lisp verify` for step proofs.
Usage: python tools/verify_step.py [--security-bits N] step0.bin [step1.bin ...]      (files written with StepProof.to_bytes())
--security-bits: minimum conjectured security the proofs' own options must reach (default 128, the reference's release default,
zk-lisp-proof/src/lib.rs:59)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

zkl = graft.import_zkl()
args = sys.argv[1:]
min_bits = 128
if args[:1] == ["--security-bits"]:
    min_bits, args = int(args[1]), args[2:]
if not args:
    sys.exit(__doc__)
ctx = zkl.Context(0)
steps = []
for path in args:
    step = zkl.step.StepProof.from_bytes(open(path, "rb").read())
    t = time.time()
    fs = zkl.verify.verify_step(ctx, step, min_bits)
    v = step.proof.inner
    print(f"{path}: ok  segment {step.proof.pi.segment_index + 1}/{step.proof.pi.segments_total}, n={v.trace_len}, W={v.main_width}, "
          f"{len(fs.query_positions)} positions, digest {step.digest().hex()[:32]}  ({1e3 * (time.time() - t):.0f} ms)")
    steps.append(step)
if len(steps) > 1:
    zkl.replay.replay_children_native(ctx, steps)          # the whole batch once more through the native entry point
    zkl.step.verify_chain(steps)
    print("chain ok, children root", zkl.step.children_root_from_steps(steps).hex()[:32])
