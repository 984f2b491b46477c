#!/usr/bin/env python
"""Permutation throughput of the two sponge kernels vs the measured IMAD.WIDE peak."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

zkl = graft.import_zkl()
ctx = zkl.Context(0)
peak = C.c_double()
ctx._check(ctx.lib.zkl_imad_peak(ctx.h, C.byref(peak)))
print(f"IMAD.WIDE peak {peak.value/1e12:.2f} T/s -> permutation floor {peak.value/72576/1e6:.1f} M perm/s")
for kind, name in ((0, "thread-per-state"), (1, "coop-16")):
    for count in (1 << 16, 1 << 18, 1 << 20):
        ms = C.c_float()
        ctx._check(ctx.lib.zkl_bench_permute(ctx.h, kind, count, 3, C.byref(ms)))
        rate = count / (ms.value * 1e-3)
        print(f"{name:18s} count={count:8d} {ms.value:8.3f} ms  {rate/1e6:7.1f} M perm/s  frac of IMAD floor {rate*72576/peak.value:.3f}")
