#!/usr/bin/env python
"""Latency of ONE dependent permutation step at small batch sizes: back-to-back launches of `count` permutations on one
stream (each launch depends on the previous one), for the 16-thread cooperative, the 64-thread-group and the tensor-core
kernels. This is the quantity that bounds a single segment proof (Merkle levels, FRI tails, the coin)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

zkl = graft.import_zkl()
ctx = zkl.Context(0)
for count in (1, 8, 64, 256, 512, 1024, 2048, 4096):
    row = [f"count {count:5d}:"]
    for kind, name in ((1, "coop16"), (3, "quad64"), (2, "tensor")):
        ms = C.c_float()
        ctx._check(ctx.lib.zkl_bench_permute(ctx.h, kind, count, 50, C.byref(ms)))
        row.append(f"{name} {ms.value * 1e3:7.1f} us")
    print("  ".join(row))
