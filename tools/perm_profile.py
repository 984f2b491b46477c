#!/usr/bin/env python
"""Minimal target for `ncu --set full -k regex:k_permute_batch`: one launch of the thread-per-state permutation."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

zkl = graft.import_zkl()
ctx = zkl.Context(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ms = C.c_float()
ctx._check(ctx.lib.zkl_bench_permute(ctx.h, kind, count, 1, C.byref(ms)))
print(f"kind={kind} count={count} {ms.value:.3f} ms")
