#!/usr/bin/env python
"""Profiling target: interpolation + blowup-16 coset LDE of a few 2^logn-row columns through the C ABI (multi-level tensor-core
transforms from 2^14 rows up). `ncu -k regex:^k_tc_cross16 -s 2 -c 1 python tools/lde_once.py --logn 20` captures the LDE's top cross
stage (the first two matches belong to the interpolation)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logn", type=int, default=20)
ap.add_argument("--cols", type=int, default=8)
ap.add_argument("--blowup", type=int, default=16)
a = ap.parse_args()
zkl = graft.import_zkl()
from util import fast_random_trace  # noqa: E402

ctx = zkl.Context(0)
cols = fast_random_trace(a.cols, 1 << a.logn, seed=a.logn)
t0 = time.time(); polys = ctx.interpolate_columns(cols); t1 = time.time()
lde = ctx.evaluate_columns(polys, a.blowup); t2 = time.time()
print(f"n=2^{a.logn} cols={a.cols}: interpolate {1e3 * (t1 - t0):.1f} ms, LDE x{a.blowup} {1e3 * (t2 - t1):.1f} ms (host wall incl. PCIe), checksum {int(lde[:, ::4097].sum() & 0xFFFFFFFF):08x}")
