import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft
zkl = graft.import_zkl()
from util import fast_random_trace
ctx = zkl.Context(0)
n = 4096
cols = fast_random_trace(5, n, seed=1)
ctx._check(ctx.lib.zkl_set_ntt_path(ctx.h, 2)); a = ctx.interpolate_columns(cols)
ctx._check(ctx.lib.zkl_set_ntt_path(ctx.h, 0)); b = ctx.interpolate_columns(cols)
print("inverse mismatches:", int((a != b).any(axis=2).sum()))
for blowup in (2, 16):
    ctx._check(ctx.lib.zkl_set_ntt_path(ctx.h, 2)); a = ctx.evaluate_columns(cols, blowup)
    ctx._check(ctx.lib.zkl_set_ntt_path(ctx.h, 0)); b = ctx.evaluate_columns(cols, blowup)
    print("forward blowup", blowup, "mismatches:", int((a != b).any(axis=2).sum()))
