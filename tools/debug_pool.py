import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft
zkl = graft.import_zkl()
import synth
extra = [zkl.Context(0) for _ in range(3)]
a = zkl.Context(0); b = zkl.Context(0)
a._check(a.lib.zkl_set_bulk_threshold(a.h, 1)); a._check(a.lib.zkl_set_hash_path(a.h, 1)); a._check(a.lib.zkl_set_coop_path(a.h, 1))
for n in (128, 256, 4096):
    trace, air_pi, opts = synth.valid_segment(n=n, queries=8, blowup=16, grind=4, seed=21)
    pa = zkl.ZkProver(a, opts, air_pi); pb = zkl.ZkProver(b, opts, air_pi)
    x = pa.prove(trace); sa = pa.last_summary
    y = pb.prove(trace); sb = pb.last_summary
    print("n", n, "equal", x == y, "len", len(x), len(y))
    if x != y:
        print(" trace_root eq", bytes(sa.trace_root) == bytes(sb.trace_root), "constraint_root eq", bytes(sa.constraint_root) == bytes(sb.constraint_root))
        print(" fri roots eq", bytes(sa.fri_roots)[:32 * sa.num_fri_roots] == bytes(sb.fri_roots)[:32 * sb.num_fri_roots], "z eq", bytes(sa.z) == bytes(sb.z) if hasattr(sa.z, '__len__') else (sa.z.lo, sa.z.hi) == (sb.z.lo, sb.z.hi))
        print(" nonce", sa.pow_nonce, sb.pow_nonce, "nq", sa.num_unique_queries, sb.num_unique_queries)
        print(" pos a", list(sa.positions[:sa.num_unique_queries])); print(" pos b", list(sb.positions[:sb.num_unique_queries]))
        for i in range(len(sa.fri_roots) // 32):
            if i < sa.num_fri_roots and bytes(sa.fri_roots)[32*i:32*i+32] != bytes(sb.fri_roots)[32*i:32*i+32]: print(" first differing fri root", i); break
