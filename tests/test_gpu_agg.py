"""GPU: the aggregation proof (ZlAggAir, 31 columns, one row per child) on the device — byte for byte the checker's proof under
FieldExtension::None and under Quadratic (prove.rs:647-651), through the same C-ABI entry point as a step proof
(air_kind = ZKL_AIR_AGG, zkl_options.field_extension), and the whole `prove_agg_proof` flow with the children verified on the device hasher."""
import ctypes as C

import numpy as np
import pytest

from test_agg import batch  # noqa: F401  (fixture: four chained step proofs made by the checker)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("min_bits", [64, 128])
def test_aggregation_proof_bytes_match_the_checker(ctx, zkl, oracle, batch, min_bits):  # noqa: F811
    G, RC = zkl.agg, zkl.recursion
    core, steps = batch
    agg_pi = RC.build_recursion_public(core, steps)
    trace = G.build_agg_trace(ctx, agg_pi, steps)               # children re-hashed and verified on the device hasher
    opts = G.agg_proof_options(queries=16, blowup=16, grind=6, min_security_bits=min_bits, n=trace.shape[1])
    air = G.AggAir(trace.shape[1], agg_pi, opts)
    ac, oc, seed = air.to_c(), opts.to_c(), zkl.prover.ints_to_fe(air.seed_elements())
    ref = oracle.prove_segment(trace, ac, seed, oc)
    buf = np.zeros(1 << 20, dtype=np.uint8)
    cap = C.c_size_t(buf.shape[0])
    summ = zkl.capi.ProofSummary()
    ctx._check(ctx.lib.zkl_prove_segment(ctx.h, trace.ctypes.data, 31, trace.shape[1], C.byref(ac), seed.ctypes.data, seed.shape[0], C.byref(oc),
                                         buf.ctypes.data, C.byref(cap), C.byref(summ)))
    proof = buf[:cap.value].tobytes()
    assert summ.constraints_violated == 0
    assert proof == ref
    rc, msg = oracle.verify_segment(proof, 31, trace.shape[1], ac, seed, oc)
    assert rc == 0, msg
    # the hook-level entry points see the same AIR: constraint evaluations of the committed aggregation trace equal the checker's
    pv = zkl.ZkProver(ctx, opts, None)
    lde, _ = pv.new_trace_lde(trace)
    cc = zkl.prover.ints_to_fe([(7919 * (i + 1)) ** 3 % zkl.field.P for i in range(air.num_constraints)])
    out = np.zeros((trace.shape[1] * air.ce_blowup, 2), dtype=np.uint64)
    ctx._check(ctx.lib.zkl_eval_constraints(ctx.h, lde.h, C.byref(ac), cc.ctypes.data, out.ctypes.data))
    assert np.array_equal(out, oracle.eval_constraints(trace, oc, ac, cc))
    lde.free()
    # a broken accumulator chain: the device refuses to hand out a proof
    bad = trace.copy()
    bad[G.COL["child_count_acc"], 1, 0] += np.uint64(1)
    cap = C.c_size_t(buf.shape[0])
    rc = ctx.lib.zkl_prove_segment(ctx.h, bad.ctypes.data, 31, trace.shape[1], C.byref(ac), seed.ctypes.data, seed.shape[0], C.byref(oc),
                                   buf.ctypes.data, C.byref(cap), C.byref(summ))
    assert rc == zkl.capi.ERR_CONSTRAINTS_VIOLATED and summ.constraints_violated == 1


def test_prove_agg_proof_end_to_end(ctx, zkl, oracle, batch):  # noqa: F811
    """prove.rs:629-702 as one call: children -> aggregation trace -> proof -> recursion artifact; the default (release) security level
    asks for the quadratic extension and for options that reach 128 bits."""
    if ctx.hash_mode != "default":
        pytest.skip("one hashing mode is enough here")
    G, RC = zkl.agg, zkl.recursion
    core, steps = batch
    agg_pi = RC.build_recursion_public(core, steps)
    proof = G.prove_agg_proof(ctx, agg_pi, steps, queries=64, blowup=16, grind=16, min_security_bits=128)
    inner = zkl.step.InnerProof.parse(proof)
    assert (inner.field_extension, inner.num_queries, inner.grinding, inner.trace_len, inner.main_width) == (2, 64, 16, 8, 31)
    opts = G.agg_proof_options(64, 16, 16, 128, 8)
    air = G.AggAir(8, agg_pi, opts)
    rc, msg = oracle.verify_segment(proof, 31, 8, air.to_c(), zkl.prover.ints_to_fe(air.seed_elements()), opts.to_c())
    assert rc == 0, msg
    art = RC.encode(proof, agg_pi)
    assert RC.decode(art)[0] == proof
    with pytest.raises(zkl.step.InvalidInput, match="min_security_bits"):
        G.prove_agg_proof(ctx, agg_pi, steps, queries=16, blowup=4, grind=0, min_security_bits=128)
