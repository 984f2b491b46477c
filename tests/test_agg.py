"""The aggregation layer (zk-lisp_b200/agg.py): ZlAggAir on Python integers, the aggregation trace built from a chain of step proofs
(agg/trace.rs:155-700), and the aggregation proof itself under both field-extension settings of the reference (prove.rs:647-651),
proven and verified by the CPU checker here; tests/test_gpu_agg.py proves the same trace on the device and compares bytes."""
from dataclasses import replace

import numpy as np
import pytest

import synth
from test_child_replay import _PlanHasher
from test_segment_driver_host import _StandInContext


@pytest.fixture(scope="module")
def batch(zkl, oracle):
    """Four chained 64-row segments of one program, proven by the checker behind the stand-in context (prove.rs:954-1051)."""
    SG = zkl.segments
    n_full, seg_rows = 256, 64
    full, air_pi_full, opts = synth.valid_segment(n=n_full, queries=8, blowup=16, grind=4, seed=31, full_layout=True)
    core = air_pi_full.core
    segs = SG.plan_fixed_segments(n_full, seg_rows)
    feats = [SG.SegmentFeatures(vm=True, sponge=True)] * len(segs)
    ctxs = [_StandInContext(oracle, air_pi_full.vm_usage_mask, device=0)]
    steps = [s.step for s in SG.prove_program(ctxs, full, segs, core, feats, opts, rom_acc=(0, 0, 0), min_bits=40)]
    return core, steps


def test_agg_trace_satisfies_the_agg_air(zkl, oracle, batch):
    G, RC = zkl.agg, zkl.recursion
    core, steps = batch
    agg_pi = RC.build_recursion_public(core, steps)
    trace = G.build_agg_trace(_PlanHasher(oracle), agg_pi, steps)
    assert trace.shape == (31, 8, 2)                                      # 4 children -> max(4, 8) rows (agg/trace.rs:35,396-399)
    col = lambda name: [int(trace[G.COL[name], r, 0]) | (int(trace[G.COL[name], r, 1]) << 64) for r in range(8)]
    v = steps[0].proof.meta.v_units
    assert col("seg_first") == [1, 1, 1, 1, 0, 0, 0, 0] and col("child_count_acc") == [0, 1, 2, 3, 4, 4, 4, 4]
    assert col("v_units_acc") == [0, v, 2 * v, 3 * v] + [4 * v] * 4 and agg_pi.v_units_total == 4 * v
    for name in ("ok", "trace_root_err", "constraint_root_err", "comp_sum", "alpha_div_zm_sum", "map_l0_sum", "final_llast_sum",
                 "vm_chain_err", "ram_u_chain_err", "ram_s_chain_err", "rom_chain_err_0", "rom_chain_err_1", "rom_chain_err_2", "r", "alpha"):
        assert not any(col(name)), name
    assert all(col("fri_x0_child")[:4]) and all(col("fri_alpha_child")[:4]) and col("fri_vnext_child")[:4] == col("fri_q1_child")[:4]
    air = G.AggAir(8, agg_pi, G.agg_proof_options(16, 16, 4, 64, 8))
    assert (air.num_transition, air.ce_blowup, air.num_composition_cols, len(air.assertions)) == (24, 4, 2, 5)
    G.check_trace(trace, air)
    # a child out of order breaks the VM / ROM chain; a wrong children_root or work total is refused before any row is written
    swapped = [steps[1], steps[0]] + steps[2:]
    with pytest.raises(zkl.step.InvalidInput, match="children_root"):
        G.build_agg_trace(_PlanHasher(oracle), replace(agg_pi, children_root=bytes(32)), steps)
    with pytest.raises(zkl.step.InvalidInput, match="v_units_total"):
        G.build_agg_trace(_PlanHasher(oracle), replace(agg_pi, v_units_total=agg_pi.v_units_total + 1), steps)
    bad_pi = RC.build_recursion_public(core, swapped)
    bad = G.build_agg_trace(_PlanHasher(oracle), bad_pi, swapped)
    with pytest.raises(zkl.step.InvalidInput, match="violates constraint (18|21)"):
        G.check_trace(bad, G.AggAir(8, bad_pi, air.opts))
    # a child whose FRI layer does not fold: refused while the children are verified, before the trace exists
    v0 = steps[2].proof.inner
    vals, paths = v0.fri_layers[1]
    tampered = replace(steps[2], proof=replace(steps[2].proof, inner=replace(v0, fri_layers=[v0.fri_layers[0], (vals[:16] + bytes(16) + vals[32:], paths)] + v0.fri_layers[2:])))
    with pytest.raises(zkl.step.InvalidInput):
        G.build_agg_trace(_PlanHasher(oracle), agg_pi, steps[:2] + [tampered] + steps[3:])


@pytest.mark.parametrize("min_bits", [64, 128])
def test_checker_proves_and_verifies_the_aggregation_proof(zkl, oracle, batch, min_bits):
    """FieldExtension::None below 128 bits, Quadratic from 128 up (prove.rs:647-651). With this hasher every drawn extension element
    is (a, 0): the quadratic proof carries the same field values with 32-byte elements and other digests."""
    G, RC = zkl.agg, zkl.recursion
    core, steps = batch
    agg_pi = RC.build_recursion_public(core, steps)
    trace = G.build_agg_trace(_PlanHasher(oracle), agg_pi, steps)
    opts = G.agg_proof_options(queries=16, blowup=16, grind=6, min_security_bits=min_bits, n=8)
    assert opts.field_extension == (2 if min_bits >= 128 else 1) and (opts.num_partitions, opts.hash_rate) == (1, 8)
    air = G.AggAir(8, agg_pi, opts)
    ac, oc, seed = air.to_c(), opts.to_c(), zkl.prover.ints_to_fe(air.seed_elements())
    summ = zkl.capi.ProofSummary()
    proof = oracle.prove_segment(trace, ac, seed, oc, summ)
    rc, msg = oracle.verify_segment(proof, 31, 8, ac, seed, oc)
    assert rc == 0, msg
    inner = zkl.step.InnerProof.parse(proof)
    assert (inner.main_width, inner.trace_len, inner.field_extension, inner.num_constraints) == (31, 8, opts.field_extension, 29)
    if opts.field_extension == 2:        # every E element on the wire is (a, 0)
        for blob in (inner.constraint_query_values, inner.ood_trace, inner.ood_constraints, inner.fri_remainder):
            assert len(blob) % 32 == 0 and not any(blob[i + 16:i + 32] != bytes(16) for i in range(0, len(blob), 32))
    # other public inputs, other seed: rejected
    other = G.AggAir(8, replace(agg_pi, batch_id=bytes([1]) * 32), opts)
    rc, _ = oracle.verify_segment(proof, 31, 8, other.to_c(), zkl.prover.ints_to_fe(other.seed_elements()), oc)
    assert rc != 0
    # a trace that breaks the work accumulator: the proof of it does not verify
    bad = trace.copy()
    bad[G.COL["v_units_acc"], 2, 0] += np.uint64(1)
    rc, _ = oracle.verify_segment(oracle.prove_segment(bad, ac, seed, oc), 31, 8, ac, seed, oc)
    assert rc != 0
    # the recursion artifact around the proof (lib.rs:483-774)
    art = RC.encode(proof, agg_pi)
    back, pi2 = RC.decode(art)
    assert back == proof and pi2.children_root == agg_pi.children_root and pi2.children_count == 4
