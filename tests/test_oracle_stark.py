"""CPU tests of the restated prover/verifier (oracle) and of the host-side AIR bookkeeping."""
import numpy as np
import pytest

import synth
from util import P, from_fe, to_fe


def _air(zkl, trace, air_pi, opts):
    a = zkl.air.Air(trace.shape[0], trace.shape[1], air_pi, opts)
    return a, a.to_c(), opts.to_c(), zkl.prover.ints_to_fe(a.seed_elements())


def test_layout_widths_match_reference(zkl):
    A = zkl.air
    # 219 (all), 212 (no merkle), 211 (no ram), 204 (neither)
    assert A.Columns.baseline().width == 219
    assert A.Columns(A.LayoutConfig(merkle=False)).width == 212
    assert A.Columns(A.LayoutConfig(ram=False)).width == 211
    assert A.Columns(A.LayoutConfig(ram=False, merkle=False)).width == 204
    c = A.Columns.baseline()
    assert (c.g_map, c.g_final, c.g_r_start, c.mask, c.r_start, c.op_start) == (12, 13, 14, 41, 42, 50)
    assert (c.sel_dst0_start, c.sel_s_bits_start, c.sel_s_active_start, c.imm, c.eq_inv) == (67, 107, 137, 147, 148)
    assert (c.ram_start, c.merkle_start, c.pi_prog, c.pc, c.rom_op_start) == (149, 157, 164, 165, 166)
    assert (c.pose_active, c.gadget_b_start, c.rom_s_start) == (183, 184, 216)


def test_constraint_counts_and_degrees(zkl):
    A = zkl.air
    fm = A.FM_POSEIDON | A.FM_VM | A.FM_SPONGE | A.FM_RAM | A.FM_MERKLE
    d = A.transition_degrees(fm, True, 0, 0)
    # poseidon 336, ctrl 91, alu 16, ram 7, merkle 7, rom 86
    assert len(d) == 336 + 91 + 16 + 7 + 7 + 86
    d_all = A.transition_degrees(fm, True, 0x1FF, 0xFFFFFFFF)
    assert len(d_all) == 336 + 10 + 91 + 40 + 16 + 42 + 7 + 33 + 7 + 86
    assert A.ce_blowup_factor(d) == 8
    assert A.num_composition_columns(d, 4096) == 7
    assert A.select_partitions_for_trace(204, 4096) == (1, 16)
    assert A.select_partitions_for_trace(219, 1 << 22) == (16, 16)


def test_assertions_match_declared_count(zkl):
    trace, air_pi, opts = synth.valid_segment(n=128)
    a = zkl.air.Air(trace.shape[0], 128, air_pi, opts)
    # 141 per level (schedule.rs:57-139) + pc + pi_prog + 6 rom
    assert len(a.assertions) == 141 * 4 + 2 + 6
    assert len(a.seed_elements()) == 9 + 5 + 13


def test_valid_trace_satisfies_air(zkl, oracle):
    n = 128
    trace, air_pi, opts = synth.valid_segment(n=n)
    a, ac, oc, seed = _air(zkl, trace, air_pi, opts)
    t = np.array([[int(lo) | (int(hi) << 64) for lo, hi in col] for col in trace], dtype=object)
    g = zkl.field.root_of_unity(7)
    x = 1
    for row in range(n - 1):  # the last row is exempt (one transition exemption)
        per = oracle.periodic_at(n, x)
        cur = to_fe([t[c][row] for c in range(trace.shape[0])])
        nxt = to_fe([t[c][row + 1] for c in range(trace.shape[0])])
        res = from_fe(oracle.eval_transition_frame(ac, cur, nxt, per))
        assert all(v == 0 for v in res), (row, [i for i, v in enumerate(res) if v][:5])
        x = x * g % P
    for col, step, val in a.assertions:
        assert t[col][step] == val


def test_periodic_columns_on_trace_domain(zkl, oracle):
    n = 64
    g = zkl.field.root_of_unity(6)
    for row in (0, 1, 27, 28, 29, 31, 32, 63):
        per = from_fe(oracle.periodic_at(n, pow(g, row, P)))
        pos = row % 32
        exp = [0] * 32
        exp[0] = int(pos == 0)
        for j in range(27):
            exp[1 + j] = int(pos == 1 + j)
        exp[28] = int(pos == 28)
        exp[29] = int(pos in (29, 30, 31))
        exp[30] = int(pos == 31)
        exp[31] = int(row == n - 1)
        assert per == exp


@pytest.mark.parametrize("n,blowup,queries,grind", [(64, 16, 8, 4), (128, 8, 12, 6)])
def test_oracle_proof_verifies(zkl, oracle, n, blowup, queries, grind):
    trace, air_pi, opts = synth.valid_segment(n=n, queries=queries, blowup=blowup, grind=grind)
    a, ac, oc, seed = _air(zkl, trace, air_pi, opts)
    summ = zkl.capi.ProofSummary()
    proof = oracle.prove_segment(trace, ac, seed, oc, summ)
    rc, msg = oracle.verify_segment(proof, trace.shape[0], n, ac, seed, oc)
    assert rc == 0, msg
    assert summ.num_unique_queries <= queries
    # tampering anywhere must be rejected
    for off in (len(proof) // 3, len(proof) // 2, len(proof) - 9):
        bad = bytearray(proof)
        bad[off] ^= 1
        rc2, _ = oracle.verify_segment(bytes(bad), trace.shape[0], n, ac, seed, oc)
        assert rc2 != 0
    # a trace that violates the AIR must not verify (composition degree overflow shows up in OOD/FRI checks)
    bad_trace = trace.copy()
    bad_trace[a.cols.r_start + 1, 40, 0] ^= np.uint64(1)
    proof2 = oracle.prove_segment(bad_trace, ac, seed, oc)
    rc3, _ = oracle.verify_segment(proof2, trace.shape[0], n, ac, seed, oc)
    assert rc3 != 0


def test_boundary_identity_matches_naive_groups(zkl, oracle):
    """Python big-int check of the identity the GPU uses for boundary constraints:
    sum_k cc_k (T(x) - v_k)/(x - a_k) == n/(x^n-1) * (T(x) P(x) - P_S(x)) for a_k in the trace domain."""
    import random
    rng = random.Random(4)
    n = 32
    g = zkl.field.root_of_unity(5)
    steps = [0, 3, 17, 31]
    cc = [rng.randrange(P) for _ in steps]
    vals = [rng.randrange(P) for _ in steps]
    x, tx = rng.randrange(P), rng.randrange(P)
    naive = sum(c * (tx - v) * pow((x - pow(g, s, P)) % P, P - 2, P) for c, v, s in zip(cc, vals, steps)) % P

    def interp_eval(u):  # evaluate the interpolant of the sparse vector u (dict step -> value) at x
        tot = 0
        for s, val in u.items():
            a = pow(g, s, P)
            L = a * pow(n, P - 2, P) % P * (pow(x, n, P) - 1) % P * pow((x - a) % P, P - 2, P) % P
            tot += val * L
        return tot % P
    pc = interp_eval({s: c * pow(g, -s % n, P) % P for c, s in zip(cc, steps)})
    ps = interp_eval({s: c * v % P * pow(g, -s % n, P) % P for c, v, s in zip(cc, vals, steps)})
    ident = n * pow((pow(x, n, P) - 1) % P, P - 2, P) % P * ((tx * pc - ps) % P) % P
    assert naive == ident
