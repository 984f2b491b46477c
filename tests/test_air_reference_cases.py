"""The reference's own AIR unit tests, restated case by case ("hand-built EvaluationFrame + periodic vector ... directly
portable as CPU checks for the constraint kernel"):

    vm/air/poseidon.rs:171-496   round_constraints_zero_on_valid_row, hold_constraints_hold_on_pad, round_gated_by_pose_active,
                                 sponge_binding_mux_positive, sponge_binding_mux_negative
    vm/air/ctrl.rs:399-648       sponge_selectors_gated_off_without_feature, sponge_selectors_on_feature_affect_eval,
                                 two_ops_set_violation, packed_bits_booleanity_violation
    vm/air/alu.rs:363-421        assert_enforcer_violates_when_false
    vm/air/rom.rs:155-288        round_constraints_zero_on_valid_row, map_enc_bindings_hold

Every case builds the reference's frame (baseline 219-column layout, zero elsewhere) and periodic vector and makes the reference's
assertion on BOTH restatements: the Python-integer block evaluator (zk-lisp_b200/air_eval.py) and the C checker's transition vector
(oracle/air.h, the slice of the block under test) — the checker is what the CUDA constraint kernel is compared with index by index
(tests/test_gpu_parity.py: test_constraint_evaluation)."""
from types import SimpleNamespace

import pytest

from zk_lisp_b200 import air_eval as pyair
import synth
from util import P, from_fe, to_fe

R = 27
PER_LEN = 1 + R + 1 + 1 + 1 + 1          # p_map, p_round[27], p_final, p_pad, p_pad_last, p_last (poseidon.rs:174)


def _air(zkl, sponge=True, usage_mask=0):
    """A baseline-layout AIR instance (all 219 columns) with the given feature / usage gating; constants are overridden per case."""
    trace, air_pi, opts = synth.random_segment(n=64, seed=3, sponge=sponge, usage_mask=usage_mask, ram_bits=0)
    a = zkl.air.Air(trace.shape[0], trace.shape[1], air_pi, opts)
    assert a.width == 219
    return a


def _zero_suites(a):
    a.pose = SimpleNamespace(dom=[0, 0], mds=[[0] * 12 for _ in range(12)], rc=[[0] * 12 for _ in range(R)])
    a.rom_mds, a.rom_rc = [[0] * 3 for _ in range(3)], [[0] * 3 for _ in range(R)]


def _blocks(zkl, oracle, a, cur, nxt, per):
    """(name -> values) of every block, from the Python evaluator and from the checker's vector cut at the same places."""
    w0, w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0), zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    root_fe = zkl.field.be_from_le8(a.air_pi.core.merkle_root)
    py = {"poseidon": pyair.poseidon_block(a, cur, nxt, per), "ctrl": pyair.ctrl_block(a, cur, nxt, per, 1), "alu": pyair.alu_block(a, cur, nxt, per),
          "ram": pyair.ram_block(a, cur, nxt, per), "merkle": pyair.merkle_block(a, cur, nxt, per, root_fe), "rom": pyair.rom_block(a, cur, nxt, per, w0, w1)}
    vec = from_fe(oracle.eval_transition_frame(a.to_c(), to_fe(cur), to_fe(nxt), to_fe(per)))
    assert len(vec) == sum(len(v) for v in py.values()) == len(a.degrees)
    chk, at = {}, 0
    for name in ("poseidon", "ctrl", "alu", "ram", "merkle", "rom"):
        chk[name] = vec[at:at + len(py[name])]; at += len(py[name])
        assert chk[name] == py[name], name            # the two restatements agree on the crafted frame, value by value
    return py


def _frame(a):
    return [0] * a.width, [0] * a.width, [0] * 32


# ---- vm/air/poseidon.rs:171-496 ------------------------------------------------------------------------------------
def test_poseidon_round_constraints_zero_on_valid_row(zkl, oracle):
    a = _air(zkl)
    a.pose = zkl.params.get_poseidon_suite(bytes([5]) * 32)          # sid = [5u8; 32]
    c = a.cols
    cur, nxt, per = _frame(a)
    j = 3
    per[1 + j] = 1
    for i in range(12):
        cur[c.lane_index(i)] = i + 1
    cur[c.pose_active] = 1
    s3 = [pow(cur[c.lane_index(k)], 3, P) for k in range(12)]
    for i in range(12):
        nxt[c.lane_index(i)] = (sum(a.pose.mds[i][k] * s3[k] for k in range(12)) + a.pose.rc[j][i]) % P
    assert all(v == 0 for v in _blocks(zkl, oracle, a, cur, nxt, per)["poseidon"])
    nxt[c.lane_index(4)] = (nxt[c.lane_index(4)] + 1) % P            # and one wrong lane is caught by exactly its constraint
    bad = _blocks(zkl, oracle, a, cur, nxt, per)["poseidon"]
    assert [i for i, v in enumerate(bad) if v] == [12 * j + 4]


def test_poseidon_hold_constraints_hold_on_pad(zkl, oracle):
    a = _air(zkl); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    per[1 + R + 1], per[1 + R + 2] = 1, 0                            # p_pad, not the last pad row
    for i in range(12):
        cur[c.lane_index(i)] = nxt[c.lane_index(i)] = i + 3
    assert all(v == 0 for v in _blocks(zkl, oracle, a, cur, nxt, per)["poseidon"])


def test_poseidon_round_gated_by_pose_active(zkl, oracle):
    a = _air(zkl); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    per[1 + 1] = 1                                                   # round 1
    for i in range(12):
        cur[c.lane_index(i)], nxt[c.lane_index(i)] = i + 11, i + 7
    cur[c.pose_active] = 0
    assert all(v == 0 for v in _blocks(zkl, oracle, a, cur, nxt, per)["poseidon"])


@pytest.mark.parametrize("bits,lane_value,holds", [((1, 0, 1), 15, True), ((1, 1, 0), 25, False)])
def test_poseidon_sponge_binding_mux(zkl, oracle, bits, lane_value, holds):
    """sponge_binding_mux_positive (index 5 -> r5 = 15) and _negative (index 3 -> r3 = 23, lane holds r5 = 25)."""
    a = _air(zkl, sponge=True, usage_mask=1 << 7); _zero_suites(a)   # FM_VM | FM_SPONGE, VM_USAGE_SPONGE
    c = a.cols
    cur, nxt, per = _frame(a)
    per[0] = 1
    cur[c.pose_active] = cur[c.op_sponge] = 1
    for i in range(8):
        cur[c.r_index(i)] = (10 if holds else 20) + i
    for k, b in enumerate(bits):
        cur[c.sel_s_bits_start + k] = b                             # lane 0: b0, b1, b2
    cur[c.sel_s_active_start] = 1
    cur[c.lane_index(0)] = lane_value
    pose = _blocks(zkl, oracle, a, cur, nxt, per)["poseidon"]
    assert len(pose) == 12 * R + 12 + 10
    assert all(v == 0 for v in pose) if holds else any(v != 0 for v in pose)


# ---- vm/air/ctrl.rs:399-648 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("sponge_on", [False, True])
def test_ctrl_sponge_selectors_gated_by_feature(zkl, oracle, sponge_on):
    """sponge_selectors_gated_off_without_feature / sponge_selectors_on_feature_affect_eval: two frames that differ only in one
    packed sponge-selector bit (2 = invalid against 1 = valid) evaluate alike without the feature and differently with it."""
    a = _air(zkl, sponge=sponge_on, usage_mask=(1 << 7) if sponge_on else 0); _zero_suites(a)
    c = a.cols
    out = []
    for bit in (2, 1):
        cur, nxt, per = _frame(a)
        per[0] = 1
        cur[c.op_sponge] = 1
        cur[c.sel_s_bits_start] = bit
        cur[c.sel_s_active_start] = 1
        out.append(_blocks(zkl, oracle, a, cur, nxt, per)["ctrl"])
    assert len(out[0]) == len(out[1]) == (131 if sponge_on else 91)
    assert (out[0] != out[1]) if sponge_on else (out[0] == out[1])


def test_ctrl_two_ops_set_violation(zkl, oracle):
    a = _air(zkl, sponge=False); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    per[0] = 1
    cur[c.op_add] = cur[c.op_sub] = 1
    assert any(v != 0 for v in _blocks(zkl, oracle, a, cur, nxt, per)["ctrl"])


def test_ctrl_packed_bits_booleanity_violation(zkl, oracle):
    a = _air(zkl, sponge=True, usage_mask=1 << 7); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    per[0] = 1
    cur[c.op_sponge] = 1
    cur[c.sel_s_bits_start] = 2                                      # b0 = 2
    cur[c.sel_s_active_start] = 1
    ctrl = _blocks(zkl, oracle, a, cur, nxt, per)["ctrl"]
    assert any(v != 0 for v in ctrl)
    assert ctrl[53] == 2                                             # 40 selector + 5 role-sum + 8 overlap constraints come first: b0 (b0 - 1) = 2


# ---- vm/air/alu.rs:363-421 -----------------------------------------------------------------------------------------
def test_alu_assert_enforcer_violates_when_false(zkl, oracle):
    a = _air(zkl, sponge=False, usage_mask=1 << 0); _zero_suites(a)  # VM_USAGE_ASSERT
    c = a.cols
    cur, nxt, per = _frame(a)
    per[1 + R] = 1                                                   # final row
    cur[c.op_assert] = 1
    cur[c.sel_c_start + 0] = 1
    cur[c.r_index(0)] = 0                                            # c_val = 0: the asserted condition is false
    alu = _blocks(zkl, oracle, a, cur, nxt, per)["alu"]
    assert any(v != 0 for v in alu)
    assert alu[16] == P - 1                                          # 8 carry + 8 write constraints first: b_assert (c - 1) = -1
    cur[c.r_index(0)] = 1                                            # and with c = 1 the enforcer is satisfied
    nxt[c.r_index(0)] = 1
    assert _blocks(zkl, oracle, a, cur, nxt, per)["alu"][16] == 0


# ---- vm/air/rom.rs:155-288 -----------------------------------------------------------------------------------------
def test_rom_round_constraints_zero_on_valid_row(zkl, oracle):
    a = _air(zkl, sponge=False); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    j = 2
    per[1 + j] = 1
    for i in range(3):
        cur[c.rom_s_index(i)] = i + 3
        a.rom_rc[j][i] = i + 1
        a.rom_mds[i][i] = 1                                          # identity MDS
    for i in range(3):
        nxt[c.rom_s_index(i)] = (pow(cur[c.rom_s_index(i)], 3, P) + a.rom_rc[j][i]) % P
    rom = _blocks(zkl, oracle, a, cur, nxt, per)["rom"]
    assert len(rom) == 3 * R + 3 + 2 and all(v == 0 for v in rom)


def test_rom_map_enc_bindings_hold(zkl, oracle):
    a = _air(zkl, sponge=False); _zero_suites(a)
    c = a.cols
    cur, nxt, per = _frame(a)
    per[0] = 1                                                       # map row
    cur[c.op_add] = 1
    cur[c.sel_dst0_start + 1] = cur[c.sel_a_start + 2] = cur[c.sel_b_start + 3] = cur[c.sel_c_start + 4] = 1
    cur[c.imm], cur[c.eq_inv] = 7, 9
    w0, w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0), zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    cur[c.rom_s_index(1)] = pyair.rom_encode(a, cur, w0)             # utils::rom_linear_encode_from_slice
    cur[c.rom_s_index(2)] = pyair.rom_encode(a, cur, w1)
    rom = _blocks(zkl, oracle, a, cur, nxt, per)["rom"]
    assert all(v == 0 for v in rom)
    cur[c.sel_b_start + 3] = 0                                       # an encoded selector changes: the binding of rom_s[1], rom_s[2] breaks
    cur[c.sel_b_start + 5] = 1
    rom = _blocks(zkl, oracle, a, cur, nxt, per)["rom"]
    assert rom[-2] != 0 and rom[-1] != 0 and all(v == 0 for v in rom[:-2])
