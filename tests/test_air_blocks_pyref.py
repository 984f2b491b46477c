"""CPU test: four blocks of the constraint system restated independently in Python (zk-lisp_b200/air_eval.py, from the reference text)
against the checker's evaluate_transition on random frames with random periodic values — value by value, index by index."""
import random

import pytest

from zk_lisp_b200 import air_eval as pyair
import synth
from util import P, from_fe, rand_fe, to_fe


@pytest.mark.parametrize("cfg", [dict(), dict(usage_mask=0x7F, ram_bits=0), dict(ram=False, merkle=False), dict(merkle=False, ram_bits=0x80000001)])
def test_poseidon_ram_merkle_rom_blocks_match_python_restatement(zkl, oracle, cfg):
    trace, air_pi, opts = synth.random_segment(seed=13, **cfg)
    a = zkl.air.Air(trace.shape[0], trace.shape[1], air_pi, opts)
    ac = a.to_c()
    rng = random.Random(17)
    w0, w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0), zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    root_fe = zkl.field.be_from_le8(air_pi.core.merkle_root)
    for trial in range(3):
        cur, nxt, per = rand_fe(rng, a.width), rand_fe(rng, a.width), rand_fe(rng, 32)
        if trial == 2:       # structured values: selector / flag columns as bits, a quiet periodic row
            cur = [v & 1 if i % 3 else v for i, v in enumerate(cur)]
            per = [1] + [0] * 31
        got = from_fe(oracle.eval_transition_frame(ac, to_fe(cur), to_fe(nxt), to_fe(per)))
        assert len(got) == len(a.degrees)
        rom = pyair.rom_block(a, cur, nxt, per, w0, w1)
        assert got[-len(rom):] == rom and len(rom) == 86
        end = len(got) - len(rom)
        if a.fm & zkl.air.FM_MERKLE:
            mk = pyair.merkle_block(a, cur, nxt, per, root_fe)
            assert got[end - 7:end] == mk
            end -= 7
        if a.fm & zkl.air.FM_RAM:
            ram = pyair.ram_block(a, cur, nxt, per)
            assert got[end - len(ram):end] == ram, [i for i, (x, y) in enumerate(zip(got[end - len(ram):end], ram)) if x != y]
        pose = pyair.poseidon_block(a, cur, nxt, per)
        assert got[:len(pose)] == pose and len(pose) in (336, 346)
    assert all(0 <= x < P for x in got)


@pytest.mark.parametrize("cfg", [dict(), dict(usage_mask=0x7F, ram_bits=0), dict(usage_mask=0, ram=False, merkle=False),
                                 dict(usage_mask=0x155, merkle=False, ram_bits=0x80000001), dict(usage_mask=0xAA, ram=False),
                                 dict(sponge=False, usage_mask=0x1FF), dict(poseidon=False, sponge=False, merkle=False, usage_mask=0x3F)])
def test_whole_transition_vector_matches_python_restatement(zkl, oracle, cfg):
    """All six blocks (Poseidon, VM control, VM ALU, RAM, Merkle, ROM) in evaluate_transition order, every usage-mask-dependent
    constraint set, against the checker — the complete constraint vector, index by index."""
    trace, air_pi, opts = synth.random_segment(seed=29, **cfg)
    a = zkl.air.Air(trace.shape[0], trace.shape[1], air_pi, opts)
    ac = a.to_c()
    rng = random.Random(31)
    w0, w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0), zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    root_fe = zkl.field.be_from_le8(air_pi.core.merkle_root)
    for trial in range(3):
        cur, nxt, per = rand_fe(rng, a.width), rand_fe(rng, a.width), rand_fe(rng, 32)
        if trial == 1:
            per = [0] * 28 + [1, 0, 0, 0]       # the final row of a level: the write and gadget constraints are live
        got = from_fe(oracle.eval_transition_frame(ac, to_fe(cur), to_fe(nxt), to_fe(per)))
        exp = pyair.evaluate_transition(a, cur, nxt, per, w0, w1, root_fe)
        assert len(got) == len(exp) == len(a.degrees)
        assert got == exp, [i for i, (x, y) in enumerate(zip(got, exp)) if x != y][:8]


@pytest.mark.parametrize("cfg", [dict(), dict(pc_init=5), dict(vm_expect=True, main_args=(("u64", 7), ("bytes32", bytes(range(32))))),
                                 dict(vm=False, sponge=False, ram=False, merkle=False), dict(ram=False, merkle=False, n=128)])
def test_assertions_match_python_restatement(zkl, cfg):
    """Boundary assertions of the host AIR mirror (air.py) against a second restatement of get_assertions: same (column, row,
    value) set, same count — the count enters the proof context and the coin seed, the sorted order the coefficient indices."""
    trace, air_pi, opts = synth.random_segment(seed=41, **cfg)
    a = zkl.air.Air(trace.shape[0], trace.shape[1], air_pi, opts)
    exp = pyair.assertions(a.cols, trace.shape[1], air_pi, a.fm)
    key = lambda t: (t[1], t[0])
    assert sorted(a.assertions, key=key) == sorted(exp, key=key)
    assert len(a.assertions) == len({(c, r) for c, r, _ in a.assertions}) == len(exp)
    assert a.num_constraints == len(a.degrees) + len(exp)
    per_level = 4 + 27 + 1 + 27 + 1 + 27 + 54
    assert len(exp) >= per_level * (trace.shape[1] // 32)


@pytest.mark.parametrize("full_layout", [False, True])
def test_synthetic_trace_satisfies_python_air(zkl, full_layout):
    """The AIR-valid synthetic trace (tests/synth.py) under the Python-integer AIR: every transition constraint vanishes on every
    row but the exempt last one, every assertion holds, and the periodic columns are the 0/1 schedule on the trace domain."""
    n = 64
    trace, air_pi, opts = synth.valid_segment(n=n, seed=3, full_layout=full_layout, start_level=0)
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    t = [[int(trace[c, r, 0]) | (int(trace[c, r, 1]) << 64) for c in range(trace.shape[0])] for r in range(n)]
    w0, w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0), zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    root_fe = zkl.field.be_from_le8(air_pi.core.merkle_root)
    g = zkl.field.root_of_unity(6)
    for row in range(n - 1):
        per = zkl.verify.periodic_at(n, pow(g, row, P))
        assert set(per) <= {0, 1} and per[0] == int(row % 32 == 0) and per[31] == 0
        res = pyair.evaluate_transition(a, t[row], t[row + 1], per, w0, w1, root_fe)
        assert not any(res), (row, [i for i, v in enumerate(res) if v][:5])
    assert zkl.verify.periodic_at(n, pow(g, n - 1, P))[31] == 1
    for col, row, val in pyair.assertions(a.cols, n, air_pi, a.fm):
        assert t[row][col] == val, (col, row)
    # and a broken trace is noticed: a register changed on a carry row
    bad = list(t[5])
    bad[a.cols.r_index(2)] = (bad[a.cols.r_index(2)] + 1) % P
    assert any(pyair.evaluate_transition(a, t[4], bad, zkl.verify.periodic_at(n, pow(g, 4, P)), w0, w1, root_fe))


def test_preflight_mode_of_the_prover_mirror(zkl, capsys):
    """prove.rs:186-211 / preflight.rs: a valid trace passes, a broken one aborts with the reference's message, the row, the
    constraint index and the block it belongs to; off by default."""
    PF = zkl.preflight
    n = 64
    trace, air_pi, opts = synth.valid_segment(n=n, seed=3)
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    assert PF.run(PF.OFF, a, trace) is None and PF.run(PF.CONSOLE, a, trace) is None
    assert "preflight: OK (rows=64" in capsys.readouterr().out
    bad = trace.copy()
    bad[a.cols.r_index(2), 37, 0] ^= 1                      # register r2 changes between two carry rows of level 1
    with pytest.raises(PF.PreflightError, match=r"preflight: constraint \d+ non-zero at row 36") as e:
        PF.run(PF.JSON, a, bad)
    rep = e.value.report
    assert rep.row == 36 and rep.pos == 4 and rep.block == "vm-alu" and not rep.gates_map and rep.rounds[3]
    assert '"row": 36' in capsys.readouterr().out
    lanes = trace.copy()
    lanes[a.cols.lane_index(5), 32 + 7, 0] ^= 1             # a Poseidon lane inside the active level 1 (round rows 33..59)
    with pytest.raises(PF.PreflightError) as e:
        PF.run(PF.CONSOLE, a, lanes)
    assert e.value.report.block == "poseidon" and e.value.report.row == 32 + 6 and e.value.report.lanes_exp is not None
    assert "FAILED row 38" in capsys.readouterr().out
    # the prover mirror: off by default, rejects unknown modes, refuses device-resident traces in preflight mode
    class NoCtx:
        pass
    pv = zkl.ZkProver(NoCtx(), opts, air_pi)
    assert pv.preflight == "off" and pv.with_preflight_mode("json") is pv
    with pytest.raises(ValueError):
        pv.with_preflight_mode("verbose")
    with pytest.raises(PF.PreflightError):
        pv.prove(bad)                                        # aborts before any device call (NoCtx has none)
    with pytest.raises(ValueError, match="host trace"):
        pv.prove(None, dev_ptr=1, shape=(trace.shape[0], n))
