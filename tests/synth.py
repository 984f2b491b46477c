"""Synthetic inputs for tests/bench (test infrastructure, host-side Python ints).

`valid_segment` builds a trace that SATISFIES the zk-lisp VM AIR for a small straight-line program
(const/mov/add/sub/mul/neg/nop levels, optional Poseidon-active levels, ROM accumulator), laid out like
the reference trace builder does for the 204-column segment layout (vm + rom, no ram, no merkle):
  zk-lisp-proof-winterfell/src/vm/trace/mod.rs:386-424 (schedule gates), vm/trace/vm.rs:131-275 (ops on
  the map and final rows, registers carried), vm/trace/rom.rs:34-108 (t=3 accumulator), prove.rs:292-423
  (boundary values). It is NOT the reference trace builder (out of scope); it exists so
  that full proofs can be checked by the restated verifier.
`random_segment` fills every column with random field elements (schedule columns excepted) to exercise every
term of the constraint kernel against the oracle; such a trace does not verify, by construction.
"""
import random

import numpy as np

from zk_lisp_b200 import air as A
from zk_lisp_b200 import params
from zk_lisp_b200.blake3 import blake3
from zk_lisp_b200.field import P

OPS = A.OPS


class _Table:
    """Sparse column-major table: dict[(col,row)] -> int, materialised once into a (W, n, 2) uint64 array."""

    def __init__(self, cols: A.Columns, n: int):
        self.W, self.n, self.d = cols.width, n, {}
        for row in range(n):
            pos = row % 32
            if pos == 0:
                self.d[(cols.g_map, row)] = 1
            if pos == 28:
                self.d[(cols.g_final, row)] = 1
            if 1 <= pos <= 27:
                self.d[(cols.g_r_index(pos - 1), row)] = 1

    def __getitem__(self, c):
        return _Col(self, c)

    def to_np(self):
        out = np.zeros((self.W, self.n, 2), dtype=np.uint64)
        if self.d:
            keys = np.array(list(self.d.keys()), dtype=np.int64)
            vals = list(self.d.values())
            lo = np.array([v & 0xFFFFFFFFFFFFFFFF for v in vals], dtype=np.uint64)
            hi = np.array([v >> 64 for v in vals], dtype=np.uint64)
            out[keys[:, 0], keys[:, 1], 0] = lo
            out[keys[:, 0], keys[:, 1], 1] = hi
        return out


class _Col:
    def __init__(self, t, c):
        self.t, self.c = t, c

    def __getitem__(self, r):
        return self.t.d.get((self.c, r), 0)

    def __setitem__(self, r, v):
        if v:
            self.t.d[(self.c, r)] = v
        else:
            self.t.d.pop((self.c, r), None)


def _empty(cols: A.Columns, n: int):
    return _Table(cols, n)


def _to_np(t, n):
    return t.to_np()


def valid_segment(n=64, queries=8, blowup=16, grind=4, seed=1, poseidon_levels=True, start_level=0, full_layout=False, program_seed=None):
    """`program_seed` names the program (its id and commitment; default: `seed`): segments of ONE program — the rollup-bench shape —
    share it and differ in `seed` (their data) and `start_level` (their first pc)."""
    rng = random.Random(seed)
    program_seed = seed if program_seed is None else program_seed
    program_id = blake3(b"zkl-b200/synth/program-id/%d" % program_seed)
    program_commitment = blake3(b"zkl-b200/synth/program-commitment/%d" % program_seed)
    fm = A.FM_POSEIDON | A.FM_VM | A.FM_SPONGE
    cfg = A.LayoutConfig() if full_layout else A.LayoutConfig(vm=True, ram=False, sponge=True, merkle=False, rom=True)
    cols = A.Columns(cfg)
    t = _empty(cols, n)
    suite = params.get_poseidon_suite(program_id)
    rom_mds, rom_rc = params.get_rom_suite(program_id)
    w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
    levels = n // 32
    regs = [rng.randrange(P) for _ in range(8)] if start_level else [0] * 8
    s0_prev = rng.randrange(P) if start_level else 0
    kinds = ["const", "mov", "add", "sub", "mul", "neg", "nop"]
    for lvl in range(levels):
        base = lvl * 32
        for r in range(base, base + 32):
            t[cols.pc][r] = start_level + lvl
        t[cols.lane_c0][base], t[cols.lane_c1][base] = suite.dom
        if lvl == 0 and start_level == 0:
            t[cols.pi_prog][base] = A.be_from_le8(program_commitment)
        kind = rng.choice(kinds)
        dst, a, b = rng.randrange(8), rng.randrange(8), rng.randrange(8)
        imm = rng.randrange(1 << 64)
        new_regs = list(regs)
        for row in (base, base + 28):  # map and final rows carry the op and its selectors (vm.rs:200-275)
            if kind != "nop":
                t[getattr(cols, "op_" + kind)][row] = 1
                t[cols.sel_dst0_start + dst][row] = 1
                if kind in ("mov", "add", "sub", "mul", "neg"):
                    t[cols.sel_a_start + a][row] = 1
                if kind in ("add", "sub", "mul"):
                    t[cols.sel_b_start + b][row] = 1
                if kind == "const":
                    t[cols.imm][row] = imm
        if kind != "nop":
            t[cols.rom_op_start + OPS.index(kind)][base] = 1
            res = {"const": imm, "mov": regs[a], "add": regs[a] + regs[b], "sub": regs[a] - regs[b],
                   "mul": regs[a] * regs[b], "neg": -regs[a]}[kind] % P
            new_regs[dst] = res
        for r in range(base, base + 29):
            for i in range(8):
                t[cols.r_start + i][r] = regs[i]
        for r in range(base + 29, base + 32):
            for i in range(8):
                t[cols.r_start + i][r] = new_regs[i]
        regs = new_regs
        # Poseidon lanes: an active level runs the 27 rounds of the program suite (poseidon.rs:83-118)
        if poseidon_levels and lvl % 3 == 1:
            st = [rng.randrange(P) for _ in range(10)] + list(suite.dom)
            for i in range(12):
                t[cols.lanes_start + i][base] = st[i]
                t[cols.lanes_start + i][base + 1] = st[i]
            for r in range(base, base + 32):
                t[cols.pose_active][r] = 1
            for j in range(27):
                c3 = [pow(v, 3, P) for v in st]
                st = [(sum(suite.mds[i][k] * c3[k] for k in range(12)) + suite.rc[j][i]) % P for i in range(12)]
                for i in range(12):
                    t[cols.lanes_start + i][base + 2 + j] = st[i]
            for r in range(base + 29, base + 32):
                for i in range(12):
                    t[cols.lanes_start + i][r] = st[i]
        # ROM accumulator (vm/trace/rom.rs:34-108)
        enc_cols = [cols.op_start + k for k in range(17)]
        for start in (cols.sel_dst0_start, cols.sel_a_start, cols.sel_b_start, cols.sel_c_start, cols.sel_dst1_start):
            enc_cols += [start + i for i in range(8)]
        s1 = sum(t[c][base] * w0[k] for k, c in enumerate(enc_cols)) % P
        s2 = sum(t[c][base] * w1[k] for k, c in enumerate(enc_cols)) % P
        s = [s0_prev, s1, s2]
        for i in range(3):
            t[cols.rom_s_start + i][base] = s[i]
        for j in range(27):
            r = base + 1 + j
            for i in range(3):
                t[cols.rom_s_start + i][r] = s[i]
            c3 = [pow(v, 3, P) for v in s]
            s = [(sum(rom_mds[i][k] * c3[k] for k in range(3)) + rom_rc[j][i]) % P for i in range(3)]
            for i in range(3):
                t[cols.rom_s_start + i][r + 1] = s[i]
        for r in range(base + 29, base + 32):
            for i in range(3):
                t[cols.rom_s_start + i][r] = s[i]
        s0_prev = s[0]
    core = A.PublicInputs(program_id=program_id, program_commitment=program_commitment, feature_mask=fm)
    air_pi = A.AirPublicInputs(
        core=core, segment_feature_mask=fm, pc_init=start_level,
        rom_s_in=tuple(t[cols.rom_s_start + i][0] for i in range(3)),
        rom_s_out=tuple(t[cols.rom_s_start + i][n - 1] for i in range(3)),
        vm_usage_mask=0, ram_delta_clk_bits=0)
    num_partitions, hash_rate = A.select_partitions_for_trace(cols.width, n)
    opts = A.ProofOptions(num_queries=queries, blowup=blowup, grinding=grind, num_partitions=num_partitions,
                          hash_rate=hash_rate)
    return _to_np(t, n), air_pi, opts


def random_segment(n=64, seed=2, ram=True, merkle=True, usage_mask=0x1FF, ram_bits=0xF00F00FF, queries=8, blowup=16,
                   grind=4, poseidon=True, vm=True, sponge=True, pc_init=0, main_args=(), vm_expect=False,
                   num_partitions=None, hash_rate=None):
    """Random columns (every constraint term non-trivial). Returns (trace, air_pi, opts)."""
    from util import fast_random_trace
    program_id = blake3(b"zkl-b200/synth/random/%d" % seed)
    program_commitment = blake3(b"zkl-b200/synth/random-commit/%d" % seed)
    fm = (A.FM_POSEIDON if poseidon else 0) | (A.FM_VM if vm else 0) | (A.FM_SPONGE if sponge else 0)
    fm |= (A.FM_RAM if ram else 0) | (A.FM_MERKLE if merkle else 0) | (A.FM_VM_EXPECT if vm_expect else 0)
    cfg = A.LayoutConfig(vm=vm, ram=ram, sponge=sponge, merkle=merkle, rom=True)
    cols = A.Columns(cfg)
    trace = fast_random_trace(cols.width, n, seed)
    rng = random.Random(seed)
    core = A.PublicInputs(program_id=program_id, program_commitment=program_commitment, feature_mask=fm,
                          merkle_root=blake3(b"root%d" % seed), main_args=list(main_args),
                          vm_out_reg=3, vm_out_row=min(29, n - 1), vm_expected_bytes=blake3(b"exp%d" % seed))
    air_pi = A.AirPublicInputs(core=core, segment_feature_mask=fm, pc_init=pc_init,
                               rom_s_in=tuple(rng.randrange(P) for _ in range(3)),
                               rom_s_out=tuple(rng.randrange(P) for _ in range(3)),
                               vm_usage_mask=usage_mask, ram_delta_clk_bits=ram_bits)
    np_, hr = A.select_partitions_for_trace(cols.width, n)
    opts = A.ProofOptions(num_queries=queries, blowup=blowup, grinding=grind,
                          num_partitions=num_partitions if num_partitions is not None else np_,
                          hash_rate=hash_rate if hash_rate is not None else hr)
    return trace, air_pi, opts


# ---------------------------------------------------------------------------------------------------------------
# rich_segment: an AIR-valid trace that exercises every gadget inside an ACCEPTED proof — RAM (load / store, the sorted
# table with last-write witnesses and the delta_clk range gadget), eq / select / assert / assert_bit, the 32- and 64-bit
# range checks, divmod / mulwide / div128, sponge absorb + squeeze and (full layout) a Merkle path. Rows are written the
# way the reference trace builder writes them: vm/trace/vm.rs:199-843 (one op per 32-row level: decode on the map and
# final rows, registers before the write up to the final row, after it on the pad rows), vm/trace/poseidon.rs:24-96
# (apply_level_absorb), vm/trace/ram.rs:44-270 (sorted events on pad rows, mirrored between same-address events, eq_inv =
# inverse of the address step, delta_clk bits), vm/trace/rom.rs:34-108. The usage masks are computed like
# prove.rs:1289-1392 scans them. Test infrastructure, not the reference trace builder.
# ---------------------------------------------------------------------------------------------------------------
def _inv(x):
    return pow(x % P, P - 2, P) if x % P else 0


def _program(rng, levels, merkle):
    """One op per level. Registers r0..r7; the script keeps the facts the gadgets need (booleans, small values, non-zero
    divisors) in known registers and sprinkles plain ALU ops between the gadget blocks."""
    prog = []
    emit = prog.append
    # r0 = 64-bit value, r1 = 32-bit value, r2 = boolean, r3 = non-zero 64-bit divisor, r4/r5 scratch, r6 = address, r7 scratch
    emit(("const", 0, rng.randrange(1 << 63, 1 << 64)))
    emit(("const", 1, rng.randrange(1 << 31, 1 << 32)))
    emit(("const", 2, 1))
    emit(("const", 3, rng.randrange(2, 1 << 40)))
    emit(("const", 6, rng.randrange(1, 1 << 20)))
    # every address is written before it is read: the RAM AIR seeds `last_write` of a new address group with the PREVIOUS group's
    # last write (vm/air/ram.rs:176-180), so a first-event read of 0 only passes when that happens to be a read too
    emit(("store", 6, 0)); emit(("store", 3, 1)); emit(("store", 1, 0))
    essential = [
        [("store", 6, 0), ("load", 4, 6), ("eq", 5, 4, 0), ("assert", 7, 5)],                      # a store read back, checked equal
        [("assert_bit", 7, 2), ("select", 4, 2, 0, 1)],
        [("assert_range", 7, 1, 32), ("range_lo", 7, 0), ("range_hi", 7, 0)],
        [("divmod", 4, 5, 0, 3), ("mulwide", 5, 4, 0, 1), ("div128", 7, 4, 1, 0, 3)],
        [("absorb", [0, 1, 3]), ("absorb", [6]), ("squeeze", 4)],
    ]
    if merkle:
        essential.append([("sub", 7, 2, 2), ("merkle_first", 0, 2, 1), ("merkle_step", 7, 3), ("merkle_step", 2, 4), ("merkle_last", 7, 0)])   # r7 = 0: both directions
    optional = [
        [("load", 5, 3), ("store", 3, 1), ("load", 7, 3)],                                        # read, overwrite, read back
        [("add", 6, 6, 2), ("store", 6, 4), ("sub", 6, 6, 2), ("load", 5, 6)],                    # neighbouring addresses
    ]
    plain = ["const", "mov", "add", "sub", "mul", "neg", "nop"]
    rng.shuffle(essential)
    rng.shuffle(optional)
    for blk in essential + optional:
        if len(prog) + len(blk) <= levels:
            prog.extend(blk)
    merkle_done = merkle and any(op[0] == "merkle_last" for op in prog)
    while len(prog) < levels:
        kind = rng.choice(plain + ["store", "load"])
        if kind == "const":
            emit(("const", rng.choice([4, 5, 7]), rng.randrange(1 << 64)))
        elif kind in ("mov", "neg"):
            emit((kind, rng.choice([4, 5, 7]), rng.randrange(8)))
        elif kind in ("add", "sub", "mul"):
            emit((kind, rng.choice([4, 5, 7]), rng.randrange(8), rng.randrange(8)))
        elif kind == "store":
            emit(("store", rng.choice([6, 3, 1]), rng.randrange(8)))
        elif kind == "load":
            emit(("load", rng.choice([4, 5, 7]), rng.choice([6, 3, 1])))
        else:
            emit(("nop",))
    return prog, merkle_done


def rich_segment(n=1024, queries=8, blowup=16, grind=4, seed=1, merkle=False, start_level=0, program_seed=None):
    """(trace, air_pi, opts) of an AIR-valid segment with RAM on (W = 212) or, with `merkle`, the full 219-column layout."""
    rng = random.Random(seed)
    program_seed = seed if program_seed is None else program_seed
    program_id = blake3(b"zkl-b200/synth/rich-program-id/%d" % program_seed)
    program_commitment = blake3(b"zkl-b200/synth/rich-program-commitment/%d" % program_seed)
    fm = A.FM_POSEIDON | A.FM_VM | A.FM_SPONGE | A.FM_RAM | (A.FM_MERKLE if merkle else 0)
    cols = A.Columns(A.LayoutConfig(vm=True, ram=True, sponge=True, merkle=merkle, rom=True))
    t = _empty(cols, n)
    suite = params.get_poseidon_suite(program_id)
    rom_mds, rom_rc = params.get_rom_suite(program_id)
    w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
    levels = n // 32
    assert levels >= 32, "rich_segment needs at least 32 levels (n >= 1024) to fit every gadget block"
    prog, merkle_done = _program(rng, levels, merkle)
    assert not merkle or merkle_done
    regs = [0] * 8
    mem, events, pending = {}, [], []
    merkle_prev_out, merkle_root_fe = None, 0
    ram = cols.ram_start      # sorted, s_addr, s_clk, s_val, s_is_write, s_last_write, gp_unsorted, gp_sorted
    mk = cols.merkle_start    # g, dir, sib, acc, first, last, leaf
    usage = 0
    s0_prev = rng.randrange(P) if start_level else 0

    def absorb_level(base, inputs):                     # vm/trace/poseidon.rs:24-96
        st = [v % P for v in inputs[:10]] + [0] * (10 - len(inputs[:10])) + list(suite.dom)
        for i in range(12):
            t[cols.lanes_start + i][base] = st[i]
            t[cols.lanes_start + i][base + 1] = st[i]
        for j in range(27):
            c3 = [pow(v, 3, P) for v in st]
            st = [(sum(suite.mds[i][k] * c3[k] for k in range(12)) + suite.rc[j][i]) % P for i in range(12)]
            for i in range(12):
                t[cols.lanes_start + i][base + 2 + j] = st[i]
        for r in range(base + 29, base + 32):
            for i in range(12):
                t[cols.lanes_start + i][r] = st[i]
        return st[0]

    for lvl, op in enumerate(prog):
        base, row_map, row_final = lvl * 32, lvl * 32, lvl * 32 + 28
        pc = start_level + lvl
        kind = op[0]
        nxt = list(regs)
        pose_active = 0
        t[cols.lane_c0][row_map], t[cols.lane_c1][row_map] = suite.dom
        if lvl == 0 and start_level == 0:
            t[cols.pi_prog][row_map] = A.be_from_le8(program_commitment)

        def both(col, v):
            t[col][row_map] = v
            t[col][row_final] = v

        def sel(start, i):
            both(start + i, 1)

        opbit = {"range_lo": "assert_range", "range_hi": "assert_range", "absorb": "sponge", "squeeze": "sponge"}.get(kind, kind)
        if kind not in ("nop", "merkle_first", "merkle_step", "merkle_last"):
            both(getattr(cols, "op_" + opbit), 1)
            t[cols.rom_op_start + OPS.index(opbit)][row_map] = 1
        if kind == "const":
            sel(cols.sel_dst0_start, op[1]); both(cols.imm, op[2]); nxt[op[1]] = op[2] % P
        elif kind in ("mov", "neg"):
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_a_start, op[2])
            nxt[op[1]] = regs[op[2]] if kind == "mov" else (-regs[op[2]]) % P
        elif kind in ("add", "sub", "mul"):
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_a_start, op[2]); sel(cols.sel_b_start, op[3])
            a, b = regs[op[2]], regs[op[3]]
            nxt[op[1]] = {"add": a + b, "sub": a - b, "mul": a * b}[kind] % P
        elif kind == "eq":
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_a_start, op[2]); sel(cols.sel_b_start, op[3])
            diff = (regs[op[2]] - regs[op[3]]) % P
            both(cols.eq_inv, _inv(diff)); nxt[op[1]] = int(diff == 0)
            usage |= 1 << A.VM_USAGE_EQ
        elif kind == "select":
            _, dst, c, a, b = op
            sel(cols.sel_dst0_start, dst); sel(cols.sel_c_start, c); sel(cols.sel_a_start, a); sel(cols.sel_b_start, b)
            assert regs[c] in (0, 1)
            nxt[dst] = regs[a] if regs[c] else regs[b]
            usage |= 1 << A.VM_USAGE_ASSERT
        elif kind == "assert":
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_c_start, op[2])
            assert regs[op[2]] == 1
            nxt[op[1]] = 1; usage |= 1 << A.VM_USAGE_ASSERT
        elif kind == "assert_bit":
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_c_start, op[2])
            assert regs[op[2]] in (0, 1)
            nxt[op[1]] = 1; usage |= 1 << A.VM_USAGE_ASSERT_BIT
        elif kind in ("assert_range", "range_lo", "range_hi"):
            dst, r = op[1], op[2]
            sel(cols.sel_dst0_start, dst); sel(cols.sel_c_start, r)
            x = regs[r]
            both(cols.imm, 0 if kind == "range_lo" else 1)
            both(cols.eq_inv, 0 if kind == "assert_range" else 1)           # mode64
            word = (x >> 32) if kind == "range_hi" else x
            if kind == "assert_range":
                assert x < 1 << 32
            if kind == "range_hi":
                assert x < 1 << 64 and regs[dst] == x & 0xFFFFFFFF          # the lo stage left sum_lo in dst
            for i in range(32):
                both(cols.gadget_b_start + i, (word >> i) & 1)
            nxt[dst] = (x & 0xFFFFFFFF) if kind == "range_lo" else 1
            usage |= 1 << A.VM_USAGE_ASSERT_RANGE
        elif kind in ("divmod", "mulwide", "div128"):
            if kind == "divmod":
                _, dq, dr, a, b = op
                av, bv = regs[a], regs[b]
                assert 0 < bv < 1 << 64 and av < 1 << 64
                q, r = divmod(av, bv)
                both(cols.eq_inv, _inv(bv)); usage |= 1 << A.VM_USAGE_DIVMOD
            elif kind == "mulwide":
                _, dq, dr, a, b = op                                       # dst0 = lo, dst1 = hi
                av, bv = regs[a] & (2**64 - 1), regs[b] & (2**64 - 1)
                assert regs[a] < 1 << 64 and regs[b] < 1 << 64
                q, r = (av * bv) & (2**64 - 1), (av * bv) >> 64
                usage |= 1 << A.VM_USAGE_MULWIDE
            else:
                _, dq, dr, a, alo, b = op                                  # (a_hi * 2^64 + a_lo) / b
                hi, lo, bv = regs[a], regs[alo], regs[b]
                assert 0 < bv < 1 << 64 and hi < bv and lo < 1 << 64      # quotient fits 64 bits
                q, r = divmod((hi << 64) | lo, bv)
                both(cols.imm, lo); both(cols.eq_inv, _inv(bv)); usage |= 1 << A.VM_USAGE_DIV128
            assert dq != dr and q < 1 << 64 and r < 1 << 64
            sel(cols.sel_dst0_start, dq); sel(cols.sel_dst1_start, dr); sel(cols.sel_a_start, a); sel(cols.sel_b_start, b)
            nxt[dq], nxt[dr] = q, r
        elif kind in ("absorb", "squeeze"):
            usage |= 1 << A.VM_USAGE_SPONGE
            lanes = list(op[1]) if kind == "absorb" else list(pending)
            for i, r in enumerate(lanes):
                for k in range(3):
                    both(cols.sel_s_bits_start + 3 * i + k, (r >> k) & 1)
                both(cols.sel_s_active_start + i, 1)
            if kind == "absorb":
                pending += lanes
                assert len(pending) <= 10
            else:
                t[cols.sel_dst0_start + op[1]][row_final] = 1                # dst selector on the final row only (vm.rs:570)
                pose_active = 1
                nxt[op[1]] = absorb_level(base, [regs[r] for r in lanes])
                pending = []
        elif kind in ("merkle_first", "merkle_step", "merkle_last"):
            pose_active = 1
            for r in range(base, base + 32):
                t[mk][r] = 1
            if kind == "merkle_first":
                leaf, d, s = regs[op[1]], regs[op[2]], regs[op[3]]
                t[mk + 4][row_map] = 1; t[mk + 6][row_map] = leaf
                acc = leaf
            else:
                d, s = regs[op[1]], regs[op[2]]
                acc = merkle_prev_out
            assert d in (0, 1)
            for r in range(row_map, row_final):
                t[mk + 3][r] = acc
            t[mk + 1][row_map], t[mk + 2][row_map] = d, s
            out = absorb_level(base, [s, acc] if d else [acc, s])
            if kind == "merkle_last":
                t[mk + 5][row_final] = 1
                merkle_root_fe = out
            for r in range(row_final, base + 32):
                t[mk + 3][r] = out
            merkle_prev_out = out
        elif kind == "load":
            sel(cols.sel_dst0_start, op[1]); sel(cols.sel_a_start, op[2])
            addr = regs[op[2]]
            val = mem.get(addr, 0)
            both(cols.imm, val); nxt[op[1]] = val
            events.append((addr, pc, val, 0))
        elif kind == "store":
            sel(cols.sel_a_start, op[1]); sel(cols.sel_b_start, op[2])
            addr, val = regs[op[1]], regs[op[2]]
            mem[addr] = val
            events.append((addr, pc, val, 1))
        # registers: the pre-write state through the final row, the written state on the pad rows; pc; pose_active
        for r in range(base, base + 29):
            for i in range(8):
                t[cols.r_start + i][r] = regs[i]
        for r in range(base + 29, base + 32):
            for i in range(8):
                t[cols.r_start + i][r] = nxt[i]
        for r in range(base, base + 32):
            t[cols.pc][r] = pc
            t[cols.pose_active][r] = pose_active
        regs = nxt
        # ROM accumulator (vm/trace/rom.rs:34-108)
        enc_cols = [cols.op_start + k for k in range(17)]
        for start in (cols.sel_dst0_start, cols.sel_a_start, cols.sel_b_start, cols.sel_c_start, cols.sel_dst1_start):
            enc_cols += [start + i for i in range(8)]
        s = [s0_prev, sum(t[c][base] * w0[k] for k, c in enumerate(enc_cols)) % P, sum(t[c][base] * w1[k] for k, c in enumerate(enc_cols)) % P]
        for i in range(3):
            t[cols.rom_s_start + i][base] = s[i]
        for j in range(27):
            r = base + 1 + j
            for i in range(3):
                t[cols.rom_s_start + i][r] = s[i]
            c3 = [pow(v, 3, P) for v in s]
            s = [(sum(rom_mds[i][k] * c3[k] for k in range(3)) + rom_rc[j][i]) % P for i in range(3)]
            for i in range(3):
                t[cols.rom_s_start + i][r + 1] = s[i]
        for r in range(base + 29, base + 32):
            for i in range(3):
                t[cols.rom_s_start + i][r] = s[i]
        s0_prev = s[0]

    # ---- RAM tables (vm/trace/ram.rs:44-270) ----
    pi0 = params.program_field_commitment(program_id)[0]
    r1, r2, r3 = (pi0 * pi0 + 1) % P, (pow(pi0, 3, P) + pi0) % P, (pow(pi0, 5, P) + 7) % P
    sorted_events = sorted(events, key=lambda e: (e[0], e[1]))
    pad_rows = [r for r in range(n) if r % 32 >= 29]
    assert len(sorted_events) <= len(pad_rows)
    event_rows = pad_rows[:len(sorted_events)]
    for row, (addr, clk, val, w) in zip(event_rows, sorted_events):
        t[ram][row] = 1
        t[ram + 1][row], t[ram + 2][row], t[ram + 3][row], t[ram + 4][row] = addr, clk, val, w
    for i in range(len(sorted_events) - 1):           # mirror the witness between two events of one address
        if sorted_events[i][0] != sorted_events[i + 1][0]:
            continue
        addr, clk, val, w = sorted_events[i]
        for row in range(event_rows[i] + 1, event_rows[i + 1]):
            if t[ram][row] == 0:
                t[ram + 1][row], t[ram + 2][row], t[ram + 3][row], t[ram + 4][row] = addr, clk, val, w
    gs, last = rng.randrange(P) if start_level else 0, 0
    gu = gs
    gp_in = gs
    ram_bits = 0
    for row in range(n):
        if row > 0 and t[ram][row - 1] == 1:
            pa, pclk, pval, pw = (t[ram + k][row - 1] for k in (1, 2, 3, 4))
            gs = (gs + pa + r1 * pclk + r2 * pval + r3 * pw) % P
            last = ((1 - pw) * last + pw * pval) % P if t[ram + 1][row] == pa else pw * pval % P
        if row > 0 and (row - 1) % 32 == 28 and (t[cols.op_load][row - 1] or t[cols.op_store][row - 1]):
            prev = row - 1
            a_ev = sum(t[cols.sel_a_start + i][prev] * t[cols.r_start + i][prev] for i in range(8)) % P
            b_ev = sum(t[cols.sel_b_start + i][prev] * t[cols.r_start + i][prev] for i in range(8)) % P
            w_ev = t[cols.op_store][prev]
            val_ev = b_ev if w_ev else t[cols.imm][prev]
            gu = (gu + a_ev + r1 * t[cols.pc][prev] + r2 * val_ev + r3 * w_ev) % P
        t[ram + 7][row], t[ram + 5][row], t[ram + 6][row] = gs, last, gu
    for row in range(n - 1):
        if t[ram][row] == 1:
            d_addr = (t[ram + 1][row + 1] - t[ram + 1][row]) % P
            t[cols.eq_inv][row] = _inv(d_addr)
            if t[ram][row + 1] == 1 and d_addr == 0:
                delta = max(t[ram + 2][row + 1] - t[ram + 2][row], 0)
                for i in range(32):
                    t[cols.gadget_b_start + i][row] = (delta >> i) & 1
                usage |= 1 << A.VM_USAGE_RAM_DELTA_CLK
                ram_bits |= delta & 0xFFFFFFFF
    assert t[ram + 6][n - 1] == t[ram + 7][n - 1], "every RAM event must be consumed by the sorted table before the last row"

    core = A.PublicInputs(program_id=program_id, program_commitment=program_commitment, feature_mask=fm,
                          merkle_root=merkle_root_fe.to_bytes(16, "little") + bytes(16))
    air_pi = A.AirPublicInputs(
        core=core, segment_feature_mask=fm, pc_init=start_level,
        ram_gp_unsorted_in=gp_in, ram_gp_unsorted_out=t[ram + 6][n - 1], ram_gp_sorted_in=gp_in, ram_gp_sorted_out=t[ram + 7][n - 1],
        rom_s_in=tuple(t[cols.rom_s_start + i][0] for i in range(3)),
        rom_s_out=tuple(t[cols.rom_s_start + i][n - 1] for i in range(3)),
        vm_usage_mask=usage, ram_delta_clk_bits=ram_bits)
    num_partitions, hash_rate = A.select_partitions_for_trace(cols.width, n)
    opts = A.ProofOptions(num_queries=queries, blowup=blowup, grinding=grind, num_partitions=num_partitions, hash_rate=hash_rate)
    return _to_np(t, n), air_pi, opts
