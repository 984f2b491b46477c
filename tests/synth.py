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
