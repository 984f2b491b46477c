"""Host-side mirror of `ZkLispAir::new` / `get_assertions` / public-input encoding.

Everything here is cheap integer bookkeeping the reference also does on the host once per
segment; it produces the plain `zkl_air_params` struct the constraint kernel consumes.
Reference: zk-lisp-proof-winterfell/src/vm/layout.rs (Columns), src/vm/air/mod.rs:114-318,380-504,
src/vm/air/{poseidon,ctrl,alu,ram,merkle,rom}.rs `push_degrees`, src/vm/air/schedule.rs:42-140,
src/vm/air/rom.rs:122-147, src/lib.rs:116-160 (AirPublicInputs::to_elements),
zk-lisp-proof/src/pi.rs (feature mask bits).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import capi, params
from .field import P, be_from_le8

POSEIDON_ROUNDS = 27
STEPS_PER_LEVEL_P2 = 32
NR = 8
SPONGE_IDX_BITS = 3

# zk-lisp-proof/src/pi.rs
FM_POSEIDON, FM_VM, FM_VM_EXPECT, FM_SPONGE, FM_MERKLE, FM_RAM = 1 << 0, 1 << 1, 1 << 4, 1 << 5, 1 << 6, 1 << 7
# vm/layout.rs VM_USAGE_*
(VM_USAGE_ASSERT, VM_USAGE_ASSERT_BIT, VM_USAGE_ASSERT_RANGE, VM_USAGE_DIVMOD, VM_USAGE_MULWIDE,
 VM_USAGE_DIV128, VM_USAGE_EQ, VM_USAGE_SPONGE, VM_USAGE_RAM_DELTA_CLK) = range(9)

OPS = ["const", "mov", "add", "sub", "mul", "neg", "eq", "select", "sponge", "assert", "assert_bit",
       "assert_range", "divmod", "div128", "mulwide", "load", "store"]


@dataclass
class LayoutConfig:  # vm/layout.rs:56-74
    vm: bool = True
    ram: bool = True
    sponge: bool = True
    merkle: bool = True
    rom: bool = True


class Columns:
    """vm/layout.rs:183-314 `Columns::for_config`."""

    def __init__(self, cfg: LayoutConfig):
        cur = 0
        self.lanes_start = cur
        self.lane_l, self.lane_r, self.lane_c0, self.lane_c1 = 0, 1, 10, 11
        cur += 12
        self.g_map = cur
        self.g_final = cur + 1
        self.g_r_start = cur + 2
        cur = self.g_r_start + POSEIDON_ROUNDS
        self.mask = cur
        cur += 1
        self.r_start = cur
        cur += NR
        self.op_start = cur
        for k, name in enumerate(OPS):
            setattr(self, "op_" + name, cur + k)
        cur += len(OPS)
        self.sel_dst0_start = cur
        self.sel_a_start = cur + NR
        self.sel_b_start = cur + 2 * NR
        self.sel_c_start = cur + 3 * NR
        self.sel_dst1_start = cur + 4 * NR
        cur += 5 * NR
        self.sel_s_bits_start = cur
        self.sel_s_active_start = cur + 10 * SPONGE_IDX_BITS
        cur = self.sel_s_active_start + 10
        self.imm = cur
        self.eq_inv = cur + 1
        cur += 2
        self.ram_start = cur  # sorted, s_addr, s_clk, s_val, s_is_write, s_last_write, gp_unsorted, gp_sorted
        if cfg.ram:
            cur += 8
        self.merkle_start = cur  # g, dir, sib, acc, first, last, leaf
        if cfg.merkle:
            cur += 7
        self.pi_prog = cur
        cur += 1
        self.pc = cur
        cur += 1
        self.rom_op_start = cur
        if cfg.rom:
            cur += 17
        self.pose_active = cur
        cur += 1
        self.gadget_b_start = cur
        cur += 32
        self.rom_s_start = cur  # == gadget_b_start + 32 (layout.rs:449-453)
        if cfg.rom:
            cur += 3
        self.width = cur if cfg.rom else self.pc + 1  # layout.rs:310-313

    @staticmethod
    def baseline() -> "Columns":
        return Columns(LayoutConfig())

    def r_index(self, i): return self.r_start + i
    def g_r_index(self, j): return self.g_r_start + j
    def rom_s_index(self, i): return self.rom_s_start + i
    def lane_index(self, i): return self.lanes_start + i

    def to_c(self) -> capi.Columns:
        c = capi.Columns()
        for n in capi.COLUMN_FIELDS:
            setattr(c, n, getattr(self, n))
        return c


@dataclass
class PublicInputs:
    """zk-lisp-proof/src/pi.rs PublicInputs (fields the hot path reads)."""
    program_id: bytes = bytes(32)
    program_commitment: bytes = bytes(32)
    merkle_root: bytes = bytes(32)
    main_args: List[Tuple[str, object]] = field(default_factory=list)  # ("u64", v) | ("u128", v) | ("bytes32", b)
    vm_out_reg: int = 0
    vm_out_row: int = 0
    vm_expected_bytes: bytes = bytes(32)
    feature_mask: int = 0


def encode_main_args_to_slots(args) -> List[int]:
    """utils.rs:79-109."""
    out = []
    for kind, v in args:
        if kind == "u64":
            out.append(v % P)
        elif kind == "u128":
            out.append(v % P)
        elif kind == "bytes32":
            out.append(int.from_bytes(v[0:16], "little") % P)
            out.append(int.from_bytes(v[16:32], "little") % P)
        else:
            raise ValueError(kind)
    return out


@dataclass
class AirPublicInputs:
    """src/lib.rs:74-96."""
    core: PublicInputs
    segment_feature_mask: int = 0
    rom_acc: Tuple[int, int, int] = (0, 0, 0)
    pc_init: int = 0
    ram_gp_unsorted_in: int = 0
    ram_gp_unsorted_out: int = 0
    ram_gp_sorted_in: int = 0
    ram_gp_sorted_out: int = 0
    rom_s_in: Tuple[int, int, int] = (0, 0, 0)
    rom_s_out: Tuple[int, int, int] = (0, 0, 0)
    vm_usage_mask: int = 0
    ram_delta_clk_bits: int = 0

    def to_elements(self) -> List[int]:
        """src/lib.rs:116-160."""
        c = self.core
        out = [c.feature_mask % P, be_from_le8(c.program_commitment), be_from_le8(c.merkle_root)]
        if any(c.program_commitment):
            out += params.program_field_commitment(c.program_commitment)
        else:
            out += [0, 0]
        out += encode_main_args_to_slots(c.main_args)
        out += [self.pc_init, self.ram_gp_unsorted_in, self.ram_gp_unsorted_out, self.ram_gp_sorted_in,
                self.ram_gp_sorted_out, *self.rom_s_in, *self.rom_s_out, self.vm_usage_mask, self.ram_delta_clk_bits]
        return out


@dataclass
class ProofOptions:
    """winterfell::ProofOptions as built at prove.rs:963-972 (+ with_partitions, prove.rs:1106-1107)."""
    num_queries: int = 64
    blowup: int = 16
    grinding: int = 16
    num_partitions: int = 1
    hash_rate: int = 16
    fri_folding: int = 2
    fri_remainder_max_degree: int = 1
    field_extension: int = 1  # FieldExtension::None

    def to_c(self) -> capi.Options:
        return capi.Options(self.num_queries, self.blowup, self.grinding, self.num_partitions, self.hash_rate,
                            self.fri_folding, self.fri_remainder_max_degree, self.field_extension)

    def to_elements(self) -> List[int]:
        """winter-air ProofOptions::to_elements [WF-mem]."""
        buf = self.field_extension
        buf = (buf << 8) | self.fri_folding
        buf = (buf << 8) | self.fri_remainder_max_degree
        return [buf, self.grinding, self.blowup, self.num_queries]


def select_partitions_for_trace(width: int, length: int) -> Tuple[int, int]:
    """utils.rs:394-409."""
    hash_rate = 8 if width <= 32 else 16
    if length >= 1 << 20:
        n = 16
    elif length >= 1 << 18:
        n = 8
    elif length >= 1 << 16:
        n = 4
    elif length >= 1 << 14:
        n = 2
    else:
        n = 1
    return n, hash_rate


def context_to_elements(width: int, length: int, num_constraints: int, opts: ProofOptions) -> List[int]:
    """winter-air Context::to_elements [WF-mem]: trace info, modulus halves, constraint count, options."""
    out = [(width << 8) | 0, length]
    mod = P.to_bytes(16, "little")
    out += [int.from_bytes(mod[:8], "little"), int.from_bytes(mod[8:], "little")]
    out.append(num_constraints)
    out += opts.to_elements()
    return out


# ---- transition constraint degrees (push_degrees of every module) ------------------------------
def transition_degrees(fm, rom_enabled: bool, vm_usage_mask: int, ram_delta_clk_bits: int):
    """Returns [(base, num_cycles)] in `evaluate_transition` block order (vm/air/mod.rs:216-238)."""
    d = []
    cyc = 1
    sponge_used = bool(vm_usage_mask >> VM_USAGE_SPONGE & 1)
    if fm & FM_POSEIDON:  # poseidon.rs:26-65
        d += [(4, cyc)] * (POSEIDON_ROUNDS * 12) + [(1, cyc)] * 12
        if fm & FM_VM and fm & FM_SPONGE and sponge_used:
            d += [(b, cyc) for b in (6, 6, 3, 3, 3, 3, 3, 3, 3, 3)]
    if fm & FM_VM:
        # ctrl.rs:29-112
        d += [(2, cyc)] * (5 * NR) + [(1, cyc)] * 5 + [(2, cyc)] * NR
        if fm & FM_SPONGE and sponge_used:
            d += [(2, cyc)] * (10 * (SPONGE_IDX_BITS + 1))
        d += [(2, cyc)] + [(2, cyc)] * 17 + [(2, cyc)] + [(2, cyc)] * 17 + [(1, cyc)] * 2
        # alu.rs:30-105
        m = vm_usage_mask
        d += [(1, cyc)] * NR + [(7, cyc)] * NR
        if m >> VM_USAGE_EQ & 1:
            d += [(5, cyc)] * 2
        if m >> VM_USAGE_DIVMOD & 1:
            d += [(5, cyc)] * 2
        if m >> VM_USAGE_ASSERT & 1:
            d += [(5, cyc)]
        if m >> VM_USAGE_ASSERT_BIT & 1:
            d += [(5, cyc)]
        if m >> VM_USAGE_ASSERT_RANGE & 1:
            d += [(5, cyc)] * 33
        if m >> VM_USAGE_MULWIDE & 1:
            d += [(5, cyc)]
        if m >> VM_USAGE_DIV128 & 1:
            d += [(5, cyc)] * 2
    if fm & FM_RAM:  # ram.rs:26-79
        d += [(4, cyc), (2, 0), (5, 0), (3, 0), (6, 0), (5, 0)]
        if vm_usage_mask >> VM_USAGE_RAM_DELTA_CLK & 1:
            d += [(5, 0)] * bin(ram_delta_clk_bits & 0xFFFFFFFF).count("1") + [(5, 0)]
        d += [(2, 0)]
    if fm & FM_MERKLE:  # merkle.rs:28-64
        d += [(3, cyc), (3, cyc), (3, cyc), (2, cyc), (3, cyc), (3, cyc), (3, cyc)]
    if rom_enabled:  # rom.rs:27-55
        d += [(3, cyc)] * (POSEIDON_ROUNDS * 3) + [(1, cyc)] * 3 + [(1, cyc)] * 2
    if not d:
        d = [(1, 0)]  # vm/air/mod.rs:302-306
    return d


def ce_blowup_factor(degrees) -> int:
    """AirContext::new [WF-mem]: max over constraints of next_pow2(base + num_cycles), at least 2."""
    def np2(x):
        r = 1
        while r < x:
            r <<= 1
        return r
    return max(max(np2(b + c) for b, c in degrees), 2)


def num_composition_columns(degrees, n: int) -> int:
    """AirContext::num_constraint_composition_columns [WF-mem]; cycles are all of length 32 here."""
    highest = max(b * (n - 1) + c * (n // STEPS_PER_LEVEL_P2) * (STEPS_PER_LEVEL_P2 - 1) for b, c in degrees)
    div_degree = n - 1  # trace_length - num_transition_exemptions(=1)
    return max((highest - div_degree + n - 1) // n, 1)


# ---- assertions ---------------------------------------------------------------------------------
def get_assertions(cols: Columns, fm: int, air_pi: AirPublicInputs, n: int, pose_dom):
    """vm/air/mod.rs:380-504 (ScheduleAir::append_assertions schedule.rs:42-140, RomAir rom.rs:122-147,
    de-duplication by (column, step) keeping the first)."""
    core = air_pi.core
    last = n - 1
    out = []
    vm = bool(fm & FM_VM)
    has_commit = any(core.program_commitment)
    lvls = (last + 1) // STEPS_PER_LEVEL_P2
    for lvl in range(lvls):
        base = lvl * STEPS_PER_LEVEL_P2
        row_map, row_final = base, base + 1 + POSEIDON_ROUNDS
        out.append((cols.lane_c0, row_map, pose_dom[0]))
        out.append((cols.lane_c1, row_map, pose_dom[1]))
        out.append((cols.g_map, row_map, 1))
        out.append((cols.g_final, row_final, 1))
        for j in range(POSEIDON_ROUNDS):
            out.append((cols.g_r_index(j), base + 1 + j, 1))
        out.append((cols.g_final, row_map, 0))
        for j in range(POSEIDON_ROUNDS):
            out.append((cols.g_r_index(j), row_map, 0))
        out.append((cols.g_map, row_final, 0))
        for j in range(POSEIDON_ROUNDS):
            out.append((cols.g_r_index(j), row_final, 0))
        for j in range(POSEIDON_ROUNDS):
            out.append((cols.g_map, base + 1 + j, 0))
            out.append((cols.g_final, base + 1 + j, 0))
        if lvl == 0 and vm:
            if air_pi.pc_init == 0 and has_commit:
                out.append((cols.pi_prog, row_map, be_from_le8(core.program_commitment)))
            out.append((cols.pc, row_map, air_pi.pc_init))
    if vm:
        if fm & FM_VM_EXPECT:
            row = min(core.vm_out_row, last)
            reg = min(core.vm_out_reg, NR - 1)
            out.append((cols.r_index(reg), row, be_from_le8(core.vm_expected_bytes)))
        slots = encode_main_args_to_slots(core.main_args)
        if air_pi.pc_init == 0 and slots:
            assert len(slots) <= NR, "main_args must fit into NR registers"
            for j, v in enumerate(slots):
                out.append((cols.r_index(NR - len(slots) + j), 0, v))
    if has_commit:
        for i in range(3):
            out.append((cols.rom_s_index(i), 0, air_pi.rom_s_in[i]))
        for i in range(3):
            out.append((cols.rom_s_index(i), last, air_pi.rom_s_out[i]))
    if not out:
        out.append((cols.mask, last, 0))
    seen, dedup = set(), []
    for a in out:
        if (a[0], a[1]) in seen:
            continue
        seen.add((a[0], a[1]))
        dedup.append(a)
    return dedup


def declared_num_assertions(fm: int, air_pi: AirPublicInputs, n: int) -> int:
    """vm/air/mod.rs:240-290 (the count handed to AirContext::new)."""
    levels = max(n // STEPS_PER_LEVEL_P2, 1)
    R = POSEIDON_ROUNDS
    num = (2 + R) * levels + (4 * R + 2) * levels + 2 * levels
    core = air_pi.core
    if fm & FM_VM:
        num += 1
        if air_pi.pc_init == 0 and any(core.program_commitment):
            num += 1
        slots = len(encode_main_args_to_slots(core.main_args))
        if slots > 0 and air_pi.pc_init == 0:
            num += slots
        if fm & FM_VM_EXPECT:
            num += 1
    if any(core.program_id):
        num += 6
    return max(num, 1)


class Air:
    """The data `ZkLispAir::new(info, pub_inputs, options)` derives (vm/air/mod.rs:114-318)."""

    def __init__(self, width: int, n: int, air_pi: AirPublicInputs, opts: ProofOptions):
        core = air_pi.core
        self.width, self.n, self.air_pi, self.opts = width, n, air_pi, opts
        self.fm = air_pi.segment_feature_mask if air_pi.segment_feature_mask else core.feature_mask
        fm = self.fm
        rom_layout = any(core.program_id)          # mod.rs:127 (layout gate)
        self.rom_enabled = any(core.program_commitment)  # mod.rs:355 / ctrl.rs:332 (constraint gate)
        baseline_w = Columns.baseline().width
        if width < baseline_w:  # mod.rs:144-166
            cfg = LayoutConfig(vm=bool(fm & FM_VM), ram=bool(fm & FM_RAM), sponge=bool(fm & FM_SPONGE),
                               merkle=bool(fm & FM_MERKLE), rom=rom_layout)
        else:
            cfg = LayoutConfig(rom=rom_layout)
        self.cols = Columns(cfg)
        if self.cols.width != width:
            raise ValueError(f"trace width {width} does not match layout width {self.cols.width}")
        # mod.rs:236-238 pushes ROM degrees under `rom_enabled` (program_id) while eval gates on
        # program_commitment (mod.rs:355); the CLI always sets both (zk-lisp-proof/src/pi.rs:96-105).
        if rom_layout != self.rom_enabled:
            raise ValueError("program_id and program_commitment must be both zero or both non-zero")
        self.degrees = transition_degrees(fm, self.rom_enabled, air_pi.vm_usage_mask, air_pi.ram_delta_clk_bits)
        self.ce_blowup = ce_blowup_factor(self.degrees)
        self.num_composition_cols = num_composition_columns(self.degrees, n)
        self.pose = params.get_poseidon_suite(bytes(core.program_id))
        self.rom_mds, self.rom_rc = params.get_rom_suite(bytes(core.program_id))
        self.program_fe = params.program_field_commitment(core.program_id) if any(core.program_id) else [0, 0]
        self.assertions = get_assertions(self.cols, fm, air_pi, n, self.pose.dom)
        declared = declared_num_assertions(fm, air_pi, n)
        if declared != len(self.assertions):  # Winterfell panics on this mismatch
            raise ValueError(f"declared {declared} assertions but get_assertions yields {len(self.assertions)}")
        self.num_constraints = len(self.degrees) + len(self.assertions)

    def seed_elements(self) -> List[int]:
        """Public-coin seed: Context::to_elements() ++ pub_inputs.to_elements() (agg/fs.rs:69-73)."""
        return context_to_elements(self.width, self.n, self.num_constraints, self.opts) + self.air_pi.to_elements()

    def to_c(self) -> capi.AirParams:
        a = capi.AirParams()
        a.cols = self.cols.to_c()
        fm = self.fm
        a.f_poseidon, a.f_vm, a.f_sponge = int(bool(fm & FM_POSEIDON)), int(bool(fm & FM_VM)), int(bool(fm & FM_SPONGE))
        a.f_merkle, a.f_ram, a.rom_enabled = int(bool(fm & FM_MERKLE)), int(bool(fm & FM_RAM)), int(self.rom_enabled)
        a.vm_usage_mask, a.ram_delta_clk_bits = self.air_pi.vm_usage_mask, self.air_pi.ram_delta_clk_bits
        a.num_transition, a.ce_blowup = len(self.degrees), self.ce_blowup
        a.num_composition_cols, a.num_assertions = self.num_composition_cols, len(self.assertions)
        arr = (capi.Assertion * len(self.assertions))()
        for i, (col, step, val) in enumerate(self.assertions):
            arr[i].column, arr[i].step, arr[i].value = col, step, capi.Fe.of(val)
        self._assertion_array = arr  # keep alive
        a.assertions = C.cast(arr, C.POINTER(capi.Assertion))
        a.pose = capi.suite_to_c(self.pose)
        a.rom = capi.rom_suite_to_c(self.rom_mds, self.rom_rc)
        w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
        for k in range(59):
            a.rom_w_enc0[k], a.rom_w_enc1[k] = capi.Fe.of(w0[k]), capi.Fe.of(w1[k])
        a.program_fe0 = capi.Fe.of(self.program_fe[0])
        a.merkle_root_fe = capi.Fe.of(be_from_le8(self.air_pi.core.merkle_root))
        return a
