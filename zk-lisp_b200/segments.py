"""Segment driver: host mirror of `prove_program` / `prove_segment`
(zk-lisp-proof-winterfell/src/prove.rs:954-1175) on top of the GPU hot path.

The reference builds the FULL trace once (prove.rs:983), then per segment slices it to the segment layout
(vm/trace/mod.rs:64-226, 367-384), reads the boundary cells (prove.rs:1197-1287), scans the rows for gadget usage
(prove.rs:1289-1392) and proves the segment; segments are independent and run on a rayon pool (prove.rs:1020-1048).
Here the full trace is uploaded once, slicing / scanning / cell reads are device kernels, and segments are proven on
a pool of contexts (one CUDA stream each, possibly on different GPUs) — no collective on the data path.
The planner policy lives in planner.py, the zl1 step container in step.py; the compiler's per-level op kinds stay inputs.
"""
import ctypes as C
import threading
from dataclasses import dataclass, replace
from typing import List, Optional, Sequence

import numpy as np

from . import air as A
from . import capi
from . import step as S
from .prover import Context, ZkProver, as_fe_array, fe_to_ints


@dataclass
class Segment:
    """zk-lisp-proof/src/segment.rs: half-open row range, level aligned."""
    r_start: int
    r_end: int


@dataclass
class SegmentFeatures:
    """segment_planner.rs:29-38 — derived from the program ops of the levels a segment covers."""
    vm: bool = True
    ram: bool = False
    sponge: bool = False
    merkle: bool = False


def compute_segment_feature_mask(core_mask: int, f: SegmentFeatures) -> int:
    """segment_planner.rs:302-334."""
    m = 0
    if core_mask & A.FM_VM:
        m |= A.FM_VM
    if core_mask & A.FM_VM_EXPECT:
        m |= A.FM_VM_EXPECT
    if core_mask & A.FM_RAM and f.ram:
        m |= A.FM_RAM
    if core_mask & A.FM_MERKLE and f.merkle:
        m |= A.FM_MERKLE
    if core_mask & A.FM_SPONGE and f.sponge:
        m |= A.FM_SPONGE
    if core_mask & A.FM_POSEIDON and (f.sponge or f.merkle):
        m |= A.FM_POSEIDON
    return m


def plan_fixed_segments(total_rows: int, max_segment_rows: int = 4096) -> List[Segment]:
    """Level-aligned fixed-size segments (the shape WinterfellSegmentPlanner produces for a uniform program,
    segment_planner.rs:24,92; planner.plan_segments is the general policy)."""
    assert total_rows % 32 == 0 and max_segment_rows % 32 == 0
    return [Segment(r, min(r + max_segment_rows, total_rows)) for r in range(0, total_rows, max_segment_rows)]


class SegmentLayout:
    """vm/trace/mod.rs:64-226 `SegmentLayout::from_full_columns`."""

    def __init__(self, full_cols: A.Columns, cfg: A.LayoutConfig):
        self.cfg, self.cols = cfg, A.Columns(cfg)
        s, f = self.cols, full_cols
        m = {}

        def blk(name, count):
            for i in range(count):
                m[getattr(s, name) + i] = getattr(f, name) + i

        blk("lanes_start", 12)
        m[s.g_map], m[s.g_final] = f.g_map, f.g_final
        blk("g_r_start", A.POSEIDON_ROUNDS)
        m[s.mask] = f.mask
        blk("r_start", A.NR)
        blk("op_start", 17)
        for name in ("sel_dst0_start", "sel_a_start", "sel_b_start", "sel_c_start", "sel_dst1_start"):
            blk(name, A.NR)
        blk("sel_s_bits_start", 30)
        blk("sel_s_active_start", 10)
        m[s.imm], m[s.eq_inv] = f.imm, f.eq_inv
        if cfg.ram:
            blk("ram_start", 8)
        if cfg.merkle:
            blk("merkle_start", 7)
        m[s.pi_prog], m[s.pc] = f.pi_prog, f.pc
        if cfg.rom:
            blk("rom_op_start", 17)
        m[s.pose_active] = f.pose_active
        blk("gadget_b_start", 32)
        if cfg.rom:
            blk("rom_s_start", 3)
        self.seg_to_full = np.array([m[i] for i in range(s.width)], dtype=np.uint32)
        assert len(m) == s.width or not cfg.rom


@dataclass
class StepProof:
    """What `prove_segment` returns (proof bytes + the per-segment public data of proof/step.rs)."""
    segment_index: int
    segments_total: int
    proof: bytes
    air_pi: A.AirPublicInputs
    width: int
    rows: int
    options: A.ProofOptions
    step: Optional[S.StepProof] = None   # zl1 container: digest, state hashes, ZKLSTP1 bytes (proof/step.rs)


class DeviceTrace:
    """The full unified trace (219 columns, baseline layout) resident on one device."""

    def __init__(self, ctx: Context, full_trace: np.ndarray):
        t = as_fe_array(full_trace)
        self.ctx, self.width, self.rows = ctx, t.shape[0], t.shape[1]
        self.dev = C.c_void_p()
        ctx._check(ctx.lib.zkl_dev_alloc(ctx.h, t.nbytes, C.byref(self.dev)))
        ctx._check(ctx.lib.zkl_dev_upload(ctx.h, self.dev, t.ctypes.data, t.nbytes))

    def free(self):
        if self.dev:
            self.ctx.lib.zkl_dev_free(self.ctx.h, self.dev)
            self.dev = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def apply_levels_absorb(ctx: Context, trace_dev, n_rows: int, suite_id: bytes, levels: Sequence[int], inputs: Sequence[Sequence[int]],
                        lanes_start: int = 0) -> None:
    """vm/trace/poseidon.rs:24-96 `apply_level_absorb` for a batch of levels of a trace resident on the device: the 27-round
    permutation rows of every listed level in one launch (the trace builder's only dense arithmetic). `inputs[i]` holds up to 10
    absorbed elements of level `levels[i]`; the suite is the program's (suite_id = program_id)."""
    from . import params
    from .prover import ints_to_fe
    if len(levels) != len(inputs):
        raise ValueError("one input list per level")
    if not levels:
        return
    lv = np.array(levels, dtype=np.uint32)
    flat = []
    for row in inputs:
        row = list(row)[:10]
        flat += row + [0] * (10 - len(row))
    fe = ints_to_fe(flat)
    suite = capi.suite_to_c(params.get_poseidon_suite(bytes(suite_id)))
    ctx._check(ctx.lib.zkl_poseidon_levels_dev(ctx.h, trace_dev, n_rows, lanes_start, C.byref(suite), lv.ctypes.data, fe.ctypes.data, len(levels)))


def read_cells(ctx: Context, dev, n_rows: int, cells) -> List[int]:
    cols = np.array([c for c, _ in cells], dtype=np.uint32)
    rows = np.array([r for _, r in cells], dtype=np.uint32)
    out = np.zeros((len(cells), 2), dtype=np.uint64)
    ctx._check(ctx.lib.zkl_read_cells_dev(ctx.h, dev, n_rows, cols.ctypes.data, rows.ctypes.data, len(cells), out.ctypes.data))
    return fe_to_ints(out)


def prove_segment(ctx: Context, full: DeviceTrace, seg: Segment, index: int, total: int, core: A.PublicInputs,
                  features: SegmentFeatures, base_opts: A.ProofOptions, rom_acc=(0, 0, 0), min_bits: int = 64) -> StepProof:
    """prove.rs:1057-1175 for one segment; the full trace must live on ctx's device. `rom_acc` is the program's ROM
    accumulator (romacc.rs, an input like the program itself), `min_bits` the requested security level echoed in the
    step metadata (prove.rs:959,1135)."""
    full_cols = A.Columns.baseline()
    base_mask = core.feature_mask
    seg_mask = compute_segment_feature_mask(base_mask, features)
    eff_mask = seg_mask if (seg_mask != 0 and seg_mask != base_mask) else base_mask      # prove.rs:1082-1084
    rom = any(core.program_id)
    cfg = A.LayoutConfig(vm=bool(eff_mask & A.FM_VM), ram=bool(eff_mask & A.FM_RAM), sponge=bool(eff_mask & A.FM_SPONGE),
                         merkle=bool(eff_mask & A.FM_MERKLE), rom=rom)
    layout = SegmentLayout(full_cols, cfg)
    n = seg.r_end - seg.r_start
    W = layout.cols.width
    # slice on the device (vm/trace/mod.rs:367-384)
    seg_dev = C.c_void_p()
    ctx._check(ctx.lib.zkl_dev_alloc(ctx.h, W * n * 16, C.byref(seg_dev)))
    try:
        ctx._check(ctx.lib.zkl_slice_segment_dev(ctx.h, full.dev, full.rows, layout.seg_to_full.ctypes.data, W, seg.r_start, n, seg_dev))
        # boundary cells from the FULL trace (prove.rs:1197-1287)
        lvl_first, lvl_last = seg.r_start // 32, (seg.r_end - 1) // 32
        row_map_first, row_final_last = lvl_first * 32, lvl_last * 32 + 28
        fc = full_cols
        cells = [(fc.pc, row_map_first),
                 (fc.ram_start + 6, seg.r_start), (fc.ram_start + 6, seg.r_end - 1),
                 (fc.ram_start + 7, seg.r_start), (fc.ram_start + 7, seg.r_end - 1)]
        cells += [(fc.rom_s_start + i, row_map_first) for i in range(3)] + [(fc.rom_s_start + i, row_final_last) for i in range(3)]
        # register file at the segment's first map row and at its last row: VM state hashes (vm/trace/mod.rs:344-351)
        cells += [(fc.r_index(i), seg.r_start) for i in range(A.NR)] + [(fc.r_index(i), seg.r_end - 1) for i in range(A.NR)]
        v = read_cells(ctx, full.dev, full.rows, cells)
        # gadget usage scan over the segment trace (prove.rs:1289-1392)
        mask, bits = C.c_uint32(), C.c_uint32()
        cc = layout.cols.to_c()
        ctx._check(ctx.lib.zkl_usage_mask_dev(ctx.h, seg_dev, W, n, C.byref(cc), int(cfg.ram), C.byref(mask), C.byref(bits)))
        air_pi = A.AirPublicInputs(core=core, segment_feature_mask=eff_mask, pc_init=v[0], ram_gp_unsorted_in=v[1],
                                   ram_gp_unsorted_out=v[2], ram_gp_sorted_in=v[3], ram_gp_sorted_out=v[4],
                                   rom_s_in=tuple(v[5:8]), rom_s_out=tuple(v[8:11]), vm_usage_mask=mask.value,
                                   ram_delta_clk_bits=bits.value, rom_acc=tuple(rom_acc))
        num_partitions, hash_rate = A.select_partitions_for_trace(W, n)                  # prove.rs:1106
        opts = replace(base_opts, num_partitions=num_partitions, hash_rate=hash_rate)
        prover = ZkProver(ctx, opts, air_pi)
        proof = prover.prove(None, dev_ptr=seg_dev, shape=(W, n))
    finally:
        ctx.lib.zkl_dev_free(ctx.h, seg_dev)
    # zl1 container (prove.rs:1135-1172): suite_id = program_id (prove.rs:985)
    meta = S.StepMeta.from_env(n, opts.blowup, opts.num_queries, min_bits, len(air_pi.to_elements()))
    names = ("ram_gp_unsorted_in", "ram_gp_unsorted_out", "ram_gp_sorted_in", "ram_gp_sorted_out", "rom_s_in_0", "rom_s_in_1",
             "rom_s_in_2", "rom_s_out_0", "rom_s_out_1", "rom_s_out_2")
    boundaries = {name: S.fe_to_bytes_fold(x) for name, x in zip(names, v[1:11])}
    boundaries["state_in_hash"] = S.vm_state_hash(v[11:11 + A.NR])
    boundaries["state_out_hash"] = S.vm_state_hash(v[11 + A.NR:11 + 2 * A.NR])
    zl1 = S.Zl1Proof.new_multi_segment(core.program_id, meta, core, index, total, S.fe_to_bytes_fold(v[0]), boundaries, proof)
    step = S.StepProof(zl1, core, tuple(rom_acc), mask.value, bits.value)
    return StepProof(index, total, proof, air_pi, W, n, opts, step)


def prove_program(contexts: Sequence[Context], full_trace: np.ndarray, segments: Sequence[Segment], core: A.PublicInputs,
                  features: Sequence[SegmentFeatures], base_opts: A.ProofOptions,
                  device_traces: Optional[dict] = None, rom_acc=(0, 0, 0), min_bits: int = 64) -> List[StepProof]:
    """prove.rs:954-1051: one worker per context (the rayon pool of `max_concurrent_segments` threads); segment i is
    proven by worker i mod len(contexts). Contexts may live on different GPUs: the full trace is uploaded once per device."""
    if not segments:
        raise ValueError("WinterfellSegmentPlanner returned no segments")
    traces = device_traces if device_traces is not None else {}
    lock = threading.Lock()
    for c in contexts:
        if c.device not in traces:
            traces[c.device] = DeviceTrace(c, full_trace)
    out: List[Optional[StepProof]] = [None] * len(segments)
    errors = []

    def worker(w):
        try:
            for i in range(w, len(segments), len(contexts)):
                out[i] = prove_segment(contexts[w], traces[contexts[w].device], segments[i], i, len(segments), core, features[i], base_opts,
                                       rom_acc, min_bits)
        except Exception as e:  # surfaced to the caller like the reference's Error::Backend
            with lock:
                errors.append(e)

    th = [threading.Thread(target=worker, args=(w,)) for w in range(len(contexts))]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        raise errors[0]
    return out
