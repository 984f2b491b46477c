"""Execution segment planner policy: which level-aligned row ranges become segments.

Reference: zk-lisp-proof-winterfell/src/segment_planner.rs (MAX_SEGMENT_ROWS :24, SegmentFeatures::from_ops :40-87,
WinterfellSegmentPlanner::plan_segments :92-284, compute_segment_features_for_levels :288-303) and
zk-lisp-proof/src/segment.rs:29-57 (Segment). The compiled program is described by what the planner reads of it: the
number of op levels, the op kind of each level (for the per-segment features) and the block table.

With the O(n^2) boundary table gone on the GPU path the better operating point is fewer, larger segments; the policy is
unchanged, only `max_segment_rows` moves (65 536-row segments prove at 635 k rows/s).
"""
import os
from typing import List, Optional, Sequence, Tuple

from .air import STEPS_PER_LEVEL_P2
from .segments import Segment, SegmentFeatures
from .step import InvalidInput

MAX_SEGMENT_ROWS = 1 << 12  # segment_planner.rs:24

ALU_OPS = {"Const", "Mov", "Add", "Sub", "Mul", "Neg", "Eq", "Select", "Assert", "AssertBit", "AssertRange", "AssertRangeLo",
           "AssertRangeHi", "DivMod", "MulWide", "DivMod128"}
RAM_OPS = {"Load", "Store"}
SPONGE_OPS = {"SAbsorbN", "SSqueeze"}
MERKLE_OPS = {"MerkleStepFirst", "MerkleStep", "MerkleStepLast"}


def new_segment(r_start: int, r_end: int) -> Segment:
    """segment.rs:38-46."""
    if r_start >= r_end:
        raise InvalidInput("segment r_start must be < r_end")
    return Segment(r_start, r_end)


def features_from_ops(ops: Sequence[str]) -> SegmentFeatures:
    """segment_planner.rs:40-87 over op names of `zk_lisp_compiler::builder::Op`."""
    f = SegmentFeatures(vm=False)
    for op in ops:
        if op in ALU_OPS:
            f.vm = True
        elif op in RAM_OPS:
            f.vm = f.ram = True
        elif op in SPONGE_OPS:
            f.vm = f.sponge = True
        elif op in MERKLE_OPS:
            f.merkle = True
        elif op != "End":
            raise InvalidInput("unknown op kind")
    return f


def compute_segment_features_for_levels(ops: Sequence[str], lvl_start: int, lvl_end: int) -> SegmentFeatures:
    """segment_planner.rs:288-303: out-of-range levels (the padded tail) are clamped away."""
    start, end = min(lvl_start, len(ops)), min(lvl_end, len(ops))
    return SegmentFeatures(vm=False) if start >= end else features_from_ops(ops[start:end])


def next_power_of_two(x: int) -> int:
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


def plan_segments(base_levels: int, blocks: Sequence[Tuple[int, int]] = (), max_segment_rows: Optional[int] = None) -> List[Segment]:
    """The segments of a program with `base_levels` op levels: the padded level range [0, next_pow2(base_levels)) cut every
    `max_segment_rows // 32` levels (`ZKL_MAX_SEGMENT_ROWS`, default 4096 rows).

    The reference (segment_planner.rs:92-284) walks the block table — blocks merged where they touch, the gaps between them, the
    padded tail — and fills segments greedily, closing one exactly when it holds the level budget. Those ranges tile the padded
    level range without holes and the fill carries an open segment across range borders, so a segment can only ever start where
    the previous one ended: the block table never moves a boundary. What it does contribute is its validation (a block past the
    last op level is an error). The equivalence with the walk is property-tested over random block tables
    (tests/test_step_container.py: the walk is restated there as the checker)."""
    for start, length in blocks:
        if length and start + length > base_levels:
            raise InvalidInput("WinterfellSegmentPlanner block out of bounds for program levels")
    if max_segment_rows is None:
        env = os.environ.get("ZKL_MAX_SEGMENT_ROWS", "")
        max_segment_rows = int(env) if env.isdigit() else MAX_SEGMENT_ROWS
    total_levels = next_power_of_two(base_levels)
    per_segment = total_levels if total_levels * STEPS_PER_LEVEL_P2 <= max_segment_rows else max(max_segment_rows // STEPS_PER_LEVEL_P2, 1)
    cuts = list(range(0, total_levels, per_segment)) + [total_levels]
    return [new_segment(a * STEPS_PER_LEVEL_P2, b * STEPS_PER_LEVEL_P2) for a, b in zip(cuts, cuts[1:])]
