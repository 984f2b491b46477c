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
    """segment_planner.rs:92-284. `blocks` = (level_start, level_len) of the program's block table."""
    total_levels = max(next_power_of_two(base_levels), 1)
    steps = STEPS_PER_LEVEL_P2
    n_rows_full = total_levels * steps
    if max_segment_rows is None:
        env = os.environ.get("ZKL_MAX_SEGMENT_ROWS", "")
        max_segment_rows = int(env) if env.isdigit() else MAX_SEGMENT_ROWS
    if n_rows_full <= max_segment_rows:
        return [new_segment(0, n_rows_full)]
    max_levels = max(max_segment_rows // steps, 1)

    # disjoint level ranges: blocks (merged when overlapping or adjacent), the gaps between them, the padded tail
    ranges: List[List[int]] = []
    block_ranges = []
    for start, length in blocks:
        if length == 0:
            continue
        if start + length > base_levels:
            raise InvalidInput("WinterfellSegmentPlanner block out of bounds for program levels")
        block_ranges.append((start, start + length))
    if not block_ranges:
        if base_levels > 0:
            ranges.append([0, base_levels])
    else:
        block_ranges.sort(key=lambda r: r[0])
        cursor = 0
        for bs, be in block_ranges:
            if cursor < bs:
                ranges.append([cursor, bs])
            if ranges and bs <= ranges[-1][1]:
                ranges[-1][1] = max(ranges[-1][1], be)
            else:
                ranges.append([bs, be])
            cursor = ranges[-1][1]
        if cursor < base_levels:
            ranges.append([cursor, base_levels])
    if base_levels < total_levels:
        ranges.append([base_levels, total_levels])

    # greedy fill: a segment closes exactly when it holds max_levels levels
    seg_levels, cur_start, cur_end = [], None, 0
    for range_start, range_end in ranges:
        lvl = range_start
        while lvl < range_end:
            if cur_start is None:
                cur_start = cur_end = lvl
            take = min(max_levels - (cur_end - cur_start), range_end - lvl)
            cur_end += take
            lvl += take
            if cur_end - cur_start == max_levels:
                seg_levels.append((cur_start, cur_end))
                cur_start, cur_end = None, 0
    if cur_start is not None and cur_start < cur_end:
        seg_levels.append((cur_start, cur_end))

    segments = [new_segment(a * steps, b * steps) for a, b in seg_levels]
    if segments and segments[-1].r_end != n_rows_full:
        raise InvalidInput("WinterfellSegmentPlanner produced segments not covering full trace length")
    return segments
