"""Segment-to-rank sharding and cross-rank metric aggregation.

The reference proves execution segments independently (prove.rs:1020-1048: rayon `par_iter` over segments, chain
check disabled in parallel mode at :1042), so the multi-GPU path has NO data-path collective: segment i goes to
GPU i mod G and only timings are reduced (MAX) for reporting.
"""
from typing import List, Tuple


def segments_for_rank(num_segments: int, rank: int, world: int) -> List[int]:
    """Round-robin: segment i -> rank i % world (strong-scaling mode: one fixed job split over the ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, num_segments, world))


def weak_scaling_segments(segments_per_rank: int, rank: int) -> List[int]:
    """Weak scaling: every rank proves its own `segments_per_rank` segments (global ids)."""
    return list(range(rank * segments_per_rank, (rank + 1) * segments_per_rank))


def aggregate(rows_local: int, ms_local: float, dist=None, device=None) -> Tuple[int, float]:
    """(total rows over all ranks, max time over ranks). `dist` = torch.distributed when initialised."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rows_local, ms_local
    import torch
    t_rows = torch.tensor([float(rows_local)], dtype=torch.float64, device=device)
    t_ms = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t_rows, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    return int(t_rows.item()), float(t_ms.item())
