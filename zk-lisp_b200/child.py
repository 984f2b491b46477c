"""GPU versions of the hashing helpers the aggregation layer replays per child proof
(zk-lisp-proof-winterfell/src/agg/child.rs:1025-1068): leaf digests of queried rows and Merkle path folding.
SURVEY.md §8f N2: with proving down to milliseconds, re-hashing every queried row / path of all children on the CPU
hasher becomes the wall-time floor; these run them as batches on the device through the C ABI."""
from typing import List, Sequence

import numpy as np

from .prover import Context, as_fe_array


def hash_row_poseidon(ctx: Context, rows: np.ndarray, partition_size: int) -> np.ndarray:
    """agg/child.rs:1025-1045 for a batch of rows (k, W, 2) -> (k, 32) digests."""
    r = as_fe_array(rows)
    k, W = r.shape[0], r.shape[1]
    if partition_size == 0:
        return np.tile(np.frombuffer(ctx.hash_bytes(b""), dtype=np.uint8), (k, 1))
    chunks = [r[:, i:i + partition_size] for i in range(0, W, partition_size)]
    digests = [ctx.hash_elements(np.ascontiguousarray(c)) for c in chunks]
    if len(digests) == 1:
        return digests[0]
    out = np.zeros((k, 32), dtype=np.uint8)
    for q in range(k):
        out[q] = np.frombuffer(ctx.merge_many(np.stack([d[q] for d in digests])), dtype=np.uint8)
    return out


def merkle_roots_from_leaves(ctx: Context, leaves: np.ndarray, indexes: Sequence[int], siblings: List[np.ndarray]) -> np.ndarray:
    """agg/child.rs:1049-1068 `merkle_root_from_leaf` for a batch of equal-depth paths: one merge batch per level.
    leaves (k,32); siblings[q] (depth,32)."""
    acc = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32).copy()
    idx = [int(i) for i in indexes]
    depth = siblings[0].shape[0] if len(siblings) else 0
    for lvl in range(depth):
        pairs = np.zeros((acc.shape[0], 64), dtype=np.uint8)
        for q in range(acc.shape[0]):
            sib = siblings[q][lvl]
            if idx[q] & 1 == 0:
                pairs[q, :32], pairs[q, 32:] = acc[q], sib
            else:
                pairs[q, :32], pairs[q, 32:] = sib, acc[q]
            idx[q] >>= 1
        acc = ctx.merge(pairs)
    return acc
