"""Child views the aggregation layer builds from a step proof (data model).

Reference: zk-lisp-proof-winterfell/src/agg/child.rs — ZlChildCompact :25-80 and `from_step` :162-526, ZlChildTranscript :528-575
and `from_step` :600-849 (FRI final degree :616-633, openings :635-720, OOD rows), `verify_child_transcript` :905-1022.
Everything is derived from the `ZKLSTP1` bytes; the hashing (transcript replay, leaf digests, authentication paths) goes through
the device hasher (`replay.py`), the rest is byte slicing.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import replay as R
from .blake3 import blake3
from .step import InvalidInput, StepMeta, StepProof


@dataclass
class FriFinal:
    """agg/child.rs:577-583 ZlFriFinal: index of the highest non-zero coefficient, coefficients as stored in the proof."""
    deg: int
    coeffs: List[int]


@dataclass
class ChildCompact:
    """agg/child.rs:25-80."""
    suite_id: bytes
    meta: StepMeta
    pi_core: object
    rom_acc: Tuple[int, int, int]
    step_digest: bytes
    trace_root: bytes               # the zl1 `root_trace` (blake3 over all inner commitments), NOT the trace Merkle root
    segment_index: int
    segments_total: int
    state_in_hash: bytes
    state_out_hash: bytes
    ram_gp_unsorted_in: bytes
    ram_gp_unsorted_out: bytes
    ram_gp_sorted_in: bytes
    ram_gp_sorted_out: bytes
    rom_s_in: Tuple[bytes, bytes, bytes]
    rom_s_out: Tuple[bytes, bytes, bytes]
    trace_roots: List[bytes]
    constraint_root: bytes
    fri_roots: List[bytes]
    pow_nonce: int
    fs_challenges: Optional[R.FsChallenges] = None
    merkle_proofs: Optional[R.MerkleProofs] = None


@dataclass
class ChildTranscript:
    """agg/child.rs:528-575."""
    compact: ChildCompact
    trace_roots: List[bytes]
    constraint_root: bytes
    fri_roots: List[bytes]
    fri_final: FriFinal
    num_queries: int
    trace_main_openings: List[List[int]]
    constraint_openings: List[List[int]]
    fri_layers: List[List[Tuple[int, int]]]
    ood_main_current: List[int]
    ood_main_next: List[int]
    ood_quotient_current: List[int]
    ood_quotient_next: List[int]
    main_trace_width: int
    constraint_frame_width: int
    ood_aux_current: List[int] = field(default_factory=list)
    ood_aux_next: List[int] = field(default_factory=list)
    aux_trace_width: int = 0


def child_compact_from_step(hasher, step: StepProof) -> ChildCompact:
    """agg/child.rs:162-526: commitments, replayed Fiat-Shamir challenges and per-query Merkle data (absent when the proof
    opened no position)."""
    v, pi = step.proof.inner, step.proof.pi
    try:
        trace_roots, constraint_root, fri_roots = v.parse_commitments()
    except InvalidInput:
        raise InvalidInput("invalid Winterfell commitments layout in ZlChildCompact")
    fs = mp = None
    if v.num_unique_queries:
        fs = R.replay_fs_from_step(hasher, step)
        if len(fs.query_positions) != v.num_unique_queries:
            raise InvalidInput("ZlChildCompact.fs_challenges.query_positions length is inconsistent with num_queries")
        mp = R.child_merkle_proofs(hasher, step, fs)
    return ChildCompact(step.proof.header.suite_id, step.proof.meta, step.pi_core, tuple(step.rom_acc), step.digest(), step.proof.root_trace,
                        pi.segment_index, pi.segments_total, pi.state_in_hash, pi.state_out_hash, pi.ram_gp_unsorted_in,
                        pi.ram_gp_unsorted_out, pi.ram_gp_sorted_in, pi.ram_gp_sorted_out, (pi.rom_s_in_0, pi.rom_s_in_1, pi.rom_s_in_2),
                        (pi.rom_s_out_0, pi.rom_s_out_1, pi.rom_s_out_2), trace_roots, constraint_root, fri_roots, v.pow_nonce, fs, mp)


def child_transcript_from_step(hasher, step: StepProof, compact: Optional[ChildCompact] = None) -> ChildTranscript:
    """agg/child.rs:600-849."""
    compact = compact or child_compact_from_step(hasher, step)
    v = step.proof.inner
    coeffs = v.fri_remainder_coeffs()
    deg = max((i for i, c in enumerate(coeffs) if c), default=0)            # :622-628
    W = v.main_width
    fe = lambda b, i: int.from_bytes(b[16 * i:16 * i + 16], "little")
    C = len(v.ood_constraints) // 32
    if len(v.ood_trace) != 32 * W or len(v.ood_constraints) != 32 * C:
        raise InvalidInput("invalid OOD frame layout in ZlChildTranscript")
    q = v.num_unique_queries
    trace_rows = v.trace_rows() if q else []
    cons_rows = v.constraint_rows() if q else []
    layers = []
    for values, _ in v.fri_layers:
        if len(values) % 32:
            raise InvalidInput("invalid FRI layer layout in ZlChildTranscript")
        layers.append([(fe(values, 2 * i), fe(values, 2 * i + 1)) for i in range(len(values) // 32)])
    return ChildTranscript(compact, list(compact.trace_roots), compact.constraint_root, list(compact.fri_roots), FriFinal(deg, coeffs), q,
                           trace_rows, cons_rows, layers if q else [], [fe(v.ood_trace, c) for c in range(W)],
                           [fe(v.ood_trace, W + c) for c in range(W)], [fe(v.ood_constraints, c) for c in range(C)],
                           [fe(v.ood_constraints, C + c) for c in range(C)], W, len(cons_rows[0]) if cons_rows else 0)


def verify_child_transcript(tx: ChildTranscript) -> None:
    """agg/child.rs:905-1022: the commitments recombine into the compact `trace_root`; openings and FRI data are shape-consistent."""
    root = blake3(b"zkl/step/root_trace" + tx.compact.suite_id + b"".join(tx.trace_roots) + tx.constraint_root + b"".join(tx.fri_roots))
    if root != tx.compact.trace_root:
        raise InvalidInput("ZlChildTranscript commitments do not match compact trace_root")
    q = tx.num_queries
    if q == 0:
        if tx.trace_main_openings:
            raise InvalidInput("ZlChildTranscript must not contain trace openings when num_queries is zero")
        if tx.constraint_openings:
            raise InvalidInput("ZlChildTranscript must not contain constraint openings when num_queries is zero")
        if tx.fri_layers:
            raise InvalidInput("ZlChildTranscript must not contain FRI layers when num_queries is zero")
    else:
        if len(tx.trace_main_openings) != q:
            raise InvalidInput("ZlChildTranscript.trace_main_openings length is inconsistent with num_queries")
        if any(not row for row in tx.trace_main_openings):
            raise InvalidInput("ZlChildTranscript contains an empty trace opening row")
        if len(tx.constraint_openings) != q:
            raise InvalidInput("ZlChildTranscript.constraint_openings length is inconsistent with num_queries")
        if tx.constraint_frame_width == 0:
            raise InvalidInput("ZlChildTranscript.constraint_frame_width must be non-zero when num_queries is non-zero")
        for row in tx.constraint_openings:
            if not row:
                raise InvalidInput("ZlChildTranscript contains an empty constraint opening row")
            if len(row) != tx.constraint_frame_width:
                raise InvalidInput("ZlChildTranscript.constraint_openings row width is inconsistent with constraint_frame_width")
        if any(not layer for layer in tx.fri_layers):
            raise InvalidInput("ZlChildTranscript contains an empty FRI layer")
    if not tx.fri_final.coeffs:
        raise InvalidInput("ZlChildTranscript.fri_final.coeffs must be non-empty")
    if tx.fri_final.deg > len(tx.fri_final.coeffs) - 1:
        raise InvalidInput("ZlChildTranscript.fri_final.deg exceeds coeffs.len() - 1")
