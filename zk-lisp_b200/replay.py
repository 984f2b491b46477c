"""Child-proof replay for the aggregation layer: Fiat-Shamir challenges re-derived from a step
proof, and every Merkle opening of the proof re-hashed against its commitment — with the hashing done as batches on
the device through the C-ABI primitives instead of one CPU sponge call at a time.

Reference: zk-lisp-proof-winterfell/src/agg/fs.rs:37-245 (replay_fs_from_step, mirrors winter_verifier's transcript),
src/agg/child.rs:162-526 (ZlChildCompact::from_step: roots, FS challenges, Merkle data), :1025-1045 (row leaf hashing
with partitions), :1049-1068 (path folding), :1072-1099 (FRI position folding); winter-crypto DefaultRandomCoin and
BatchMerkleProof::get_root [WF-mem].

`hasher` is anything with the batch methods of `prover.Context`: hash_elements((k, len, 2) u64) -> (k, 32) u8,
merge((k, 64) u8) -> (k, 32) u8, merge_with_int(seed32, values) -> (k, 32) u8, merge_many((k, 32) u8) -> 32 bytes.
"""
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import air as A
from .field import P
from .step import InnerProof, InvalidInput, StepProof, _Reader, fe_from_bytes_fold


def _ints_to_fe(xs) -> np.ndarray:
    out = np.empty((len(xs), 2), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i, 0], out[i, 1] = x & 0xFFFFFFFFFFFFFFFF, x >> 64
    return out


def _digest_fe(d: bytes) -> int:
    """draw::<BaseElement>: the first 16 digest bytes; always canonical for this hasher (utils.rs:375-381)."""
    v = int.from_bytes(d[:16], "little")
    if v >= P:
        raise InvalidInput("failed to draw a field element in FS replay")
    return v


class Coin:
    """winter-crypto DefaultRandomCoin over the zk-lisp hasher [WF-mem]."""

    def __init__(self, hasher, seed_elems: Sequence[int]):
        self.h = hasher
        self.seed = self.h.hash_elements(_ints_to_fe(seed_elems).reshape(1, -1, 2))[0].tobytes()
        self.counter = 0

    def reseed(self, digest: bytes):
        self.seed = self.h.merge(np.frombuffer(self.seed + bytes(digest), dtype=np.uint8).reshape(1, 64))[0].tobytes()
        self.counter = 0

    def draw_many(self, k: int) -> List[int]:
        """k consecutive `draw()` calls as ONE batch of merge_with_int(seed, counter + 1 .. counter + k)."""
        if k == 0:
            return []
        ctrs = np.arange(self.counter + 1, self.counter + 1 + k, dtype=np.uint64)
        self.counter += k
        return [_digest_fe(d.tobytes()) for d in self.h.merge_with_int(self.seed, ctrs)]

    def draw(self) -> int:
        return self.draw_many(1)[0]

    def check_leading_zeros(self, value: int) -> int:
        d = self.h.merge_with_int(self.seed, np.array([value], dtype=np.uint64))[0].tobytes()
        head = int.from_bytes(d[:8], "little")
        return 64 if head == 0 else (head & -head).bit_length() - 1

    def draw_integers(self, k: int, domain: int, nonce: int) -> List[int]:
        self.seed = self.h.merge_with_int(self.seed, np.array([nonce], dtype=np.uint64))[0].tobytes()
        self.counter = k
        out = self.h.merge_with_int(self.seed, np.arange(1, k + 1, dtype=np.uint64))
        return [int.from_bytes(d.tobytes()[:8], "little") & (domain - 1) for d in out]


@dataclass
class FsChallenges:
    """agg/child.rs ZlFsChallenges."""
    z_ood: int
    deep_coeffs: List[int]
    fri_alphas: List[int]
    query_positions: List[int]
    pow_nonce: int


def air_public_inputs_of(step: StepProof) -> A.AirPublicInputs:
    """agg/fs.rs:43-66."""
    pi = step.proof.pi
    f = fe_from_bytes_fold
    return A.AirPublicInputs(core=step.pi_core, segment_feature_mask=0, rom_acc=tuple(step.rom_acc), pc_init=f(pi.pc_init),
                             ram_gp_unsorted_in=f(pi.ram_gp_unsorted_in), ram_gp_unsorted_out=f(pi.ram_gp_unsorted_out),
                             ram_gp_sorted_in=f(pi.ram_gp_sorted_in), ram_gp_sorted_out=f(pi.ram_gp_sorted_out),
                             rom_s_in=(f(pi.rom_s_in_0), f(pi.rom_s_in_1), f(pi.rom_s_in_2)),
                             rom_s_out=(f(pi.rom_s_out_0), f(pi.rom_s_out_1), f(pi.rom_s_out_2)),
                             vm_usage_mask=step.vm_usage_mask, ram_delta_clk_bits=step.ram_delta_clk_bits)


def _options_of(v: InnerProof) -> A.ProofOptions:
    return A.ProofOptions(num_queries=v.num_queries, blowup=v.blowup, grinding=v.grinding, num_partitions=v.num_partitions,
                          hash_rate=v.hash_rate, fri_folding=v.fri_folding, fri_remainder_max_degree=v.fri_remainder_max_degree,
                          field_extension=v.field_extension)


def replay_fs_from_step(hasher, step: StepProof, capture: dict = None) -> FsChallenges:
    """agg/fs.rs:37-245. Everything comes from the step-proof bytes: the context elements from the proof's own context,
    the number of composition columns from the width of the opened constraint rows. `capture`, when given, receives the coin
    seed the prover ground against ("grinding_seed")."""
    v = step.proof.inner
    if v.aux_width != 0:
        raise InvalidInput("multi-segment trace must provide at least two trace commitments")
    seed = A.context_to_elements(v.main_width, v.trace_len, v.num_constraints, _options_of(v)) + air_public_inputs_of(step).to_elements()
    coin = Coin(hasher, seed)
    try:
        trace_roots, constraint_root, fri_roots = v.parse_commitments()
    except InvalidInput:
        raise InvalidInput("invalid Winterfell commitments layout in FS replay")
    # 1) main trace commitment (the composition coefficients drawn here do not survive the next reseed)
    coin.reseed(trace_roots[0])
    # 2) constraint commitment, OOD point
    coin.reseed(constraint_root)
    z = coin.draw()
    # 3) OOD frame: current(z) || quotient(z) || next(zg) || quotient(zg), hashed and absorbed
    W = v.main_width
    if len(v.ood_trace) != 2 * 16 * W or len(v.ood_constraints) % 32 or not v.ood_constraints:
        raise InvalidInput("invalid OOD frame layout in FS replay")
    C = len(v.ood_constraints) // 32
    ood = v.ood_trace[:16 * W] + v.ood_constraints[:16 * C] + v.ood_trace[16 * W:] + v.ood_constraints[16 * C:]
    elems = np.frombuffer(ood, dtype=np.uint64).reshape(1, 2 * (W + C), 2)
    coin.reseed(hasher.hash_elements(elems)[0].tobytes())
    # 4) DEEP coefficients: one per trace column, then one per composition column
    deep = coin.draw_many(W + C)
    # 5) FRI layer alphas (the last root is the remainder commitment)
    alphas, max_degree_plus_1 = [], v.trace_len
    for depth, root in enumerate(fri_roots):
        coin.reseed(root)
        alphas.append(coin.draw())
        if depth != len(fri_roots) - 1 and max_degree_plus_1 % v.fri_folding:
            raise InvalidInput("FRI degree truncation detected while replaying FS challenges")
        max_degree_plus_1 //= v.fri_folding
    # 6) proof of work, query positions
    if capture is not None:
        capture["grinding_seed"] = coin.seed
    if coin.check_leading_zeros(v.pow_nonce) < v.grinding:
        raise InvalidInput("pow_nonce does not satisfy grinding factor in FS replay")
    positions = sorted(set(coin.draw_integers(v.num_queries, v.lde_domain_size, v.pow_nonce)))
    if len(positions) != v.num_unique_queries:
        raise InvalidInput("unique query positions count does not match proof.num_unique_queries")
    return FsChallenges(z, deep, alphas, positions, v.pow_nonce)


class CoinBatch:
    """`Coin` for k transcripts advanced in lockstep: every transcript step is ONE device batch over all children
    (`merge` for the reseeds, `merge_with_int_pairs` — one seed per value — for the draws)."""

    def __init__(self, hasher, seed_elems: Sequence[Sequence[int]]):
        self.h, self.k = hasher, len(seed_elems)
        by_len = {}
        for i, e in enumerate(seed_elems):
            by_len.setdefault(len(e), []).append(i)
        self.seeds: List[bytes] = [b""] * self.k
        for ln, members in by_len.items():
            batch = np.stack([_ints_to_fe(seed_elems[i]) for i in members])
            for i, d in zip(members, self.h.hash_elements(batch)):
                self.seeds[i] = d.tobytes()

    def reseed(self, digests: Sequence[bytes]):
        pairs = np.frombuffer(b"".join(s + bytes(d) for s, d in zip(self.seeds, digests)), dtype=np.uint8).reshape(self.k, 64)
        self.seeds = [d.tobytes() for d in self.h.merge(pairs)]

    def _mwi(self, values_per_child: Sequence[Sequence[int]]) -> List[List[bytes]]:
        seeds = b"".join(self.seeds[i] * len(v) for i, v in enumerate(values_per_child))
        flat = np.array([x for v in values_per_child for x in v], dtype=np.uint64)
        if flat.size == 0:
            return [[] for _ in values_per_child]
        out = self.h.merge_with_int_pairs(np.frombuffer(seeds, dtype=np.uint8).reshape(-1, 32), flat)
        res, at = [], 0
        for v in values_per_child:
            res.append([out[at + j].tobytes() for j in range(len(v))])
            at += len(v)
        return res

    def draw_many(self, counts: Sequence[int]) -> List[List[int]]:
        """counts[i] consecutive draws of transcript i (counter restarts at 1 after every reseed)."""
        return [[_digest_fe(d) for d in ds] for ds in self._mwi([range(1, c + 1) for c in counts])]

    def check_leading_zeros(self, values: Sequence[int]) -> List[int]:
        heads = [int.from_bytes(ds[0][:8], "little") for ds in self._mwi([[v] for v in values])]
        return [64 if h == 0 else (h & -h).bit_length() - 1 for h in heads]

    def draw_integers(self, counts: Sequence[int], domains: Sequence[int], nonces: Sequence[int]) -> List[List[int]]:
        self.seeds = [ds[0] for ds in self._mwi([[n] for n in nonces])]
        return [[int.from_bytes(d[:8], "little") & (dom - 1) for d in ds]
                for ds, dom in zip(self._mwi([range(1, c + 1) for c in counts]), domains)]


def replay_fs_from_steps(hasher, steps: Sequence[StepProof]) -> List[FsChallenges]:
    """`replay_fs_from_step` for a batch of children in lockstep (agg/fs.rs:37-245 per child): about 2 x (FRI layers + 4) device
    submissions for ALL children instead of that many per child. Children may differ in width and trace length; transcripts with
    fewer FRI layers simply sit out the later reseeds."""
    k = len(steps)
    if k == 0:
        return []
    views = [s.proof.inner for s in steps]
    for v in views:
        if v.aux_width != 0:
            raise InvalidInput("multi-segment trace must provide at least two trace commitments")
        if len(v.ood_trace) != 2 * 16 * v.main_width or len(v.ood_constraints) % 32 or not v.ood_constraints:
            raise InvalidInput("invalid OOD frame layout in FS replay")
    seeds = [A.context_to_elements(v.main_width, v.trace_len, v.num_constraints, _options_of(v)) + air_public_inputs_of(s).to_elements()
             for s, v in zip(steps, views)]
    coin = CoinBatch(hasher, seeds)
    try:
        roots = [v.parse_commitments() for v in views]
    except InvalidInput:
        raise InvalidInput("invalid Winterfell commitments layout in FS replay")
    coin.reseed([r[0][0] for r in roots])
    coin.reseed([r[1] for r in roots])
    z = [d[0] for d in coin.draw_many([1] * k)]
    # OOD frames: equal shapes hash as one batch
    ood_digest, by_shape = [b""] * k, {}
    for i, v in enumerate(views):
        W, C = v.main_width, len(v.ood_constraints) // 32
        ood = v.ood_trace[:16 * W] + v.ood_constraints[:16 * C] + v.ood_trace[16 * W:] + v.ood_constraints[16 * C:]
        by_shape.setdefault(2 * (W + C), []).append((i, np.frombuffer(ood, dtype=np.uint64).reshape(2 * (W + C), 2)))
    for _, members in by_shape.items():
        for (i, _), d in zip(members, hasher.hash_elements(np.stack([m[1] for m in members]))):
            ood_digest[i] = d.tobytes()
    coin.reseed(ood_digest)
    deep = coin.draw_many([v.main_width + len(v.ood_constraints) // 32 for v in views])
    # FRI alphas: transcript i takes part in round d while it still has a root
    alphas = [[] for _ in range(k)]
    degree = [v.trace_len for v in views]
    for d in range(max(len(r[2]) for r in roots)):
        live = [i for i in range(k) if d < len(roots[i][2])]
        sub = CoinBatch.__new__(CoinBatch)
        sub.h, sub.k, sub.seeds = hasher, len(live), [coin.seeds[i] for i in live]
        sub.reseed([roots[i][2][d] for i in live])
        for i, a, sd in zip(live, sub.draw_many([1] * len(live)), sub.seeds):
            coin.seeds[i] = sd
            alphas[i].append(a[0])
            if d != len(roots[i][2]) - 1 and degree[i] % views[i].fri_folding:
                raise InvalidInput("FRI degree truncation detected while replaying FS challenges")
            degree[i] //= views[i].fri_folding
    zeros = coin.check_leading_zeros([v.pow_nonce for v in views])
    if any(zz < v.grinding for zz, v in zip(zeros, views)):
        raise InvalidInput("pow_nonce does not satisfy grinding factor in FS replay")
    drawn = coin.draw_integers([v.num_queries for v in views], [v.lde_domain_size for v in views], [v.pow_nonce for v in views])
    out = []
    for i, v in enumerate(views):
        positions = sorted(set(drawn[i]))
        if len(positions) != v.num_unique_queries:
            raise InvalidInput("unique query positions count does not match proof.num_unique_queries")
        out.append(FsChallenges(z[i], deep[i], alphas[i], positions, v.pow_nonce))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Merkle openings
# ---------------------------------------------------------------------------------------------------------------
def read_batch_proof(data: bytes):
    """BatchMerkleProof::read_from [WF-mem]: depth u8, number of lists, then per list its length and digests."""
    r = _Reader(bytes(data))
    depth = r.u8()
    lists = [[r.take(32) for _ in range(r.usize())] for _ in range(r.usize())]
    if r.i != len(r.b):
        raise InvalidInput("trailing bytes in batch Merkle proof")
    return depth, lists


def fold_positions(positions: Sequence[int], source_domain: int, folding: int = 2) -> List[int]:
    """agg/child.rs:1072-1099: positions mod the folded domain, first occurrence order, no duplicates."""
    target = source_domain // folding
    out = []
    for p in positions:
        q = p % target
        if q not in out:
            out.append(q)
    return out


def partition_size(num_partitions: int, hash_rate: int, cols: int) -> int:
    """PartitionOptions::partition_size [WF-mem] (csrc: partition_size_of)."""
    if num_partitions == 1 and hash_rate == 0:
        return cols
    return max(-(-cols // num_partitions), hash_rate)


def hash_rows(hasher, rows: np.ndarray, psize: int) -> List[bytes]:
    """agg/child.rs:1025-1045 for a batch of rows (k, W, 2): one hash_elements batch per partition, then merge_many."""
    k, W = rows.shape[0], rows.shape[1]
    digests = [hasher.hash_elements(np.ascontiguousarray(rows[:, i:i + psize])) for i in range(0, W, psize)]
    if len(digests) == 1:
        return [d.tobytes() for d in digests[0]]
    return [hasher.merge_many(np.stack([d[q] for d in digests])) for q in range(k)]


def _walk_batch_opening(hasher, depth: int, lists: List[List[bytes]], leaves: Sequence[bytes], positions: Sequence[int]):
    """BatchMerkleProof::get_root [WF-mem] with every digest it touches kept under its heap index (leaf i = 2^depth + i, root = 1);
    the merges of one tree level are ONE batch."""
    if depth < 1 or len(leaves) != len(positions) or not positions:
        raise InvalidInput("malformed batch Merkle opening")
    n = 1 << depth
    known = {p: bytes(l) for p, l in zip(positions, leaves)}
    norm = sorted({p & ~1 for p in positions})
    if len(norm) != len(lists) or max(positions) >= n:
        raise InvalidInput("malformed batch Merkle opening")
    ptr = [0] * len(lists)
    nodes = {}

    def take(s):
        if ptr[s] >= len(lists[s]):
            raise InvalidInput("malformed batch Merkle opening")
        ptr[s] += 1
        return lists[s][ptr[s] - 1]

    def merge(pairs):
        return [d.tobytes() for d in hasher.merge(np.frombuffer(b"".join(pairs), dtype=np.uint8).reshape(-1, 64))]

    pairs = []
    for i, e in enumerate(norm):
        lo = known[e] if e in known else take(i)
        hi = known[e + 1] if e + 1 in known else take(i)
        nodes[n + e], nodes[n + e + 1] = lo, hi
        pairs.append(lo + hi)
    cur, vals = [(e + n) >> 1 for e in norm], merge(pairs)
    for _ in range(1, depth):
        ncur, pairs, i = [], [], 0
        while i < len(cur):
            me, sib, used = vals[i], cur[i] ^ 1, i
            if i + 1 < len(cur) and cur[i + 1] == sib:
                sv = vals[i + 1]
                i += 1
            else:
                sv = take(used)
            nodes[cur[used]], nodes[sib] = me, sv
            pairs.append(sv + me if cur[used] & 1 else me + sv)
            ncur.append(sib >> 1)
            i += 1
        cur, vals = ncur, merge(pairs)
    if cur != [1] or any(ptr[s] != len(lists[s]) for s in range(len(lists))):
        raise InvalidInput("malformed batch Merkle opening")
    nodes[1] = vals[0]
    return nodes


def batch_root(hasher, depth: int, lists: List[List[bytes]], leaves: Sequence[bytes], positions: Sequence[int]) -> bytes:
    """The root implied by `leaves` opened at `positions`."""
    return _walk_batch_opening(hasher, depth, lists, leaves, positions)[1]


def opening_paths(hasher, depth: int, lists: List[List[bytes]], leaves: Sequence[bytes], positions: Sequence[int]):
    """BatchMerkleProof::into_openings [WF-mem]: (root, [siblings of position p, leaf level first]) — the individual authentication
    paths `ZlMerklePath` carries (agg/child.rs:84-88, extracted at :250-292, folded by `merkle_root_from_leaf` :1049-1068)."""
    nodes = _walk_batch_opening(hasher, depth, lists, leaves, positions)
    paths = []
    for p in positions:
        idx, sibs = (1 << depth) + p, []
        while idx > 1:
            sibs.append(nodes[idx ^ 1])
            idx >>= 1
        paths.append(sibs)
    return nodes[1], paths


@dataclass
class MerklePath:
    """agg/child.rs:84-88 ZlMerklePath."""
    leaf: bytes
    siblings: List[bytes]


@dataclass
class FriLayerProof:
    """agg/child.rs:92-97 ZlFriLayerProof: folded position, the opened pair, its path in the layer tree."""
    idx: int
    values: tuple
    path: MerklePath


@dataclass
class MerkleProofs:
    """agg/child.rs:101-110 ZlMerkleProofs."""
    trace_paths: List[MerklePath]
    constraint_paths: List[MerklePath]
    fri_layers: List[List[FriLayerProof]]


def merkle_root_from_leaf(hasher, leaf: bytes, index: int, siblings: Sequence[bytes]) -> bytes:
    """agg/child.rs:1049-1068."""
    acc = bytes(leaf)
    for s in siblings:
        pair = (s + acc) if index & 1 else (acc + s)
        acc = hasher.merge(np.frombuffer(pair, dtype=np.uint8).reshape(1, 64))[0].tobytes()
        index >>= 1
    return acc


def child_merkle_proofs(hasher, step: StepProof, fs: FsChallenges) -> MerkleProofs:
    """agg/child.rs:230-400: per-query authentication paths of the trace, constraint and FRI layer trees, each checked against
    its commitment while it is extracted."""
    v = step.proof.inner
    trace_roots, constraint_root, fri_roots = v.parse_commitments()
    pos = fs.query_positions
    k = len(pos)

    def row_paths(values, width, paths_bytes, expected, what):
        rows = np.frombuffer(values, dtype=np.uint64)
        if k == 0 or width == 0 or rows.size != k * width * 2:
            raise InvalidInput("invalid trace queries layout in ZlChildTranscript")
        leaves = hash_rows(hasher, rows.reshape(k, width, 2), partition_size(v.num_partitions, v.hash_rate, width))
        root, paths = opening_paths(hasher, *read_batch_proof(paths_bytes), leaves, pos)
        if root != expected:
            raise InvalidInput(what)
        return [MerklePath(l, p) for l, p in zip(leaves, paths)]

    trace = row_paths(v.trace_query_values, v.main_width, v.trace_query_paths, trace_roots[0], "trace openings do not match the trace commitment")
    cons = row_paths(v.constraint_query_values, len(v.constraint_query_values) // (16 * k) if k else 0, v.constraint_query_paths,
                     constraint_root, "constraint openings do not match the constraint commitment")
    layers, positions, domain = [], list(pos), v.lde_domain_size
    for d, (values, paths_bytes) in enumerate(v.fri_layers):
        positions = fold_positions(positions, domain, v.fri_folding)
        pairs = np.frombuffer(values, dtype=np.uint64)
        if pairs.size != len(positions) * 4:
            raise InvalidInput("FRI layer value count does not match the folded positions")
        pairs = pairs.reshape(len(positions), 2, 2)
        leaves = [x.tobytes() for x in hasher.hash_elements(pairs)]
        root, paths = opening_paths(hasher, *read_batch_proof(paths_bytes), leaves, positions)
        if root != fri_roots[d]:
            raise InvalidInput(f"FRI layer {d} openings do not match the layer commitment")
        ints = lambda e: int(e[0]) | (int(e[1]) << 64)
        layers.append([FriLayerProof(q, (ints(pr[0]), ints(pr[1])), MerklePath(l, p)) for q, pr, l, p in zip(positions, pairs, leaves, paths)])
        domain //= v.fri_folding
    return MerkleProofs(trace, cons, layers)


class MergePlan:
    """The merge schedule of any number of batch openings, laid out for `zkl_merge_plan`: a flat digest table (opened leaves
    and proof nodes first), per tree level the (left, right) table indices of every merge of every tree, and the table index
    of each tree's root. The schedule follows from the query positions alone, so it is built before any hashing happens."""

    def __init__(self):
        self.known: List[bytes] = []
        self.levels: List[List[tuple]] = []      # levels[l] = [(left, right), ...] with symbolic node ids
        self.roots: List[tuple] = []             # (symbolic id, expected root, message)

    def _known(self, digest: bytes):
        self.known.append(bytes(digest))
        return ("k", len(self.known) - 1)

    def _merge(self, level: int, left, right):
        while len(self.levels) <= level:
            self.levels.append([])
        self.levels[level].append((left, right))
        return ("m", level, len(self.levels[level]) - 1)

    def add_opening(self, depth: int, lists: List[List[bytes]], leaves: Sequence[bytes], positions: Sequence[int], expected: bytes, what: str):
        """Same walk as `batch_root`, recording merges instead of performing them."""
        if depth < 1 or len(leaves) != len(positions) or not positions:
            raise InvalidInput("malformed batch Merkle opening")
        n = 1 << depth
        norm = sorted({p & ~1 for p in positions})
        if len(norm) != len(lists) or max(positions) >= n:
            raise InvalidInput("malformed batch Merkle opening")
        known = {p: self._known(l) for p, l in zip(positions, leaves)}
        ptr = [0] * len(lists)

        def take(s):
            if ptr[s] >= len(lists[s]):
                raise InvalidInput("malformed batch Merkle opening")
            ptr[s] += 1
            return self._known(lists[s][ptr[s] - 1])

        cur, vals = [], []
        for i, e in enumerate(norm):
            lo = known[e] if e in known else take(i)
            hi = known[e + 1] if e + 1 in known else take(i)
            cur.append((e + n) >> 1)
            vals.append(self._merge(0, lo, hi))
        for lvl in range(1, depth):
            ncur, nvals, i = [], [], 0
            while i < len(cur):
                me, sib, used = vals[i], cur[i] ^ 1, i
                if i + 1 < len(cur) and cur[i + 1] == sib:
                    sv = vals[i + 1]
                    i += 1
                else:
                    sv = take(used)
                nvals.append(self._merge(lvl, sv, me) if cur[used] & 1 else self._merge(lvl, me, sv))
                ncur.append(sib >> 1)
                i += 1
            cur, vals = ncur, nvals
        if cur != [1] or any(ptr[s] != len(lists[s]) for s in range(len(lists))):
            raise InvalidInput("malformed batch Merkle opening")
        self.roots.append((vals[0], bytes(expected), what))

    def flatten(self):
        """(known (k, 32) u8, left u32[], right u32[], level_counts u32[], want u32[])."""
        base, acc = [], len(self.known)
        for lv in self.levels:
            base.append(acc)
            acc += len(lv)

        def index(node):
            return node[1] if node[0] == "k" else base[node[1]] + node[2]

        left = [index(a) for lv in self.levels for a, _ in lv]
        right = [index(b) for lv in self.levels for _, b in lv]
        known = np.frombuffer(b"".join(self.known), dtype=np.uint8).reshape(-1, 32)
        return (known, np.array(left, dtype=np.uint32), np.array(right, dtype=np.uint32),
                np.array([len(lv) for lv in self.levels], dtype=np.uint32), np.array([index(r) for r, _, _ in self.roots], dtype=np.uint32))

    def run(self, hasher) -> None:
        """One device submission for all trees; raises on the first root that does not match its commitment."""
        if not self.roots:
            return
        got = hasher.merge_plan(*self.flatten())
        for (_, expected, what), g in zip(self.roots, got):
            if g.tobytes() != expected:
                raise InvalidInput(what)


def verify_children_openings(hasher, steps: Sequence[StepProof], challenges: Sequence[FsChallenges]) -> None:
    """`verify_child_openings` for a whole batch of children in FOUR device submissions: the leaves of all opened trace rows
    (one hash_elements batch per partition), of all constraint rows, of all FRI layer pairs and remainders, then ONE merge plan
    that authenticates every tree of every child (the reference walks each path with one CPU sponge call per node,
    agg/child.rs:1049-1068)."""
    plan = MergePlan()
    trace_rows, cons_rows, fri_pairs, rems = [], [], [], []
    jobs = []          # (kind, child, rows-slice, ...) resolved after the batched leaf hashing
    for ci, (step, fs) in enumerate(zip(steps, challenges)):
        v = step.proof.inner
        k = len(fs.query_positions)
        rows = np.frombuffer(v.trace_query_values, dtype=np.uint64)
        if k == 0 or rows.size != k * v.main_width * 2:
            raise InvalidInput("invalid trace queries layout in ZlChildTranscript")
        crow = np.frombuffer(v.constraint_query_values, dtype=np.uint64)
        C = len(v.constraint_query_values) // (16 * k)
        if C == 0 or crow.size != k * C * 2:
            raise InvalidInput("failed to decode constraint query values in ZlChildTranscript")
        if len(v.fri_layers) != v.num_fri_layers:
            raise InvalidInput("FRI layer count does not match the proof options")
        trace_rows.append((ci, rows.reshape(k, v.main_width, 2)))
        cons_rows.append((ci, crow.reshape(k, C, 2)))
        domain, positions, layers = v.lde_domain_size, list(fs.query_positions), []
        for values, _ in v.fri_layers:
            positions = fold_positions(positions, domain, v.fri_folding)
            pairs = np.frombuffer(values, dtype=np.uint64)
            if v.fri_folding != 2 or pairs.size != len(positions) * 4:
                raise InvalidInput("FRI layer value count does not match the folded positions")
            layers.append((positions, pairs.reshape(len(positions), 2, 2)))
            domain //= v.fri_folding
        fri_pairs.append((ci, layers))
        rem = np.frombuffer(v.fri_remainder, dtype=np.uint64)
        if rem.size == 0 or rem.size % 2:
            raise InvalidInput("FRI remainder does not match its commitment")
        rems.append(rem.reshape(1, -1, 2))

    def batched(groups, key):
        """Hash row groups that share a shape (width, partition size) in one call; returns per-group digest lists."""
        out, by_shape = {}, {}
        for gi, (ci, rows) in enumerate(groups):
            by_shape.setdefault(key(ci, rows), []).append(gi)
        for (width, psize), members in by_shape.items():
            digests = hash_rows(hasher, np.concatenate([groups[g][1] for g in members]), psize)
            at = 0
            for g in members:
                cnt = groups[g][1].shape[0]
                out[g] = digests[at:at + cnt]
                at += cnt
        return [out[g] for g in range(len(groups))]

    def psize_of(ci, rows):
        v = steps[ci].proof.inner
        return rows.shape[1], partition_size(v.num_partitions, v.hash_rate, rows.shape[1])

    t_leaves, c_leaves = batched(trace_rows, psize_of), batched(cons_rows, psize_of)
    all_pairs = [pairs for _, layers in fri_pairs for _, pairs in layers]
    f_leaves = [d.tobytes() for d in hasher.hash_elements(np.concatenate(all_pairs))] if all_pairs else []
    at = 0
    for ci, (step, fs) in enumerate(zip(steps, challenges)):
        v = step.proof.inner
        trace_roots, constraint_root, fri_roots = v.parse_commitments()
        tag = f"child {ci}: "
        plan.add_opening(*read_batch_proof(v.trace_query_paths), t_leaves[ci], fs.query_positions, trace_roots[0],
                         tag + "trace openings do not match the trace commitment")
        plan.add_opening(*read_batch_proof(v.constraint_query_paths), c_leaves[ci], fs.query_positions, constraint_root,
                         tag + "constraint openings do not match the constraint commitment")
        for d, ((positions, pairs), (_, paths)) in enumerate(zip(fri_pairs[ci][1], v.fri_layers)):
            plan.add_opening(*read_batch_proof(paths), f_leaves[at:at + len(positions)], positions, fri_roots[d],
                             tag + f"FRI layer {d} openings do not match the layer commitment")
            at += len(positions)
    # remainders: equal lengths hash as one batch
    by_len = {}
    for ci, rem in enumerate(rems):
        by_len.setdefault(rem.shape[1], []).append(ci)
    for _, members in by_len.items():
        digests = hasher.hash_elements(np.concatenate([rems[ci] for ci in members]))
        for ci, d in zip(members, digests):
            if d.tobytes() != steps[ci].proof.inner.parse_commitments()[2][-1]:
                raise InvalidInput(f"child {ci}: FRI remainder does not match its commitment")
    plan.run(hasher)


def verify_child_deep_and_fri(step: StepProof, fs: FsChallenges) -> None:
    """The arithmetic half of the child check the aggregation trace replays (agg/trace.rs:1076-1193 DEEP composition at the
    query positions, :752-950 FRI folding and remainder): the values opened in FRI layer 0 must be the DEEP composition of
    the opened trace / constraint rows and the OOD frame, every layer must fold into the next with the replayed alphas, and
    the last layer must agree with the remainder polynomial. Host integers — 64 positions x (W + 7) products and 11 small
    folds per child; the hashing half is `verify_child(ren)_openings`."""
    from .field import GENERATOR, inv, root_of_unity
    v = step.proof.inner
    W, n, N = v.main_width, v.trace_len, v.lde_domain_size
    fe = lambda b, i: int.from_bytes(b[16 * i:16 * i + 16], "little")
    C = len(v.ood_constraints) // 32
    if len(v.ood_trace) != 32 * W or C == 0 or len(fs.deep_coeffs) != W + C:
        raise InvalidInput("invalid OOD frame layout in FS replay")
    t_z, t_zg = [fe(v.ood_trace, c) for c in range(W)], [fe(v.ood_trace, W + c) for c in range(W)]
    h_z, h_zg = [fe(v.ood_constraints, c) for c in range(C)], [fe(v.ood_constraints, C + c) for c in range(C)]
    dc, z = fs.deep_coeffs, fs.z_ood
    zg = z * root_of_unity(n.bit_length() - 1) % P
    a_z = (sum(d * t for d, t in zip(dc, t_z)) + sum(d * h for d, h in zip(dc[W:], h_z))) % P
    a_zg = (sum(d * t for d, t in zip(dc, t_zg)) + sum(d * h for d, h in zip(dc[W:], h_zg))) % P
    trows, crows = v.trace_rows(), v.constraint_rows()
    pos = fs.query_positions
    if len(trows) != len(pos) or len(crows) != len(pos) or len(crows[0]) != C:
        raise InvalidInput("invalid trace queries layout in ZlChildTranscript")
    w_n = root_of_unity(N.bit_length() - 1)
    evals = []
    for p, tr, cr in zip(pos, trows, crows):
        x = GENERATOR * pow(w_n, p, P) % P
        a = (sum(d * t for d, t in zip(dc, tr)) + sum(d * h for d, h in zip(dc[W:], cr))) % P
        evals.append(((a - a_z) * inv((x - z) % P) + (a - a_zg) * inv((x - zg) % P)) % P)
    if len(fs.fri_alphas) != len(v.fri_layers) + 1:
        raise InvalidInput("FRI layer count does not match the proof options")
    positions, domain, w_d = list(pos), N, w_n
    for d, (values, _) in enumerate(v.fri_layers):
        folded = fold_positions(positions, domain, 2)
        half = domain // 2
        if len(values) != 32 * len(folded):
            raise InvalidInput("FRI layer value count does not match the folded positions")
        vals = [fe(values, i) for i in range(2 * len(folded))]
        for p, e in zip(positions, evals):
            if vals[2 * folded.index(p % half) + p // half] != e:
                raise InvalidInput(f"FRI layer {d} values do not continue the previous layer" if d else
                                   "FRI layer 0 values are not the DEEP composition of the opened rows")
        alpha, nxt = fs.fri_alphas[d], []
        for i, q in enumerate(folded):      # agg/trace.rs:803-820
            x0 = GENERATOR * pow(w_d, q, P) % P
            x1 = P - x0
            v0, v1 = vals[2 * i], vals[2 * i + 1]
            nxt.append((v1 * (alpha - x0) - v0 * (alpha - x1)) % P * inv((x1 - x0) % P) % P)
        positions, evals, domain, w_d = folded, nxt, half, w_d * w_d % P
    rem = v.fri_remainder_coeffs()          # highest degree first (agg/trace.rs:941-945)
    if not rem or len(rem) > (n >> len(v.fri_layers)):
        raise InvalidInput("FRI remainder degree exceeds the bound")
    for p, e in zip(positions, evals):
        x, acc = GENERATOR * pow(w_d, p, P) % P, 0
        for c in rem:
            acc = (acc * x + c) % P
        if acc != e:
            raise InvalidInput("FRI remainder does not continue the last layer")


def verify_child(hasher, step: StepProof) -> FsChallenges:
    """Everything the aggregation layer re-checks of one child: transcript, every Merkle opening, DEEP and FRI arithmetic."""
    fs = replay_fs_from_step(hasher, step)
    verify_child_openings(hasher, step, fs)
    verify_child_deep_and_fri(step, fs)
    return fs


def replay_children_native(ctx, steps: Sequence[StepProof], verify_openings: bool = True) -> List[FsChallenges]:
    """`replay_fs_from_steps` + `verify_children_openings` in ONE C-ABI call (`zkl_replay_children`, csrc/api_replay.inl): proof
    parsing, transcript bookkeeping and plan building run in native code, the hashing on the device. `ctx` must be a GPU Context."""
    import ctypes as C
    from . import capi
    k = len(steps)
    if k == 0:
        return []
    kids = (capi.ChildProof * k)()
    keep = []
    for i, s in enumerate(steps):
        proof = np.frombuffer(s.proof.inner_bytes, dtype=np.uint8)
        pi = _ints_to_fe(air_public_inputs_of(s).to_elements())
        keep += [proof, pi]
        kids[i].proof, kids[i].proof_len = proof.ctypes.data, proof.shape[0]
        kids[i].pi_elems, kids[i].num_pi = pi.ctypes.data, pi.shape[0]
    out = (capi.ChildChallenges * k)()
    stride = 256 + 16
    deep = np.zeros((k, stride, 2), dtype=np.uint64)
    ctx._check(ctx.lib.zkl_replay_children(ctx.h, kids, k, int(verify_openings), out, deep.ctypes.data, stride))
    res = []
    for i in range(k):
        o = out[i]
        if o.status:
            layer = f" (layer {o.status >> 8})" if (o.status & 0xFF) == 12 else ""
            raise InvalidInput(f"child {i}: {ctx.lib.zkl_replay_status_name(o.status & 0xFF).decode()}{layer}")
        res.append(FsChallenges(o.z.int(), [int(lo) | (int(hi) << 64) for lo, hi in deep[i, :o.num_deep]],
                                [o.alphas[j].int() for j in range(o.num_alphas)], list(o.positions[:o.num_positions]), int(o.pow_nonce)))
    return res


def verify_child_openings(hasher, step: StepProof, fs: FsChallenges) -> None:
    """Every opening of the child proof re-hashed against its commitment: trace rows and constraint rows at the query
    positions, FRI layer pairs at the folded positions, the remainder against the last FRI commitment."""
    v = step.proof.inner
    trace_roots, constraint_root, fri_roots = v.parse_commitments()
    pos = fs.query_positions
    k = len(pos)
    rows = np.frombuffer(v.trace_query_values, dtype=np.uint64)
    if rows.size != k * v.main_width * 2:
        raise InvalidInput("invalid trace queries layout in ZlChildTranscript")
    leaves = hash_rows(hasher, rows.reshape(k, v.main_width, 2), partition_size(v.num_partitions, v.hash_rate, v.main_width))
    if batch_root(hasher, *read_batch_proof(v.trace_query_paths), leaves, pos) != trace_roots[0]:
        raise InvalidInput("trace openings do not match the trace commitment")
    C = len(v.constraint_query_values) // (16 * k)
    crow = np.frombuffer(v.constraint_query_values, dtype=np.uint64)
    if C == 0 or crow.size != k * C * 2:
        raise InvalidInput("failed to decode constraint query values in ZlChildTranscript")
    leaves = hash_rows(hasher, crow.reshape(k, C, 2), partition_size(v.num_partitions, v.hash_rate, C))
    if batch_root(hasher, *read_batch_proof(v.constraint_query_paths), leaves, pos) != constraint_root:
        raise InvalidInput("constraint openings do not match the constraint commitment")
    if len(v.fri_layers) != v.num_fri_layers:
        raise InvalidInput("FRI layer count does not match the proof options")
    domain, positions = v.lde_domain_size, list(pos)
    for d, (values, paths) in enumerate(v.fri_layers):
        positions = fold_positions(positions, domain, v.fri_folding)
        pairs = np.frombuffer(values, dtype=np.uint64)
        if pairs.size != len(positions) * v.fri_folding * 2:
            raise InvalidInput("FRI layer value count does not match the folded positions")
        leaves = [x.tobytes() for x in hasher.hash_elements(pairs.reshape(len(positions), v.fri_folding, 2))]
        if batch_root(hasher, *read_batch_proof(paths), leaves, positions) != fri_roots[d]:
            raise InvalidInput(f"FRI layer {d} openings do not match the layer commitment")
        domain //= v.fri_folding
    rem = np.frombuffer(v.fri_remainder, dtype=np.uint64)
    if rem.size == 0 or rem.size % 2 or hasher.hash_elements(rem.reshape(1, -1, 2))[0].tobytes() != fri_roots[-1]:
        raise InvalidInput("FRI remainder does not match its commitment")
