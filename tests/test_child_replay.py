"""CPU tests of the child-proof replay logic (zk-lisp_b200/replay.py) with the CPU checker standing in
for the device hasher: the Fiat-Shamir replay written after agg/fs.rs must land on the challenges the prover used
(tests/fs_replay.rs:109-271 does this against a live coin), and every opening must re-hash to its commitment
(tests/agg_child.rs:60-146). The GPU suite runs the same functions over the C-ABI primitives (test_gpu_segments.py)."""
from dataclasses import replace

import numpy as np
import pytest

import synth
from util import single_segment_step


def _step(zkl, oracle, n, seed, queries=8, parts=None, grind=4):
    trace, air_pi, opts = synth.valid_segment(n=n, queries=queries, blowup=16, grind=grind, seed=seed)
    if parts:
        opts = replace(opts, num_partitions=parts[0], hash_rate=parts[1])
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    summ = zkl.capi.ProofSummary()
    proof = oracle.prove_segment(trace, a.to_c(), zkl.prover.ints_to_fe(a.seed_elements()), opts.to_c(), summ)
    return single_segment_step(zkl, air_pi, opts, n, proof), summ, a


@pytest.fixture(scope="module")
def child(zkl, oracle):
    return _step(zkl, oracle, 64, 7)


def test_fs_replay_lands_on_the_prover_challenges(zkl, oracle, child):
    R = zkl.replay
    step, summ, a = child
    # the replay goes through the ZKLSTP1 bytes, like an aggregator that received the proof
    step = zkl.step.StepProof.from_bytes(step.to_bytes())
    assert R.air_public_inputs_of(step).to_elements() == a.air_pi.to_elements()
    fs = R.replay_fs_from_step(oracle, step)
    assert fs.z_ood == int.from_bytes(bytes(summ.z), "little")
    assert fs.pow_nonce == summ.pow_nonce
    assert fs.query_positions == list(summ.positions[:summ.num_unique_queries])
    v = step.proof.inner
    assert len(fs.deep_coeffs) == v.main_width + len(v.constraint_rows()[0])
    assert len(fs.fri_alphas) == v.num_fri_layers + 1
    assert len(set(fs.deep_coeffs + fs.fri_alphas + [fs.z_ood])) == len(fs.deep_coeffs) + len(fs.fri_alphas) + 1
    R.verify_child_openings(oracle, step, fs)


def test_replay_with_partitions_and_more_queries(zkl, oracle):
    """Partitioned row leaves (merge_many of partition digests, agg/child.rs:1031-1044) and duplicate query draws."""
    R = zkl.replay
    step, summ, _ = _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)
    fs = R.replay_fs_from_step(oracle, step)
    assert fs.query_positions == list(summ.positions[:summ.num_unique_queries])
    R.verify_child_openings(oracle, step, fs)


def test_replay_rejects_tampered_proofs(zkl, oracle, child):
    R, S = zkl.replay, zkl.step
    step, _, _ = child
    v = step.proof.inner
    fs = R.replay_fs_from_step(oracle, step)

    def with_inner(**kw):
        return replace(step, proof=replace(step.proof, inner=replace(v, **kw)))

    # a different nonce fails the grinding check or moves the query positions
    with pytest.raises(S.InvalidInput):
        bad = with_inner(pow_nonce=v.pow_nonce + 1)
        fs2 = R.replay_fs_from_step(oracle, bad)
        R.verify_child_openings(oracle, bad, fs2)
    # public inputs are part of the coin seed
    moved = replace(step, vm_usage_mask=step.vm_usage_mask ^ 1)
    with pytest.raises(S.InvalidInput):
        R.verify_child_openings(oracle, moved, R.replay_fs_from_step(oracle, moved))
    # one flipped byte in an opened trace row, a constraint row, a FRI value, a Merkle node, the remainder
    def flip(b, i):
        return b[:i] + bytes([b[i] ^ 1]) + b[i + 1:]
    layers = list(v.fri_layers)
    layers[1] = (flip(layers[1][0], 5), layers[1][1])
    cases = [("trace openings", dict(trace_query_values=flip(v.trace_query_values, 100))),
             ("constraint openings", dict(constraint_query_values=flip(v.constraint_query_values, 3))),
             ("FRI layer 1", dict(fri_layers=layers)),
             ("trace openings", dict(trace_query_paths=flip(v.trace_query_paths, len(v.trace_query_paths) - 20))),
             ("remainder", dict(fri_remainder=flip(v.fri_remainder, 0)))]
    for msg, kw in cases:
        with pytest.raises(S.InvalidInput, match=msg):
            R.verify_child_openings(oracle, with_inner(**kw), fs)
    with pytest.raises(S.InvalidInput):
        R.verify_child_openings(oracle, step, replace(fs, query_positions=fs.query_positions[:-1]))
    # commitments are absorbed by the coin: another constraint root gives another OOD point
    c = bytearray(v.commitments)
    c[40] ^= 1
    try:
        assert R.replay_fs_from_step(oracle, with_inner(commitments=bytes(c), grinding=0)).z_ood != fs.z_ood
    except S.InvalidInput:
        pass   # the moved query positions no longer give the recorded number of unique queries


def test_batch_root_and_fold_positions(zkl, oracle):
    """BatchMerkleProof::get_root against the checker's own implementation on the same openings."""
    R = zkl.replay
    rng = np.random.default_rng(5)
    leaves = rng.integers(0, 256, size=(64, 32), dtype=np.uint8)
    leaves[:, 16:] = 0
    root, nodes = oracle.merkle_build(leaves)
    for pos in ([5], [0, 1], [2, 3, 9, 63], [7, 6, 40], list(range(0, 64, 3))):
        pos = sorted(pos)
        # authentication lists built the way MerkleTree::prove_batch lays them out
        norm = sorted({p & ~1 for p in pos})
        lists, cur = [], []
        for e in norm:
            lists.append([leaves[i].tobytes() for i in (e, e + 1) if i not in pos])
            cur.append((e + 64) >> 1)
        for _ in range(1, 6):
            nxt, i = [], 0
            while i < len(cur):
                sib = cur[i] ^ 1
                if i + 1 < len(cur) and cur[i + 1] == sib:
                    i += 1
                else:
                    lists[i].append(nodes[sib].tobytes())
                nxt.append(sib >> 1)
                i += 1
            cur = nxt
        got = R.batch_root(oracle, 6, lists, [leaves[p].tobytes() for p in pos], pos)
        assert got == root
        assert got == oracle.batch_root({"depth": 6, "nodes": lists}, np.stack([leaves[p] for p in pos]), pos)
        lists[0] = lists[0][:-1]
        with pytest.raises(zkl.step.InvalidInput):
            R.batch_root(oracle, 6, lists, [leaves[p].tobytes() for p in pos], pos)
    assert R.fold_positions([3, 70, 67, 5, 64 + 5], 128) == [3, 6, 5]
    assert R.partition_size(1, 0, 204) == 204 and R.partition_size(1, 16, 7) == 16 and R.partition_size(16, 16, 219) == 16
    assert R.partition_size(4, 8, 204) == 51


class _PlanHasher:
    """The checker's batch methods + `merge_plan` restated with numpy (the contract of zkl_merge_plan in include/zkl_b200.h)."""

    def __init__(self, oracle):
        self.o = oracle
        self.hash_elements, self.merge = oracle.hash_elements, oracle.merge
        self.merge_with_int, self.merge_many = oracle.merge_with_int, oracle.merge_many
        self.calls = 0

    def merge_with_int_pairs(self, seeds, values):
        self.pair_calls = getattr(self, "pair_calls", 0) + 1
        return np.stack([self.o.merge_with_int(bytes(s), np.array([v], dtype=np.uint64))[0] for s, v in zip(np.asarray(seeds).reshape(-1, 32), values)])

    def merge_plan(self, known, left, right, level_counts, want):
        self.calls += 1
        table = [bytes(k) for k in np.asarray(known).reshape(-1, 32)]
        off = 0
        for cnt in level_counts:
            cnt = int(cnt)
            assert cnt > 0 and max(left[off:off + cnt].max(), right[off:off + cnt].max()) < len(table)
            pairs = np.frombuffer(b"".join(table[l] + table[r] for l, r in zip(left[off:off + cnt], right[off:off + cnt])), dtype=np.uint8)
            table += [d.tobytes() for d in self.o.merge(pairs.reshape(cnt, 64))]
            off += cnt
        assert off == len(left) == len(right)
        return np.frombuffer(b"".join(table[w] for w in want), dtype=np.uint8).reshape(-1, 32)


def test_children_openings_in_one_merge_plan(zkl, oracle, child):
    """All trees of all children authenticated by ONE merge schedule (trace, constraint, every FRI layer; children of
    different sizes and partitionings side by side)."""
    R, S = zkl.replay, zkl.step
    h = _PlanHasher(oracle)
    kids = [child[0], _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)[0], _step(zkl, oracle, 64, 11)[0]]
    fss = [R.replay_fs_from_step(h, k) for k in kids]
    # the same challenges from the lockstep replay of all children (n = 64 and n = 32 children: different widths of the OOD
    # frame are not needed here, different numbers of FRI layers are)
    assert R.replay_fs_from_steps(h, kids) == fss and R.replay_fs_from_steps(h, []) == []
    assert len({len(f.fri_alphas) for f in fss}) == 2 and h.pair_calls <= 2 * max(len(f.fri_alphas) for f in fss) + 8
    bad_nonce = replace(kids[1], proof=replace(kids[1].proof, inner=replace(kids[1].proof.inner, grinding=30)))
    with pytest.raises(S.InvalidInput, match="grinding factor"):
        R.replay_fs_from_steps(h, [kids[0], bad_nonce])
    R.verify_children_openings(h, kids, fss)
    assert h.calls == 1
    # schedule shape: indices only refer to earlier entries, one root per tree
    plan = R.MergePlan()
    v = kids[0].proof.inner
    leaves = R.hash_rows(h, np.frombuffer(v.trace_query_values, dtype=np.uint64).reshape(len(fss[0].query_positions), v.main_width, 2), v.main_width)
    plan.add_opening(*R.read_batch_proof(v.trace_query_paths), leaves, fss[0].query_positions, v.parse_commitments()[0][0], "trace")
    known, left, right, counts, want = plan.flatten()
    assert len(counts) == 10 and counts[-1] == 1 and len(want) == 1          # 1024 LDE rows -> depth 10
    assert int(counts.sum()) == len(left) and want[0] == len(known) + len(left) - 1
    plan.run(h)
    # tampering is reported with the child and the tree it belongs to
    def flip(b, i):
        return b[:i] + bytes([b[i] ^ 1]) + b[i + 1:]
    v2 = kids[2].proof.inner
    layers = list(v2.fri_layers)
    layers[3] = (layers[3][0], flip(layers[3][1], len(layers[3][1]) - 5))
    cases = [(1, dict(trace_query_values=flip(kids[1].proof.inner.trace_query_values, 7)), "child 1: trace openings"),
             (2, dict(fri_layers=layers), "child 2: FRI layer 3"),
             (0, dict(constraint_query_paths=flip(v.constraint_query_paths, 40)), "child 0: constraint openings"),
             (2, dict(fri_remainder=flip(v2.fri_remainder, 17)), "child 2: FRI remainder")]
    for ci, kw, msg in cases:
        bad = list(kids)
        bad[ci] = replace(kids[ci], proof=replace(kids[ci].proof, inner=replace(kids[ci].proof.inner, **kw)))
        with pytest.raises(S.InvalidInput, match=msg):
            R.verify_children_openings(h, bad, fss)
    R.verify_children_openings(h, [], [])


def test_deep_and_fri_arithmetic_of_a_child(zkl, oracle, child):
    """agg/trace.rs:752-950, 1076-1193 on host integers: layer 0 is the DEEP composition of the opened rows, layers fold with
    the replayed alphas, the remainder closes the chain; any moved value breaks exactly the link it belongs to."""
    R, S = zkl.replay, zkl.step
    step = child[0]
    fs = R.verify_child(oracle, step)
    for other in (_step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)[0], _step(zkl, oracle, 128, 5, queries=16)[0]):
        R.verify_child(oracle, other)
    v = step.proof.inner

    def with_inner(**kw):
        return replace(step, proof=replace(step.proof, inner=replace(v, **kw)))

    def bump(b, i):                      # +1 on element i (stays a field element)
        x = (int.from_bytes(b[16 * i:16 * i + 16], "little") + 1) % zkl.field.P
        return b[:16 * i] + x.to_bytes(16, "little") + b[16 * i + 16:]

    layers = list(v.fri_layers)
    moved0 = [(bump(layers[0][0], 0), layers[0][1])] + layers[1:]
    moved3 = layers[:3] + [(bump(layers[3][0], 1), layers[3][1])] + layers[4:]
    cases = [("DEEP composition", dict(trace_query_values=bump(v.trace_query_values, 3))),
             ("DEEP composition", dict(constraint_query_values=bump(v.constraint_query_values, 0))),
             ("DEEP composition", dict(ood_trace=bump(v.ood_trace, 7))),
             ("DEEP composition", dict(ood_constraints=bump(v.ood_constraints, 1))),
             ("DEEP composition", dict(fri_layers=moved0)),
             ("FRI layer [34] values", dict(fri_layers=moved3)),
             ("remainder does not continue", dict(fri_remainder=bump(v.fri_remainder, 1)))]
    for msg, kw in cases:
        with pytest.raises(S.InvalidInput, match=msg):
            R.verify_child_deep_and_fri(with_inner(**kw), fs)
    with pytest.raises(S.InvalidInput, match="DEEP composition"):
        R.verify_child_deep_and_fri(step, replace(fs, z_ood=fs.z_ood + 1))
    with pytest.raises(S.InvalidInput):
        R.verify_child_deep_and_fri(step, replace(fs, fri_alphas=[fs.fri_alphas[0] + 1] + fs.fri_alphas[1:]))
    with pytest.raises(S.InvalidInput, match="degree exceeds"):
        R.verify_child_deep_and_fri(with_inner(fri_remainder=v.fri_remainder + bytes(16)), fs)


def test_per_query_merkle_paths(zkl, oracle, child):
    """agg/child.rs:230-400 + tests/agg_child.rs:60-146: every extracted path folds back to its commitment through
    merkle_root_from_leaf; leaf digests follow the partition convention; FRI entries carry folded positions and the opened pairs."""
    R = zkl.replay
    for step in (child[0], _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)[0]):
        fs = R.replay_fs_from_step(oracle, step)
        v = step.proof.inner
        mp = R.child_merkle_proofs(oracle, step, fs)
        troots, croot, froots = v.parse_commitments()
        depth = v.lde_domain_size.bit_length() - 1
        assert len(mp.trace_paths) == len(mp.constraint_paths) == len(fs.query_positions)
        for p, tp, cp in zip(fs.query_positions, mp.trace_paths, mp.constraint_paths):
            assert len(tp.siblings) == len(cp.siblings) == depth
            assert R.merkle_root_from_leaf(oracle, tp.leaf, p, tp.siblings) == troots[0]
            assert R.merkle_root_from_leaf(oracle, cp.leaf, p, cp.siblings) == croot
        assert R.merkle_root_from_leaf(oracle, mp.trace_paths[0].leaf, fs.query_positions[0] ^ 1, mp.trace_paths[0].siblings) != troots[0]
        assert len(mp.fri_layers) == v.num_fri_layers
        positions, domain = list(fs.query_positions), v.lde_domain_size
        for d, layer in enumerate(mp.fri_layers):
            positions = R.fold_positions(positions, domain)
            assert [e.idx for e in layer] == positions and len(layer[0].path.siblings) == depth - 1 - d
            for e in layer:
                assert R.merkle_root_from_leaf(oracle, e.path.leaf, e.idx, e.path.siblings) == froots[d]
                assert e.path.leaf == oracle.hash_elements(zkl.prover.ints_to_fe(list(e.values)).reshape(1, 2, 2))[0].tobytes()
            domain //= 2
    bad = replace(child[0], proof=replace(child[0].proof, inner=replace(child[0].proof.inner,
                  trace_query_values=bytes([child[0].proof.inner.trace_query_values[0] ^ 1]) + child[0].proof.inner.trace_query_values[1:])))
    with pytest.raises(zkl.step.InvalidInput, match="trace openings"):
        R.child_merkle_proofs(oracle, bad, R.replay_fs_from_step(oracle, child[0]))


def test_child_compact_and_transcript_views(zkl, oracle, child):
    """tests/transcript_basic.rs:48-178 and tests/agg_child.rs restated: the transcript's commitments recombine into the compact
    root, shapes are consistent, tampering with a root is caught."""
    AC, S = zkl.agg_child, zkl.step
    step = child[0]
    compact = AC.child_compact_from_step(oracle, step)
    tx = AC.child_transcript_from_step(oracle, step, compact)
    AC.verify_child_transcript(tx)
    v = step.proof.inner
    assert compact.step_digest == step.digest() and compact.trace_root == step.proof.root_trace and compact.pow_nonce == v.pow_nonce
    assert compact.fs_challenges.query_positions == child[1].positions[:child[1].num_unique_queries]
    assert len(compact.merkle_proofs.trace_paths) == tx.num_queries == len(tx.trace_main_openings) == len(tx.constraint_openings)
    assert tx.fri_final.coeffs and tx.fri_final.deg <= len(tx.fri_final.coeffs) - 1
    assert all(len(r) == tx.main_trace_width for r in tx.trace_main_openings)
    assert tx.constraint_frame_width > 0 and all(len(r) == tx.constraint_frame_width for r in tx.constraint_openings)
    assert len(tx.ood_main_current) == len(tx.ood_main_next) == tx.main_trace_width
    assert len(tx.ood_quotient_current) == len(tx.ood_quotient_next) == tx.constraint_frame_width
    assert len(tx.fri_layers) == v.num_fri_layers and all(tx.fri_layers)
    assert [p.values for p in compact.merkle_proofs.fri_layers[0]] == tx.fri_layers[0]
    # transcript_basic.rs: a tampered root no longer recombines
    bad = replace(tx, trace_roots=[bytes([tx.trace_roots[0][0] ^ 1]) + tx.trace_roots[0][1:]])
    with pytest.raises(S.InvalidInput, match="do not match compact trace_root"):
        AC.verify_child_transcript(bad)
    for kw, msg in ((dict(trace_main_openings=tx.trace_main_openings[:-1]), "trace_main_openings length"),
                    (dict(constraint_openings=[r[:-1] for r in tx.constraint_openings]), "row width"),
                    (dict(fri_layers=[[]] + tx.fri_layers[1:]), "empty FRI layer"),
                    (dict(fri_final=AC.FriFinal(5, tx.fri_final.coeffs)), "deg exceeds"),
                    (dict(fri_final=AC.FriFinal(0, [])), "must be non-empty"),
                    (dict(num_queries=0), "must not contain trace openings")):
        with pytest.raises(S.InvalidInput, match=msg):
            AC.verify_child_transcript(replace(tx, **kw))
    assert tx.fri_final.deg == max((i for i, c in enumerate(tx.fri_final.coeffs) if c), default=0)


def test_python_verifier_ood_consistency(zkl, oracle, child):
    """The last link of an independent verifier: constraints evaluated at z from the OOD frame (Python restatement of the AIR),
    divided and combined with the composition coefficients of the replayed transcript, equal the composition columns at z —
    and any moved OOD value, public input or constraint breaks the equality."""
    V = zkl.verify
    from util import P, from_fe
    n = 64
    for x in (3, 0x1234567 << 70 | 5, P - 2):
        assert V.periodic_at(n, x) == from_fe(oracle.periodic_at(n, x))
    step, _, air = child
    lhs, rhs, z = V.ood_consistency(oracle, step, air)
    assert lhs == rhs and z == zkl.replay.replay_fs_from_step(oracle, step).z_ood
    good = rhs
    v = step.proof.inner
    bump = lambda b, i: b[:16 * i] + ((int.from_bytes(b[16 * i:16 * i + 16], "little") + 1) % P).to_bytes(16, "little") + b[16 * i + 16:]
    for kw in (dict(ood_trace=bump(v.ood_trace, 45)), dict(ood_trace=bump(v.ood_trace, v.main_width + 45)), dict(ood_constraints=bump(v.ood_constraints, 2))):
        bad = replace(step, proof=replace(step.proof, inner=replace(v, **kw)))
        lhs, rhs, _ = V.ood_consistency(oracle, bad, air)
        assert lhs != rhs
    # (a schedule column's NEXT-row value enters no transition constraint: moving it must leave the left side alone)
    quiet = replace(step, proof=replace(step.proof, inner=replace(v, ood_trace=bump(v.ood_trace, v.main_width + 12))))
    assert V.ood_consistency(oracle, quiet, air)[0] == good
    # a proof of another shape: RAM-less layout with more queries and partitions
    step2, _, air2 = _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)
    lhs, rhs, _ = V.ood_consistency(oracle, step2, air2)
    assert lhs == rhs
    zkl.replay.verify_child(oracle, step2)


def test_host_verifier_accepts_and_rejects(zkl, oracle, child):
    """verify.verify_step: the whole verification of a step proof from its ZKLSTP1 bytes (the AIR instance is rebuilt from the
    proof), accepting valid proofs of several shapes and naming the failing check for tampered ones."""
    V, S = zkl.verify, zkl.step
    step = S.StepProof.from_bytes(child[0].to_bytes())
    # 8 queries at blowup 16 = 32 query bits (grinding only counts from 80 up, prove.rs:1185) -> 31 bits of conjectured security
    assert V.conjectured_security_bits(8, 16, 4, 1, 64 * 16) == 31 and V.conjectured_security_bits(64, 16, 16, 1, 65536) == 111
    assert V.conjectured_security_bits(64, 16, 16, 2, 1 << 16) == 128 and V.conjectured_security_bits(1, 2, 0, 1, 128) == 0
    fs = V.verify_step(oracle, step, 31)
    assert fs.query_positions == child[1].positions[:child[1].num_unique_queries]
    V.verify_step(oracle, _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)[0], 100)   # 160 query bits, field-bound: 128 - 9 - 1
    # the proof's self-declared options are not trusted: below the caller's minimum nothing else is looked at
    with pytest.raises(S.InvalidInput, match="31 bits of conjectured security, below the required 64"):
        V.verify_step(oracle, step, 64)
    v = step.proof.inner
    bump = lambda b, i: b[:16 * i] + ((int.from_bytes(b[16 * i:16 * i + 16], "little") + 1) % zkl.field.P).to_bytes(16, "little") + b[16 * i + 16:]
    with_inner = lambda **kw: replace(step, proof=replace(step.proof, inner=replace(v, **kw)))
    for kw, msg in ((dict(ood_trace=bump(v.ood_trace, 45)), "out-of-domain"), (dict(ood_constraints=bump(v.ood_constraints, 0)), "out-of-domain"),
                    (dict(trace_query_values=bump(v.trace_query_values, 9)), "trace openings"),
                    (dict(pow_nonce=v.pow_nonce + 1), "grinding factor|unique query|openings"),
                    (dict(fri_remainder=bump(v.fri_remainder, 0)), "remainder")):
        with pytest.raises(S.InvalidInput, match=msg):
            V.verify_step(oracle, with_inner(**kw), 31)
    with pytest.raises(S.InvalidInput):
        V.verify_step(oracle, replace(step, vm_usage_mask=step.vm_usage_mask ^ 4), 31)     # another AIR: constraint count / seed differ
    # a forged context: fewer queries / no grinding / another extension or batching declared -> refused before any replay
    P = zkl.field.P
    noncanon = lambda b, i: b[:16 * i] + (P + 5).to_bytes(16, "little") + b[16 * i + 16:]   # an encoding winter-math's reader rejects
    for kw, msg in ((dict(num_queries=1), "below the required"), (dict(blowup=2), "below the required"), (dict(num_queries=0), "no queries"),
                    (dict(field_extension=2), "FieldExtension::None"), (dict(batching=(1, 0)), "Linear batching"), (dict(fri_folding=4), "folding factor 2"),
                    (dict(ood_trace=noncanon(v.ood_trace, 3)), "non-canonical"), (dict(trace_query_values=noncanon(v.trace_query_values, 0)), "non-canonical"),
                    (dict(fri_remainder=v.fri_remainder[:-1]), "multiple of 16")):
        with pytest.raises(S.InvalidInput, match=msg):
            V.verify_step(oracle, with_inner(**kw), 31)
