"""CPU tests of the zl1 step container, the ZKLSTP1 codec, the step digest and the segment planner.
They read like the reference's tests/step_serialization.rs, tests/transcript_basic.rs:48-115 and the unit tests at
segment_planner.rs:336-475. Proof bytes come from the CPU checker (tests only); the GPU suite repeats the container
checks on GPU proofs (test_gpu_segments.py)."""
import pytest

import synth


@pytest.fixture(scope="module")
def proven(zkl, oracle):
    """One small segment proven by the CPU checker + everything a step container needs."""
    n = 64
    trace, air_pi, opts = synth.valid_segment(n=n, queries=8, blowup=16, grind=4, seed=7)
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    seed = zkl.prover.ints_to_fe(a.seed_elements())
    summ = zkl.capi.ProofSummary()
    proof = oracle.prove_segment(trace, a.to_c(), seed, opts.to_c(), summ)
    return trace, air_pi, opts, proof, summ


def _container(zkl, proven, index=0, total=1, main_args=()):
    S = zkl.step
    trace, air_pi, opts, proof, _ = proven
    from dataclasses import replace
    core = replace(air_pi.core, main_args=list(main_args))
    pi_len = len(replace(air_pi, core=core).to_elements())      # as prove.rs:1131 does
    meta = S.StepMeta.from_env(trace.shape[1], opts.blowup, opts.num_queries, 40, pi_len)
    boundaries = {f: bytes([i + 1]) * 16 + bytes(16) for i, f in enumerate(S.BOUNDARY_FIELDS)}
    boundaries["state_in_hash"] = S.vm_state_hash([0] * 8)
    boundaries["state_out_hash"] = S.vm_state_hash(list(range(8)))
    zl1 = S.Zl1Proof.new_multi_segment(core.program_id, meta, core, index, total, S.fe_to_bytes_fold(5), boundaries, proof)
    return S.StepProof(zl1, core, (11, 22, 33), air_pi.vm_usage_mask, air_pi.ram_delta_clk_bits)


def test_inner_proof_view_matches_the_prover_summary(zkl, proven):
    S = zkl.step
    trace, air_pi, opts, proof, summ = proven
    v = S.InnerProof.parse(proof)
    assert (v.main_width, v.aux_width, v.trace_len) == (trace.shape[0], 0, trace.shape[1])
    assert (v.num_queries, v.blowup, v.grinding, v.fri_folding, v.fri_remainder_max_degree) == (8, 16, 4, 2, 1)
    assert (v.num_partitions, v.hash_rate) == (opts.num_partitions, opts.hash_rate)
    assert v.num_unique_queries == summ.num_unique_queries and v.pow_nonce == summ.pow_nonce
    trace_roots, constraint_root, fri_roots = v.parse_commitments()
    assert trace_roots == [bytes(summ.trace_root)] and constraint_root == bytes(summ.constraint_root)
    assert len(fri_roots) == summ.num_fri_roots == v.num_fri_layers + 1
    assert b"".join(fri_roots) == bytes(summ.fri_roots)[:32 * summ.num_fri_roots]
    # shapes the child transcript relies on (transcript_basic.rs:60-115)
    assert len(v.trace_rows()) == v.num_unique_queries and all(len(r) == v.main_width for r in v.trace_rows())
    assert len(v.constraint_rows()) == v.num_unique_queries and len(v.constraint_rows()[0]) > 0
    assert len(v.ood_trace) == 2 * 16 * v.main_width and len(v.ood_constraints) == 2 * 16 * len(v.constraint_rows()[0])
    coeffs = v.fri_remainder_coeffs()
    assert len(coeffs) == 2 and len(v.fri_layers) == v.num_fri_layers
    # root_trace = blake3 over the roots in commitment order (format.rs:219-243, agg/child.rs:905-930)
    expect = zkl.blake3.blake3(b"zkl/step/root_trace" + air_pi.core.program_id + bytes(summ.trace_root) + bytes(summ.constraint_root)
                               + bytes(summ.fri_roots)[:32 * summ.num_fri_roots])
    assert S.root_trace(air_pi.core.program_id, v) == expect


def test_inner_proof_view_rejects_malformed_bytes(zkl, proven):
    S = zkl.step
    proof = proven[3]
    for bad in (proof[:-1], proof + b"\x00", proof[:40], b""):
        with pytest.raises(S.InvalidInput):
            S.InnerProof.parse(bad)
    v = S.InnerProof.parse(proof)
    v.commitments = v.commitments[:-32]
    with pytest.raises(S.InvalidInput, match="commitments layout"):
        v.parse_commitments()


def test_vint_reader_matches_the_writer_format(zkl):
    """winter-utils vint64 as written by csrc/api_stark.inl byte_writer::usize."""
    S = zkl.step

    def enc(value):
        zeros = 64 if value == 0 else 64 - value.bit_length()
        ln = min(max(zeros - 1, 0) // 7, 8)
        length = 9 - ln
        if length == 9:
            return b"\x00" + value.to_bytes(8, "little")
        return ((((value << 1) | 1) << (length - 1)) & ((1 << (8 * length)) - 1)).to_bytes(length, "little")

    for value in [0, 1, 127, 128, 16383, 16384, 26112, (1 << 21) - 1, 1 << 21, (1 << 56) - 1, 1 << 56, (1 << 64) - 1]:
        r = S._Reader(enc(value) + b"\xAA")
        assert r.usize() == value and r.u8() == 0xAA


def test_step_proof_roundtrip_preserves_digest_and_meta(zkl, proven):
    """step_serialization.rs:48-108."""
    S = zkl.step
    args = [("u64", 10), ("u128", (1 << 100) + 3), ("bytes32", bytes(range(32)))]
    step = _container(zkl, proven, index=2, total=5, main_args=args)
    data = step.to_bytes()
    assert data[:7] == b"ZKLSTP1" and data.endswith(proven[3])
    back = S.StepProof.from_bytes(data)
    assert back.digest() == step.digest() and len(step.digest()) == 32 and step.digest()[16:] == bytes(16)
    assert back.state_in_hash() == step.state_in_hash() and back.state_out_hash() == step.state_out_hash()
    assert back.proof.header.suite_id == step.proof.header.suite_id == proven[1].core.program_id
    assert back.proof.meta == step.proof.meta
    assert (back.proof.meta.m, back.proof.meta.rho, back.proof.meta.q, back.proof.meta.o, back.proof.meta.lambda_bits) == (64, 16, 8, 2, 40)
    assert back.proof.meta.v_units == 64 * 8
    assert back.proof.pi == step.proof.pi and (back.proof.pi.segment_index, back.proof.pi.segments_total) == (2, 5)
    assert back.proof.root_trace == step.proof.root_trace
    for f in ("program_id", "program_commitment", "merkle_root", "feature_mask", "main_args"):
        assert getattr(back.pi_core, f) == getattr(step.pi_core, f), f
    assert back.rom_acc == (11, 22, 33)
    assert (back.vm_usage_mask, back.ram_delta_clk_bits) == (step.vm_usage_mask, step.ram_delta_clk_bits)
    assert back.to_bytes() == data


def test_step_digest_binds_every_component(zkl, proven):
    """digest.rs:15-68: meta, public inputs (incl. segment index and boundary state) and the commitment root all enter."""
    from dataclasses import replace
    S = zkl.step
    step = _container(zkl, proven, index=1, total=4)
    d0 = step.digest()
    variants = [
        replace(step.proof, meta=replace(step.proof.meta, lambda_bits=41)),
        replace(step.proof, pi=replace(step.proof.pi, segment_index=2)),
        replace(step.proof, pi=replace(step.proof.pi, rom_s_out_2=bytes(32))),
        replace(step.proof, pi=replace(step.proof.pi, state_out_hash=bytes(32))),
        replace(step.proof, root_trace=bytes(32)),
        replace(step.proof, header=replace(step.proof.header, suite_id=bytes(32))),
    ]
    digests = {S.step_digest(p) for p in variants}
    assert len(digests) == len(variants) and d0 not in digests
    # the digest is the sponge's lane 0 of three chained two-lane hashes: check the chain against the generic permutation
    sid = step.proof.header.suite_id
    suite = zkl.params.get_poseidon_suite(sid)
    l, r = 123456789, zkl.field.P - 5
    assert S.poseidon_hash_two_lanes(sid, l, r) == zkl.params.permute(suite, [l, r] + [0] * 8 + list(suite.dom))[0]


def test_two_lane_hash_matches_the_checker_permutation(zkl, oracle):
    """Same permutation as the commitment hasher when suite_id = [0;32] (poseidon/mod.rs:255-291 vs hasher.rs:173-190)."""
    S = zkl.step
    suite = zkl.params.get_poseidon_suite(bytes(32))
    l, r = 0xDEADBEEF << 64 | 7, 3
    state = zkl.prover.ints_to_fe([l, r] + [0] * 8 + list(suite.dom)).reshape(1, 12, 2)
    out = zkl.prover.fe_to_ints(oracle.permute(state)[0])
    assert S.poseidon_hash_two_lanes(bytes(32), l, r) == out[0]


def test_step_proof_decode_rejects_bad_input(zkl, proven):
    """step_serialization.rs:110-175 + the descriptive messages of step.rs:154-420."""
    S = zkl.step
    data = _container(zkl, proven, main_args=[("u64", 1)]).to_bytes()
    bad_magic = bytes([data[0] ^ 0xFF]) + data[1:]
    with pytest.raises(S.InvalidInput, match="magic"):
        S.StepProof.from_bytes(bad_magic)
    with pytest.raises(S.InvalidInput, match="too short"):
        S.StepProof.from_bytes(data[:3])
    with pytest.raises(S.InvalidInput, match="lambda_bits"):
        S.StepProof.from_bytes(data[:8])
    with pytest.raises(S.InvalidInput, match="suite_id bytes"):
        S.StepProof.from_bytes(data[:20])
    with pytest.raises(S.InvalidInput, match="inner proof bytes"):
        S.StepProof.from_bytes(data[:-1])
    off_args = 7 + 4 + 4 * 32 + 8
    assert data[off_args:off_args + 5] == b"\x01\x00\x00\x00\x00"
    with pytest.raises(S.InvalidInput, match="VmArg::U64"):
        S.StepProof.from_bytes(data[:off_args + 8])
    with pytest.raises(S.InvalidInput, match="invalid VmArg tag"):
        S.StepProof.from_bytes(data[:off_args + 4] + b"\x09" + data[off_args + 5:])
    with pytest.raises(S.InvalidInput, match="32-byte field"):
        S.StepProof.from_bytes(data[:off_args + 4 + 9 + 8 + 40])      # inside rom_acc_1
    # every strict prefix is rejected, never mis-parsed
    for cut in range(0, len(data) - len(proven[3]), 7):
        with pytest.raises(S.InvalidInput):
            S.StepProof.from_bytes(data[:cut])


def test_public_input_digest_and_state_hash(zkl):
    S, A = zkl.step, zkl.air
    core = A.PublicInputs(program_id=bytes([1]) * 32, feature_mask=0x33, main_args=[("u64", 2), ("bytes32", bytes([9]) * 32)])
    expect = zkl.blake3.blake3(b"zkl/pi/v1" + bytes([1]) * 32 + bytes(64) + (0x33).to_bytes(8, "little") + (2).to_bytes(4, "little")
                               + b"\x00" + (2).to_bytes(8, "little") + b"\x02" + bytes([9]) * 32)
    assert S.public_inputs_digest(core) == expect
    regs = [1, 2, 3, 4, 5, 6, 7, zkl.field.P - 1]
    assert S.vm_state_hash(regs) == zkl.blake3.blake3(b"zkl/vm/state-v1" + b"".join(r.to_bytes(16, "little") for r in regs))


def test_children_root(zkl):
    """agg/child.rs:853-895: order independent (leaves are sorted), odd layers pair the last node with itself."""
    S = zkl.step
    sid = bytes([3]) * 32
    kids = [(bytes([i]) * 16 + bytes(16), bytes([0x40 + i]) * 32) for i in range(5)]
    root = S.children_root(sid, kids)
    assert S.children_root(sid, kids[::-1]) == root and root[16:] == bytes(16)
    assert S.children_root(sid, []) == bytes(32)
    h = lambda a, b: S.poseidon_hash_two_lanes(sid, a, b)
    f = zkl.field.fold_bytes32_to_fe
    leaves = sorted(S.fe_to_bytes_fold(h(f(d), f(t))) for d, t in kids)
    a, b, c, d, e = (f(x) for x in leaves)
    assert root == S.fe_to_bytes_fold(h(h(h(a, b), h(c, d)), h(h(e, e), h(e, e))))
    assert S.children_root(sid, kids[:1]) == leaves_single(S, sid, kids[0], f)
    assert S.children_root(sid, kids[:4]) != root


def leaves_single(S, sid, kid, f):
    return S.fe_to_bytes_fold(S.poseidon_hash_two_lanes(sid, f(kid[0]), f(kid[1])))


def test_verify_chain(zkl, proven):
    from dataclasses import replace
    S = zkl.step
    a, b = _container(zkl, proven, 0, 2), _container(zkl, proven, 1, 2)
    with pytest.raises(S.InvalidInput, match="state_out_hash"):
        S.verify_chain([a, b])
    b.proof.pi = replace(b.proof.pi, state_in_hash=a.state_out_hash())
    S.verify_chain([a, b])
    assert S.children_root_from_steps([a, b]) == S.children_root(a.proof.header.suite_id,
                                                                 [(a.digest(), a.proof.root_trace), (S.step_digest(b.proof), b.proof.root_trace)])


# ---- segment planner (segment_planner.rs:336-475) -------------------------------------------------------------------
def test_single_segment_covers_full_trace(zkl):
    PL = zkl.planner
    segs = PL.plan_segments(4)
    assert len(segs) == 1 and (segs[0].r_start, segs[0].r_end) == (0, 4 * 32)
    assert PL.plan_segments(0) == [zkl.segments.Segment(0, 32)] and PL.plan_segments(3)[0].r_end == 128


def test_large_program_splits_into_multiple_segments(zkl):
    PL = zkl.planner
    levels = PL.MAX_SEGMENT_ROWS // 32 + 1 + 1          # target_levels Const ops + End
    segs = PL.plan_segments(levels)
    assert len(segs) > 1
    acc = 0
    for s in segs:
        assert s.r_start == acc and s.r_end > s.r_start and s.r_start % 32 == 0 and s.r_end % 32 == 0
        assert s.r_end - s.r_start <= PL.MAX_SEGMENT_ROWS
        acc = s.r_end
    assert acc == PL.next_power_of_two(levels) * 32
    # rollup-bench shape: 2048 levels -> 16 segments of 4096 rows; a larger budget gives fewer, larger segments
    assert [(s.r_start, s.r_end) for s in PL.plan_segments(2048)] == [(i * 4096, (i + 1) * 4096) for i in range(16)]
    assert len(PL.plan_segments(2048, max_segment_rows=65536)) == 1


def test_planner_walks_blocks_gaps_and_padded_tail(zkl, monkeypatch):
    PL = zkl.planner
    # 10 code levels -> 16 total; blocks [2,5) and [4,7) merge, gap [0,2), tail [7,10), padding [10,16); 3 levels per segment
    segs = PL.plan_segments(10, blocks=[(4, 3), (2, 3), (9, 0)], max_segment_rows=96)
    assert [(s.r_start // 32, s.r_end // 32) for s in segs] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 16)]
    with pytest.raises(zkl.step.InvalidInput, match="out of bounds"):
        PL.plan_segments(10, blocks=[(8, 3)], max_segment_rows=96)
    # a budget below one level still advances one level at a time
    assert len(PL.plan_segments(4, max_segment_rows=8)) == 4
    monkeypatch.setenv("ZKL_MAX_SEGMENT_ROWS", "64")
    assert len(PL.plan_segments(8)) == 4
    monkeypatch.setenv("ZKL_MAX_SEGMENT_ROWS", "not-a-number")
    assert len(PL.plan_segments(8)) == 1
    with pytest.raises(zkl.step.InvalidInput):
        PL.new_segment(5, 5)


def test_segment_features_and_masks(zkl):
    """segment_planner.rs:416-475 + from_ops."""
    A, PL, SG = zkl.air, zkl.planner, zkl.segments
    ops = ["Const", "Add", "Load", "SAbsorbN", "MerkleStep", "End"]
    f = PL.features_from_ops(ops)
    assert (f.vm, f.ram, f.sponge, f.merkle) == (True, True, True, True)
    f = PL.compute_segment_features_for_levels(ops, 0, 2)
    assert (f.vm, f.ram, f.sponge, f.merkle) == (True, False, False, False)
    f = PL.compute_segment_features_for_levels(ops, 4, 5)
    assert (f.vm, f.ram, f.sponge, f.merkle) == (False, False, False, True)
    f = PL.compute_segment_features_for_levels(ops, 6, 64)   # padded tail
    assert (f.vm, f.ram, f.sponge, f.merkle) == (False, False, False, False)
    full = A.FM_VM | A.FM_VM_EXPECT | A.FM_RAM | A.FM_MERKLE | A.FM_SPONGE | A.FM_POSEIDON
    m = SG.compute_segment_feature_mask(full, SG.SegmentFeatures(vm=True, ram=False, sponge=True, merkle=False))
    assert m & ~full == 0 and not m & A.FM_RAM and not m & A.FM_MERKLE
    assert m & A.FM_VM and m & A.FM_VM_EXPECT and m & A.FM_SPONGE and m & A.FM_POSEIDON
    m = SG.compute_segment_feature_mask(full, SG.SegmentFeatures(vm=False))
    assert not m & (A.FM_RAM | A.FM_MERKLE | A.FM_SPONGE | A.FM_POSEIDON) and m & A.FM_VM and m & A.FM_VM_EXPECT
    core = A.FM_VM | A.FM_VM_EXPECT
    assert SG.compute_segment_feature_mask(core, SG.SegmentFeatures(vm=True, ram=True, sponge=True, merkle=True)) == core


def _planner_walk_restated(base_levels, blocks, max_segment_rows):
    """segment_planner.rs:92-284 restated as the checker of planner.plan_segments: level ranges in level units."""
    total_levels = max(1 << max(base_levels - 1, 0).bit_length() if base_levels > 1 else 1, 1)
    if total_levels * 32 <= max_segment_rows:
        return [(0, total_levels)]
    max_levels = max(max_segment_rows // 32, 1)
    ranges, block_ranges = [], sorted((s, s + ln) for s, ln in blocks if ln)
    if not block_ranges:
        if base_levels > 0:
            ranges.append([0, base_levels])
    else:
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
    out, cur_start, cur_end = [], None, 0
    for lo, hi in ranges:
        lvl = lo
        while lvl < hi:
            if cur_start is None:
                cur_start = cur_end = lvl
            take = min(max_levels - (cur_end - cur_start), hi - lvl)
            cur_end += take
            lvl += take
            if cur_end - cur_start == max_levels:
                out.append((cur_start, cur_end))
                cur_start, cur_end = None, 0
    if cur_start is not None and cur_start < cur_end:
        out.append((cur_start, cur_end))
    return out


def test_planner_properties_random_programs(zkl):
    """Invariants the reference's planner tests assert on one program (segment_planner.rs:370-414), over random block tables:
    contiguous, level aligned, covering the padded trace, never above the row budget, and all full except possibly the last."""
    from hypothesis import given, settings, strategies as st
    PL = zkl.planner

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 3000), st.lists(st.tuples(st.integers(0, 2999), st.integers(0, 400)), max_size=6), st.integers(1, 300))
    def check(base_levels, raw_blocks, budget_levels):
        blocks = [(s, min(ln, base_levels - s)) for s, ln in raw_blocks if s < base_levels]
        max_rows = budget_levels * 32
        segs = PL.plan_segments(base_levels, blocks, max_segment_rows=max_rows)
        total = PL.next_power_of_two(base_levels) * 32
        assert segs[0].r_start == 0 and segs[-1].r_end == total
        for a, b in zip(segs, segs[1:]):
            assert a.r_end == b.r_start
        if len(segs) == 1:
            assert total <= max_rows or budget_levels >= total // 32
        else:
            assert all(s.r_end - s.r_start == max_rows for s in segs[:-1]) and 0 < segs[-1].r_end - segs[-1].r_start <= max_rows
        assert all(s.r_start % 32 == 0 and s.r_end % 32 == 0 for s in segs)
        assert [(s.r_start // 32, s.r_end // 32) for s in segs] == _planner_walk_restated(base_levels, blocks, max_rows)

    check()


def test_recursion_public_inputs_and_zklrc1_codec(zkl, proven):
    """lib.rs:395-480 (build_public) and the ZKLRC1 artifact codec lib.rs:483-774 over a chain of step proofs."""
    from dataclasses import replace
    S, RC = zkl.step, zkl.recursion
    steps = [_container(zkl, proven, i, 3) for i in range(3)]
    core = steps[0].pi_core
    rc = RC.build_recursion_public(core, steps)
    assert rc.children_count == 3 and rc.children_ms == [64, 64, 64] and rc.v_units_total == 3 * 64 * 8
    assert rc.pi_digest == S.public_inputs_digest(core) and rc.suite_id == core.program_id
    assert rc.children_root == S.children_root_from_steps(steps)
    assert (rc.vm_state_initial, rc.vm_state_final) == (steps[0].state_in_hash(), steps[-1].state_out_hash())
    assert rc.rom_s_final == (steps[-1].proof.pi.rom_s_out_0, steps[-1].proof.pi.rom_s_out_1, steps[-1].proof.pi.rom_s_out_2)
    assert (rc.profile_meta.m, rc.profile_meta.q, rc.profile_fri.lde_blowup, rc.profile_fri.num_layers) == (64, 8, 16, 1)
    one = RC.from_step_proof(steps[1])
    assert one.children_count == 1 and one.profile_fri.num_layers == len(steps[1].proof.inner.fri_layers)
    assert one.children_root == S.children_root(core.program_id, [(steps[1].digest(), steps[1].proof.root_trace)])
    els = rc.to_elements()
    assert len(els) == 32 and els[18] == 3 and els[0] == zkl.field.fold_bytes32_to_fe(core.program_id)
    assert els != replace(rc, vm_state_final=bytes(32)).to_elements()
    with pytest.raises(RC.RecursionInvalid, match="at least one step"):
        RC.build_recursion_public(core, [])
    with pytest.raises(RC.RecursionInvalid, match="share the same program_id"):
        RC.build_recursion_public(replace(core, main_args=[("u64", 1)]), steps)
    blob = RC.encode(b"\x01\x02\x03 aggregation proof bytes", rc)
    assert blob[:6] == b"ZKLRC1"
    proof, back = RC.decode(blob)
    assert proof == b"\x01\x02\x03 aggregation proof bytes" and back == rc and RC.encode(proof, back) == blob
    with pytest.raises(S.InvalidInput, match="magic"):
        RC.decode(b"ZKLRC2" + blob[6:])
    with pytest.raises(S.InvalidInput, match="proof bytes truncated"):
        RC.decode(blob[:-1])
    with pytest.raises(S.InvalidInput, match="pi_digest truncated"):
        RC.decode(blob[:6 + 70])
    for cut in range(0, len(blob), 5):
        with pytest.raises(S.InvalidInput):
            RC.decode(blob[:cut])


def test_native_two_lane_hash_matches_python(zkl):
    """zkl_hash_two_lanes_host (the kernels' limb arithmetic compiled for the host) against the Python-integer permutation,
    under the hasher suite and a program suite, incl. edge values; batches keep their order."""
    import random
    import time
    S, P = zkl.step, zkl.field.P
    rng = random.Random(3)
    for sid in (bytes(32), bytes([7]) * 32):
        pairs = [(0, 0), (1, 0), (0, 1), (P - 1, P - 1), (P - 1, 1), ((1 << 127) + 12345, (1 << 64) - 1)]
        pairs += [(rng.randrange(P), rng.randrange(P)) for _ in range(20)]
        got = S.hash_two_lanes_batch(sid, pairs)
        assert got == [S.poseidon_hash_two_lanes_py(sid, l, r) for l, r in pairs]
        assert S.poseidon_hash_two_lanes(sid, *pairs[7]) == got[7] and all(0 <= x < P for x in got)
        assert S.hash_two_lanes_batch(sid, [(l + P, r + 2 * P) for l, r in pairs[:3]]) == got[:3]      # reduced before the call
    assert S.hash_two_lanes_batch(bytes(32), []) == []
    t = time.time()
    S.hash_two_lanes_batch(bytes(32), [(i, i + 1) for i in range(200)])
    assert time.time() - t < 2.0
