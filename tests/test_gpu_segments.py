"""GPU tests of the segment driver (mirror of prove.rs:954-1175): device-side slicing, boundary cells, usage-mask
scan and multi-context proving, checked against numpy restatements and the restated verifier."""
import ctypes as C

import numpy as np
import pytest

import synth
from util import fast_random_trace, from_fe

pytestmark = pytest.mark.gpu


def _usage_mask_numpy(t, cols, has_ram):
    """prove.rs:1289-1392 restated on a (W, n, 2) uint64 array."""
    nz = lambda c: (t[c, :, 0] != 0) | (t[c, :, 1] != 0)
    n = t.shape[1]
    fin = (np.arange(n) % 32) == 28
    o = cols.op_start
    mask = 0
    for bit, cs in ((0, (o + 9, o + 7)), (1, (o + 10,)), (2, (o + 11,)), (3, (o + 12,)), (4, (o + 14,)), (5, (o + 13,)), (6, (o + 6,))):
        if any((nz(c) & fin).any() for c in cs):
            mask |= 1 << bit
    if nz(o + 8).any():
        mask |= 1 << 7
    bits = 0
    if has_ram:
        s = nz(cols.ram_start)
        same = (t[cols.ram_start + 1, :-1] == t[cols.ram_start + 1, 1:]).all(axis=1)
        pair = s[:-1] & s[1:] & same
        if pair.any():
            mask |= 1 << 8
            for i in range(32):
                if (nz(cols.gadget_b_start + i)[:-1] & pair).any():
                    bits |= 1 << i
    return mask, bits


def test_slice_and_usage_mask_match_numpy(ctx, zkl):
    A, S = zkl.air, zkl.segments
    full_cols = A.Columns.baseline()
    n_full = 256
    full = fast_random_trace(full_cols.width, n_full, seed=3)
    # sparsify the columns the scan looks at so that both outcomes occur
    for c in list(range(full_cols.op_start, full_cols.op_start + 17)) + [full_cols.ram_start]:
        full[c, ::3] = 0
    full[full_cols.ram_start + 1, 10:14] = full[full_cols.ram_start + 1, 10]
    full[full_cols.ram_start, 8:16] = 1
    dt = S.DeviceTrace(ctx, full)
    for cfg in (A.LayoutConfig(ram=False, merkle=False), A.LayoutConfig(merkle=False), A.LayoutConfig()):
        lay = S.SegmentLayout(full_cols, cfg)
        W, r0, n = lay.cols.width, 64, 128
        dev = C.c_void_p()
        ctx._check(ctx.lib.zkl_dev_alloc(ctx.h, W * n * 16, C.byref(dev)))
        ctx._check(ctx.lib.zkl_slice_segment_dev(ctx.h, dt.dev, n_full, lay.seg_to_full.ctypes.data, W, r0, n, dev))
        exp = full[lay.seg_to_full.astype(np.int64), r0:r0 + n]
        cells = [(c, r) for c in (0, 5, W - 1) for r in (0, 7, n - 1)]
        got = S.read_cells(ctx, dev, n, cells)
        assert got == [int(exp[c, r, 0]) | (int(exp[c, r, 1]) << 64) for c, r in cells]
        m, b = C.c_uint32(), C.c_uint32()
        cc = lay.cols.to_c()
        ctx._check(ctx.lib.zkl_usage_mask_dev(ctx.h, dev, W, n, C.byref(cc), int(cfg.ram), C.byref(m), C.byref(b)))
        assert (m.value, b.value) == _usage_mask_numpy(exp, lay.cols, cfg.ram)
        ctx._check(ctx.lib.zkl_dev_free(ctx.h, dev))
    dt.free()


def test_prove_program_multi_segment(ctx, zkl, oracle):
    """A 4-segment program proven on two contexts (streams); every step proof verifies and equals the proof of the
    host-sliced segment (agg_multiseg.rs-style: max_segment_rows shrunk, max_concurrent_segments = 2)."""
    A, S = zkl.air, zkl.segments
    n_full, seg_rows = 512, 128
    full, air_pi_full, opts = synth.valid_segment(n=n_full, queries=8, blowup=16, grind=4, seed=21, full_layout=True)
    core = air_pi_full.core
    segs = S.plan_fixed_segments(n_full, seg_rows)
    feats = [S.SegmentFeatures(vm=True, sponge=True)] * len(segs)
    ctx2 = zkl.Context(0)
    steps = S.prove_program([ctx, ctx2], full, segs, core, feats, opts)
    assert [s.segment_index for s in steps] == [0, 1, 2, 3]
    full_cols = A.Columns.baseline()
    summaries = {}
    for s, seg in zip(steps, segs):
        assert s.width == 204 and s.rows == seg_rows
        a = A.Air(s.width, s.rows, s.air_pi, s.options)
        ac, oc, seed = a.to_c(), s.options.to_c(), zkl.prover.ints_to_fe(a.seed_elements())
        rc, msg = oracle.verify_segment(s.proof, s.width, s.rows, ac, seed, oc)
        assert rc == 0, (s.segment_index, msg)
        # boundary values: pc at the first map row, ROM state in/out (prove.rs:1197-1287)
        fi = lambda c, r: int(full[c, r, 0]) | (int(full[c, r, 1]) << 64)
        assert s.air_pi.pc_init == fi(full_cols.pc, seg.r_start) == seg.r_start // 32
        assert list(s.air_pi.rom_s_in) == [fi(full_cols.rom_s_start + i, seg.r_start) for i in range(3)]
        assert list(s.air_pi.rom_s_out) == [fi(full_cols.rom_s_start + i, seg.r_end - 4) for i in range(3)]
        # same bytes as proving the host-sliced segment through the host-buffer entry point
        lay = S.SegmentLayout(full_cols, A.LayoutConfig(ram=False, merkle=False))
        host_seg = np.ascontiguousarray(full[lay.seg_to_full.astype(np.int64), seg.r_start:seg.r_end])
        pv = zkl.ZkProver(ctx, s.options, s.air_pi)
        again = pv.prove(host_seg)
        if again != s.proof:   # say what kind of difference it is: a racy kernel yields proofs that still verify
            diff = [i for i in range(min(len(again), len(s.proof))) if again[i] != s.proof[i]]
            rc2, msg2 = oracle.verify_segment(again, s.width, s.rows, ac, seed, oc)
            raise AssertionError(f"segment {s.segment_index}: {len(diff)} bytes differ from byte {diff[:1]}, lengths {len(again)} / {len(s.proof)}; "
                                 f"host-path proof verifies: {rc2 == 0} ({msg2})")
        summaries[s.segment_index] = pv.last_summary
    # zl1 step containers (proof/step.rs, proof/format.rs, proof/digest.rs): state hashes from the full trace, ZKLSTP1
    # round trip preserves the digest (step_serialization.rs:48-108), commitments recombine into root_trace
    # (transcript_basic.rs:48-60), consecutive segments chain (vm/trace/mod.rs:353-359)
    ST = zkl.step
    regs = lambda r: [int(full[full_cols.r_index(i), r, 0]) | (int(full[full_cols.r_index(i), r, 1]) << 64) for i in range(8)]
    challenges = []
    for s, seg in zip(steps, segs):
        c = s.step
        assert c.proof.inner_bytes == s.proof and c.proof.header.suite_id == core.program_id
        assert (c.proof.pi.segment_index, c.proof.pi.segments_total) == (s.segment_index, len(segs))
        assert c.state_in_hash() == ST.vm_state_hash(regs(seg.r_start)) and c.state_out_hash() == ST.vm_state_hash(regs(seg.r_end - 1))
        assert c.proof.pi.pc_init == ST.fe_to_bytes_fold(seg.r_start // 32)
        assert c.proof.pi.rom_s_out_1 == ST.fe_to_bytes_fold(s.air_pi.rom_s_out[1])
        v = c.proof.inner
        assert (v.main_width, v.trace_len, v.blowup, v.num_queries) == (s.width, s.rows, 16, 8)
        assert (c.proof.meta.m, c.proof.meta.o, c.proof.meta.pi_len) == (seg_rows, 2, len(s.air_pi.to_elements()))
        troots, croot, froots = v.parse_commitments()
        assert c.proof.root_trace == zkl.blake3.blake3(b"zkl/step/root_trace" + core.program_id + troots[0] + croot + b"".join(froots))
        back = ST.StepProof.from_bytes(c.to_bytes())
        assert back.digest() == c.digest() and back.proof.pi == c.proof.pi and back.proof.meta == c.proof.meta
        assert back.to_bytes() == c.to_bytes()
        # child replay on the device hasher (agg/fs.rs:37-245, agg/child.rs:1025-1099): the challenges re-derived from the
        # ZKLSTP1 bytes are the ones the prover used, and every opening re-hashes to its commitment
        fs = zkl.replay.replay_fs_from_step(ctx, back)
        summ = summaries[s.segment_index]
        assert fs.z_ood == int.from_bytes(bytes(summ.z), "little") and fs.pow_nonce == summ.pow_nonce
        assert fs.query_positions == list(summ.positions[:summ.num_unique_queries])
        zkl.replay.verify_child_openings(ctx, back, fs)
        zkl.replay.verify_child_deep_and_fri(back, fs)         # DEEP composition and FRI folding arithmetic (agg/trace.rs:752-1193)
        challenges.append(fs)
    # ... and all four children at once: leaves in three hashing batches, every tree in ONE device merge plan (zkl_merge_plan)
    before = ctx.launches
    assert zkl.replay.replay_fs_from_steps(ctx, [s.step for s in steps]) == challenges      # all transcripts in lockstep
    zkl.replay.verify_children_openings(ctx, [s.step for s in steps], challenges)
    assert ctx.launches > before
    tampered = [s.step for s in steps]
    inner = tampered[2].proof.inner
    bad_rows = inner.trace_query_values[:9] + bytes([inner.trace_query_values[9] ^ 4]) + inner.trace_query_values[10:]
    from dataclasses import replace
    tampered[2] = replace(tampered[2], proof=replace(tampered[2].proof, inner=replace(inner, trace_query_values=bad_rows)))
    with pytest.raises(ST.InvalidInput, match="child 2: trace openings"):
        zkl.replay.verify_children_openings(ctx, tampered, challenges)
    # the same replay with the host side in native code (zkl_replay_children): identical challenges, same verdicts
    assert zkl.replay.replay_children_native(ctx, [s.step for s in steps]) == challenges
    raw = bytearray(steps[2].proof)           # the native path reads the proof BYTES: flip the same bit there
    raw[steps[2].proof.index(inner.trace_query_values[:48]) + 9] ^= 4
    tampered[2] = replace(tampered[2], proof=replace(tampered[2].proof, inner_bytes=bytes(raw), inner=ST.InnerProof.parse(bytes(raw))))
    with pytest.raises(ST.InvalidInput, match="child 2: trace openings"):
        zkl.replay.replay_children_native(ctx, tampered)
    assert zkl.replay.replay_children_native(ctx, tampered, verify_openings=False) == challenges
    rebuilt = bytearray(steps[1].proof)
    rebuilt[-20] ^= 1                          # a byte of the FRI remainder (the proof ends: 32 remainder bytes, u8 partitions, u64 nonce)
    bad_rem = replace(steps[1].step, proof=replace(steps[1].step.proof, inner_bytes=bytes(rebuilt), inner=ST.InnerProof.parse(bytes(rebuilt))))
    with pytest.raises(ST.InvalidInput, match="child 1: FRI remainder"):
        zkl.replay.replay_children_native(ctx, [steps[0].step, bad_rem])
    ST.verify_chain([s.step for s in steps])
    assert len({s.step.digest() for s in steps}) == len(steps)
    assert ST.children_root_from_steps([s.step for s in steps]) == ST.children_root_from_steps([s.step for s in steps[::-1]])
    # the pi_prog binding exists only in the segment that starts at pc = 0 (schedule.rs:117-131)
    assert len(A.Air(204, seg_rows, steps[0].air_pi, steps[0].options).assertions) == len(A.Air(204, seg_rows, steps[1].air_pi, steps[1].options).assertions) + 1
    ctx2.close()


def test_child_replay_helpers(ctx, zkl, oracle):
    """agg/child.rs:1025-1068 on the GPU: leaf digests (with and without partitions) and path folding reproduce the
    committed root, like tests/agg_child.rs:60-146 does against live proofs."""
    import pyref
    W, n = 40, 64
    trace = fast_random_trace(W, n, seed=5)
    for parts in ((1, 16), (4, 8)):
        opts = zkl.air.ProofOptions(num_queries=8, blowup=16, grinding=0, num_partitions=parts[0], hash_rate=parts[1])
        prover = zkl.ZkProver(ctx, opts, None)
        lde, _ = prover.new_trace_lde(trace)
        root, _, olde = oracle.trace_lde_commit(trace, opts.to_c(), want_lde=True)
        N = n * 16
        psize = W if parts == (1, 16) else max(-(-W // parts[0]), parts[1])
        pos = [1, 2, 77, N - 1]
        rows = np.stack([olde[:, p] for p in pos])
        leaves = zkl.replay.hash_rows(ctx, rows, psize)
        # full tree from all leaves (oracle) gives the sibling paths
        all_leaves = np.stack([np.frombuffer(d, dtype=np.uint8) for d in zkl.replay.hash_rows(ctx, np.ascontiguousarray(olde.transpose(1, 0, 2)), psize)])
        oroot, nodes = oracle.merkle_build(all_leaves)
        assert oroot == root == lde.get_main_trace_commitment()
        sibs = []
        for p in pos:
            path, i = [all_leaves[p ^ 1]], (p + N) >> 1
            while i > 1:
                path.append(nodes[i ^ 1]); i >>= 1
            sibs.append([bytes(x) for x in path])
        assert all(zkl.replay.merkle_root_from_leaf(ctx, leaves[q], pos[q], sibs[q]) == root for q in range(len(pos)))
        lde.free()


def test_poseidon_levels_match_trace_builder(ctx, zkl):
    """zkl_poseidon_levels_dev = vm/trace/poseidon.rs:24-96 `apply_level_absorb` batched over levels: re-filling the blanked
    Poseidon lanes of an AIR-valid trace reproduces it cell for cell; the reference's own unit cases (0, 2, 10 inputs,
    poseidon.rs:199-230) give the sponge value in lane 0 of the final row."""
    if ctx.hash_mode != "default":
        pytest.skip("trace-builder kernel, independent of the hashing mode")
    A, SG = zkl.air, zkl.segments
    n = 512
    full, air_pi, _ = synth.valid_segment(n=n, seed=33, full_layout=True)
    cols = A.Columns.baseline()
    active = [lvl for lvl in range(n // 32) if lvl % 3 == 1]
    cell = lambda t, c, r: int(t[c, r, 0]) | (int(t[c, r, 1]) << 64)
    inputs = [[cell(full, cols.lanes_start + i, 32 * lvl) for i in range(10)] for lvl in active]
    blank = full.copy()
    for lvl in active:
        blank[cols.lanes_start:cols.lanes_start + 12, 32 * lvl:32 * lvl + 32] = 0
    assert not np.array_equal(blank, full)
    dt = SG.DeviceTrace(ctx, blank)
    SG.apply_levels_absorb(ctx, dt.dev, n, air_pi.core.program_id, active, inputs, cols.lanes_start)
    out = np.zeros_like(blank)
    ctx._check(ctx.lib.zkl_dev_download(ctx.h, out.ctypes.data, dt.dev, out.nbytes))
    assert np.array_equal(out, full)
    with pytest.raises(zkl.capi.ZklError, match="level outside the trace"):
        SG.apply_levels_absorb(ctx, dt.dev, n, air_pi.core.program_id, [n // 32], [[1]], cols.lanes_start)
    dt.free()
    # poseidon.rs:199-230
    sid = bytes([3]) * 32
    suite = zkl.params.get_poseidon_suite(sid)
    cases = [[], [1, 2], list(range(10))]
    empty = np.zeros((cols.width, 32 * len(cases), 2), dtype=np.uint64)
    dt = SG.DeviceTrace(ctx, empty)
    SG.apply_levels_absorb(ctx, dt.dev, empty.shape[1], sid, list(range(len(cases))), cases, cols.lanes_start)
    out = np.zeros_like(empty)
    ctx._check(ctx.lib.zkl_dev_download(ctx.h, out.ctypes.data, dt.dev, out.nbytes))
    for lvl, ins in enumerate(cases):
        st = list(ins) + [0] * (10 - len(ins)) + list(suite.dom)
        assert [cell(out, i, 32 * lvl) for i in range(12)] == st == [cell(out, i, 32 * lvl + 1) for i in range(12)]
        fin = zkl.params.permute(suite, st)
        for r in (28, 29, 30, 31):
            assert [cell(out, i, 32 * lvl + r) for i in range(12)] == fin
    assert not out[12:].any()
    dt.free()


def test_native_replay_with_partitions_and_mixed_children(ctx, zkl):
    """zkl_replay_children on children of different sizes and partitionings in one batch: partitioned row leaves
    (merge_many of partition digests, agg/child.rs:1031-1044), duplicate query draws (40 queries over 512 rows), different
    numbers of FRI layers. Python replay on the same device hasher is the reference."""
    if ctx.hash_mode not in ("default", "int-pipe"):
        pytest.skip("covered by two hashing modes")
    from dataclasses import replace
    from util import single_segment_step
    R = zkl.replay
    kids = []
    for n, seed, queries, parts, grind in ((32, 9, 40, (4, 8), 0), (64, 7, 8, None, 4), (128, 5, 16, (2, 16), 3)):
        trace, air_pi, opts = synth.valid_segment(n=n, queries=queries, blowup=16, grind=grind, seed=seed)
        if parts:
            opts = replace(opts, num_partitions=parts[0], hash_rate=parts[1])
        proof = zkl.ZkProver(ctx, opts, air_pi).prove(trace)
        kids.append(single_segment_step(zkl, air_pi, opts, n, proof))
    seeds = [{} for _ in kids]
    fss = [R.replay_fs_from_step(ctx, k, cap) for k, cap in zip(kids, seeds)]
    assert len(fss[0].query_positions) < 40                      # duplicates were drawn and removed
    # zkl_grind on the replayed coin seed finds the prover's nonce (the smallest one), for every grinding factor incl. 0
    assert [ctx.grind(cap["grinding_seed"], k.proof.inner.grinding) for k, cap in zip(kids, seeds)] == [f.pow_nonce for f in fss]
    n12 = ctx.grind(seeds[1]["grinding_seed"], 12)
    head = lambda v: int.from_bytes(ctx.merge_with_int(seeds[1]["grinding_seed"], [v])[0].tobytes()[:8], "little")
    assert n12 >= 1 and head(n12) % (1 << 12) == 0 and all(head(v) % (1 << 12) for v in range(max(1, n12 - 300), n12))
    R.verify_children_openings(ctx, kids, fss)
    assert R.replay_fs_from_steps(ctx, kids) == fss
    assert R.replay_children_native(ctx, kids) == fss
    # a flipped bit in a partitioned child's opened constraint row
    v = kids[0].proof.inner
    raw = bytearray(kids[0].proof.inner_bytes)
    raw[kids[0].proof.inner_bytes.index(v.constraint_query_values[:32]) + 5] ^= 2
    bad = replace(kids[0], proof=replace(kids[0].proof, inner_bytes=bytes(raw), inner=zkl.step.InnerProof.parse(bytes(raw))))
    with pytest.raises(zkl.step.InvalidInput, match="child 1: constraint openings"):
        R.replay_children_native(ctx, [kids[1], bad, kids[2]])
    with pytest.raises(zkl.step.InvalidInput, match="child 1: constraint openings"):
        R.verify_children_openings(ctx, [kids[1], bad, kids[2]], [fss[1], fss[0], fss[2]])
    # garbage instead of a proof is a per-child verdict, not a crash
    junk = replace(kids[1], proof=replace(kids[1].proof, inner_bytes=bytes(kids[1].proof.inner_bytes[:100])))
    with pytest.raises(zkl.step.InvalidInput, match="child 0: failed to decode"):
        R.replay_children_native(ctx, [junk, kids[2]])
