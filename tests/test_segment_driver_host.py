"""CPU test of the segment driver's HOST logic (zk-lisp_b200/segments.py: layouts, boundary cells, step containers, the
worker pool) with a stand-in context: the device entry points are replaced by numpy restatements and the CPU checker's
prover. The product never does this — it is how the host code is exercised without a GPU; test_gpu_segments.py runs the
same flow on the device."""
import ctypes as C

import numpy as np

import synth


class _StandInLib:
    """The C-ABI entry points segments.py calls, restated on host arrays (device pointers become dictionary keys)."""

    def __init__(self, oracle, usage_mask):
        self.oracle, self.usage_mask, self.bufs, self.next = oracle, usage_mask, {}, 0x1000

    def zkl_dev_alloc(self, h, nbytes, out):
        self.next += 0x1000000
        out._obj.value = self.next
        return 0

    def zkl_dev_upload(self, h, dev, host, nbytes):
        self.bufs[dev.value] = np.ctypeslib.as_array((C.c_uint64 * (nbytes // 8)).from_address(host)).copy()
        return 0

    def zkl_dev_free(self, h, dev):
        self.bufs.pop(dev.value, None)
        return 0

    def zkl_slice_segment_dev(self, h, full, full_rows, map_ptr, W, r0, n, out):
        m = np.ctypeslib.as_array((C.c_uint32 * W).from_address(map_ptr)).astype(np.int64)
        self.bufs[out.value] = np.ascontiguousarray(self.bufs[full.value].reshape(-1, full_rows, 2)[m, r0:r0 + n])
        return 0

    def zkl_read_cells_dev(self, h, dev, n_rows, cols, rows, k, out):
        c = np.ctypeslib.as_array((C.c_uint32 * k).from_address(cols)).astype(np.int64)
        r = np.ctypeslib.as_array((C.c_uint32 * k).from_address(rows)).astype(np.int64)
        np.ctypeslib.as_array((C.c_uint64 * (2 * k)).from_address(out)).reshape(k, 2)[:] = self.bufs[dev.value].reshape(-1, n_rows, 2)[c, r]
        return 0

    def zkl_usage_mask_dev(self, h, dev, W, n, cols, ram, mask, bits):
        mask._obj.value, bits._obj.value = self.usage_mask, 0
        return 0

    def zkl_prove_segment_dev(self, h, dev, W, n, air, seed, num_seed, opts, buf, cap, summ):
        s = np.ctypeslib.as_array((C.c_uint64 * (2 * num_seed)).from_address(seed)).reshape(num_seed, 2).copy()
        p = self.oracle.prove_segment(self.bufs[dev.value].reshape(W, n, 2), air._obj, s, opts._obj, summ._obj)
        C.memmove(buf, p, len(p))
        cap._obj.value = len(p)
        return 0


    def zkl_prove_segment_pi_dev(self, h, dev, W, n, program, pi, opts, buf, cap, summ):
        """The AIR instance comes from the library's own host-side builder (zkl_air_build needs no device), the proof from the checker."""
        from zk_lisp_b200 import capi
        from zk_lisp_b200.prover import NativeAir
        from zk_lisp_b200.air import ProofOptions
        o = opts._obj
        po = ProofOptions(o.num_queries, o.blowup, o.grinding, o.num_partitions, o.hash_rate, o.fri_folding, o.fri_remainder_max_degree)
        a = NativeAir(program._obj, pi._obj, W, n, po)
        p = self.oracle.prove_segment(self.bufs[dev.value].reshape(W, n, 2), a.params, np.ascontiguousarray(a.seed), o, summ._obj)
        C.memmove(buf, p, len(p))
        cap._obj.value = len(p)
        return 0


class _StandInContext:
    def __init__(self, oracle, usage_mask, device=0):
        self.lib, self.h, self.device = _StandInLib(oracle, usage_mask), None, device
        self.hash_elements, self.merge = oracle.hash_elements, oracle.merge
        self.merge_with_int, self.merge_many = oracle.merge_with_int, oracle.merge_many

    def _check(self, rc):
        assert rc == 0


def test_prove_program_host_logic(zkl, oracle):
    A, SG, ST, R = zkl.air, zkl.segments, zkl.step, zkl.replay
    n_full, seg_rows = 256, 64
    full, air_pi_full, opts = synth.valid_segment(n=n_full, queries=8, blowup=16, grind=4, seed=21, full_layout=True)
    core = air_pi_full.core
    segs = [SG.Segment(s.r_start, s.r_end) for s in zkl.planner.plan_segments(n_full // 32, max_segment_rows=seg_rows)]
    assert segs == SG.plan_fixed_segments(n_full, seg_rows)
    feats = [SG.SegmentFeatures(vm=True, sponge=True)] * len(segs)
    ctxs = [_StandInContext(oracle, air_pi_full.vm_usage_mask, device=d) for d in (0, 1)]      # two "devices": one upload each
    steps = SG.prove_program(ctxs, full, segs, core, feats, opts, rom_acc=(5, 6, 7), min_bits=40)
    assert [s.segment_index for s in steps] == list(range(len(segs))) and len(segs) == 4
    cols = A.Columns.baseline()
    cell = lambda c, r: int(full[c, r, 0]) | (int(full[c, r, 1]) << 64)
    for s, seg in zip(steps, segs):
        assert (s.width, s.rows) == (204, seg_rows) and s.air_pi.rom_acc == (5, 6, 7)
        a = A.Air(s.width, s.rows, s.air_pi, s.options)
        rc, msg = oracle.verify_segment(s.proof, s.width, s.rows, a.to_c(), zkl.prover.ints_to_fe(a.seed_elements()), s.options.to_c())
        assert rc == 0, (s.segment_index, msg)
        c = s.step
        assert c.rom_acc == (5, 6, 7) and c.proof.meta.lambda_bits == 40 and c.proof.meta.m == seg_rows
        assert c.proof.pi.pc_init == ST.fe_to_bytes_fold(cell(cols.pc, seg.r_start))
        assert c.proof.pi.rom_s_in_2 == ST.fe_to_bytes_fold(cell(cols.rom_s_index(2), seg.r_start))
        assert c.proof.pi.rom_s_out_0 == ST.fe_to_bytes_fold(cell(cols.rom_s_index(0), seg.r_end - 4))
        assert c.proof.pi.ram_gp_sorted_out == ST.fe_to_bytes_fold(cell(cols.ram_start + 7, seg.r_end - 1))
        assert c.state_in_hash() == ST.vm_state_hash([cell(cols.r_index(i), seg.r_start) for i in range(8)])
        assert c.state_out_hash() == ST.vm_state_hash([cell(cols.r_index(i), seg.r_end - 1) for i in range(8)])
        back = ST.StepProof.from_bytes(c.to_bytes())
        assert back.digest() == c.digest() and back.proof.pi == c.proof.pi and back.proof.meta == c.proof.meta
        fs = R.replay_fs_from_step(ctxs[0], back)
        R.verify_child_openings(ctxs[0], back, fs)
        assert zkl.verify.verify_step(ctxs[0], back, 16) == fs          # the whole host verifier on the ZKLSTP1 bytes of a driver-made step
    ST.verify_chain([s.step for s in steps])
    assert len({s.step.digest() for s in steps}) == len(steps)
    # a failing worker surfaces its error to the caller (prove.rs: Error::Backend)
    ctxs[1].lib.zkl_prove_segment_pi_dev = lambda *a: 1
    try:
        SG.prove_program(ctxs, full, segs, core, feats, opts)
        raise RuntimeError("expected a failure")
    except AssertionError:
        pass
    try:
        SG.prove_program(ctxs, full, [], core, [], opts)
        raise RuntimeError("expected a failure")
    except ValueError as e:
        assert "no segments" in str(e)


def test_apply_levels_absorb_marshalling(zkl):
    """segments.apply_levels_absorb: inputs are zero padded to 10 per level, the PROGRAM suite (suite_id = program_id) is the one
    handed to the C ABI; checked by restating zkl_poseidon_levels_dev in Python behind the stand-in library."""
    SG, P = zkl.segments, zkl.field.P
    seen = {}

    class Lib:
        def zkl_poseidon_levels_dev(self, h, trace_dev, n_rows, lanes_start, suite, levels, inputs, count):
            lv = np.ctypeslib.as_array((C.c_uint32 * count).from_address(levels))
            ins = np.ctypeslib.as_array((C.c_uint64 * (20 * count)).from_address(inputs)).reshape(count, 10, 2)
            s = suite._obj
            seen.update(levels=list(map(int, lv)), n_rows=n_rows, lanes_start=lanes_start, dom=[s.dom[0].int(), s.dom[1].int()],
                        mds00=s.mds[0][0].int(), rc_last=s.rc[26][11].int(),
                        inputs=[[int(a) | (int(b) << 64) for a, b in row] for row in ins])
            return 0

    class Ctx:
        lib, h = Lib(), None

        def _check(self, rc):
            assert rc == 0

    sid = bytes([3]) * 32
    SG.apply_levels_absorb(Ctx(), 0xBEEF, 256, sid, [0, 5], [[1, 2], list(range(100, 112))], lanes_start=0)
    suite = zkl.params.get_poseidon_suite(sid)
    assert seen["levels"] == [0, 5] and seen["n_rows"] == 256 and seen["dom"] == list(suite.dom)
    assert seen["mds00"] == suite.mds[0][0] and seen["rc_last"] == suite.rc[26][11]
    assert seen["inputs"] == [[1, 2] + [0] * 8, list(range(100, 110))]          # padded to 10, extra inputs dropped (poseidon.rs:38-49)
    assert suite.dom != zkl.params.get_poseidon_suite(bytes(32)).dom              # not the commitment hasher's suite
    seen.clear()
    SG.apply_levels_absorb(Ctx(), 0, 32, sid, [], [])
    assert not seen
    try:
        SG.apply_levels_absorb(Ctx(), 0, 32, sid, [0], [])
        raise RuntimeError("expected a failure")
    except ValueError:
        pass
    assert all(0 <= x < P for x in seen.get("inputs", [[0]])[0])
