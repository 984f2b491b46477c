"""The reference's remaining host-side unit tests on this path, restated on the Python mirror AND on the native library where the
logic lives there too (zkl_air_build / zkl_suite_derive are pure host code of libzkl_b200.so):

    vm/layout.rs:470-533          layout_width_shrinks_when_disabling_blocks, core_indices_are_stable_across_layout_configs
    vm/schedule.rs:87-97          schedule_shapes (the periodic selectors of one level)
    vm/air/schedule.rs:152-199    schedule_strict_boundary (the one-hot schedule columns are asserted to ZERO off their row)
    prove.rs:1415-1447            seg_air_mask_zero_when_equal_to_global, seg_air_mask_nonzero_when_optional_feature_dropped
    poseidon/mod.rs:451-479       mds_12x12_basic, rc_12_len_matches_rounds, suite_determinism_and_cache
(the AIR block tests are in test_air_reference_cases.py, the planner's in test_step_container.py, vm/trace/poseidon.rs's in
test_gpu_segments.py)."""
import ctypes as C

import __graft_entry__ as graft
import synth
from util import P, from_fe

R = 27


def _native_air(zkl, trace, air_pi, opts):
    from zk_lisp_b200 import prover as PV
    graft.build()
    return PV.ZkProver(None, opts, air_pi).native_air(trace.shape[0], trace.shape[1])


def test_layout_width_shrinks_when_disabling_blocks(zkl):
    A = zkl.air
    w_all = A.Columns(A.LayoutConfig(vm=True, ram=True, sponge=True, merkle=True, rom=True)).width
    w_no_rom = A.Columns(A.LayoutConfig(vm=True, ram=True, sponge=True, merkle=True, rom=False)).width
    w_vm_only = A.Columns(A.LayoutConfig(vm=True, ram=False, sponge=False, merkle=False, rom=False)).width
    assert w_no_rom < w_all and w_vm_only <= w_no_rom
    assert (w_all, A.Columns.baseline().width) == (219, 219)
    # the four widths seen in practice, from the native column map as well
    for kw, width in ((dict(), 219), (dict(merkle=False), 212), (dict(ram=False), 211), (dict(ram=False, merkle=False), 204)):
        trace, air_pi, opts = synth.random_segment(n=64, seed=4, usage_mask=0, ram_bits=0, **kw)
        assert trace.shape[0] == width
        assert _native_air(zkl, trace, air_pi, opts).params.cols.width == width


def test_core_indices_are_stable_across_layout_configs(zkl):
    A = zkl.air
    all_, vm = A.Columns(A.LayoutConfig(vm=True, ram=True, sponge=True, merkle=True, rom=True)), A.Columns(A.LayoutConfig(vm=True, ram=False, sponge=False, merkle=False, rom=False))
    assert [all_.r_index(i) for i in range(A.NR)] == [vm.r_index(i) for i in range(A.NR)]
    assert (all_.op_const, all_.op_store) == (vm.op_const, vm.op_store)
    full = _native_air(zkl, *synth.random_segment(n=64, seed=4, usage_mask=0, ram_bits=0)).params.cols
    slim = _native_air(zkl, *synth.random_segment(n=64, seed=4, usage_mask=0, ram_bits=0, ram=False, merkle=False)).params.cols
    assert (full.r_start, full.op_start) == (slim.r_start, slim.op_start) == (all_.r_index(0), all_.op_const)
    assert full.pc != slim.pc                       # dropping the RAM / Merkle blocks moves everything behind them


def test_schedule_shapes_and_strict_boundary(zkl, oracle):
    """One level (n = 32): every periodic selector has one value per row, and the assertion list pins the one-hot schedule columns to
    zero where they are off (g_final = 0 and g_map = 0 at each other's row and on round rows), both in the mirror and natively."""
    n = 32
    trace, air_pi, opts = synth.random_segment(n=n, seed=6, usage_mask=0, ram_bits=0)
    g = zkl.field.root_of_unity(5)
    rows = [from_fe(oracle.periodic_at(n, pow(g, r, P))) for r in range(n)]      # the periodic columns on the trace domain
    assert all(len(r) == 32 for r in rows)
    for k in range(32):
        col = [rows[r][k] for r in range(n)]
        assert len(col) == n and set(col) <= {0, 1}
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    nat = _native_air(zkl, trace, air_pi, opts)
    native = {(nat.assertions[i].column, nat.assertions[i].step): nat.assertions[i].value.int() for i in range(nat.num_assertions)}
    mirror = {(col, step): val for col, step, val in a.assertions}
    cols = a.cols
    row_map, row_fin, r1 = 0, 1 + R, 1
    for listing in (mirror, native):
        assert listing[(cols.g_final, row_map)] == 0
        assert listing[(cols.g_map, row_fin)] == 0
        assert listing[(cols.g_map, r1)] == 0 and listing[(cols.g_final, r1)] == 0
        assert listing[(cols.g_map, row_map)] == 1 and listing[(cols.g_final, row_fin)] == 1


def _derive_seg_mask_for_air(zkl, base_mask, seg):
    """prove.rs:1397-1412 (the helper of the two tests) over segments.compute_segment_feature_mask."""
    seg_mask = zkl.segments.compute_segment_feature_mask(base_mask, seg)
    use = seg_mask != 0 and seg_mask != base_mask
    return seg_mask, (seg_mask if use else 0)


def test_seg_air_mask_zero_when_equal_to_global(zkl):
    A, SG = zkl.air, zkl.segments
    base = A.FM_VM | A.FM_VM_EXPECT | A.FM_RAM
    seg_mask, seg_air = _derive_seg_mask_for_air(zkl, base, SG.SegmentFeatures(vm=True, ram=True, sponge=False, merkle=False))
    assert seg_mask == base and seg_air == 0


def test_seg_air_mask_nonzero_when_optional_feature_dropped(zkl):
    A, SG = zkl.air, zkl.segments
    base = A.FM_VM | A.FM_VM_EXPECT | A.FM_RAM | A.FM_SPONGE | A.FM_POSEIDON
    seg_mask, seg_air = _derive_seg_mask_for_air(zkl, base, SG.SegmentFeatures(vm=True, ram=False, sponge=False, merkle=False))
    assert seg_mask != 0 and seg_mask != base and seg_air == seg_mask
    assert seg_air & ~base == 0
    assert seg_air & (A.FM_RAM | A.FM_SPONGE | A.FM_POSEIDON) == 0
    assert seg_air & (A.FM_VM | A.FM_VM_EXPECT) != 0


def test_poseidon_suite_shape_and_determinism(zkl):
    """mds_12x12_basic (sid = [42; 32]), rc_12_len_matches_rounds (sid = [7; 32]), suite_determinism_and_cache (sid = [1; 32]) — on the
    Python derivation and on the native one (`zkl_suite_derive`), which must also agree with each other."""
    lib = zkl.capi.load_library()
    for sid_byte in (42, 7, 1):
        sid = bytes([sid_byte]) * 32
        s = zkl.params.get_poseidon_suite(sid)
        assert len(s.mds) == 12 and len(s.mds[0]) == 12 and any(v != 0 for row in s.mds for v in row)
        assert len(s.rc) == R and s.rc[0][0] != 0 and s.rc[R - 1][11] != 0
        again = zkl.params.get_poseidon_suite(sid)
        assert (again.dom, again.mds, again.rc) == (s.dom, s.mds, s.rc)
        nat = zkl.capi.Suite()
        assert lib.zkl_suite_derive((C.c_uint8 * 32)(*sid), C.byref(nat)) == 0
        assert [nat.dom[i].int() for i in range(2)] == list(s.dom)
        assert [[nat.mds[i][j].int() for j in range(12)] for i in range(12)] == [list(r) for r in s.mds]
        assert [[nat.rc[r][j].int() for j in range(12)] for r in range(R)] == [list(r) for r in s.rc]
