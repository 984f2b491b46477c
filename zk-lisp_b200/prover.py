"""Host-side mirror of the reference's prover interface over the C ABI (include/zkl_b200.h).

`ZkProver.prove(trace)` mirrors zk-lisp-proof-winterfell/src/prove.rs:174-242 (ZkProver::prove ->
winterfell::Prover::prove); the three hook-level methods mirror prove.rs:454-516. All arithmetic
happens on the GPU; this module only marshals buffers (numpy) and builds plain parameter structs.
"""
import ctypes as C
import threading

import numpy as np

from . import air as air_mod
from . import capi, params


def hasher_suite():
    """Commitment/coin suite, suite_id = [0;32] (poseidon/hasher.rs:23), Python-integer derivation (params.py)."""
    return params.get_poseidon_suite(params.HASHER_SUITE_ID)


_HASHER_SUITE_C = None
_PROGRAMS = {}
_PROGRAMS_LOCK = threading.Lock()


def hasher_suite_c() -> capi.Suite:
    """The same suite derived in native code (`zkl_suite_derive`); computed once per process."""
    global _HASHER_SUITE_C
    if _HASHER_SUITE_C is None:
        s = capi.Suite()
        if capi.load_library().zkl_suite_derive((C.c_uint8 * 32)(), C.byref(s)):
            raise capi.ZklError("zkl_suite_derive failed")
        _HASHER_SUITE_C = s
    return _HASHER_SUITE_C


def program_constants(program_id: bytes, program_commitment: bytes, cache: bool = True) -> capi.Program:
    """`zkl_program_derive`: the per-program constants ZkLispAir::new needs (vm/air/mod.rs:129-176). One derivation per
    program, shared by all of its segments (the reference keeps the suites in a process-global cache, poseidon/mod.rs:47)."""
    key = (bytes(program_id), bytes(program_commitment))
    if cache:
        with _PROGRAMS_LOCK:
            if key in _PROGRAMS:
                return _PROGRAMS[key]
    prog = capi.Program()
    a = (C.c_uint8 * 32).from_buffer_copy(key[0])
    b = (C.c_uint8 * 32).from_buffer_copy(key[1])
    if capi.load_library().zkl_program_derive(a, b, C.byref(prog)):
        raise capi.ZklError("zkl_program_derive failed (no valid 12x12 MDS for this program id)")
    if cache:
        with _PROGRAMS_LOCK:
            _PROGRAMS[key] = prog
    return prog


def public_inputs_to_c(air_pi: "air_mod.AirPublicInputs") -> capi.PublicInputsC:
    """AirPublicInputs (lib.rs:74-96) as the plain struct the C ABI takes."""
    core = air_pi.core
    p = capi.PublicInputsC()
    p.feature_mask, p.segment_feature_mask = core.feature_mask, air_pi.segment_feature_mask
    p.merkle_root[:] = bytes(core.merkle_root)
    p.vm_expected_bytes[:] = bytes(core.vm_expected_bytes)
    p.vm_out_reg, p.vm_out_row = core.vm_out_reg, core.vm_out_row
    slots = air_mod.encode_main_args_to_slots(core.main_args)
    if len(slots) > capi.MAX_MAIN_SLOTS:
        raise ValueError("too many main_args slots")
    p.num_main_slots = len(slots)
    for j, v in enumerate(slots):
        p.main_slots[j] = capi.Fe.of(v)
    p.vm_usage_mask, p.ram_delta_clk_bits = air_pi.vm_usage_mask, air_pi.ram_delta_clk_bits
    p.pc_init = capi.Fe.of(air_pi.pc_init)
    p.ram_gp_unsorted_in, p.ram_gp_unsorted_out = capi.Fe.of(air_pi.ram_gp_unsorted_in), capi.Fe.of(air_pi.ram_gp_unsorted_out)
    p.ram_gp_sorted_in, p.ram_gp_sorted_out = capi.Fe.of(air_pi.ram_gp_sorted_in), capi.Fe.of(air_pi.ram_gp_sorted_out)
    for i in range(3):
        p.rom_s_in[i], p.rom_s_out[i] = capi.Fe.of(air_pi.rom_s_in[i]), capi.Fe.of(air_pi.rom_s_out[i])
    return p


class NativeAir:
    """`zkl_air_build`: the data ZkLispAir::new derives, built by the library from the public inputs."""

    def __init__(self, program: capi.Program, pi: capi.PublicInputsC, width: int, n: int, options: "air_mod.ProofOptions"):
        lib = capi.load_library()
        cap_a = (n // 32) * 141 + 64
        self.assertions = (capi.Assertion * cap_a)()
        self.seed = np.zeros((64, 2), dtype=np.uint64)
        self.params = capi.AirParams()
        na, ns = C.c_size_t(cap_a), C.c_size_t(64)
        err = C.create_string_buffer(256)
        o = options.to_c()
        if lib.zkl_air_build(C.byref(program), C.byref(pi), width, n, C.byref(o), C.byref(self.params), self.assertions, C.byref(na),
                             self.seed.ctypes.data, C.byref(ns), err, 256):
            raise ValueError(err.value.decode())
        self.num_assertions, self.seed = na.value, self.seed[:ns.value]
        self.num_transition, self.ce_blowup = self.params.num_transition, self.params.ce_blowup
        self.num_composition_cols = self.params.num_composition_cols
        self.num_constraints = self.num_transition + self.num_assertions


def as_fe_array(a) -> np.ndarray:
    """Field elements as a C-contiguous uint64 array with a trailing axis of 2 (lo, hi)."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 2
    return a


def ints_to_fe(xs) -> np.ndarray:
    out = np.empty((len(xs), 2), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i, 0] = x & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = x >> 64
    return out


def fe_to_ints(a: np.ndarray):
    a = a.reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


class Context:
    """One CUDA stream pinned to one device (SURVEY.md §8b thread rule: one ctx per host thread)."""

    def __init__(self, device: int = 0, load_suite: bool = True):
        self.lib = capi.load_library()
        h = C.c_void_p()
        if self.lib.zkl_ctx_create(device, C.byref(h)):
            raise capi.ZklError(self.lib.zkl_last_error(None).decode())
        self.h = h
        self.device = device
        self._lock = threading.Lock()
        if load_suite:
            self._check(self.lib.zkl_suite_load(self.h, C.byref(hasher_suite_c())))

    def _check(self, rc):
        if rc:
            raise capi.ZklError(self.lib.zkl_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.zkl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.zkl_launch_count(self.h))

    # ---- hasher primitives (poseidon/hasher.rs:57-140) ----------------------------------------
    def permute(self, states: np.ndarray) -> np.ndarray:
        st = as_fe_array(states).copy()
        self._check(self.lib.zkl_permute_batch(self.h, st.ctypes.data, st.shape[0]))
        return st

    def hash_elements(self, elems: np.ndarray) -> np.ndarray:
        e = as_fe_array(elems)
        count, length = e.shape[0], e.shape[1]
        out = np.zeros((count, 32), dtype=np.uint8)
        self._check(self.lib.zkl_hash_elements_batch(self.h, e.ctypes.data, length, count, out.ctypes.data))
        return out

    def merge(self, pairs: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(pairs, dtype=np.uint8).reshape(-1, 64)
        out = np.zeros((p.shape[0], 32), dtype=np.uint8)
        self._check(self.lib.zkl_merge_batch(self.h, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def merge_with_int(self, seed: bytes, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        out = np.zeros((v.shape[0], 32), dtype=np.uint8)
        self._check(self.lib.zkl_merge_with_int_batch(self.h, s.ctypes.data, v.ctypes.data, v.shape[0], out.ctypes.data))
        return out

    def merge_with_int_pairs(self, seeds: np.ndarray, values) -> np.ndarray:
        """out[i] = merge_with_int(seeds[i], values[i]): one seed per value (`zkl_merge_with_int_pairs`)."""
        s = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 32)
        v = np.ascontiguousarray(values, dtype=np.uint64)
        assert s.shape[0] == v.shape[0]
        out = np.zeros((v.shape[0], 32), dtype=np.uint8)
        self._check(self.lib.zkl_merge_with_int_pairs(self.h, s.ctypes.data, v.ctypes.data, v.shape[0], out.ctypes.data))
        return out

    def grind(self, seed: bytes, bits: int) -> int:
        """ProverChannel::grind: the smallest nonce >= 1 whose merge_with_int(seed, nonce) has `bits` trailing zero bits."""
        s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        nonce = C.c_uint64()
        self._check(self.lib.zkl_grind(self.h, s.ctypes.data, bits, C.byref(nonce)))
        return int(nonce.value)

    def hash_bytes(self, data: bytes) -> bytes:
        d = np.frombuffer(bytes(data), dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
        out = np.zeros(32, dtype=np.uint8)
        self._check(self.lib.zkl_hash_bytes(self.h, d.ctypes.data, len(data), out.ctypes.data))
        return out.tobytes()

    def merge_many(self, digests: np.ndarray) -> bytes:
        d = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros(32, dtype=np.uint8)
        self._check(self.lib.zkl_merge_many(self.h, d.ctypes.data if d.size else None, d.shape[0], out.ctypes.data))
        return out.tobytes()

    def merge_plan(self, known: np.ndarray, left, right, level_counts, want) -> np.ndarray:
        """Batched Merkle authentication (`zkl_merge_plan`): a digest table seeded with `known` (k, 32) grows level by level,
        entry = merge(table[left], table[right]); returns the `want`ed entries as (len(want), 32)."""
        kn = np.ascontiguousarray(known, dtype=np.uint8).reshape(-1, 32)
        lf, rt = np.ascontiguousarray(left, dtype=np.uint32), np.ascontiguousarray(right, dtype=np.uint32)
        lc, wt = np.ascontiguousarray(level_counts, dtype=np.uint32), np.ascontiguousarray(want, dtype=np.uint32)
        assert lf.shape == rt.shape and int(lc.sum()) == lf.shape[0]
        out = np.zeros((wt.shape[0], 32), dtype=np.uint8)
        self._check(self.lib.zkl_merge_plan(self.h, kn.ctypes.data, kn.shape[0], lf.ctypes.data, rt.ctypes.data, lc.ctypes.data,
                                            lc.shape[0], wt.ctypes.data, wt.shape[0], out.ctypes.data))
        return out

    def merkle_build(self, leaves: np.ndarray):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
        root = np.zeros(32, dtype=np.uint8)
        nodes = np.zeros((lv.shape[0], 32), dtype=np.uint8)
        self._check(self.lib.zkl_merkle_build(self.h, lv.ctypes.data, lv.shape[0], root.ctypes.data, nodes.ctypes.data))
        return root.tobytes(), nodes

    # ---- NTT primitives ---------------------------------------------------------------------------
    def interpolate_columns(self, cols: np.ndarray) -> np.ndarray:
        c = as_fe_array(cols).copy()
        self._check(self.lib.zkl_interpolate_columns(self.h, c.ctypes.data, c.shape[0], c.shape[1]))
        return c

    def evaluate_columns(self, polys: np.ndarray, blowup: int, offset: int = 3) -> np.ndarray:
        p = as_fe_array(polys)
        out = np.zeros((p.shape[0], p.shape[1] * blowup, 2), dtype=np.uint64)
        self._check(self.lib.zkl_evaluate_columns(self.h, p.ctypes.data, p.shape[0], p.shape[1], blowup,
                                                  capi.Fe.of(offset), out.ctypes.data))
        return out


class TraceLde:
    """Replacement for winterfell's DefaultTraceLde returned by `new_trace_lde` (prove.rs:432-433)."""

    def __init__(self, ctx: Context, handle, width, n, blowup, root: bytes, polys: np.ndarray):
        self.ctx, self.h, self.width, self.n, self.blowup = ctx, handle, width, n, blowup
        self.root, self.polys = root, polys

    def trace_len(self):
        return self.n

    def get_main_trace_commitment(self) -> bytes:
        return self.root

    def read_main_trace_frame(self, lde_step: int):
        cur = np.zeros((self.width, 2), dtype=np.uint64)
        nxt = np.zeros((self.width, 2), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkl_trace_lde_read_frame(self.ctx.h, self.h, lde_step, cur.ctypes.data, nxt.ctypes.data))
        return cur, nxt

    def query(self, positions):
        return _query(self.ctx, self.ctx.lib.zkl_trace_query, self.h, positions, self.width)

    def free(self):
        if self.h:
            self.ctx.lib.zkl_trace_lde_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ConstraintCommitment:
    """Replacement for DefaultConstraintCommitment (prove.rs:436-437)."""

    def __init__(self, ctx, handle, num_cols, root, poly):
        self.ctx, self.h, self.num_cols, self.root, self.poly = ctx, handle, num_cols, root, poly

    def commitment(self) -> bytes:
        return self.root

    def query(self, positions):
        return _query(self.ctx, self.ctx.lib.zkl_constraint_query, self.h, positions, self.num_cols)

    def free(self):
        if self.h:
            self.ctx.lib.zkl_constraint_commitment_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _query(ctx, fn, handle, positions, width):
    pos = np.ascontiguousarray(positions, dtype=np.uint32)
    k = pos.shape[0]
    rows = np.zeros((k, width, 2), dtype=np.uint64)
    list_len = np.zeros(max(k, 1), dtype=np.uint32)
    nodes = np.zeros((max(k, 1) * 64, 32), dtype=np.uint8)
    bp = capi.BatchProof()
    bp.list_len = list_len.ctypes.data_as(C.POINTER(C.c_uint32))
    bp.nodes = nodes.ctypes.data_as(C.POINTER(C.c_uint8))
    ctx._check(fn(ctx.h, handle, pos.ctypes.data, k, rows.ctypes.data, C.byref(bp)))
    lists, off = [], 0
    for i in range(bp.num_lists):
        ln = int(list_len[i])
        lists.append([nodes[off + j].tobytes() for j in range(ln)])
        off += ln
    return rows, {"depth": int(bp.depth), "nodes": lists}


class ZkProver:
    """Mirror of `ZkProver` / `ZkWinterfellProver` (prove.rs:104-242, 425-517) over the GPU path."""

    def __init__(self, ctx: Context, options: air_mod.ProofOptions, air_pi: air_mod.AirPublicInputs, cache_program: bool = True):
        """`cache_program=False` re-derives the per-program constants on every call instead of sharing them across the
        program's segments (measurement of a first-ever prove, bench.py `e2e_cold`)."""
        self.ctx, self.options, self.air_pi = ctx, options, air_pi
        self.last_summary = None
        self.preflight = "off"
        self._buf = None
        self._cache_program = cache_program

    def with_preflight_mode(self, mode: str) -> "ZkProver":
        """prove.rs:160-163: "off" | "console" | "json" (zk-lisp-proof/src/frontend.rs:23-27)."""
        if mode not in ("off", "console", "json"):
            raise ValueError("preflight mode must be off, console or json")
        self.preflight = mode
        return self

    def _program(self) -> capi.Program:
        core = self.air_pi.core
        return program_constants(core.program_id, core.program_commitment, cache=self._cache_program)

    def native_air(self, width, n) -> NativeAir:
        """`ZkLispAir::new` for this trace shape through `zkl_air_build` (hook-level calls; `prove` builds it inside the library)."""
        return NativeAir(self._program(), public_inputs_to_c(self.air_pi), width, n, self.options)

    # hook 1 (prove.rs:454-474)
    def new_trace_lde(self, trace: np.ndarray):
        t = as_fe_array(trace)
        width, n = t.shape[0], t.shape[1]
        root = np.zeros(32, dtype=np.uint8)
        polys = np.zeros((width, n, 2), dtype=np.uint64)
        h = C.c_void_p()
        o = self.options.to_c()
        self.ctx._check(self.ctx.lib.zkl_trace_lde_commit(self.ctx.h, t.ctypes.data, width, n, C.byref(o), root.ctypes.data,
                                                          polys.ctypes.data, C.byref(h)))
        return TraceLde(self.ctx, h, width, n, self.options.blowup, root.tobytes(), polys), polys

    # hook 2 (prove.rs:476-492) + ConstraintEvaluator::evaluate
    def evaluate_constraints(self, trace_lde: TraceLde, coeffs: np.ndarray) -> np.ndarray:
        a = self.native_air(trace_lde.width, trace_lde.n)
        cc = as_fe_array(coeffs)
        assert cc.shape[0] == a.num_constraints
        out = np.zeros((trace_lde.n * a.ce_blowup, 2), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkl_eval_constraints(self.ctx.h, trace_lde.h, C.byref(a.params), cc.ctypes.data, out.ctypes.data))
        return out

    # hook 3 (prove.rs:494-516)
    def build_constraint_commitment(self, ce_evals: np.ndarray, n: int, ce_blowup: int, num_cols: int):
        e = as_fe_array(ce_evals)
        root = np.zeros(32, dtype=np.uint8)
        poly = np.zeros((num_cols, n, 2), dtype=np.uint64)
        h = C.c_void_p()
        o = self.options.to_c()
        self.ctx._check(self.ctx.lib.zkl_constraint_commit(self.ctx.h, e.ctypes.data, n, ce_blowup, num_cols, C.byref(o),
                                                           root.ctypes.data, poly.ctypes.data, C.byref(h)))
        return ConstraintCommitment(self.ctx, h, num_cols, root.tobytes(), poly), poly

    # whole proof (prove.rs:174-242)
    def prove(self, trace, dev_ptr=None, shape=None, allow_invalid: bool = False) -> bytes:
        """`trace`: (W, n, 2) uint64 host array, column-major like winterfell's ColMatrix; or pass
        `dev_ptr` + `shape=(W, n)` for a trace already resident in device memory. The AIR instance and the coin seed are
        built inside the library from the public inputs, per call (`zkl_prove_segment_pi`). A trace that violates the AIR
        raises (the reference's debug builds panic in Winterfell's degree check) unless `allow_invalid`."""
        if dev_ptr is None:
            t = as_fe_array(trace)
            width, n = t.shape[0], t.shape[1]
        else:
            width, n = shape
        if self.preflight != "off":                  # prove.rs:186-211: check the trace against the AIR first
            if dev_ptr is not None:
                raise ValueError("preflight walks the host trace: pass the trace array, not a device pointer")
            from . import preflight
            preflight.run(self.preflight, air_mod.Air(width, n, self.air_pi, self.options), t)
        prog, pi, o = self._program(), public_inputs_to_c(self.air_pi), self.options.to_c()
        if self._buf is None:
            self._buf = np.empty(1 << 20, dtype=np.uint8)   # reused across calls; grown when the library asks for more
        summ = capi.ProofSummary()
        lib = self.ctx.lib
        for _ in range(2):
            cap = C.c_size_t(self._buf.shape[0])
            if dev_ptr is None:
                rc = lib.zkl_prove_segment_pi(self.ctx.h, t.ctypes.data, width, n, C.byref(prog), C.byref(pi), C.byref(o),
                                              self._buf.ctypes.data, C.byref(cap), C.byref(summ))
            else:
                rc = lib.zkl_prove_segment_pi_dev(self.ctx.h, dev_ptr, width, n, C.byref(prog), C.byref(pi), C.byref(o),
                                                  self._buf.ctypes.data, C.byref(cap), C.byref(summ))
            if rc != capi.ERR_BUFFER_TOO_SMALL:
                break
            self._buf = np.empty(cap.value + (cap.value >> 2), dtype=np.uint8)
        self.last_summary = summ
        if rc == capi.ERR_CONSTRAINTS_VIOLATED and allow_invalid:
            rc = 0
        self.ctx._check(rc)
        return self._buf[:cap.value].tobytes()
