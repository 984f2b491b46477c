"""ctypes mirror of include/zkl_b200.h and the loader of the CUDA library.

There is NO CPU fallback: if libzkl_b200.so is missing or no CUDA device is present, loading /
context creation raises. The CPU checker used by tests/, smoke() and bench.py's cpu_baseline legs lives
outside this package and is never loaded from here.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libzkl_b200.so")

ROUNDS, NR, NUM_OPS = 27, 8, 17


class Fe(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]

    @staticmethod
    def of(x: int) -> "Fe":
        return Fe(x & 0xFFFFFFFFFFFFFFFF, x >> 64)

    def int(self) -> int:
        return self.lo | (self.hi << 64)


class Suite(C.Structure):
    _fields_ = [("dom", Fe * 2), ("mds", (Fe * 12) * 12), ("rc", (Fe * 12) * ROUNDS)]


class RomSuite(C.Structure):
    _fields_ = [("mds", (Fe * 3) * 3), ("rc", (Fe * 3) * ROUNDS)]


class Assertion(C.Structure):
    _fields_ = [("column", C.c_uint32), ("step", C.c_uint32), ("value", Fe)]


COLUMN_FIELDS = [
    "width", "lanes_start", "g_map", "g_final", "g_r_start", "mask", "r_start", "op_start",
    "sel_dst0_start", "sel_a_start", "sel_b_start", "sel_c_start", "sel_dst1_start",
    "sel_s_bits_start", "sel_s_active_start", "imm", "eq_inv", "ram_start", "merkle_start",
    "pi_prog", "pc", "rom_op_start", "pose_active", "gadget_b_start", "rom_s_start",
]


class Columns(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in COLUMN_FIELDS]


class AirParams(C.Structure):
    _fields_ = [
        ("cols", Columns),
        ("f_poseidon", C.c_uint32), ("f_vm", C.c_uint32), ("f_sponge", C.c_uint32),
        ("f_merkle", C.c_uint32), ("f_ram", C.c_uint32), ("rom_enabled", C.c_uint32),
        ("vm_usage_mask", C.c_uint32), ("ram_delta_clk_bits", C.c_uint32),
        ("num_transition", C.c_uint32), ("ce_blowup", C.c_uint32),
        ("num_composition_cols", C.c_uint32), ("num_assertions", C.c_uint32),
        ("assertions", C.POINTER(Assertion)),
        ("pose", Suite), ("rom", RomSuite),
        ("rom_w_enc0", Fe * 59), ("rom_w_enc1", Fe * 59),
        ("program_fe0", Fe), ("merkle_root_fe", Fe), ("air_kind", C.c_uint32), ("pad_", C.c_uint32),
    ]


AIR_ZKLISP, AIR_AGG, AGG_WIDTH, AGG_TRANSITIONS = 0, 1, 31, 24


class Options(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_queries", "blowup", "grinding", "num_partitions", "hash_rate", "fri_folding",
        "fri_remainder_max_degree", "field_extension")]


class BatchProof(C.Structure):
    _fields_ = [("depth", C.c_uint8), ("num_lists", C.c_uint32), ("list_len", C.POINTER(C.c_uint32)),
                ("nodes", C.POINTER(C.c_uint8)), ("total_nodes", C.c_uint32)]


class ProofSummary(C.Structure):
    _fields_ = [
        ("trace_root", C.c_uint8 * 32), ("constraint_root", C.c_uint8 * 32),
        ("fri_roots", C.c_uint8 * (32 * 40)), ("num_fri_roots", C.c_uint32),
        ("z", Fe), ("pow_nonce", C.c_uint64), ("num_unique_queries", C.c_uint32),
        ("positions", C.c_uint32 * 256), ("stage_ms", C.c_float * 12), ("constraints_violated", C.c_uint32),
    ]


ERR_BUFFER_TOO_SMALL, ERR_CONSTRAINTS_VIOLATED = 2, 3

COMM_MOVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)      # all_gather / all_to_all(user, send, recv, bytes)
COMM_SUM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)                   # all_reduce_sum_u64(user, buf, count)


class Comm(C.Structure):
    """zkl_comm: the caller's collectives for the column-sharded prover."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("user", C.c_void_p), ("all_gather", COMM_MOVE), ("all_to_all", COMM_MOVE),
                ("all_reduce_sum_u64", COMM_SUM)]
MAX_MAIN_SLOTS = 16


class Program(C.Structure):
    """zkl_program: per-program constants (zkl_program_derive)."""
    _fields_ = [("program_id", C.c_uint8 * 32), ("program_commitment", C.c_uint8 * 32), ("pose", Suite), ("rom", RomSuite),
                ("rom_w_enc0", Fe * 59), ("rom_w_enc1", Fe * 59), ("program_fe", Fe * 2), ("commitment_fe", Fe * 2)]


class PublicInputsC(C.Structure):
    """zkl_public_inputs: AirPublicInputs (lib.rs:74-96) + the core fields the AIR reads."""
    _fields_ = [("feature_mask", C.c_uint64), ("segment_feature_mask", C.c_uint64), ("merkle_root", C.c_uint8 * 32),
                ("vm_expected_bytes", C.c_uint8 * 32), ("vm_out_reg", C.c_uint32), ("vm_out_row", C.c_uint32),
                ("num_main_slots", C.c_uint32), ("vm_usage_mask", C.c_uint32), ("ram_delta_clk_bits", C.c_uint32), ("pad_", C.c_uint32),
                ("main_slots", Fe * MAX_MAIN_SLOTS), ("pc_init", Fe), ("ram_gp_unsorted_in", Fe), ("ram_gp_unsorted_out", Fe),
                ("ram_gp_sorted_in", Fe), ("ram_gp_sorted_out", Fe), ("rom_s_in", Fe * 3), ("rom_s_out", Fe * 3)]


class ChildProof(C.Structure):
    _fields_ = [("proof", C.c_void_p), ("proof_len", C.c_size_t), ("pi_elems", C.c_void_p), ("num_pi", C.c_size_t)]


class ChildChallenges(C.Structure):
    _fields_ = [("status", C.c_int32), ("num_positions", C.c_uint32), ("num_alphas", C.c_uint32), ("num_deep", C.c_uint32),
                ("pow_nonce", C.c_uint64), ("z", Fe), ("alphas", Fe * 40), ("positions", C.c_uint32 * 256)]


def suite_to_c(suite) -> Suite:
    s = Suite()
    for i in range(2):
        s.dom[i] = Fe.of(suite.dom[i])
    for i in range(12):
        for j in range(12):
            s.mds[i][j] = Fe.of(suite.mds[i][j])
    for r in range(ROUNDS):
        for j in range(12):
            s.rc[r][j] = Fe.of(suite.rc[r][j])
    return s


def rom_suite_to_c(mds, rc) -> RomSuite:
    s = RomSuite()
    for i in range(3):
        for j in range(3):
            s.mds[i][j] = Fe.of(mds[i][j])
    for r in range(ROUNDS):
        for j in range(3):
            s.rc[r][j] = Fe.of(rc[r][j])
    return s


_lib = None


class ZklError(RuntimeError):
    pass


def load_library():
    """Load the CUDA shared library; fail loudly when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZklError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback for the proving path.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, u8p, fep = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(Fe)

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype, f.argtypes = res, list(args)

    sig("zkl_ctx_create", C.c_int, C.c_int, C.POINTER(vp))
    sig("zkl_ctx_destroy", None, vp)
    sig("zkl_last_error", C.c_char_p, vp)
    sig("zkl_suite_load", C.c_int, vp, C.POINTER(Suite))
    sig("zkl_launch_count", C.c_uint64, vp)
    sig("zkl_permute_batch", C.c_int, vp, vp, sz)
    sig("zkl_hash_elements_batch", C.c_int, vp, vp, sz, sz, vp)
    sig("zkl_merge_batch", C.c_int, vp, vp, sz, vp)
    sig("zkl_merge_with_int_batch", C.c_int, vp, vp, vp, sz, vp)
    sig("zkl_hash_bytes", C.c_int, vp, vp, sz, vp)
    sig("zkl_merge_many", C.c_int, vp, vp, sz, vp)
    sig("zkl_merkle_build", C.c_int, vp, vp, sz, vp, vp)
    sig("zkl_interpolate_columns", C.c_int, vp, vp, sz, sz)
    sig("zkl_evaluate_columns", C.c_int, vp, vp, sz, sz, sz, Fe, vp)
    sig("zkl_trace_lde_commit", C.c_int, vp, vp, sz, sz, C.POINTER(Options), vp, vp, C.POINTER(vp))
    sig("zkl_trace_lde_free", None, vp, vp)
    sig("zkl_trace_lde_read_frame", C.c_int, vp, vp, sz, vp, vp)
    sig("zkl_trace_query", C.c_int, vp, vp, vp, sz, vp, C.POINTER(BatchProof))
    sig("zkl_eval_constraints", C.c_int, vp, vp, C.POINTER(AirParams), vp, vp)
    sig("zkl_constraint_commit", C.c_int, vp, vp, sz, sz, sz, C.POINTER(Options), vp, vp, C.POINTER(vp))
    sig("zkl_constraint_commitment_free", None, vp, vp)
    sig("zkl_constraint_query", C.c_int, vp, vp, vp, sz, vp, C.POINTER(BatchProof))
    sig("zkl_prove_segment", C.c_int, vp, vp, sz, sz, C.POINTER(AirParams), vp, sz, C.POINTER(Options), vp,
        C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_prove_segment_dev", C.c_int, vp, vp, sz, sz, C.POINTER(AirParams), vp, sz, C.POINTER(Options), vp,
        C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_stage_name", C.c_char_p, C.c_int)
    sig("zkl_program_derive", C.c_int, vp, vp, C.POINTER(Program))
    sig("zkl_suite_derive", C.c_int, vp, C.POINTER(Suite))
    sig("zkl_air_build", C.c_int, C.POINTER(Program), C.POINTER(PublicInputsC), sz, sz, C.POINTER(Options), C.POINTER(AirParams), vp,
        C.POINTER(sz), vp, C.POINTER(sz), C.c_char_p, sz)
    sig("zkl_shard_columns", C.c_int, C.POINTER(Options), sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz))
    sig("zkl_prove_segment_sharded_dev", C.c_int, vp, vp, sz, sz, sz, sz, C.POINTER(AirParams), vp, sz, C.POINTER(Options), C.POINTER(Comm), vp,
        C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_prove_segment_sharded_pi_dev", C.c_int, vp, vp, sz, sz, sz, sz, C.POINTER(Program), C.POINTER(PublicInputsC), C.POINTER(Options),
        C.POINTER(Comm), vp, C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_comm_loopback", C.POINTER(Comm))
    sig("zkl_comm_local_create", C.c_int, C.c_int, C.POINTER(vp))
    sig("zkl_comm_local_rank", C.POINTER(Comm), vp, C.c_int)
    sig("zkl_comm_local_destroy", None, vp)
    sig("zkl_set_shard_min_leaves", C.c_int, vp, sz)
    sig("zkl_agg_air_params", C.c_int, sz, C.c_uint64, C.c_uint32, C.POINTER(AirParams), C.POINTER(Assertion))
    sig("zkl_prove_segment_pi", C.c_int, vp, vp, sz, sz, C.POINTER(Program), C.POINTER(PublicInputsC), C.POINTER(Options), vp,
        C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_prove_segment_pi_dev", C.c_int, vp, vp, sz, sz, C.POINTER(Program), C.POINTER(PublicInputsC), C.POINTER(Options), vp,
        C.POINTER(sz), C.POINTER(ProofSummary))
    sig("zkl_dev_alloc", C.c_int, vp, sz, C.POINTER(vp))
    sig("zkl_dev_free", C.c_int, vp, vp)
    sig("zkl_dev_upload", C.c_int, vp, vp, vp, sz)
    sig("zkl_sync", C.c_int, vp)
    sig("zkl_slice_segment_dev", C.c_int, vp, vp, sz, vp, sz, sz, sz, vp)
    sig("zkl_usage_mask_dev", C.c_int, vp, vp, sz, sz, C.POINTER(Columns), C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
    sig("zkl_read_cells_dev", C.c_int, vp, vp, sz, vp, vp, sz, vp)
    sig("zkl_timer_start", C.c_int, vp)
    sig("zkl_timer_stop", C.c_int, vp, C.POINTER(C.c_float))
    sig("zkl_imad_peak", C.c_int, vp, C.POINTER(C.c_double))
    sig("zkl_tc_i8_peak", C.c_int, vp, C.POINTER(C.c_double))
    sig("zkl_partition_size", sz, C.POINTER(Options), sz)
    sig("zkl_lde_dev", C.c_int, vp, vp, sz, sz, sz, vp, vp)
    sig("zkl_hash_partitions_dev", C.c_int, vp, vp, sz, sz, sz, sz, vp)
    sig("zkl_merge_partition_digests_dev", C.c_int, vp, vp, sz, sz, sz, sz, vp)
    sig("zkl_merkle_dev", C.c_int, vp, vp, sz, vp)
    sig("zkl_dev_download", C.c_int, vp, vp, vp, sz)
    sig("zkl_poseidon_levels_dev", C.c_int, vp, vp, sz, C.c_uint32, C.POINTER(Suite), vp, vp, sz)
    sig("zkl_merge_with_int_pairs", C.c_int, vp, vp, vp, sz, vp)
    sig("zkl_replay_children", C.c_int, vp, C.POINTER(ChildProof), sz, C.c_int, C.POINTER(ChildChallenges), vp, sz)
    sig("zkl_replay_status_name", C.c_char_p, C.c_int)
    sig("zkl_hash_two_lanes_host", C.c_int, C.POINTER(Suite), vp, vp, sz, vp)
    sig("zkl_grind", C.c_int, vp, vp, C.c_uint32, C.POINTER(C.c_uint64))
    sig("zkl_merge_plan", C.c_int, vp, vp, sz, vp, vp, vp, sz, vp, sz, vp)
    sig("zkl_set_bulk_threshold", C.c_int, vp, C.c_long)
    sig("zkl_set_hash_path", C.c_int, vp, C.c_int)
    sig("zkl_set_ntt_path", C.c_int, vp, C.c_int)
    sig("zkl_set_coop_path", C.c_int, vp, C.c_int)
    sig("zkl_bench_permute", C.c_int, vp, C.c_int, sz, C.c_int, C.POINTER(C.c_float))
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "zkl_ctx_create", "zkl_ctx_destroy", "zkl_last_error", "zkl_suite_load", "zkl_launch_count",
    "zkl_permute_batch", "zkl_hash_elements_batch", "zkl_merge_batch", "zkl_merge_with_int_batch",
    "zkl_hash_bytes", "zkl_merge_many", "zkl_merkle_build", "zkl_interpolate_columns",
    "zkl_evaluate_columns", "zkl_trace_lde_commit", "zkl_trace_lde_free", "zkl_trace_lde_read_frame",
    "zkl_trace_query", "zkl_eval_constraints", "zkl_constraint_commit",
    "zkl_constraint_commitment_free", "zkl_constraint_query", "zkl_prove_segment",
    "zkl_prove_segment_dev", "zkl_stage_name", "zkl_dev_alloc", "zkl_dev_free", "zkl_dev_upload",
    "zkl_sync", "zkl_slice_segment_dev", "zkl_usage_mask_dev", "zkl_read_cells_dev", "zkl_timer_start", "zkl_timer_stop", "zkl_imad_peak", "zkl_set_bulk_threshold", "zkl_set_hash_path", "zkl_set_ntt_path", "zkl_set_coop_path", "zkl_bench_permute", "zkl_partition_size", "zkl_lde_dev", "zkl_hash_partitions_dev", "zkl_merge_partition_digests_dev", "zkl_merkle_dev", "zkl_dev_download", "zkl_merge_plan", "zkl_poseidon_levels_dev", "zkl_merge_with_int_pairs", "zkl_replay_children", "zkl_replay_status_name", "zkl_grind", "zkl_hash_two_lanes_host",
    "zkl_program_derive", "zkl_suite_derive", "zkl_air_build", "zkl_prove_segment_pi", "zkl_prove_segment_pi_dev", "zkl_tc_i8_peak", "zkl_agg_air_params",
    "zkl_shard_columns", "zkl_prove_segment_sharded_dev", "zkl_prove_segment_sharded_pi_dev", "zkl_comm_loopback",
    "zkl_comm_local_create", "zkl_comm_local_rank", "zkl_comm_local_destroy", "zkl_set_shard_min_leaves",
]
