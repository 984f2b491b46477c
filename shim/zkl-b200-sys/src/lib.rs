//! `extern "C"` view of `include/zkl_b200.h`. Field elements are winter-math f128 `BaseElement`s: a canonical `u128` in
//! little-endian memory order, so `&[BaseElement]` is passed as `*const ZklFe` without conversion on little-endian hosts.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const ZKL_ROUNDS: usize = 27;
pub const ZKL_MAX_MAIN_SLOTS: usize = 16;
pub const ZKL_ERR_BUFFER_TOO_SMALL: c_int = 2;
pub const ZKL_ERR_CONSTRAINTS_VIOLATED: c_int = 3;

#[repr(C)]
#[derive(Clone, Copy, Default, Debug, PartialEq, Eq)]
pub struct ZklFe {
    pub lo: u64,
    pub hi: u64,
}
impl From<u128> for ZklFe {
    fn from(v: u128) -> Self {
        ZklFe { lo: v as u64, hi: (v >> 64) as u64 }
    }
}
impl From<ZklFe> for u128 {
    fn from(v: ZklFe) -> u128 {
        (v.lo as u128) | ((v.hi as u128) << 64)
    }
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklSuite {
    pub dom: [ZklFe; 2],
    pub mds: [[ZklFe; 12]; 12],
    pub rc: [[ZklFe; 12]; ZKL_ROUNDS],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklRomSuite {
    pub mds: [[ZklFe; 3]; 3],
    pub rc: [[ZklFe; 3]; ZKL_ROUNDS],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklAssertion {
    pub column: u32,
    pub step: u32,
    pub value: ZklFe,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct ZklColumns {
    pub width: u32,
    pub lanes_start: u32,
    pub g_map: u32,
    pub g_final: u32,
    pub g_r_start: u32,
    pub mask: u32,
    pub r_start: u32,
    pub op_start: u32,
    pub sel_dst0_start: u32,
    pub sel_a_start: u32,
    pub sel_b_start: u32,
    pub sel_c_start: u32,
    pub sel_dst1_start: u32,
    pub sel_s_bits_start: u32,
    pub sel_s_active_start: u32,
    pub imm: u32,
    pub eq_inv: u32,
    pub ram_start: u32,
    pub merkle_start: u32,
    pub pi_prog: u32,
    pub pc: u32,
    pub rom_op_start: u32,
    pub pose_active: u32,
    pub gadget_b_start: u32,
    pub rom_s_start: u32,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklAirParams {
    pub cols: ZklColumns,
    pub f_poseidon: u32,
    pub f_vm: u32,
    pub f_sponge: u32,
    pub f_merkle: u32,
    pub f_ram: u32,
    pub rom_enabled: u32,
    pub vm_usage_mask: u32,
    pub ram_delta_clk_bits: u32,
    pub num_transition: u32,
    pub ce_blowup: u32,
    pub num_composition_cols: u32,
    pub num_assertions: u32,
    pub assertions: *const ZklAssertion,
    pub pose: ZklSuite,
    pub rom: ZklRomSuite,
    pub rom_w_enc0: [ZklFe; 59],
    pub rom_w_enc1: [ZklFe; 59],
    pub program_fe0: ZklFe,
    pub merkle_root_fe: ZklFe,
    /// ZKL_AIR_ZKLISP (0) or ZKL_AIR_AGG (1)
    pub air_kind: u32,
    pub pad_: u32,
}
pub const ZKL_AIR_ZKLISP: u32 = 0;
pub const ZKL_AIR_AGG: u32 = 1;
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct ZklOptions {
    pub num_queries: u32,
    pub blowup: u32,
    pub grinding: u32,
    pub num_partitions: u32,
    pub hash_rate: u32,
    pub fri_folding: u32,
    pub fri_remainder_max_degree: u32,
    /// winterfell::FieldExtension as u8: 1 = None, 2 = Quadratic
    pub field_extension: u32,
}
#[repr(C)]
pub struct ZklBatchProof {
    pub depth: u8,
    pub num_lists: u32,
    pub list_len: *mut u32,
    pub nodes: *mut u8,
    pub total_nodes: u32,
}
#[repr(C)]
pub struct ZklProofSummary {
    pub trace_root: [u8; 32],
    pub constraint_root: [u8; 32],
    pub fri_roots: [u8; 32 * 40],
    pub num_fri_roots: u32,
    pub z: ZklFe,
    pub pow_nonce: u64,
    pub num_unique_queries: u32,
    pub positions: [u32; 256],
    pub stage_ms: [f32; 12],
    pub constraints_violated: u32,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklProgram {
    pub program_id: [u8; 32],
    pub program_commitment: [u8; 32],
    pub pose: ZklSuite,
    pub rom: ZklRomSuite,
    pub rom_w_enc0: [ZklFe; 59],
    pub rom_w_enc1: [ZklFe; 59],
    pub program_fe: [ZklFe; 2],
    pub commitment_fe: [ZklFe; 2],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ZklPublicInputs {
    pub feature_mask: u64,
    pub segment_feature_mask: u64,
    pub merkle_root: [u8; 32],
    pub vm_expected_bytes: [u8; 32],
    pub vm_out_reg: u32,
    pub vm_out_row: u32,
    pub num_main_slots: u32,
    pub vm_usage_mask: u32,
    pub ram_delta_clk_bits: u32,
    pub pad_: u32,
    pub main_slots: [ZklFe; ZKL_MAX_MAIN_SLOTS],
    pub pc_init: ZklFe,
    pub ram_gp_unsorted_in: ZklFe,
    pub ram_gp_unsorted_out: ZklFe,
    pub ram_gp_sorted_in: ZklFe,
    pub ram_gp_sorted_out: ZklFe,
    pub rom_s_in: [ZklFe; 3],
    pub rom_s_out: [ZklFe; 3],
}

#[repr(C)]
pub struct zkl_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkl_trace_lde {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkl_constraint_commitment {
    _private: [u8; 0],
}

extern "C" {
    pub fn zkl_ctx_create(device: c_int, out: *mut *mut zkl_ctx) -> c_int;
    pub fn zkl_ctx_destroy(ctx: *mut zkl_ctx);
    pub fn zkl_last_error(ctx: *const zkl_ctx) -> *const c_char;
    pub fn zkl_suite_load(ctx: *mut zkl_ctx, suite: *const ZklSuite) -> c_int;
    pub fn zkl_suite_derive(suite_id: *const u8, out: *mut ZklSuite) -> c_int;
    pub fn zkl_program_derive(program_id: *const u8, program_commitment: *const u8, out: *mut ZklProgram) -> c_int;
    pub fn zkl_air_build(
        program: *const ZklProgram, pi: *const ZklPublicInputs, width: usize, n: usize, opts: *const ZklOptions,
        out_params: *mut ZklAirParams, out_assertions: *mut ZklAssertion, num_assertions: *mut usize, out_seed: *mut ZklFe,
        num_seed: *mut usize, err: *mut c_char, err_cap: usize,
    ) -> c_int;

    pub fn zkl_agg_air_params(n: usize, v_units_total: u64, children_count: u32, out_params: *mut ZklAirParams, out_assertions: *mut ZklAssertion) -> c_int;

    pub fn zkl_hash_elements_batch(ctx: *mut zkl_ctx, elems: *const ZklFe, len: usize, count: usize, out: *mut u8) -> c_int;
    pub fn zkl_merge_batch(ctx: *mut zkl_ctx, pairs: *const u8, count: usize, out: *mut u8) -> c_int;
    pub fn zkl_merge_with_int_batch(ctx: *mut zkl_ctx, seed: *const u8, values: *const u64, count: usize, out: *mut u8) -> c_int;
    pub fn zkl_hash_bytes(ctx: *mut zkl_ctx, data: *const u8, len: usize, out: *mut u8) -> c_int;
    pub fn zkl_merge_many(ctx: *mut zkl_ctx, digests: *const u8, count: usize, out: *mut u8) -> c_int;
    pub fn zkl_merkle_build(ctx: *mut zkl_ctx, leaves: *const u8, count: usize, out_root: *mut u8, out_nodes: *mut u8) -> c_int;
    pub fn zkl_interpolate_columns(ctx: *mut zkl_ctx, cols: *mut ZklFe, ncols: usize, n: usize) -> c_int;
    pub fn zkl_evaluate_columns(ctx: *mut zkl_ctx, polys: *const ZklFe, ncols: usize, n: usize, blowup: usize, offset: ZklFe, out: *mut ZklFe) -> c_int;

    pub fn zkl_trace_lde_commit(
        ctx: *mut zkl_ctx, trace: *const ZklFe, width: usize, n: usize, opts: *const ZklOptions, out_root: *mut u8,
        out_polys: *mut ZklFe, out_handle: *mut *mut zkl_trace_lde,
    ) -> c_int;
    pub fn zkl_trace_lde_free(ctx: *mut zkl_ctx, h: *mut zkl_trace_lde);
    pub fn zkl_trace_lde_read_frame(ctx: *mut zkl_ctx, h: *const zkl_trace_lde, lde_step: usize, out_cur: *mut ZklFe, out_next: *mut ZklFe) -> c_int;
    pub fn zkl_trace_query(ctx: *mut zkl_ctx, h: *const zkl_trace_lde, positions: *const u32, k: usize, out_rows: *mut ZklFe, out_proof: *mut ZklBatchProof) -> c_int;
    pub fn zkl_eval_constraints(ctx: *mut zkl_ctx, h: *const zkl_trace_lde, air: *const ZklAirParams, coeffs: *const ZklFe, out_ce_evals: *mut ZklFe) -> c_int;
    pub fn zkl_constraint_commit(
        ctx: *mut zkl_ctx, ce_evals: *const ZklFe, n: usize, ce_blowup: usize, num_cols: usize, opts: *const ZklOptions,
        out_root: *mut u8, out_poly: *mut ZklFe, out_handle: *mut *mut zkl_constraint_commitment,
    ) -> c_int;
    pub fn zkl_constraint_commitment_free(ctx: *mut zkl_ctx, h: *mut zkl_constraint_commitment);
    pub fn zkl_constraint_query(ctx: *mut zkl_ctx, h: *const zkl_constraint_commitment, positions: *const u32, k: usize, out_rows: *mut ZklFe, out_proof: *mut ZklBatchProof) -> c_int;

    pub fn zkl_prove_segment(
        ctx: *mut zkl_ctx, trace: *const ZklFe, width: usize, n: usize, air: *const ZklAirParams, seed_elems: *const ZklFe,
        num_seed: usize, opts: *const ZklOptions, out_proof: *mut u8, out_len: *mut usize, summary: *mut ZklProofSummary,
    ) -> c_int;
    pub fn zkl_prove_segment_pi(
        ctx: *mut zkl_ctx, trace: *const ZklFe, width: usize, n: usize, program: *const ZklProgram, pi: *const ZklPublicInputs,
        opts: *const ZklOptions, out_proof: *mut u8, out_len: *mut usize, summary: *mut ZklProofSummary,
    ) -> c_int;
    pub fn zkl_grind(ctx: *mut zkl_ctx, seed: *const u8, bits: u32, out_nonce: *mut u64) -> c_int;
    pub fn zkl_sync(ctx: *mut zkl_ctx) -> c_int;
    pub fn zkl_dev_alloc(ctx: *mut zkl_ctx, bytes: usize, out_dev: *mut *mut c_void) -> c_int;
    pub fn zkl_dev_free(ctx: *mut zkl_ctx, dev: *mut c_void) -> c_int;
    pub fn zkl_dev_upload(ctx: *mut zkl_ctx, dev: *mut c_void, host: *const c_void, bytes: usize) -> c_int;

    // ---- one large segment column-sharded over the GPUs of a box (include/zkl_b200.h: zkl_comm) ----
    pub fn zkl_shard_columns(opts: *const ZklOptions, width: usize, rank: c_int, world: c_int, first_col: *mut usize, local_width: *mut usize) -> c_int;
    pub fn zkl_prove_segment_sharded_pi_dev(
        ctx: *mut zkl_ctx, local_trace_dev: *const c_void, first_col: usize, local_width: usize, width: usize, n: usize, program: *const ZklProgram,
        pi: *const ZklPublicInputs, opts: *const ZklOptions, comm: *const ZklComm, out_proof: *mut u8, out_len: *mut usize, summary: *mut ZklProofSummary,
    ) -> c_int;
    pub fn zkl_prove_segment_sharded_dev(
        ctx: *mut zkl_ctx, local_trace_dev: *const c_void, first_col: usize, local_width: usize, width: usize, n: usize, air: *const ZklAirParams,
        seed_elems: *const ZklFe, num_seed: usize, opts: *const ZklOptions, comm: *const ZklComm, out_proof: *mut u8, out_len: *mut usize,
        summary: *mut ZklProofSummary,
    ) -> c_int;
    pub fn zkl_comm_loopback() -> *const ZklComm;
    pub fn zkl_comm_local_create(world: c_int, out: *mut *mut zkl_local_group) -> c_int;
    pub fn zkl_comm_local_rank(g: *mut zkl_local_group, rank: c_int) -> *const ZklComm;
    pub fn zkl_comm_local_destroy(g: *mut zkl_local_group);
    pub fn zkl_set_shard_min_leaves(ctx: *mut zkl_ctx, leaves: usize) -> c_int;
}

/// The caller's collectives for the column-sharded prover: callbacks on DEVICE pointers of the calling rank, collective over all
/// ranks, returning 0 once the result is complete (the library drains its stream before every call).
#[repr(C)]
pub struct ZklComm {
    pub rank: c_int,
    pub world: c_int,
    pub user: *mut c_void,
    /// recv = [world][bytes_per_rank]
    pub all_gather: Option<unsafe extern "C" fn(user: *mut c_void, send_dev: *const c_void, recv_dev: *mut c_void, bytes_per_rank: usize) -> c_int>,
    /// send, recv = [world][bytes_per_pair]
    pub all_to_all: Option<unsafe extern "C" fn(user: *mut c_void, send_dev: *const c_void, recv_dev: *mut c_void, bytes_per_pair: usize) -> c_int>,
    /// in place, exact modulo 2^64
    pub all_reduce_sum_u64: Option<unsafe extern "C" fn(user: *mut c_void, buf_dev: *mut c_void, count: usize) -> c_int>,
}
#[repr(C)]
pub struct zkl_local_group {
    _private: [u8; 0],
}
