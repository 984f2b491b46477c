/*
 * zkl_b200.h — C ABI of the B200-native zk-lisp / Winterfell proving hot path.
 *
 * Drop-in boundary. The reference has no FFI: its seam is the Rust trait
 * `winterfell::Prover` implemented at
 *   /root/reference/zk-lisp-proof-winterfell/src/prove.rs:425-517   (ZkWinterfellProver)
 * whose three hooks forward to Winterfell's CPU defaults:
 *   new_trace_lde                 prove.rs:454-474  -> zkl_trace_lde_commit
 *   new_evaluator (+ evaluate)    prove.rs:476-492  -> zkl_eval_constraints
 *   build_constraint_commitment   prove.rs:494-516  -> zkl_constraint_commit
 * plus the hook types' query methods (TraceLde::query, ConstraintCommitment::query)
 *                                                    -> zkl_trace_query / zkl_constraint_query
 * and, because OOD/DEEP/FRI/grinding live in the provided method `Prover::prove`
 * (prove.rs:225 calls it), a whole-proof entry point          -> zkl_prove_segment.
 * INTEGRATION.md shows the Rust-side binding (cc/bindgen shim) a maintainer would add.
 *
 * Conventions
 *  - every field element is winter-math f128 `BaseElement`: canonical value < p,
 *    p = 2^128 - 45*2^40 + 1, stored as 16 little-endian bytes (`zkl_fe`);
 *  - digests are the reference's 32-byte PoseidonDigest (poseidon/hasher.rs:25-26): the
 *    16 LE bytes of one field element followed by 16 zero bytes (utils.rs:375-381);
 *  - all pointers are HOST pointers unless the name ends in `_dev`;
 *  - every call returns 0 on success, non-zero on failure; zkl_last_error() gives the text.
 *    There is NO CPU fallback: without a CUDA device zkl_ctx_create fails.
 *  - threading: one zkl_ctx per host thread (prove.rs:1020-1048 runs one prover per rayon
 *    worker); a ctx owns one CUDA stream and is pinned to the device given at creation.
 */
#ifndef ZKL_B200_H
#define ZKL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t lo, hi; } zkl_fe; /* canonical u128, little-endian limbs */

#define ZKL_ROUNDS 27        /* vm/layout.rs:22 POSEIDON_ROUNDS */
#define ZKL_NR 8             /* vm/layout.rs NR */
#define ZKL_NUM_OPS 17       /* op_const .. op_store, vm/layout.rs:208-226 */
#define ZKL_NUM_PERIODIC 32  /* vm/air/mod.rs:526 */

/* Poseidon-style 12-lane suite (poseidon/mod.rs:33-38) */
typedef struct {
  zkl_fe dom[2];
  zkl_fe mds[12][12];
  zkl_fe rc[ZKL_ROUNDS][12];
} zkl_suite;

/* t=3 ROM suite (poseidon/mod.rs:186-197,219-251) */
typedef struct {
  zkl_fe mds[3][3];
  zkl_fe rc[ZKL_ROUNDS][3];
} zkl_rom_suite;

/* One `Assertion::single(column, step, value)` (vm/air/mod.rs:380-504). */
typedef struct {
  uint32_t column;
  uint32_t step;
  zkl_fe value;
} zkl_assertion;

/* Column map = vm/layout.rs `Columns` (for_config, :183-314) */
typedef struct {
  uint32_t width;
  uint32_t lanes_start, g_map, g_final, g_r_start, mask, r_start;
  uint32_t op_start;                 /* op_const; the 17 op columns are consecutive */
  uint32_t sel_dst0_start, sel_a_start, sel_b_start, sel_c_start, sel_dst1_start;
  uint32_t sel_s_bits_start, sel_s_active_start;
  uint32_t imm, eq_inv;
  uint32_t ram_start;                /* ram_sorted; 8 consecutive columns */
  uint32_t merkle_start;             /* merkle_g; 7 consecutive columns */
  uint32_t pi_prog, pc, rom_op_start, pose_active, gadget_b_start, rom_s_start;
} zkl_columns;

/* Everything ZkLispAir::new (vm/air/mod.rs:114-318) derives from AirPublicInputs that the
 * constraint kernel needs. Built by the host layer (zk-lisp_b200/air.py mirrors the Rust). */
typedef struct {
  zkl_columns cols;
  /* FeaturesMap (zk-lisp-proof/src/pi.rs) + rom gate (program_commitment != 0) */
  uint32_t f_poseidon, f_vm, f_sponge, f_merkle, f_ram, rom_enabled;
  uint32_t vm_usage_mask;            /* vm/layout.rs VM_USAGE_* bits */
  uint32_t ram_delta_clk_bits;
  uint32_t num_transition;           /* len(degrees), vm/air/mod.rs:216-238 */
  uint32_t ce_blowup;                /* AirContext::ce_blowup_factor  (8 for this AIR) */
  uint32_t num_composition_cols;     /* AirContext::num_constraint_composition_columns (7) */
  uint32_t num_assertions;
  const zkl_assertion *assertions;   /* get_assertions() order, already de-duplicated */
  zkl_suite pose;                    /* suite_id = program_id (vm/air/mod.rs:129-138) */
  zkl_rom_suite rom;                 /* vm/air/mod.rs:169-176 */
  zkl_fe rom_w_enc0[59], rom_w_enc1[59]; /* utils.rs:114-141 */
  zkl_fe program_fe0;                /* RAM compressor pi (vm/air/ram.rs:109) */
  zkl_fe merkle_root_fe;             /* be_from_le8(merkle_root) (vm/air/merkle.rs:121) */
  uint32_t air_kind;                 /* ZKL_AIR_ZKLISP (0): ZkLispAir, everything above; ZKL_AIR_AGG (1): ZlAggAir (agg/air.rs:22-332), only
                                        cols.width (31), num_transition (24), ce_blowup, num_composition_cols and the assertions are read */
  uint32_t pad_;
} zkl_air_params;
#define ZKL_AIR_ZKLISP 0
#define ZKL_AIR_AGG 1
#define ZKL_AGG_WIDTH 31             /* agg/layout.rs:97-180 AggColumns::baseline().width() */
#define ZKL_AGG_TRANSITIONS 24       /* agg/air.rs:69-94 */

/* winterfell::ProofOptions as zk-lisp builds them (prove.rs:963-972, :1106-1107) */
typedef struct {
  uint32_t num_queries;      /* 64 */
  uint32_t blowup;           /* 16 */
  uint32_t grinding;         /* 16 */
  uint32_t num_partitions;   /* utils.rs:394-409 */
  uint32_t hash_rate;
  uint32_t fri_folding;      /* 2  (only 2 is supported) */
  uint32_t fri_remainder_max_degree; /* 1 */
  uint32_t field_extension;  /* winterfell::FieldExtension: 0 or 1 = None (step proofs, prove.rs:967), 2 = Quadratic (the aggregation
                                proof when min_security_bits >= 128, prove.rs:647-651). With this hasher every drawn extension element
                                has a zero second coordinate (a digest is 16 bytes + 16 zero bytes, utils.rs:375-381, and
                                RandomCoin::draw reads E::ELEMENT_BYTES = 32 of them [WF-mem]), so all arithmetic stays in the base
                                field and the extension only changes what is hashed and serialized: E elements are (a, 0). */
} zkl_options;

typedef struct zkl_ctx zkl_ctx;
typedef struct zkl_trace_lde zkl_trace_lde;               /* replaces DefaultTraceLde (prove.rs:432) */
typedef struct zkl_constraint_commitment zkl_constraint_commitment; /* DefaultConstraintCommitment (:436) */

/* ---- context ------------------------------------------------------------------------- */
int zkl_ctx_create(int device, zkl_ctx **out);
void zkl_ctx_destroy(zkl_ctx *ctx);
const char *zkl_last_error(const zkl_ctx *ctx); /* ctx may be NULL: last create error */
/* Commitment/coin hasher constants: suite_id = [0;32] (poseidon/hasher.rs:23). The host computes them exactly as
 * poseidon/mod.rs does (or calls zkl_suite_derive) and hands them over. The suite is PER DEVICE, shared by every context
 * on it (the reference has one process-wide hasher suite too): loading the same constants again is a no-op, loading
 * different ones while another context on the device holds a suite fails. zkl_ctx_create refuses to run when the
 * reference's ZKL_POSEIDON_HASHER_ROUNDS (hasher.rs:235-241) selects anything but the 27 rounds implemented here. */
int zkl_suite_load(zkl_ctx *ctx, const zkl_suite *suite);
/* number of kernels launched on this ctx since creation (bench.py's gpu_launches) */
uint64_t zkl_launch_count(const zkl_ctx *ctx);

/* ---- hasher primitives (poseidon/hasher.rs:57-140), batched ----------------------------
 * Exposed so parity tests can check each PoseidonHasher entry point separately. */
int zkl_permute_batch(zkl_ctx *ctx, zkl_fe *states /* count*12, in/out */, size_t count);
int zkl_hash_elements_batch(zkl_ctx *ctx, const zkl_fe *elems /* count*len */, size_t len,
                            size_t count, uint8_t *out_digests /* count*32 */);
int zkl_merge_batch(zkl_ctx *ctx, const uint8_t *pairs /* count*64 */, size_t count,
                    uint8_t *out_digests);
int zkl_merge_with_int_batch(zkl_ctx *ctx, const uint8_t seed[32], const uint64_t *values,
                             size_t count, uint8_t *out_digests);
/* Batched Merkle authentication for the child-proof replay (agg/child.rs:1049-1068 `merkle_root_from_leaf` per path and
 * winter-crypto BatchMerkleProof::get_root per batch opening, here for every tree of every child at once).
 * A flat digest table starts with the `num_known` given digests (opened leaves and the proofs' nodes); level l appends
 * level_counts[l] entries, entry i of that level = merge(table[left[k+i]], table[right[k+i]]) with k the running offset
 * into left/right; indices may only refer to entries that exist before the level starts. The levels run back to back
 * on the device (one gather + one merge launch each), no host round trip in between; `want` selects the table entries
 * copied back (the recomputed roots). */
int zkl_merge_plan(zkl_ctx *ctx, const uint8_t *known /* num_known*32 */, size_t num_known, const uint32_t *left,
                   const uint32_t *right, const uint32_t *level_counts, size_t num_levels, const uint32_t *want,
                   size_t num_want, uint8_t *out_digests /* num_want*32 */);
/* out[i] = merge_with_int(seeds[i], values[i]) (hasher.rs:107-120): the random coins of many child proofs advanced in lockstep
 * (agg/fs.rs:37-245 replayed for a whole batch of children, one launch per transcript step instead of one per child). */
int zkl_merge_with_int_pairs(zkl_ctx *ctx, const uint8_t *seeds /* count*32 */, const uint64_t *values, size_t count,
                             uint8_t *out_digests /* count*32 */);
int zkl_hash_bytes(zkl_ctx *ctx, const uint8_t *data, size_t len, uint8_t out[32]);
int zkl_merge_many(zkl_ctx *ctx, const uint8_t *digests, size_t count, uint8_t out[32]);
/* MerkleTree::new over `count` leaves (power of two); out_nodes[i], i in 1..count, is the
 * Winterfell node array (nodes[1] = root); out_nodes may be NULL. */
int zkl_merkle_build(zkl_ctx *ctx, const uint8_t *leaves, size_t count, uint8_t out_root[32],
                     uint8_t *out_nodes /* count*32 or NULL */);

/* ---- NTT primitives (winter-math fft) --------------------------------------------------
 * cols: `ncols` columns of `n` elements, column-major. */
int zkl_interpolate_columns(zkl_ctx *ctx, zkl_fe *cols, size_t ncols, size_t n);
/* out[c][i] = poly_c(offset * w_N^i), N = n*blowup, natural order */
int zkl_evaluate_columns(zkl_ctx *ctx, const zkl_fe *polys, size_t ncols, size_t n,
                         size_t blowup, zkl_fe offset, zkl_fe *out /* ncols*N */);

/* ---- hook 1: new_trace_lde (prove.rs:454-474 -> DefaultTraceLde::new) -------------------
 * trace: W columns x n rows, column-major (winterfell ColMatrix). Computes the trace
 * polynomials, the blowup-`blowup` LDE over the coset GENERATOR*<w_N>, row hashes with the
 * given partition options and the Merkle tree. out_polys (W*n, column-major) may be NULL. */
int zkl_trace_lde_commit(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n,
                         const zkl_options *opts, uint8_t out_root[32], zkl_fe *out_polys,
                         zkl_trace_lde **out_handle);
void zkl_trace_lde_free(zkl_ctx *ctx, zkl_trace_lde *h);
/* TraceLde::read_main_trace_frame_into(lde_step): current = LDE row lde_step,
 * next = row (lde_step + blowup) mod N. out_cur/out_next: width elements each. */
int zkl_trace_lde_read_frame(zkl_ctx *ctx, const zkl_trace_lde *h, size_t lde_step,
                             zkl_fe *out_cur, zkl_fe *out_next);
/* TraceLde::query(positions): rows (k*width elements) and the Winterfell BatchMerkleProof
 * node lists flattened as out_nodes with per-position counts (see zkl_batch_proof). */
typedef struct {
  uint8_t depth;
  uint32_t num_lists;      /* number of node vectors */
  uint32_t *list_len;      /* num_lists entries (caller-allocated, capacity >= k) */
  uint8_t *nodes;          /* concatenated 32-byte digests (capacity >= k*depth*32) */
  uint32_t total_nodes;
} zkl_batch_proof;
int zkl_trace_query(zkl_ctx *ctx, const zkl_trace_lde *h, const uint32_t *positions, size_t k,
                    zkl_fe *out_rows, zkl_batch_proof *out_proof);

/* ---- hook 2: new_evaluator + ConstraintEvaluator::evaluate (prove.rs:476-492) -----------
 * coeffs: ConstraintCompositionCoefficients, `num_transition` transition coefficients
 * followed by `num_assertions` boundary coefficients in Winterfell's sorted-assertion order.
 * out: the CompositionPolyTrace, n*ce_blowup combined evaluations over GENERATOR*<w_{8n}>. */
int zkl_eval_constraints(zkl_ctx *ctx, const zkl_trace_lde *h, const zkl_air_params *air,
                         const zkl_fe *coeffs, zkl_fe *out_ce_evals);

/* ---- hook 3: build_constraint_commitment (prove.rs:494-516) ------------------------------
 * ce_evals: n*ce_blowup evaluations. Interpolates, splits into num_cols columns of n
 * coefficients (out_poly, column-major num_cols*n, may be NULL), evaluates over the LDE
 * domain, hashes rows and builds the Merkle tree. */
int zkl_constraint_commit(zkl_ctx *ctx, const zkl_fe *ce_evals, size_t n, size_t ce_blowup,
                          size_t num_cols, const zkl_options *opts, uint8_t out_root[32],
                          zkl_fe *out_poly, zkl_constraint_commitment **out_handle);
void zkl_constraint_commitment_free(zkl_ctx *ctx, zkl_constraint_commitment *h);
int zkl_constraint_query(zkl_ctx *ctx, const zkl_constraint_commitment *h,
                         const uint32_t *positions, size_t k, zkl_fe *out_rows,
                         zkl_batch_proof *out_proof);

/* ---- whole segment proof: winterfell::Prover::prove (prove.rs:225) ----------------------
 * trace: W x n column-major; seed_elems: Context::to_elements() ++ AirPublicInputs::to_elements()
 * (lib.rs:116-160), i.e. the public-coin seed (agg/fs.rs:69-73). Runs stages 1-11 of
 * SURVEY.md §3.2 on the device and writes the serialized winterfell::Proof.
 * out_proof/out_len: caller buffer and capacity in, bytes written out. */
typedef struct {
  uint8_t trace_root[32];
  uint8_t constraint_root[32];
  uint8_t fri_roots[32 * 40];   /* num_fri_layers + 1 (remainder commitment) */
  uint32_t num_fri_roots;
  zkl_fe z;                     /* OOD point */
  uint64_t pow_nonce;
  uint32_t num_unique_queries;
  uint32_t positions[256];
  float stage_ms[12];           /* [0..8] device ms per stage, [9] trace k_lde ms, [10] trace k_hash_rows ms,
                                   [11] host ms spent building the AIR parameters (zkl_prove_segment_pi*); see zkl_stage_name */
  uint32_t constraints_violated; /* 1 = the trace does not satisfy the AIR (call returned ZKL_ERR_CONSTRAINTS_VIOLATED) */
} zkl_proof_summary;
int zkl_prove_segment(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n,
                      const zkl_air_params *air, const zkl_fe *seed_elems, size_t num_seed,
                      const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                      zkl_proof_summary *summary /* may be NULL */);
/* Same, with the trace already resident in device memory (bench.py `value` leg). */
int zkl_prove_segment_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n,
                          const zkl_air_params *air, const zkl_fe *seed_elems, size_t num_seed,
                          const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                          zkl_proof_summary *summary);
const char *zkl_stage_name(int stage);
/* zkl_prove_segment* return codes beyond 0 / 1: the proof buffer was too small (*out_len holds the size needed) and the trace
 * does not satisfy the AIR (the composition polynomial overflowed its degree: Winterfell panics on this in debug builds and
 * emits an unverifiable proof in release builds; here the bytes are still written, summary->constraints_violated is set and
 * the call fails so that no caller ships such a proof by accident). */
#define ZKL_ERR_BUFFER_TOO_SMALL 2
#define ZKL_ERR_CONSTRAINTS_VIOLATED 3

/* ---- ZkLispAir::new on this side of the boundary (vm/air/mod.rs:114-318, 380-504; lib.rs:116-160) ------------------
 * The per-segment AIR data is formulaic: 141 schedule assertions per 32-row level (vm/air/schedule.rs:57-139), at most 16
 * that depend on the public inputs (vm/air/mod.rs:388-446, rom.rs:122-147), the constraint degrees of the enabled modules
 * and the public-coin seed Context::to_elements() ++ AirPublicInputs::to_elements(). The library derives all of it from the
 * public inputs, so a caller hands over what `prove_segment` (prove.rs:1057-1175) has at hand and nothing else. */
/* Per-program constants (one derivation per program, shared by all of its segments; the reference keeps them in a
 * process-global cache, poseidon/mod.rs:47). zkl_program_derive computes them exactly as poseidon/mod.rs:52-251,
 * utils.rs:114-141 and commit.rs:31-79 do (blake3-derived round constants, Cauchy MDS); host only, needs no context. */
typedef struct {
  uint8_t program_id[32];
  uint8_t program_commitment[32];
  zkl_suite pose;                        /* get_poseidon_suite(program_id) */
  zkl_rom_suite rom;                     /* derive_rom_mds_cauchy_3x3 / derive_rom_round_constants_3 (program_id) */
  zkl_fe rom_w_enc0[59], rom_w_enc1[59]; /* rom_weights_for_seed(ROM_W_SEED_0 / _1) */
  zkl_fe program_fe[2];                  /* program_field_commitment(program_id), zero when program_id is zero */
  zkl_fe commitment_fe[2];               /* program_field_commitment(program_commitment), zero when it is zero */
} zkl_program;
int zkl_program_derive(const uint8_t program_id[32], const uint8_t program_commitment[32], zkl_program *out);
/* get_poseidon_suite(suite_id) alone (poseidon/mod.rs:52-75); suite_id = [0;32] gives the commitment hasher's suite
 * (poseidon/hasher.rs:23) for zkl_suite_load. */
int zkl_suite_derive(const uint8_t suite_id[32], zkl_suite *out);

#define ZKL_MAX_MAIN_SLOTS 16
/* AirPublicInputs (lib.rs:74-96) + the CorePublicInputs fields the AIR reads (zk-lisp-proof/src/pi.rs) */
typedef struct {
  uint64_t feature_mask;           /* core.feature_mask */
  uint64_t segment_feature_mask;   /* 0 = use core.feature_mask (vm/air/mod.rs:119-124) */
  uint8_t merkle_root[32];
  uint8_t vm_expected_bytes[32];
  uint32_t vm_out_reg, vm_out_row;
  uint32_t num_main_slots;         /* utils::encode_main_args_to_slots(core.main_args) (utils.rs:79-109) */
  uint32_t vm_usage_mask, ram_delta_clk_bits, pad_;
  zkl_fe main_slots[ZKL_MAX_MAIN_SLOTS];
  zkl_fe pc_init;
  zkl_fe ram_gp_unsorted_in, ram_gp_unsorted_out, ram_gp_sorted_in, ram_gp_sorted_out;
  zkl_fe rom_s_in[3], rom_s_out[3];
} zkl_public_inputs;
/* Fills *out_params (its `assertions` pointer is set to out_assertions), the assertions in get_assertions() order
 * (de-duplicated by (column, step), first kept) and the public-coin seed elements. Capacities in, counts out; a too small
 * buffer fails with the needed counts stored. Host only. Errors (width/layout mismatch, declared vs produced assertion count,
 * program_id / program_commitment not both zero or both non-zero) are reported through `err` (may be NULL). */
int zkl_air_build(const zkl_program *program, const zkl_public_inputs *pi, size_t width, size_t n, const zkl_options *opts,
                  zkl_air_params *out_params, zkl_assertion *out_assertions, size_t *num_assertions /* in: capacity */,
                  zkl_fe *out_seed, size_t *num_seed /* in: capacity */, char *err, size_t err_cap);
/* zkl_prove_segment with the AIR parameters and the coin seed built inside (zkl_air_build), per call: nothing about a
 * segment is cached between calls. */
/* ZlAggAir::new (agg/air.rs:32-113) + get_assertions (:274-305): the parameters of the 31-column aggregation AIR over a trace of
 * n rows (one row per child, padded to a power of two >= 8, agg/trace.rs:396-399). out_assertions: room for 5. Host only. */
int zkl_agg_air_params(size_t n, uint64_t v_units_total, uint32_t children_count, zkl_air_params *out_params,
                       zkl_assertion *out_assertions);
int zkl_prove_segment_pi(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n, const zkl_program *program,
                         const zkl_public_inputs *pi, const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                         zkl_proof_summary *summary);
int zkl_prove_segment_pi_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n, const zkl_program *program,
                             const zkl_public_inputs *pi, const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                             zkl_proof_summary *summary);

/* device memory helpers for callers that keep traces resident (torch tensors' data_ptr()
 * can be passed instead) */
int zkl_dev_alloc(zkl_ctx *ctx, size_t bytes, void **out_dev);
int zkl_dev_free(zkl_ctx *ctx, void *dev);
int zkl_dev_upload(zkl_ctx *ctx, void *dev, const void *host, size_t bytes);
int zkl_sync(zkl_ctx *ctx);

/* ---- child-proof replay for the aggregation layer -------------------------------------
 * agg/fs.rs:37-245 (Fiat-Shamir replay of a step proof) and agg/child.rs:1025-1099 (re-hashing of every opening) for a whole
 * batch of children: transcripts advance in lockstep (one device batch per transcript step), leaves are hashed in three
 * batches and every tree of every child is authenticated by one merge plan. Per-child results carry a status
 * (0 = ok, else see zkl_replay_status_name; FRI opening failures carry the layer in bits 8..15). */
typedef struct {
  const uint8_t *proof; /* inner winterfell::Proof bytes (as written by zkl_prove_segment) */
  size_t proof_len;
  const zkl_fe *pi_elems; /* AirPublicInputs::to_elements() of the child (lib.rs:116-160), rebuilt as in agg/fs.rs:43-66 */
  size_t num_pi;
} zkl_child_proof;
typedef struct {
  int32_t status;
  uint32_t num_positions, num_alphas, num_deep;
  uint64_t pow_nonce;
  zkl_fe z;
  zkl_fe alphas[40];
  uint32_t positions[256];
} zkl_child_challenges;
/* out_deep receives the DEEP coefficients of child i at out_deep[i * deep_stride ...] (trace columns, then composition columns). */
int zkl_replay_children(zkl_ctx *ctx, const zkl_child_proof *children, size_t count, int verify_openings,
                        zkl_child_challenges *out, zkl_fe *out_deep, size_t deep_stride);
const char *zkl_replay_status_name(int status);
/* poseidon_hash_two_lanes (poseidon/mod.rs:255-291) under an arbitrary suite, computed on the host with the kernels' limb
 * arithmetic: out[i] = lane 0 of the permutation of [left[i], right[i], 0 x 8, dom0, dom1]. Used by the step digest
 * (proof/digest.rs:15-68) and the children root (agg/child.rs:853-895), which run under the PROGRAM's suite. Inputs must be
 * canonical (< p). Needs no context and no device. */
int zkl_hash_two_lanes_host(const zkl_suite *suite, const zkl_fe *left, const zkl_fe *right, size_t count, zkl_fe *out);
/* ProverChannel::grind on its own: smallest nonce >= 1 with trailing_zeros(u64_le(merge_with_int(seed, nonce)[0..8])) >= bits
 * (check restated at agg/fs.rs:202-207). For provers that stay on the CPU (the aggregation proof) but whose proof-of-work
 * search is 2^bits calls of the CPU hasher. bits = 0 returns nonce 0 like Winterfell. */
int zkl_grind(zkl_ctx *ctx, const uint8_t seed[32], uint32_t bits, uint64_t *out_nonce);

/* ---- segment driver helpers (callers of the hot path: prove.rs:954-1175) ------------
 * The reference builds the full trace once and, per segment, slices it on the host
 * (vm/trace/mod.rs:367-384), reads boundary cells (prove.rs:1197-1287) and scans every row for gadget usage
 * (prove.rs:1289-1392). With the full trace resident in HBM these are one kernel each. */
int zkl_slice_segment_dev(zkl_ctx *ctx, const void *full_dev, size_t full_rows, const uint32_t *seg_to_full,
                          size_t seg_width, size_t r_start, size_t seg_rows, void *out_dev);
/* Poseidon gadget rows of the full trace, batched over levels (vm/trace/poseidon.rs:24-96 apply_level_absorb): for level
 * L = levels[i], row 32L (map) receives inputs[i][0..10) in lanes 0..9 (zero padded by the caller) and the suite's domain tags
 * in lanes 10, 11; row 32L+1+j the state before round j (j < 27); row 32L+28 and the pad rows 29..31 the final state.
 * `trace_dev` is column-major [width][n_rows]; the lanes are columns lanes_start .. lanes_start+11; `suite` is the PROGRAM suite
 * (suite_id = program_id, prove.rs:985), not the commitment hasher's. */
int zkl_poseidon_levels_dev(zkl_ctx *ctx, void *trace_dev, size_t n_rows, uint32_t lanes_start, const zkl_suite *suite,
                            const uint32_t *levels, const zkl_fe *inputs /* count*10 */, size_t count);
int zkl_usage_mask_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n, const zkl_columns *cols,
                       int has_ram, uint32_t *out_mask, uint32_t *out_ram_bits);
int zkl_read_cells_dev(zkl_ctx *ctx, const void *trace_dev, size_t n, const uint32_t *cols, const uint32_t *rows,
                       size_t k, zkl_fe *out);

/* ---- column-sharded trace commitment: device-pointer building blocks (single huge segment) ----
 * The reference hashes a row as merge_many(hash_elements(partition_k)) when PartitionOptions asks for several
 * partitions (agg/child.rs:1025-1045). A rank that owns whole partitions therefore runs iNTT, LDE and partition
 * hashing locally; the ranks exchange one 16-byte digest per (row, partition) (NCCL all-gather in
 * zk-lisp_b200/sharded.py), then merge_many + MerkleTree::new. Elements are 16 bytes (zkl_fe layout); a digest
 * is carried as ONE element (its upper 16 bytes are zero, utils.rs:375-381). Calls enqueue on the ctx stream. */
size_t zkl_partition_size(const zkl_options *opts, size_t width);   /* PartitionOptions::partition_size */
/* cols_dev: ncols x n evaluations (column-major) -> out_polys_dev (ncols x n) and the coset-major LDE
 * out_lde_dev[c][k][i] = T_c(GENERATOR w_N^(k + blowup i)) (ncols x n*blowup) */
int zkl_lde_dev(zkl_ctx *ctx, const void *cols_dev, size_t ncols, size_t n, size_t blowup,
                void *out_polys_dev, void *out_lde_dev);
/* out_digests_dev[q*N + r] = hash_elements(LDE row r, local columns [q*psize, min(ncols,(q+1)*psize))), natural r */
int zkl_hash_partitions_dev(zkl_ctx *ctx, const void *lde_dev, size_t ncols, size_t n, size_t blowup,
                            size_t partition_size, void *out_digests_dev);
/* out_leaves_dev[r] = merge_many(digests[0][first_row + r], ..., digests[nparts-1][first_row + r]), r < rows;
 * digests[q] starts at digests_dev + q*row_stride elements */
int zkl_merge_partition_digests_dev(zkl_ctx *ctx, const void *digests_dev, size_t nparts, size_t row_stride,
                                    size_t first_row, size_t rows, void *out_leaves_dev);
/* MerkleTree::new on the device: out_nodes_dev[1] = root, out_nodes_dev[count/2 .. count) = parents of the leaves */
int zkl_merkle_dev(zkl_ctx *ctx, const void *leaves_dev, size_t count, void *out_nodes_dev);
int zkl_dev_download(zkl_ctx *ctx, void *host, const void *dev, size_t bytes);

/* ---- column-sharded proof of ONE segment across the GPUs of a box (SURVEY.md §8e "single huge segment") ---------------------------
 * Rank r owns whole column partitions (zkl_shard_columns; PartitionOptions must ask for >= world partitions), the LDE cosets
 * [r B / R, (r + 1) B / R) of everything that is not column-sharded, and the Merkle sub-tree over the rows [r N / R, (r + 1) N / R).
 * Per rank the flow is the single-GPU prover's with these exchanges:
 *  (1) trace commitment: all-to-all of the 16-byte partition digests (every rank receives all partitions of ITS rows and merges them
 *      into leaves), all-gather of the leaves, all-gather of the per-rank sub-trees;
 *  (2) constraint evaluation: all-to-all of the constraint-evaluation cosets (rank r evaluates CE cosets [r ce / R, (r + 1) ce / R) on
 *      full-width rows), all-gather of the evaluations; composition columns on every rank, their row hashes per coset range;
 *  (3) sums of values with disjoint support (out-of-domain trace evaluations; every opened element) as 64-bit lanes;
 *  (4) DEEP: partial sums over the owned columns, reduce-scatter by coset range (all-to-all + field addition);
 *  (5) FRI per coset range: leaves hashed and folded locally, leaves and sub-trees all-gathered, until a layer is smaller than
 *      zkl_set_shard_min_leaves — from there on (and for grinding and the query plan) every rank does the same work.
 * Every rank returns the same proof bytes, identical to the single-GPU proof of the same trace and options.
 * The collectives are the caller's (NCCL through torch.distributed in zk-lisp_b200/sharded.py, ncclComm in a Rust host, or
 * zkl_comm_local_* below for the threads of one process): callbacks on DEVICE pointers of the calling rank, collective over all ranks,
 * returning 0 once the result is complete in `recv` / `buf` (the library synchronises its stream before each call). */
typedef struct {
  int rank, world;             /* world must divide the constraint-evaluation blowup (8 for ZkLispAir): 1, 2, 4 or 8 */
  void *user;
  int (*all_gather)(void *user, const void *send_dev, void *recv_dev, size_t bytes_per_rank);   /* recv = [world][bytes_per_rank] */
  int (*all_to_all)(void *user, const void *send_dev, void *recv_dev, size_t bytes_per_pair);   /* send, recv = [world][bytes_per_pair] */
  int (*all_reduce_sum_u64)(void *user, void *buf_dev, size_t count);                            /* in place */
} zkl_comm;
/* [first_col, first_col + local_width) = the columns `rank` owns: contiguous blocks of ceil(P / world) whole partitions, P =
 * ceil(width / partition_size). Fails unless every rank owns at least one partition. */
int zkl_shard_columns(const zkl_options *opts, size_t width, int rank, int world, size_t *first_col, size_t *local_width);
/* local_trace_dev: this rank's columns, local_width x n, column-major, in device memory. air / seed_elems describe the WHOLE trace. */
int zkl_prove_segment_sharded_dev(zkl_ctx *ctx, const void *local_trace_dev, size_t first_col, size_t local_width, size_t width, size_t n,
                                  const zkl_air_params *air, const zkl_fe *seed_elems, size_t num_seed, const zkl_options *opts,
                                  const zkl_comm *comm, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary);
int zkl_prove_segment_sharded_pi_dev(zkl_ctx *ctx, const void *local_trace_dev, size_t first_col, size_t local_width, size_t width, size_t n,
                                     const zkl_program *program, const zkl_public_inputs *pi, const zkl_options *opts, const zkl_comm *comm,
                                     uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary);
/* world = 1 with device-to-device copies as "collectives": single-GPU runs of the sharded code path (tests) */
const zkl_comm *zkl_comm_loopback(void);
/* Collectives for `world` proving THREADS of one process, one zkl_ctx each, on one device or on several devices of a box (peer copies):
 * publish the pointer, barrier, copy out of the other ranks' buffers, barrier. zkl_comm_local_rank(g, r) is rank r's zkl_comm; a rank
 * that fails breaks the group (the other ranks' collectives return an error instead of waiting forever). */
typedef struct zkl_local_group zkl_local_group;
int zkl_comm_local_create(int world, zkl_local_group **out);
const zkl_comm *zkl_comm_local_rank(zkl_local_group *g, int rank);
void zkl_comm_local_destroy(zkl_local_group *g);
/* Trees and FRI layers with fewer leaves than this are built by every rank instead of exchanged (default 32768; ZKL_SHARD_MIN_LEAVES). */
int zkl_set_shard_min_leaves(zkl_ctx *ctx, size_t leaves);

/* ---- measurement helpers (bench.py) -------------------------------------------------------- */
int zkl_timer_start(zkl_ctx *ctx);             /* CUDA event on the ctx stream */
int zkl_timer_stop(zkl_ctx *ctx, float *ms);   /* second event, synchronised; elapsed device time */
int zkl_imad_peak(zkl_ctx *ctx, double *wide_macs_per_s); /* dependency-free IMAD.WIDE.U32 probe */
/* saturating tcgen05.mma kind::i8 loop (M 128, N 256, K 32 from shared memory, one CTA per SM): dense u8 x u8 -> s32 tensor
 * operations per second (2 per multiply-accumulate) — the measured denominator of the sponge kernels' tensor roofline */
int zkl_tc_i8_peak(zkl_ctx *ctx, double *ops_per_s);
/* sponge count from which the bulk (tensor-core) sponge kernels replace the small-batch ones; 0 = automatic
 * (4096 for a lone prover on the 64-thread groups, 1024 when several contexts share the device) */
int zkl_set_bulk_threshold(zkl_ctx *ctx, long sponges);
/* bulk hashing path: 0 = tensor-core MDS product (tcgen05 kind::i8, default), 1 = integer-pipe thread-per-sponge
 * kernels. Both are bit-identical; 1 exists for A/B measurements and parity tests. */
int zkl_set_hash_path(zkl_ctx *ctx, int path);
/* Small sponge batches (tree levels, FRI layer tails, the Fiat-Shamir coin, the OOD digest): 1 = 16-thread groups per
 * sponge (least SM time: a pool of provers sharing the GPU), 2 = 64-thread groups (lowest latency: a lone prover),
 * 0 = choose by the number of contexts alive on the device (default; fewer than 4 -> 2). Bit-identical. */
int zkl_set_coop_path(zkl_ctx *ctx, int path);
/* 4096-row coset LDE: 0 = radix-16 stages on the tensor cores with the coset shift folded into the constant matrix
 * (default), 1 = radix-2 butterflies in shared memory, 2 = tensor cores with a separate coset-shift product.
 * Bit-identical; 1 and 2 exist for A/B measurements and parity tests. */
int zkl_set_ntt_path(zkl_ctx *ctx, int path);
/* permutation throughput probe on device-resident states: kind 0 = thread-per-state kernel, 1 = cooperative,
 * 2 = tensor-core MDS kernel, 3 = 64-thread groups */
int zkl_bench_permute(zkl_ctx *ctx, int kind, size_t count, int iters, float *ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* ZKL_B200_H */
