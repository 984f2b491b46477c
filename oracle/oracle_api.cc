// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/f128.h header).
// C entry points (ctypes) over the CPU restatement. Loaded by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs ONLY; never by the product path.
#include <cstdio>
#include <thread>

#include "../include/zkl_b200.h"
#include "stark_core.h"
#include "prover.h"

using namespace zkl;

static const Hasher &H() { static Hasher h; return h; }
static inline F fe_in(const zkl_fe &x) { F r; r.v = mk128(x.hi, x.lo); return r; }
static inline zkl_fe fe_out(F x) { zkl_fe r; r.lo = (uint64_t)x.v; r.hi = (uint64_t)(x.v >> 64); return r; }

extern "C" {

// ---- field ---------------------------------------------------------------
void zo_f_binop(int op, const zkl_fe *a, const zkl_fe *b, zkl_fe *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    F x = fe_in(a[i]), y = fe_in(b[i]), r;
    switch (op) {
      case 0: r = f_add(x, y); break;
      case 1: r = f_sub(x, y); break;
      case 2: r = f_mul(x, y); break;
      case 3: r = f_inv(x); break;
      case 4: r = f_cube(x); break;
      default: r = f_zero();
    }
    out[i] = fe_out(r);
  }
}
void zo_root_of_unity(int k, zkl_fe *out) { *out = fe_out(f_root_of_unity(k)); }
void zo_from_u128_mod(const zkl_fe *raw, zkl_fe *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = fe_out(f_from_u128(mk128(raw[i].hi, raw[i].lo)));
}

// ---- blake3 / suites -------------------------------------------------------
void zo_blake3(const uint8_t *data, size_t len, uint8_t out[32]) { Blake3::hash(data, len, out); }
void zo_derive_suite(const uint8_t sid[32], zkl_suite *out) {
  Suite s = derive_suite(sid);
  for (int i = 0; i < 2; i++) out->dom[i] = fe_out(s.dom[i]);
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) out->mds[i][j] = fe_out(s.mds[i][j]);
  for (int r = 0; r < ROUNDS; r++) for (int j = 0; j < 12; j++) out->rc[r][j] = fe_out(s.rc[r][j]);
}
void zo_derive_rom_suite(const uint8_t sid[32], zkl_rom_suite *out) {
  RomSuite s = derive_rom_suite(sid);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out->mds[i][j] = fe_out(s.mds[i][j]);
  for (int r = 0; r < ROUNDS; r++) for (int j = 0; j < 3; j++) out->rc[r][j] = fe_out(s.rc[r][j]);
}

// ---- hasher ------------------------------------------------------------------
void zo_permute(zkl_fe *state, size_t count) {
  for (size_t c = 0; c < count; c++) {
    F st[12]; for (int i = 0; i < 12; i++) st[i] = fe_in(state[c * 12 + i]);
    permute(H().ps, st);
    for (int i = 0; i < 12; i++) state[c * 12 + i] = fe_out(st[i]);
  }
}
void zo_hash_elements(const zkl_fe *e, size_t len, size_t count, uint8_t *out) {
  for (size_t c = 0; c < count; c++) {
    Digest d = H().hash_elements((const F *)(e + c * len), len);
    memcpy(out + 32 * c, d.b, 32);
  }
}
void zo_merge(const uint8_t *pairs, size_t count, uint8_t *out) {
  for (size_t c = 0; c < count; c++) {
    Digest a, b; memcpy(a.b, pairs + 64 * c, 32); memcpy(b.b, pairs + 64 * c + 32, 32);
    Digest d = H().merge(a, b); memcpy(out + 32 * c, d.b, 32);
  }
}
void zo_merge_with_int(const uint8_t seed[32], const uint64_t *vals, size_t count, uint8_t *out) {
  Digest s; memcpy(s.b, seed, 32);
  for (size_t c = 0; c < count; c++) { Digest d = H().merge_with_int(s, vals[c]); memcpy(out + 32 * c, d.b, 32); }
}
void zo_hash_bytes(const uint8_t *data, size_t len, uint8_t out[32]) {
  Digest d = H().hash_bytes(data, len); memcpy(out, d.b, 32);
}
void zo_merge_many(const uint8_t *digests, size_t count, uint8_t out[32]) {
  Digest d = H().merge_many((const Digest *)digests, count); memcpy(out, d.b, 32);
}
void zo_merkle_build(const uint8_t *leaves, size_t count, uint8_t root[32], uint8_t *nodes) {
  std::vector<Digest> lv(count); memcpy(lv.data(), leaves, 32 * count);
  MerkleTree t(H(), lv);
  memcpy(root, t.root().b, 32);
  if (nodes) memcpy(nodes, t.nodes.data(), 32 * count);
}

int zo_batch_root(int depth, size_t num_lists, const uint32_t *list_len, const uint8_t *nodes, const uint8_t *leaves,
                  const uint32_t *positions, size_t k, uint8_t out[32]) {
  BatchMerkleProof bp; bp.depth = (uint8_t)depth;
  size_t off = 0;
  for (size_t i = 0; i < num_lists; i++) {
    std::vector<Digest> v(list_len[i]);
    for (auto &d : v) { memcpy(d.b, nodes + 32 * off, 32); off++; }
    bp.nodes.push_back(v);
  }
  std::vector<Digest> lv(k); std::vector<size_t> pos(k);
  for (size_t i = 0; i < k; i++) { memcpy(lv[i].b, leaves + 32 * i, 32); pos[i] = positions[i]; }
  Digest root;
  if (!batch_root(H(), bp, lv, pos, root)) return 1;
  memcpy(out, root.b, 32);
  return 0;
}

// ---- fft --------------------------------------------------------------------------
void zo_interpolate_columns(zkl_fe *cols, size_t ncols, size_t n) {
  for (size_t c = 0; c < ncols; c++) {
    std::vector<F> a((F *)cols + c * n, (F *)cols + (c + 1) * n);
    interpolate(a);
    memcpy(cols + c * n, a.data(), 16 * n);
  }
}
void zo_evaluate_columns(const zkl_fe *polys, size_t ncols, size_t n, size_t blowup, const zkl_fe *offset, zkl_fe *out) {
  for (size_t c = 0; c < ncols; c++) {
    std::vector<F> p((F *)polys + c * n, (F *)polys + (c + 1) * n);
    std::vector<F> e = evaluate_over(p, n * blowup, fe_in(*offset));
    memcpy(out + c * n * blowup, e.data(), 16 * n * blowup);
  }
}

// ---- hook 1 restatement: DefaultTraceLde::new ------------------------------------
void zo_trace_lde_commit(const zkl_fe *trace, size_t width, size_t n, const zkl_options *o,
                         uint8_t root[32], zkl_fe *out_polys, zkl_fe *out_lde /* width*N col-major or NULL */) {
  TraceCommitment tc = commit_trace(H(), (const F *)trace, width, n, o->blowup, o->num_partitions, o->hash_rate);
  memcpy(root, tc.tree.root().b, 32);
  if (out_polys) for (size_t c = 0; c < width; c++) memcpy(out_polys + c * n, tc.polys[c].data(), 16 * n);
  if (out_lde) { size_t N = n * o->blowup; for (size_t c = 0; c < width; c++) memcpy(out_lde + c * N, tc.lde[c].data(), 16 * N); }
}

// ---- hook 2 restatement: DefaultConstraintEvaluator::evaluate ----------------------
void zo_eval_constraints(const zkl_fe *trace, size_t width, size_t n, const zkl_options *o, const zkl_air_params *air,
                         const zkl_fe *coeffs, zkl_fe *out_ce) {
  TraceCommitment tc = commit_trace(H(), (const F *)trace, width, n, o->blowup, o->num_partitions, o->hash_rate, /*skip_hash=*/true);
  std::vector<F> cc((const F *)coeffs, (const F *)coeffs + air->num_transition + air->num_assertions);
  std::vector<F> ce = evaluate_constraints(*air, tc, n, o->blowup, cc);
  memcpy(out_ce, ce.data(), 16 * ce.size());
}
// transition constraints for a single frame (AIR unit-test style: vm/air/*.rs tests)
void zo_eval_transition_frame(const zkl_air_params *air, const zkl_fe *cur, const zkl_fe *next, const zkl_fe *periodic,
                              zkl_fe *out) {
  std::vector<F> r(air->num_transition);
  eval_transition(*air, (const F *)cur, (const F *)next, (const F *)periodic, r.data());
  memcpy(out, r.data(), 16 * r.size());
}
// periodic values at point x (vm/air/mod.rs:520-592 polys evaluated the Winterfell way)
void zo_periodic_at(size_t n, const zkl_fe *x, zkl_fe *out32) {
  PeriodicPolys pp(n);
  F v[32]; pp.eval_at(fe_in(*x), v);
  memcpy(out32, v, sizeof v);
}

// ---- hook 3 restatement: DefaultConstraintCommitment::new ---------------------------
void zo_constraint_commit(const zkl_fe *ce_evals, size_t n, size_t ce_blowup, size_t num_cols, const zkl_options *o,
                          uint8_t root[32], zkl_fe *out_poly) {
  std::vector<F> ce((const F *)ce_evals, (const F *)ce_evals + n * ce_blowup);
  ConstraintCommitment cm = commit_constraints(H(), ce, n, num_cols, o->blowup, o->num_partitions, o->hash_rate);
  memcpy(root, cm.tree.root().b, 32);
  if (out_poly) for (size_t c = 0; c < num_cols; c++) memcpy(out_poly + c * n, cm.polys[c].data(), 16 * n);
}

// ---- whole proof -----------------------------------------------------------------------
int zo_prove_segment(const zkl_fe *trace, size_t width, size_t n, const zkl_air_params *air, const zkl_fe *seed,
                     size_t num_seed, const zkl_options *o, uint8_t *out_proof, size_t *out_len,
                     zkl_proof_summary *summary) {
  std::vector<F> se((const F *)seed, (const F *)seed + num_seed);
  ProofOut po = prove_segment(H(), (const F *)trace, width, n, *air, se, *o);
  if (summary) *summary = po.summary;
  if (out_proof) {
    if (po.bytes.size() > *out_len) { *out_len = po.bytes.size(); return 2; }
    memcpy(out_proof, po.bytes.data(), po.bytes.size());
  }
  if (out_len) *out_len = po.bytes.size();
  return 0;
}
// restated winter-verifier over the serialized proof; returns 0 if accepted, else a stage code
int zo_verify_segment(const uint8_t *proof, size_t len, size_t width, size_t n, const zkl_air_params *air,
                      const zkl_fe *seed, size_t num_seed, const zkl_options *o, char *msg, size_t msg_cap) {
  std::vector<F> se((const F *)seed, (const F *)seed + num_seed);
  std::string m;
  int rc = verify_segment(H(), proof, len, width, n, *air, se, *o, m);
  if (msg && msg_cap) { snprintf(msg, msg_cap, "%s", m.c_str()); }
  return rc;
}
// `threads` segments in flight, one thread each (prove.rs:1020-1048 rayon pool shape)
double zo_prove_many(const zkl_fe *traces, size_t count, size_t width, size_t n, const zkl_air_params *air,
                     const zkl_fe *seed, size_t num_seed, const zkl_options *o, int threads) {
  std::vector<F> se((const F *)seed, (const F *)seed + num_seed);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  std::atomic<size_t> nexti(0);
  for (int t = 0; t < threads; t++)
    th.emplace_back([&] {
      for (;;) {
        size_t i = nexti++;
        if (i >= count) break;
        ProofOut po = prove_segment(H(), (const F *)traces + i * width * n, width, n, *air, se, *o);
        (void)po;
      }
    });
  for (auto &x : th) x.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
