// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/f128.h header).
// CPU restatement of zk-lisp's Poseidon-style parameter derivation and sponge hasher:
//   /root/reference/zk-lisp-proof-winterfell/src/poseidon/mod.rs   (suite derivation)
//   /root/reference/zk-lisp-proof-winterfell/src/poseidon/hasher.rs (PoseidonHasher)
//   /root/reference/zk-lisp-proof-winterfell/src/utils.rs:30-66,346-390 (byte folds)
// PARITY UNPINNED at the byte level (no golden vectors in the reference); cross-checked
// against the independent Python derivation in zk-lisp_b200/params.py by tests/.
#pragma once
#include <string>
#include <vector>

#include "blake3.h"
#include "f128.h"

namespace zkl {

static const int ROUNDS = 27;  // vm/layout.rs:22 POSEIDON_ROUNDS; hasher.rs:235-241 default

struct Suite {
  F dom[2];
  F mds[12][12];
  F rc[ROUNDS][12];
};
struct RomSuite {
  F mds[3][3];
  F rc[ROUNDS][3];
};

// utils.rs:34-45
static inline F pow2_64() { F r; r.v = (u128)1 << 64; return r; }

// poseidon/mod.rs:421-440  ro_from_slices
static inline F ro_from_slices(const std::string &domain, const std::vector<std::vector<uint8_t>> &parts) {
  std::vector<uint8_t> buf(domain.begin(), domain.end());
  for (auto &p : parts) buf.insert(buf.end(), p.begin(), p.end());
  uint8_t d[32];
  Blake3::hash(buf.data(), buf.size(), d);
  uint64_t lo, hi;
  memcpy(&lo, d, 8); memcpy(&hi, d + 8, 8);
  return f_add(f_from_u64(lo), f_mul(f_from_u64(hi), pow2_64()));
}

static inline std::vector<uint8_t> bytes_of(const uint8_t *p, size_t n) { return std::vector<uint8_t>(p, p + n); }
static inline std::vector<uint8_t> le32(uint32_t v) {
  return {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
}

// poseidon/mod.rs:114-131 / 220-237  derive_points
static inline std::vector<F> derive_points(const std::string &domain, const uint8_t sid[32], size_t n) {
  std::vector<F> pts;
  uint32_t ctr = 0;
  while (pts.size() < n) {
    F cand = ro_from_slices(domain, {bytes_of(sid, 32), {(uint8_t)pts.size()}, le32(ctr)});
    bool dup = false;
    for (auto &q : pts) if (f_eq(q, cand)) dup = true;
    if (cand.v != 0 && !dup) pts.push_back(cand); else ctr++;
  }
  return pts;
}

// poseidon/mod.rs:56-75, 111-217
static inline Suite derive_suite(const uint8_t sid[32]) {
  Suite s;
  s.dom[0] = ro_from_slices("zkl/poseidon2/dom/c0", {bytes_of(sid, 32)});
  s.dom[1] = ro_from_slices("zkl/poseidon2/dom/c1", {bytes_of(sid, 32)});
  std::vector<F> x = derive_points("zkl/poseidon2/mds/x", sid, 12);
  std::vector<F> y = derive_points("zkl/poseidon2/mds/y", sid, 12);
  uint32_t adj = 0;
  for (;;) {  // mod.rs:139-173: re-derive y until no x_i + y_j == 0
    bool ok = true;
    for (auto &xi : x) for (auto &yj : y) if (f_add(xi, yj).v == 0) ok = false;
    if (ok) break;
    for (size_t j = 0; j < 12; j++) {
      F cand = ro_from_slices("zkl/poseidon2/mds/y", {bytes_of(sid, 32), {(uint8_t)j}, le32(adj)});
      y[j] = cand.v == 0 ? f_one() : cand;
    }
    adj++;
  }
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++) s.mds[i][j] = f_inv(f_add(x[i], y[j]));
  for (int r = 0; r < ROUNDS; r++)
    for (int l = 0; l < 12; l++)
      s.rc[r][l] = ro_from_slices("zkl/poseidon2/rc", {bytes_of(sid, 32), {(uint8_t)r}, {(uint8_t)l}});
  return s;
}

// poseidon/mod.rs:186-197, 219-251
static inline RomSuite derive_rom_suite(const uint8_t sid[32]) {
  RomSuite s;
  std::vector<F> x = derive_points("zkl/rom3/mds/x", sid, 3);
  std::vector<F> y = derive_points("zkl/rom3/mds/y", sid, 3);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) s.mds[i][j] = f_inv(f_add(x[i], y[j]));
  for (int r = 0; r < ROUNDS; r++)
    for (int l = 0; l < 3; l++)
      s.rc[r][l] = ro_from_slices("zkl/rom3/rc", {bytes_of(sid, 32), {(uint8_t)r}, {(uint8_t)l}});
  return s;
}

// hasher.rs:173-190  one full permutation
static inline void permute(const Suite &ps, F st[12]) {
  for (int r = 0; r < ROUNDS; r++) {
    F c[12], y[12];
    for (int i = 0; i < 12; i++) c[i] = f_cube(st[i]);
    for (int i = 0; i < 12; i++) {
      F acc = f_zero();
      for (int k = 0; k < 12; k++) acc = f_add(acc, f_mul(ps.mds[i][k], c[k]));
      y[i] = f_add(acc, ps.rc[r][i]);
    }
    memcpy(st, y, sizeof y);
  }
}

// utils.rs:50-74: be_from_le_bytes16 = (u128 LE) mod p ; utils.rs:359-371 fold_bytes32_to_fe
static inline F fold_bytes32(const uint8_t b[32]) {
  u128 lo, hi; memcpy(&lo, b, 16); memcpy(&hi, b + 16, 16);
  return f_add(f_from_u128(lo), f_mul(f_from_u128(hi), pow2_64()));
}
// utils.rs:346-355 be_from_le8
static inline F be_from_le8(const uint8_t b[32]) {
  uint64_t lo, hi; memcpy(&lo, b, 8); memcpy(&hi, b + 8, 8);
  return f_add(f_from_u64(lo), f_mul(f_from_u64(hi), pow2_64()));
}

struct Digest { uint8_t b[32]; };
// utils.rs:375-381 fe_to_bytes_fold
static inline Digest digest_of(F x) { Digest d; memset(d.b, 0, 32); memcpy(d.b, &x.v, 16); return d; }
// element carried by a digest (first 16 bytes, canonical)
static inline F digest_fe(const Digest &d) { F r; memcpy(&r.v, d.b, 16); return r; }

// hasher.rs:144-231 ro_bytes_sponge_custom_rounds (rounds fixed at 27)
struct Sponge {
  const Suite &ps; F st[12]; int lane = 0;
  Sponge(const Suite &s, const char *domain) : ps(s) {
    for (int i = 0; i < 12; i++) st[i] = f_zero();
    st[10] = ps.dom[0]; st[11] = ps.dom[1];
    uint8_t dbuf[32]; memset(dbuf, 0, 32);
    size_t dl = strlen(domain); if (dl > 32) dl = 32;
    memcpy(dbuf, domain, dl);
    absorb(fold_bytes32(dbuf));
  }
  void absorb(F m) {
    st[lane] = f_add(st[lane], m);
    if (++lane == 10) { permute(ps, st); lane = 0; }
  }
  void absorb_bytes(const uint8_t *data, size_t len) {
    for (size_t i = 0; i < len; i += 32) {
      uint8_t chunk[32]; memset(chunk, 0, 32);
      size_t l = len - i < 32 ? len - i : 32;
      memcpy(chunk, data + i, l);
      absorb(fold_bytes32(chunk));
    }
  }
  F finish() { if (lane != 0) permute(ps, st); return st[0]; }
};

struct Hasher {
  Suite ps;
  Hasher() { uint8_t sid[32] = {0}; ps = derive_suite(sid); }  // hasher.rs:23 POSEIDON_SUITE_ID = [0;32]
  Digest hash_bytes(const uint8_t *d, size_t n) const {         // hasher.rs:62-70
    Sponge s(ps, "zkl/winter/hash/bytes"); s.absorb_bytes(d, n); return digest_of(s.finish());
  }
  Digest merge(const Digest &a, const Digest &b) const {        // hasher.rs:72-85
    uint8_t buf[64]; memcpy(buf, a.b, 32); memcpy(buf + 32, b.b, 32);
    Sponge s(ps, "zkl/winter/hash/merge"); s.absorb_bytes(buf, 64); return digest_of(s.finish());
  }
  Digest merge_many(const Digest *v, size_t n) const {          // hasher.rs:87-105
    if (n == 0) { Digest z; memset(z.b, 0, 32); return z; }
    Sponge s(ps, "zkl/winter/hash/merge_many");
    s.absorb_bytes((const uint8_t *)v, n * 32); return digest_of(s.finish());
  }
  Digest merge_with_int(const Digest &seed, uint64_t value) const {  // hasher.rs:107-120
    uint8_t buf[40]; memcpy(buf, seed.b, 32); memcpy(buf + 32, &value, 8);
    Sponge s(ps, "zkl/winter/hash/merge_with_int"); s.absorb_bytes(buf, 40); return digest_of(s.finish());
  }
  Digest hash_elements(const F *e, size_t n) const {            // hasher.rs:126-139
    Sponge s(ps, "winter/hash/elements");
    s.absorb_bytes((const uint8_t *)e, n * 16); return digest_of(s.finish());
  }
};

// agg/child.rs:1025-1045 hash_row_poseidon (Winterfell row hashing with partitions)
static inline Digest hash_row(const Hasher &h, const F *row, size_t w, size_t partition_size) {
  if (partition_size == 0) return h.hash_bytes(nullptr, 0);
  std::vector<Digest> ds;
  for (size_t i = 0; i < w; i += partition_size) {
    size_t l = w - i < partition_size ? w - i : partition_size;
    ds.push_back(h.hash_elements(row + i, l));
  }
  return ds.size() == 1 ? ds[0] : h.merge_many(ds.data(), ds.size());
}

// winter-air PartitionOptions::partition_size  [WF-mem, SURVEY Appendix A]
static inline size_t partition_size(size_t num_partitions, size_t hash_rate, size_t cols, size_t ext_degree = 1) {
  if (num_partitions == 1 && hash_rate == 0) return cols;
  size_t min_ps = hash_rate / ext_degree;
  size_t ps = (cols + num_partitions - 1) / num_partitions;
  return ps > min_ps ? ps : min_ps;
}

}  // namespace zkl
