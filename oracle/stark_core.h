// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/f128.h header).
// CPU restatement of the Winterfell 0.13.1 building blocks the zk-lisp prover drives
// (winter-math fft/polynom, winter-crypto MerkleTree / BatchMerkleProof / DefaultRandomCoin).
// The crates are un-vendored (Cargo.lock:1212-1305); behaviour is restated from the
// published algorithms and anchored on the reference's own restatements:
//   FS transcript order            src/agg/fs.rs:67-228
//   Merkle node order / path fold  src/agg/child.rs:1049-1068
//   FRI position folding           src/agg/child.rs:1072-1099
// Items marked [WF-mem] are recollection of the crate source.
// PARITY UNPINNED at the byte level.
#pragma once
#include <algorithm>
#include <map>
#include <vector>

#include "poseidon.h"

namespace zkl {

// ---------------------------------------------------------------- polynomials / FFT
// All transforms are exact field maps, so only the mathematical definition matters:
//   evaluate:    out[i] = p(offset * w^i)
//   interpolate: inverse of the above.
static inline int ilog2(size_t n) { int k = 0; while (((size_t)1 << k) < n) k++; return k; }

static inline void bit_reverse(std::vector<F> &a) {
  size_t n = a.size(); int lg = ilog2(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = 0; for (int b = 0; b < lg; b++) if (i >> b & 1) j |= (size_t)1 << (lg - 1 - b);
    if (i < j) std::swap(a[i], a[j]);
  }
}
// in-place radix-2 NTT, natural in / natural out, root = primitive n-th root
static inline void ntt(std::vector<F> &a, F root) {
  size_t n = a.size();
  bit_reverse(a);
  for (size_t len = 2; len <= n; len <<= 1) {
    F wl = f_pow(root, n / len);
    std::vector<F> tw(len / 2); tw[0] = f_one();
    for (size_t k = 1; k < len / 2; k++) tw[k] = f_mul(tw[k - 1], wl);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; k++) {
        F u = a[i + k], v = f_mul(a[i + k + len / 2], tw[k]);
        a[i + k] = f_add(u, v); a[i + k + len / 2] = f_sub(u, v);
      }
  }
}
// winter-math fft::interpolate_poly / interpolate_poly_with_offset
static inline void interpolate(std::vector<F> &a, F offset = f_one()) {
  size_t n = a.size();
  F root = f_root_of_unity(ilog2(n));
  ntt(a, f_inv(root));
  F ninv = f_inv(f_from_u64(n)), oinv = f_inv(offset), s = ninv;
  for (size_t i = 0; i < n; i++) { a[i] = f_mul(a[i], s); s = f_mul(s, oinv); }
}
// winter-math fft::evaluate_poly_with_offset: evaluations of p over offset*<w_N>, N = blowup * len
static inline std::vector<F> evaluate_over(const std::vector<F> &coef, size_t domain, F offset) {
  std::vector<F> a(domain, f_zero());
  F s = f_one();
  for (size_t i = 0; i < coef.size(); i++) { a[i] = f_mul(coef[i], s); s = f_mul(s, offset); }
  ntt(a, f_root_of_unity(ilog2(domain)));
  return a;
}
// winter-math polynom::eval (Horner)
static inline F poly_eval(const F *c, size_t n, F x) {
  F acc = f_zero();
  for (size_t i = n; i-- > 0;) acc = f_add(f_mul(acc, x), c[i]);
  return acc;
}

// ---------------------------------------------------------------- Merkle tree
// winter-crypto MerkleTree::new: nodes[i] = merge(nodes[2i], nodes[2i+1]), leaves are the
// bottom layer, root = nodes[1]  [WF-mem]; node order confirmed by src/agg/child.rs:1049-1068.
struct MerkleTree {
  std::vector<Digest> nodes;   // size n, nodes[0] unused
  std::vector<Digest> leaves;  // size n
  int depth = 0;
  MerkleTree() {}
  MerkleTree(const Hasher &h, std::vector<Digest> lv) : leaves(std::move(lv)) {
    size_t n = leaves.size(); depth = ilog2(n);
    nodes.assign(n, Digest());
    for (size_t i = n / 2; i < n; i++) {
      size_t j = (i - n / 2) * 2; nodes[i] = h.merge(leaves[j], leaves[j + 1]);
    }
    for (size_t i = n / 2; i-- > 1;) nodes[i] = h.merge(nodes[2 * i], nodes[2 * i + 1]);
    if (n == 1) nodes.assign(2, leaves[0]);  // degenerate; not used by the prover
  }
  const Digest &root() const { return nodes[1]; }
};

// winter-crypto MerkleTree::prove_batch -> BatchMerkleProof { nodes: Vec<Vec<Digest>>, depth } [WF-mem]
struct BatchMerkleProof {
  std::vector<std::vector<Digest>> nodes;
  uint8_t depth = 0;
};
static inline BatchMerkleProof prove_batch(const MerkleTree &t, const std::vector<size_t> &indexes) {
  size_t n = t.leaves.size();
  std::map<size_t, size_t> index_map;
  for (size_t i = 0; i < indexes.size(); i++) index_map[indexes[i]] = i;
  // normalize_indexes: BTreeSet of even-aligned pair starts -> ascending order [WF-mem]
  std::vector<size_t> norm;
  for (size_t idx : indexes) norm.push_back(idx & ~(size_t)1);
  std::sort(norm.begin(), norm.end());
  norm.erase(std::unique(norm.begin(), norm.end()), norm.end());
  BatchMerkleProof p; p.depth = (uint8_t)t.depth;
  std::vector<size_t> next;
  for (size_t e : norm) {
    std::vector<Digest> missing;
    for (size_t i = e; i < e + 2; i++)
      if (!index_map.count(i)) missing.push_back(t.leaves[i]);
    p.nodes.push_back(missing);
    next.push_back((e + n) >> 1);
  }
  for (int d = 1; d < t.depth; d++) {
    std::vector<size_t> cur = next; next.clear();
    size_t i = 0;
    while (i < cur.size()) {
      size_t sib = cur[i] ^ 1;
      if (i + 1 < cur.size() && cur[i + 1] == sib) i++;
      else p.nodes[i].push_back(t.nodes[sib]);
      next.push_back(sib >> 1);
      i++;
    }
  }
  return p;
}
// src/agg/child.rs:1049-1068 merkle_root_from_leaf (single-path check used by tests)
static inline Digest merkle_root_from_leaf(const Hasher &h, Digest acc, size_t idx, const std::vector<Digest> &sibs) {
  for (auto &s : sibs) { acc = (idx & 1) ? h.merge(s, acc) : h.merge(acc, s); idx >>= 1; }
  return acc;
}
static inline std::vector<Digest> merkle_path(const MerkleTree &t, size_t idx) {
  std::vector<Digest> sibs; size_t n = t.leaves.size();
  sibs.push_back(t.leaves[idx ^ 1]);
  size_t i = (idx + n) >> 1;
  while (i > 1) { sibs.push_back(t.nodes[i ^ 1]); i >>= 1; }
  return sibs;
}

// ---------------------------------------------------------------- random coin
// winter-crypto DefaultRandomCoin [WF-mem, SURVEY Appendix A]; call order pinned by src/agg/fs.rs.
struct Coin {
  const Hasher &h; Digest seed; uint64_t counter = 0;
  Coin(const Hasher &hh, const std::vector<F> &seed_elems) : h(hh) {
    seed = h.hash_elements(seed_elems.data(), seed_elems.size());
  }
  void reseed(const Digest &d) { seed = h.merge(seed, d); counter = 0; }
  Digest next() { counter++; return h.merge_with_int(seed, counter); }
  // draw::<BaseElement>: first 16 bytes, always canonical for this hasher
  F draw() { return digest_fe(next()); }
  uint32_t check_leading_zeros(uint64_t v) const {
    Digest d = h.merge_with_int(seed, v);
    uint64_t head; memcpy(&head, d.b, 8);
    return head == 0 ? 64 : (uint32_t)__builtin_ctzll(head);
  }
  std::vector<size_t> draw_integers(size_t k, size_t domain, uint64_t nonce) {
    seed = h.merge_with_int(seed, nonce); counter = 0;
    std::vector<size_t> out;
    for (size_t i = 0; i < k; i++) {
      Digest d = next(); uint64_t v; memcpy(&v, d.b, 8);
      out.push_back((size_t)(v & (uint64_t)(domain - 1)));
    }
    return out;
  }
};

// src/agg/child.rs:1072-1099 fold_positions_usize (winter-fri fold_positions)
static inline std::vector<size_t> fold_positions(const std::vector<size_t> &pos, size_t src_domain, size_t folding) {
  size_t tgt = src_domain / folding;
  std::vector<size_t> r;
  for (size_t p : pos) { size_t i = p % tgt; if (std::find(r.begin(), r.end(), i) == r.end()) r.push_back(i); }
  return r;
}

}  // namespace zkl
