// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may link or call anything under oracle/.
//
// CPU restatement of winter-math 0.13.1 `fields::f128::BaseElement`
// (crates.io dependency pinned at /root/reference/Cargo.lock:1212-1305; source NOT vendored,
// so the definition below is the published one: p = 2^128 - 45*2^40 + 1, GENERATOR = 3,
// TWO_ADICITY = 40, canonical little-endian u128 representation).
// Call sites in the reference: zk-lisp-proof-winterfell/src/prove.rs:20,26,
// src/poseidon/hasher.rs:21, src/vm/air/mod.rs:40.
//
// PARITY UNPINNED at the byte level: the reference ships no golden vectors for this
// path. The arithmetic here is pinned against Python big-int
// arithmetic in tests/test_oracle_field.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace zkl {

typedef unsigned __int128 u128;

struct F {
  u128 v;  // always canonical: 0 <= v < P
};

static inline u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }
// p = 0xffffffffffffffffffffd30000000001
static const u128 P = mk128(0xffffffffffffffffULL, 0xffffd30000000001ULL);
// 2^128 mod p = 45*2^40 - 1
static const u128 C128 = ((u128)45 << 40) - 1;

static inline F f_from_u128(u128 x) { F r; r.v = x >= P ? x - P : x; return r; }  // x mod p (x < 2^128 < 2p)
static inline F f_from_u64(uint64_t x) { F r; r.v = x; return r; }
static inline F f_zero() { F r; r.v = 0; return r; }
static inline F f_one() { F r; r.v = 1; return r; }
static inline bool f_eq(F a, F b) { return a.v == b.v; }

static inline F f_add(F a, F b) {
  u128 s = a.v + b.v;            // may wrap mod 2^128
  bool carry = s < a.v;
  if (carry) s += C128;          // s + 2^128 ≡ s + C128; s < 2p-2^128 so no second wrap
  if (s >= P) s -= P;
  F r; r.v = s; return r;
}
static inline F f_sub(F a, F b) {
  F r; r.v = a.v >= b.v ? a.v - b.v : a.v + (P - b.v); return r;
}
static inline F f_neg(F a) { F r; r.v = a.v ? P - a.v : 0; return r; }

// 128x128 -> 256 (hi, lo)
static inline void mul_wide(u128 a, u128 b, u128 &hi, u128 &lo) {
  uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64);
  uint64_t b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
  u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
  u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;  // < 3*2^64
  lo = (mid << 64) | (uint64_t)p00;
  hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
}

// reduce hi*2^128 + lo  (mod p) using 2^128 ≡ C128 (46 bits)
static inline F f_reduce256(u128 hi, u128 lo) {
  // hi*C128 with hi = h1*2^64 + h0: two 64x46-bit products (each < 2^110)
  const uint64_t C = (uint64_t)C128;
  u128 p0 = (u128)(uint64_t)hi * C, p1 = (u128)(uint64_t)(hi >> 64) * C;
  // value = lo + p0 + p1*2^64
  u128 s = lo + p0; u128 c = s < lo;
  u128 t = p1 << 64; u128 s1 = s + t; c += s1 < s;
  u128 h2 = (p1 >> 64) + c;             // < 2^47
  // fold again: h2*2^128 ≡ h2*C (< 2^93)
  u128 s2 = s1 + h2 * C128; bool c2 = s2 < s1;
  if (c2) s2 += C128;                   // wrapped value is < 2^93, cannot wrap again
  if (s2 >= P) s2 -= P;
  F r; r.v = s2; return r;
}

static inline F f_mul(F a, F b) {
  u128 hi, lo; mul_wide(a.v, b.v, hi, lo);
  return f_reduce256(hi, lo);
}
static inline F f_sqr(F a) { return f_mul(a, a); }
static inline F f_cube(F a) { return f_mul(f_mul(a, a), a); }

static inline F f_pow(F b, u128 e) {
  F r = f_one();
  while (e) { if (e & 1) r = f_mul(r, b); b = f_mul(b, b); e >>= 1; }
  return r;
}
static inline F f_inv(F a) { return a.v == 0 ? f_zero() : f_pow(a, P - 2); }  // winter: inv(0) = 0

static const int TWO_ADICITY = 40;
static inline F f_generator() { return f_from_u64(3); }
// root of unity of order 2^k (StarkField::get_root_of_unity): G^(2^(40-k)) with
// G = 3^((p-1)/2^40) = 23953097886125630542083529559205016746 (winter-math TWO_ADIC_ROOT_OF_UNITY;
// the decimal value is asserted in tests/test_oracle_field.py)
static inline F f_root_of_unity(int k) {
  static F g40 = f_pow(f_generator(), (P - 1) >> 40);
  F r = g40;
  for (int i = k; i < TWO_ADICITY; i++) r = f_sqr(r);
  return r;
}

static inline void f_to_le16(F a, uint8_t out[16]) { memcpy(out, &a.v, 16); }
static inline F f_from_le16_canon(const uint8_t in[16]) { F r; memcpy(&r.v, in, 16); return r; }

}  // namespace zkl
