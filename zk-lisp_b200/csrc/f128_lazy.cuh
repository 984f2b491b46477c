// f128_lazy.cuh — "lazy" f128 arithmetic for the tensor-core permutation (tcperm.cuh).
// A lazy element is ANY 128-bit integer v in [0, 2^128) standing for v mod p (p = 2^128 - 45*2^40 + 1,
// so every residue has one or two representatives). The tensor-core MDS product only needs the BYTES of
// some representative of x^3, so the full canonical reduction (compare with p + conditional subtract)
// is skipped inside the 27 rounds and done once when a digest leaves the sponge.
//
// The limb code is written against the PTX carry-flag primitives of f128.cuh (p_add_cc, ...). On the host
// the same primitives are emulated with a thread-local carry flag, so tests/host_f128.cc runs the DEVICE
// limb algorithms on the CPU against Python integers.
#pragma once
#include "f128.cuh"

#if !defined(__CUDA_ARCH__)
static thread_local uint32_t zkl_cc_ = 0;  // emulated carry/borrow flag
inline uint32_t p_add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; zkl_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t p_addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + zkl_cc_; zkl_cc_ = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t p_addc(uint32_t a, uint32_t b) { return a + b + zkl_cc_; }
inline uint32_t p_sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; zkl_cc_ = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t p_subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - zkl_cc_; zkl_cc_ = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t p_subc(uint32_t a, uint32_t b) { return a - b - zkl_cc_; }
inline uint32_t p_shf_l(uint32_t lo, uint32_t hi, uint32_t n) { return n == 0 ? hi : ((hi << n) | (lo >> (32 - n))); }
#endif

// 256-bit product of two 128-bit values (not necessarily canonical)
ZKL_HD void lazy_mul_wide(const fe &a, const fe &b, uint32_t x[8]) {
  wide_acc w; wacc_zero(w); wacc_mad(w, a, b);
  uint32_t t[9]; wacc_merge(w, t);
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = t[i];
}

// conditional "+C" after an addition that overflowed 2^128 (2^128 ≡ C = 45*2^40 - 1): the wrapped value is
// small enough in every caller (< 2^96) that the correction cannot overflow again and never reaches limb 3.
ZKL_HD void lazy_fix_carry(uint32_t &z0, uint32_t &z1, uint32_t &z2, uint32_t cout) {
  uint32_t m = 0u - cout;
  z0 = p_add_cc(z0, m);               // m & 0xFFFFFFFF
  z1 = p_addc_cc(z1, m & ZKL_C1);
  z2 = p_addc(z2, 0);
}

// x (256 bits) -> lazy element. Two folds of the high part with 2^128 ≡ 45*2^40 - 1.
//   y = lo + (hi*45 << 40) - hi            < 2^175   (never negative: hi*45*2^40 >= hi)
//   z = y_lo + (hi2*45 << 40) - hi2        hi2 = y >> 128 < 2^47 ; z < 2^128 + 2^93, one conditional +C
ZKL_HD fe lazy_reduce8(const uint32_t x[8]) {
  // t = hi * 45 : four independent wide products, one ripple add
  uint64_t P0 = (uint64_t)x[4] * 45u, P1 = (uint64_t)x[5] * 45u, P2 = (uint64_t)x[6] * 45u, P3 = (uint64_t)x[7] * 45u;
  uint32_t t0 = (uint32_t)P0;
  uint32_t t1 = p_add_cc((uint32_t)P1, (uint32_t)(P0 >> 32));
  uint32_t t2 = p_addc_cc((uint32_t)P2, (uint32_t)(P1 >> 32));
  uint32_t t3 = p_addc_cc((uint32_t)P3, (uint32_t)(P2 >> 32));
  uint32_t t4 = p_addc((uint32_t)(P3 >> 32), 0);          // < 45
  // s = t << 40 : limbs 1..5
  uint32_t s1 = t0 << 8, s2 = p_shf_l(t0, t1, 8), s3 = p_shf_l(t1, t2, 8), s4 = p_shf_l(t2, t3, 8), s5 = p_shf_l(t3, t4, 8);
  // u = s - hi (limbs 0..5), y = lo + u
  uint32_t u0 = p_sub_cc(0, x[4]), u1 = p_subc_cc(s1, x[5]), u2 = p_subc_cc(s2, x[6]), u3 = p_subc_cc(s3, x[7]),
           u4 = p_subc_cc(s4, 0), u5 = p_subc(s5, 0);
  uint32_t y0 = p_add_cc(x[0], u0), y1 = p_addc_cc(x[1], u1), y2 = p_addc_cc(x[2], u2), y3 = p_addc_cc(x[3], u3),
           y4 = p_addc_cc(u4, 0), y5 = p_addc(u5, 0);
  // second fold: hi2 = y5:y4 < 2^47 ; v = hi2*45 < 2^53 ; d = (v << 40) - hi2 in limbs 0..2 (< 2^93)
  uint64_t v = (((uint64_t)y5 << 32) | y4) * 45u;
  uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
  uint32_t d0 = p_sub_cc(0, y4), d1 = p_subc_cc(v0 << 8, y5), d2 = p_subc(p_shf_l(v0, v1, 8), 0);
  uint32_t z0 = p_add_cc(y0, d0), z1 = p_addc_cc(y1, d1), z2 = p_addc_cc(y2, d2), z3 = p_addc_cc(y3, 0), c = p_addc(0, 0);
  lazy_fix_carry(z0, z1, z2, c);
  fe r; r.l[0] = z0; r.l[1] = z1; r.l[2] = z2; r.l[3] = z3;
  return r;
}
ZKL_HD fe lazy_mul(const fe &a, const fe &b) { uint32_t x[8]; lazy_mul_wide(a, b, x); return lazy_reduce8(x); }

// a + b for lazy a, b (sponge absorb). The sum wraps at most once; the wrapped value + C stays below 2^128
// unless both inputs were within 2^46 of 2^128, in which case a second (exact) wrap is applied.
ZKL_HD fe lazy_add(const fe &a, const fe &b) {
  uint32_t s0 = p_add_cc(a.l[0], b.l[0]), s1 = p_addc_cc(a.l[1], b.l[1]), s2 = p_addc_cc(a.l[2], b.l[2]),
           s3 = p_addc_cc(a.l[3], b.l[3]), c = p_addc(0, 0);
  uint32_t m = 0u - c;
  uint32_t q0 = p_add_cc(s0, m), q1 = p_addc_cc(s1, m & ZKL_C1), q2 = p_addc_cc(s2, 0), q3 = p_addc_cc(s3, 0), c2 = p_addc(0, 0);
  uint32_t m2 = 0u - c2;   // only if s + C wrapped again: then q < C and q + C < 2^128
  fe r;
  r.l[0] = p_add_cc(q0, m2); r.l[1] = p_addc_cc(q1, m2 & ZKL_C1); r.l[2] = p_addc_cc(q2, 0); r.l[3] = p_addc(q3, 0);
  return r;
}
// lazy -> canonical (v < 2^128 < 2p: one conditional subtraction)
ZKL_HD fe lazy_canon(const fe &a) { return fe_from_u128_mod(a); }

// Tensor-core MDS epilogue: acc[u] = sum over (k, v) of byte_u(M[i][k] * 2^(8v) mod p) * byte_v(cube_k), u = 0..15,
// each < 2^25. V = sum_u acc[u] * 2^(8u) + rc  (< 2^146) is folded once to a lazy element.
ZKL_HD fe tc_epilogue(const uint32_t acc[16], const fe &rc) {
  // G_j = acc[4j] + acc[4j+1]*2^8 + acc[4j+2]*2^16 + acc[4j+3]*2^24  (< 2^50)
  uint64_t G[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint64_t g = (uint64_t)acc[4 * j + 3] * (1u << 24) + acc[4 * j];
    g = (uint64_t)acc[4 * j + 2] * (1u << 16) + g;
    g = (uint64_t)acc[4 * j + 1] * (1u << 8) + g;
    G[j] = g;
  }
  uint32_t w0 = (uint32_t)G[0];
  uint32_t w1 = p_add_cc((uint32_t)(G[0] >> 32), (uint32_t)G[1]);
  uint32_t w2 = p_addc_cc((uint32_t)(G[1] >> 32), (uint32_t)G[2]);
  uint32_t w3 = p_addc_cc((uint32_t)(G[2] >> 32), (uint32_t)G[3]);
  uint32_t w4 = p_addc((uint32_t)(G[3] >> 32), 0);
  w0 = p_add_cc(w0, rc.l[0]); w1 = p_addc_cc(w1, rc.l[1]); w2 = p_addc_cc(w2, rc.l[2]); w3 = p_addc_cc(w3, rc.l[3]);
  w4 = p_addc(w4, 0);                                   // < 2^19
  // fold: V ≡ lo + (w4*45 << 40) - w4 ; t = w4*45 < 2^24 so (t << 40) lives in limb 1 only
  uint32_t t = w4 * 45u;
  uint32_t d0 = p_sub_cc(0, w4), d1 = p_subc(t << 8, 0);
  uint32_t z0 = p_add_cc(w0, d0), z1 = p_addc_cc(w1, d1), z2 = p_addc_cc(w2, 0), z3 = p_addc_cc(w3, 0), c = p_addc(0, 0);
  lazy_fix_carry(z0, z1, z2, c);
  fe r; r.l[0] = z0; r.l[1] = z1; r.l[2] = z2; r.l[3] = z3;
  return r;
}
