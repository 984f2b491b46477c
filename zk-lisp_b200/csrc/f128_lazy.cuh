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

// 256-bit square: 4 squares + 6 cross products (10 wide multiplies instead of 16). Squares and the cross products
// of one parity sit at disjoint limb positions, so they are assembled by concatenation; x^2 = S + 2 X.
ZKL_HD void lazy_sqr_wide(const fe &a, uint32_t x[8]) {
  const uint64_t s0 = (uint64_t)a.l[0] * a.l[0], s1 = (uint64_t)a.l[1] * a.l[1], s2 = (uint64_t)a.l[2] * a.l[2], s3 = (uint64_t)a.l[3] * a.l[3];
  const uint64_t o12 = (uint64_t)a.l[0] * a.l[1], o34 = (uint64_t)a.l[0] * a.l[3], o56 = (uint64_t)a.l[2] * a.l[3];   // limbs 1, 3, 5
  const uint64_t m34 = (uint64_t)a.l[1] * a.l[2];                                                                      // limb 3
  const uint64_t e23 = (uint64_t)a.l[0] * a.l[2], e45 = (uint64_t)a.l[1] * a.l[3];                                     // limbs 2, 4
  // X = O + M + E (limbs 1..7), X < 2^255
  uint32_t x1 = (uint32_t)o12;
  uint32_t x2 = p_add_cc((uint32_t)(o12 >> 32), (uint32_t)e23);
  uint32_t x3 = p_addc_cc((uint32_t)o34, (uint32_t)(e23 >> 32));
  uint32_t x4 = p_addc_cc((uint32_t)(o34 >> 32), (uint32_t)e45);
  uint32_t x5 = p_addc_cc((uint32_t)o56, (uint32_t)(e45 >> 32));
  uint32_t x6 = p_addc_cc((uint32_t)(o56 >> 32), 0);
  uint32_t x7 = p_addc(0, 0);
  x3 = p_add_cc(x3, (uint32_t)m34); x4 = p_addc_cc(x4, (uint32_t)(m34 >> 32)); x5 = p_addc_cc(x5, 0); x6 = p_addc_cc(x6, 0); x7 = p_addc(x7, 0);
  // 2X
  uint32_t d1 = x1 << 1, d2 = p_shf_l(x1, x2, 1), d3 = p_shf_l(x2, x3, 1), d4 = p_shf_l(x3, x4, 1), d5 = p_shf_l(x4, x5, 1),
           d6 = p_shf_l(x5, x6, 1), d7 = p_shf_l(x6, x7, 1);
  x[0] = (uint32_t)s0;
  x[1] = p_add_cc((uint32_t)(s0 >> 32), d1); x[2] = p_addc_cc((uint32_t)s1, d2); x[3] = p_addc_cc((uint32_t)(s1 >> 32), d3);
  x[4] = p_addc_cc((uint32_t)s2, d4); x[5] = p_addc_cc((uint32_t)(s2 >> 32), d5); x[6] = p_addc_cc((uint32_t)s3, d6);
  x[7] = p_addc((uint32_t)(s3 >> 32), d7);
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
// (Measured on B200: the word-aligned form 2^128 ≡ 11520*2^32 - 1, which needs no funnel shifts and 7 fewer instructions per
//  reduction, runs the permutation kernel 6% SLOWER (316 vs 337 M perm/s, tools/tc_probe A/B run) — ptxas schedules it worse.)
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
// Variant of the first fold as a multiply-accumulate: y = lo + hi * C with the two-word constant C = 2^128 mod p
// = [0xFFFFFFFF, 0x2CFF], accumulated onto lo in even/odd pairs (8 wide multiplies, no shifts, no subtraction chain).
// Measured on B200 (tools/tc_probe, -DZKL_REDUCE_MAC): 326 vs 336 M perm/s for the shift/subtract form above, although it
// executes fewer instructions — carry-chained IMAD.WIDE.X are dearer than the IADD3.X they replace. Kept for reference, off.
ZKL_HD fe lazy_reduce8_mac(const uint32_t x[8]) {
  uint32_t e0 = x[0], e1 = x[1], e2 = x[2], e3 = x[3], e4 = 0, e5 = 0;     // even pairs: limbs (0,1) (2,3) (4,5)
  uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;                                 // odd pairs: limbs (1,2) (3,4)
  uint32_t ce = 0, co = 0;                                                 // carries out of limb 3 (even) / limb 4 (odd)
  uint32_t dummy = 0;
  wacc_row(e0, e1, e2, e3, ce, x[4], x[6], ZKL_C0);       // hi0*C0 @0, hi2*C0 @2
  wacc_row(o0, o1, o2, o3, co, x[5], x[7], ZKL_C0);       // hi1*C0 @1, hi3*C0 @3
  wacc_row(o0, o1, o2, o3, co, x[4], x[6], ZKL_C1);       // hi0*C1 @1, hi2*C1 @3
  wacc_row(e2, e3, e4, e5, dummy, x[5], x[7], ZKL_C1);    // hi1*C1 @2, hi3*C1 @4   (limb 5 cannot overflow: y < 2^175)
  // y = E + (O << 32) + ce * 2^128 + co * 2^160
  uint32_t y0 = e0;
  uint32_t y1 = p_add_cc(e1, o0), y2 = p_addc_cc(e2, o1), y3 = p_addc_cc(e3, o2), y4 = p_addc_cc(e4, o3), y5 = p_addc(e5, co);
  y4 = p_add_cc(y4, ce); y5 = p_addc(y5, 0);
  // second fold: hi2 = y5:y4 < 2^47 ; v = hi2*45 < 2^53 ; d = (v << 40) - hi2 in limbs 0..2 (< 2^93)
  uint64_t v = (((uint64_t)y5 << 32) | y4) * 45u;
  uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
  uint32_t d0 = p_sub_cc(0, y4), d1 = p_subc_cc(v0 << 8, y5), d2 = p_subc(p_shf_l(v0, v1, 8), 0);
  uint32_t z0 = p_add_cc(y0, d0), z1 = p_addc_cc(y1, d1), z2 = p_addc_cc(y2, d2), z3 = p_addc_cc(y3, 0), c = p_addc(0, 0);
  lazy_fix_carry(z0, z1, z2, c);
  fe r; r.l[0] = z0; r.l[1] = z1; r.l[2] = z2; r.l[3] = z3;
  return r;
}
#ifdef ZKL_REDUCE_MAC
#define LAZY_REDUCE8 lazy_reduce8_mac
#else
#define LAZY_REDUCE8 lazy_reduce8
#endif
ZKL_HD fe lazy_mul(const fe &a, const fe &b) { uint32_t x[8]; lazy_mul_wide(a, b, x); return LAZY_REDUCE8(x); }
ZKL_HD fe lazy_sqr(const fe &a) { uint32_t x[8]; lazy_sqr_wide(a, x); return LAZY_REDUCE8(x); }

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

// Tensor-core MDS epilogue: acc[u] = sum over (k, v) of byte_u(M[i][k] * 2^(8v) mod p) * byte_v(cube_k), u = 0..15.
// PRECONDITION acc[u] <= TC_ACC_MAX = 2^24 - 2^16 - 1, so that a pair acc[2j] + acc[2j+1]*2^8 fits 32 bits. The bound
// holds iff 255 * (largest row sum of the byte matrix) <= TC_ACC_MAX, which zkl_suite_load verifies for the loaded
// constants (hasher suite: 13 497 915). V = sum_u acc[u] * 2^(8u) + rc (< 2^146) is folded once to a lazy element.
//   pairs P_j at bit 16 j: even j concatenate into limbs 0..3, odd j concatenate and shift by 16 — no multiplies.
#define TC_ACC_MAX 16711679u
// same without an addend (radix-16 NTT stages, kernels_tcntt.cuh)
ZKL_HD fe tc_epilogue_plain(const uint32_t acc[16]) {
  uint32_t P[8];
#pragma unroll
  for (int j = 0; j < 8; j++) P[j] = acc[2 * j] + (acc[2 * j + 1] << 8);
  uint32_t o0 = P[1] << 16, o1 = p_shf_l(P[1], P[3], 16), o2 = p_shf_l(P[3], P[5], 16), o3 = p_shf_l(P[5], P[7], 16), o4 = P[7] >> 16;
  uint32_t w0 = p_add_cc(P[0], o0), w1 = p_addc_cc(P[2], o1), w2 = p_addc_cc(P[4], o2), w3 = p_addc_cc(P[6], o3), w4 = p_addc(o4, 0);
  uint32_t t = w4 * 45u;
  uint32_t d0 = p_sub_cc(0, w4), d1 = p_subc(t << 8, 0);
  uint32_t z0 = p_add_cc(w0, d0), z1 = p_addc_cc(w1, d1), z2 = p_addc_cc(w2, 0), z3 = p_addc_cc(w3, 0), c = p_addc(0, 0);
  lazy_fix_carry(z0, z1, z2, c);
  fe r; r.l[0] = z0; r.l[1] = z1; r.l[2] = z2; r.l[3] = z3;
  return r;
}
ZKL_HD fe tc_epilogue(const uint32_t acc[16], const fe &rc) {
  uint32_t P[8];
#pragma unroll
  for (int j = 0; j < 8; j++) P[j] = acc[2 * j] + (acc[2 * j + 1] << 8);
  uint32_t o0 = P[1] << 16, o1 = p_shf_l(P[1], P[3], 16), o2 = p_shf_l(P[3], P[5], 16), o3 = p_shf_l(P[5], P[7], 16), o4 = P[7] >> 16;
  uint32_t w0 = p_add_cc(P[0], o0), w1 = p_addc_cc(P[2], o1), w2 = p_addc_cc(P[4], o2), w3 = p_addc_cc(P[6], o3), w4 = p_addc(o4, 0);
  w0 = p_add_cc(w0, rc.l[0]); w1 = p_addc_cc(w1, rc.l[1]); w2 = p_addc_cc(w2, rc.l[2]); w3 = p_addc_cc(w3, rc.l[3]);
  w4 = p_addc(w4, 0);                                   // < 2^17
  // fold: V ≡ lo + (w4*45 << 40) - w4 ; t = w4*45 < 2^23 so (t << 40) lives in limb 1 only
  uint32_t t = w4 * 45u;
  uint32_t d0 = p_sub_cc(0, w4), d1 = p_subc(t << 8, 0);
  uint32_t z0 = p_add_cc(w0, d0), z1 = p_addc_cc(w1, d1), z2 = p_addc_cc(w2, 0), z3 = p_addc_cc(w3, 0), c = p_addc(0, 0);
  lazy_fix_carry(z0, z1, z2, c);
  fe r; r.l[0] = z0; r.l[1] = z1; r.l[2] = z2; r.l[3] = z3;
  return r;
}
