// f128.cuh — winter-math f128 arithmetic for sm_100a.
// p = 2^128 - 45*2^40 + 1 (reference field: zk-lisp-proof-winterfell/src/prove.rs:20,26).
// Elements are canonical (< p) little-endian 4 x u32 limbs; 16-byte aligned so one element is one
// 128-bit load/store. All products are 32x32->64 IMAD.WIDE chains; the special form of p makes
// reduction a multiply by 45 plus shifts (2^128 ≡ 45*2^40 - 1).
//
// The limb algorithms are written against a tiny set of carry primitives that have a PTX
// implementation on the device and a plain-C one on the host, so tests/test_f128_host.py can run
// the same limb code on the CPU (g++ -x c++) against Python integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKL_HD __host__ __device__ __forceinline__
#define ZKL_D __device__ __forceinline__
#else
#define ZKL_HD inline
#define ZKL_D inline
#endif

struct __attribute__((aligned(16))) fe {
  uint32_t l[4];
};

// ---------------------------------------------------------------- carry primitives
// add with carry-out / carry-in-out, carry held in a uint32_t (0/1).
ZKL_HD uint32_t add_co(uint32_t a, uint32_t b, uint32_t &c) {
  uint32_t r = a + b; c = r < a; return r;
}
ZKL_HD uint32_t add_cio(uint32_t a, uint32_t b, uint32_t &c) {
  uint64_t t = (uint64_t)a + b + c; c = (uint32_t)(t >> 32); return (uint32_t)t;
}
ZKL_HD uint32_t sub_bo(uint32_t a, uint32_t b, uint32_t &bw) {
  uint32_t r = a - b; bw = a < b; return r;
}
ZKL_HD uint32_t sub_bio(uint32_t a, uint32_t b, uint32_t &bw) {
  uint64_t t = (uint64_t)a - b - bw; bw = (uint32_t)(t >> 63); return (uint32_t)t;
}

#if defined(__CUDA_ARCH__)
// PTX carry-flag primitives. Consecutive asm volatile statements keep their order, and nothing the
// compiler generates itself touches CC, so a chain add_cc -> addc_cc -> ... -> addc is one IADD3/IADD3.X
// ripple on the ALU pipe (keeping carry arithmetic OFF the multiply pipe, which is the bottleneck).
__device__ __forceinline__ uint32_t p_add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t p_shf_l(uint32_t lo, uint32_t hi, uint32_t n) { uint32_t r; asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(n)); return r; }
#endif

ZKL_HD fe fe_zero() { fe r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
ZKL_HD fe fe_one() { fe r; r.l[0] = 1; r.l[1] = r.l[2] = r.l[3] = 0; return r; }
ZKL_HD fe fe_from_u64(uint64_t x) { fe r; r.l[0] = (uint32_t)x; r.l[1] = (uint32_t)(x >> 32); r.l[2] = r.l[3] = 0; return r; }
ZKL_HD bool fe_is_zero(const fe &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
ZKL_HD bool fe_eq(const fe &a, const fe &b) {
  return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3])) == 0;
}

// p limbs: [0x00000001, 0xFFFFD300, 0xFFFFFFFF, 0xFFFFFFFF]; C = 2^128 - p = [0xFFFFFFFF, 0x00002CFF, 0, 0]
#define ZKL_C0 0xFFFFFFFFu
#define ZKL_C1 0x00002CFFu

// a + b mod p (a, b canonical)
ZKL_HD fe fe_add(const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
  uint32_t s0, s1, s2, s3, c, u0, u1, u2, u3, c2;
  asm("add.cc.u32 %0, %5, %9;\n\t"
      "addc.cc.u32 %1, %6, %10;\n\t"
      "addc.cc.u32 %2, %7, %11;\n\t"
      "addc.cc.u32 %3, %8, %12;\n\t"
      "addc.u32 %4, 0, 0;"
      : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(c)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]));
  asm("add.cc.u32 %0, %5, 0xFFFFFFFF;\n\t"
      "addc.cc.u32 %1, %6, 0x2CFF;\n\t"
      "addc.cc.u32 %2, %7, 0;\n\t"
      "addc.cc.u32 %3, %8, 0;\n\t"
      "addc.u32 %4, 0, 0;"
      : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3), "=r"(c2)
      : "r"(s0), "r"(s1), "r"(s2), "r"(s3));
  bool ge = (c | c2) != 0;
  fe r; r.l[0] = ge ? u0 : s0; r.l[1] = ge ? u1 : s1; r.l[2] = ge ? u2 : s2; r.l[3] = ge ? u3 : s3;
  return r;
#else
  uint32_t c, c2; fe s, u;
  s.l[0] = add_co(a.l[0], b.l[0], c); s.l[1] = add_cio(a.l[1], b.l[1], c);
  s.l[2] = add_cio(a.l[2], b.l[2], c); s.l[3] = add_cio(a.l[3], b.l[3], c);
  u.l[0] = add_co(s.l[0], ZKL_C0, c2); u.l[1] = add_cio(s.l[1], ZKL_C1, c2);
  u.l[2] = add_cio(s.l[2], 0, c2); u.l[3] = add_cio(s.l[3], 0, c2);
  return (c | c2) ? u : s;   // v >= p  <=>  v + C >= 2^128 ; then v - p = (v + C) mod 2^128
#endif
}

// a - b mod p
ZKL_HD fe fe_sub(const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
  uint32_t d0, d1, d2, d3, bw;
  asm("sub.cc.u32 %0, %5, %9;\n\t"
      "subc.cc.u32 %1, %6, %10;\n\t"
      "subc.cc.u32 %2, %7, %11;\n\t"
      "subc.cc.u32 %3, %8, %12;\n\t"
      "subc.u32 %4, 0, 0;"
      : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(bw)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]));
  // bw = 0 or 0xFFFFFFFF; on borrow add p == subtract C (mod 2^128)
  uint32_t m0 = bw, m1 = bw & ZKL_C1;
  fe r;
  asm("sub.cc.u32 %0, %4, %8;\n\t"
      "subc.cc.u32 %1, %5, %9;\n\t"
      "subc.cc.u32 %2, %6, 0;\n\t"
      "subc.u32 %3, %7, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3])
      : "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(m0), "r"(m1));
  return r;
#else
  uint32_t bw, b2; fe d, r;
  d.l[0] = sub_bo(a.l[0], b.l[0], bw); d.l[1] = sub_bio(a.l[1], b.l[1], bw);
  d.l[2] = sub_bio(a.l[2], b.l[2], bw); d.l[3] = sub_bio(a.l[3], b.l[3], bw);
  uint32_t m = 0u - bw;
  r.l[0] = sub_bo(d.l[0], m & ZKL_C0, b2); r.l[1] = sub_bio(d.l[1], m & ZKL_C1, b2);
  r.l[2] = sub_bio(d.l[2], 0, b2); r.l[3] = sub_bio(d.l[3], 0, b2);
  return r;
#endif
}
ZKL_HD fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }
ZKL_HD fe fe_dbl(const fe &a) { return fe_add(a, a); }

// ---------------------------------------------------------------- wide products
// Unreduced accumulator for sums of up to 12 products a*b (+ one canonical addend): value < 2^260.
// Even/odd column split: the product a_j*b_i lands on limb i+j. Products with i+j even accumulate
// into the 64-bit pairs of `ev` (ev[k] = limb k), those with i+j odd into the pairs of `od`
// (od[k] = limb k+1), so every mad.lo/madc.hi pair targets an aligned register pair (one
// IMAD.WIDE.U32 with carry in SASS) and a carry only ripples along one row of two products.
// The carry that leaves a row goes into a dedicated counter for its limb (cnt[0..4] = limbs 4..8),
// never into a limb that other rows also accumulate into, so no carry is ever lost.
struct wide_acc {
  uint32_t ev[8];   // limbs 0..7
  uint32_t od[6];   // limbs 1..6
  uint32_t cnt[5];  // carry counters for limbs 4,5,6,7,8
};
ZKL_HD void wacc_zero(wide_acc &w) {
#pragma unroll
  for (int i = 0; i < 8; i++) w.ev[i] = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) w.od[i] = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) w.cnt[i] = 0;
}

// (hi:lo) += a*b + cin ; cout   (host model of mad.lo.cc / madc.hi.cc)
ZKL_HD void mad_wide_cio(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi, uint32_t &c) {
  uint64_t p = (uint64_t)a * b;
  uint64_t acc = ((uint64_t)hi << 32) | lo;
  uint64_t s = acc + p; uint32_t c1 = s < p;
  uint64_t s2 = s + c; uint32_t c2 = s2 < s;
  lo = (uint32_t)s2; hi = (uint32_t)(s2 >> 32); c = c1 | c2;
}

// one row: (e0,e1) += x0*b ; (e2,e3) += x1*b + carry ; cnt += carry
ZKL_HD void wacc_row(uint32_t &e0, uint32_t &e1, uint32_t &e2, uint32_t &e3, uint32_t &cnt, uint32_t x0, uint32_t x1,
                     uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(cnt)
      : "r"(x0), "r"(x1), "r"(b));
#else
  uint32_t c = 0;
  mad_wide_cio(x0, b, e0, e1, c); mad_wide_cio(x1, b, e2, e3, c); cnt += c;
#endif
}

// w += a * b
ZKL_HD void wacc_mad(wide_acc &w, const fe &a, const fe &b) {
  // i = 0: even j -> ev[0..3] (carry -> limb 4), odd j -> od[0..3] (limbs 1..4, carry -> limb 5)
  wacc_row(w.ev[0], w.ev[1], w.ev[2], w.ev[3], w.cnt[0], a.l[0], a.l[2], b.l[0]);
  wacc_row(w.od[0], w.od[1], w.od[2], w.od[3], w.cnt[1], a.l[1], a.l[3], b.l[0]);
  // i = 1: odd j -> ev[2..5] (carry -> limb 6), even j -> od[0..3] (carry -> limb 5)
  wacc_row(w.ev[2], w.ev[3], w.ev[4], w.ev[5], w.cnt[2], a.l[1], a.l[3], b.l[1]);
  wacc_row(w.od[0], w.od[1], w.od[2], w.od[3], w.cnt[1], a.l[0], a.l[2], b.l[1]);
  // i = 2: even j -> ev[2..5] (carry -> limb 6), odd j -> od[2..5] (limbs 3..6, carry -> limb 7)
  wacc_row(w.ev[2], w.ev[3], w.ev[4], w.ev[5], w.cnt[2], a.l[0], a.l[2], b.l[2]);
  wacc_row(w.od[2], w.od[3], w.od[4], w.od[5], w.cnt[3], a.l[1], a.l[3], b.l[2]);
  // i = 3: odd j -> ev[4..7] (carry -> limb 8), even j -> od[2..5] (carry -> limb 7)
  wacc_row(w.ev[4], w.ev[5], w.ev[6], w.ev[7], w.cnt[4], a.l[1], a.l[3], b.l[3]);
  wacc_row(w.od[2], w.od[3], w.od[4], w.od[5], w.cnt[3], a.l[0], a.l[2], b.l[3]);
}

// ---------------------------------------------------------------- reduction
// Reduce a 9-limb value x (x < 2^264) modulo p to canonical form.
//   x = lo + hi*2^128, hi = limbs 4..8 (up to 136 bits), 2^128 ≡ C = 45*2^40 - 1
//   hi*C = (hi*45 << 40) - hi                      (non-negative, < 2^182)
//   y = lo + hi*C = lo2 + hi2*2^128, hi2 < 2^55    -> second fold, then one conditional fix-up.
ZKL_HD fe reduce9(const uint32_t x[9]) {
#if defined(__CUDA_ARCH__)
  // t = hi * 45 (hi = x[4..8]): five independent wide products, then one ripple add
  uint64_t P0 = (uint64_t)x[4] * 45u, P1 = (uint64_t)x[5] * 45u, P2 = (uint64_t)x[6] * 45u, P3 = (uint64_t)x[7] * 45u,
           P4 = (uint64_t)x[8] * 45u;
  uint32_t t0 = (uint32_t)P0;
  uint32_t t1 = p_add_cc((uint32_t)P1, (uint32_t)(P0 >> 32));
  uint32_t t2 = p_addc_cc((uint32_t)P2, (uint32_t)(P1 >> 32));
  uint32_t t3 = p_addc_cc((uint32_t)P3, (uint32_t)(P2 >> 32));
  uint32_t t4 = p_addc_cc((uint32_t)P4, (uint32_t)(P3 >> 32));
  uint32_t t5 = p_addc((uint32_t)(P4 >> 32), 0);
  // s = t << 40 (limbs 1..7), u = s - hi, y = lo + u   (u >= 0; y < 2^183)
  uint32_t s1 = t0 << 8, s2 = p_shf_l(t0, t1, 8), s3 = p_shf_l(t1, t2, 8), s4 = p_shf_l(t2, t3, 8), s5 = p_shf_l(t3, t4, 8),
           s6 = p_shf_l(t4, t5, 8), s7 = t5 >> 24;
  uint32_t u0 = p_sub_cc(0, x[4]), u1 = p_subc_cc(s1, x[5]), u2 = p_subc_cc(s2, x[6]), u3 = p_subc_cc(s3, x[7]),
           u4 = p_subc_cc(s4, x[8]), u5 = p_subc_cc(s5, 0), u6 = p_subc_cc(s6, 0), u7 = p_subc(s7, 0);
  (void)u6; (void)u7;  // zero by the bound above
  uint32_t y0 = p_add_cc(x[0], u0), y1 = p_addc_cc(x[1], u1), y2 = p_addc_cc(x[2], u2), y3 = p_addc_cc(x[3], u3),
           y4 = p_addc_cc(u4, 0), y5 = p_addc(u5, 0);
  // second fold: hi2 = y5:y4 (< 2^55); v = hi2 * 45 (< 2^61); a = y_lo + (v << 40) ; z = a - hi2 (>= 0), 129 bits
  uint64_t v = (((uint64_t)y5 << 32) | y4) * 45u;
  uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
  uint32_t a0 = y0;
  uint32_t a1 = p_add_cc(y1, v0 << 8), a2 = p_addc_cc(y2, p_shf_l(v0, v1, 8)), a3 = p_addc_cc(y3, v1 >> 24), a4 = p_addc(0, 0);
  uint32_t z0 = p_sub_cc(a0, y4), z1 = p_subc_cc(a1, y5), z2 = p_subc_cc(a2, 0), z3 = p_subc_cc(a3, 0), c3 = p_subc(a4, 0);
  // Z = z + c3*2^128 ≡ z + c3*C. If c3: z is tiny and z + C is the answer; else subtract p iff z + C overflows.
  uint32_t q0 = p_add_cc(z0, ZKL_C0), q1 = p_addc_cc(z1, ZKL_C1), q2 = p_addc_cc(z2, 0), q3 = p_addc_cc(z3, 0), c5 = p_addc(0, 0);
  bool take = (c3 | c5) != 0;
  fe r; r.l[0] = take ? q0 : z0; r.l[1] = take ? q1 : z1; r.l[2] = take ? q2 : z2; r.l[3] = take ? q3 : z3;
  return r;
#else
  // t = hi * 45 : 6 limbs (hi 5 limbs -> up to 142 bits)
  uint32_t t[6];
  {
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) { acc = (uint64_t)x[4 + i] * 45u + (acc >> 32); t[i] = (uint32_t)acc; }
    t[5] = (uint32_t)(acc >> 32);
  }
  // u = (t << 40) - hi : limbs 0..7 ; (t<<40): limb k gets bits from t[k-1] (<<8) and t[k-2] (>>24)
  uint32_t u[8];
  {
    uint32_t s[8];
    s[0] = 0;
    s[1] = t[0] << 8;
#pragma unroll
    for (int k = 2; k < 7; k++) s[k] = (t[k - 1] << 8) | (t[k - 2] >> 24);
    s[7] = (t[5] >> 24);  // t[6] does not exist; t[5] < 2^14 so this is 0, kept for clarity
    uint32_t bw;
    u[0] = sub_bo(s[0], x[4], bw); u[1] = sub_bio(s[1], x[5], bw); u[2] = sub_bio(s[2], x[6], bw);
    u[3] = sub_bio(s[3], x[7], bw); u[4] = sub_bio(s[4], x[8], bw); u[5] = sub_bio(s[5], 0, bw);
    u[6] = sub_bio(s[6], 0, bw); u[7] = sub_bio(s[7], 0, bw);
  }
  // y = lo + u : 8 limbs (y < 2^183)
  uint32_t y[8], c;
  y[0] = add_co(x[0], u[0], c); y[1] = add_cio(x[1], u[1], c); y[2] = add_cio(x[2], u[2], c); y[3] = add_cio(x[3], u[3], c);
  y[4] = add_cio(0, u[4], c); y[5] = add_cio(0, u[5], c); y[6] = add_cio(0, u[6], c); y[7] = add_cio(0, u[7], c);
  // second fold: hi2 = y[4..5] (+ y[6], y[7] are zero: y < 2^183 -> hi2 < 2^55)
  // v = hi2*45 (< 2^61) ; u2 = (v << 40) - hi2  (< 2^101, limbs 0..3)
  uint64_t hi2 = ((uint64_t)y[5] << 32) | y[4];
  uint64_t v = hi2 * 45u;
  uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
  uint32_t s1 = v0 << 8, s2 = (v1 << 8) | (v0 >> 24), s3 = (v1 >> 24);
  uint32_t bw, w0, w1, w2, w3;
  w0 = sub_bo(0, y[4], bw); w1 = sub_bio(s1, y[5], bw); w2 = sub_bio(s2, 0, bw); w3 = sub_bio(s3, 0, bw);
  // z = lo2 + u2 : may carry once more out of 128 bits (then the low part is tiny and +C cannot overflow)
  fe z; uint32_t c3;
  z.l[0] = add_co(y[0], w0, c3); z.l[1] = add_cio(y[1], w1, c3); z.l[2] = add_cio(y[2], w2, c3); z.l[3] = add_cio(y[3], w3, c3);
  // fix-ups: (a) if c3: z += C ; (b) if z >= p: z -= p  (<=> z + C overflows 2^128 -> take z + C mod 2^128)
  uint32_t m = 0u - c3, c4;
  z.l[0] = add_co(z.l[0], m & ZKL_C0, c4); z.l[1] = add_cio(z.l[1], m & ZKL_C1, c4);
  z.l[2] = add_cio(z.l[2], 0, c4); z.l[3] = add_cio(z.l[3], 0, c4);
  fe q; uint32_t c5;
  q.l[0] = add_co(z.l[0], ZKL_C0, c5); q.l[1] = add_cio(z.l[1], ZKL_C1, c5);
  q.l[2] = add_cio(z.l[2], 0, c5); q.l[3] = add_cio(z.l[3], 0, c5);
  fe r;
  r.l[0] = c5 ? q.l[0] : z.l[0]; r.l[1] = c5 ? q.l[1] : z.l[1]; r.l[2] = c5 ? q.l[2] : z.l[2]; r.l[3] = c5 ? q.l[3] : z.l[3];
  return r;
#endif
}

// merge the even/odd accumulators and the carry counters into 9 plain limbs
ZKL_HD void wacc_merge(const wide_acc &w, uint32_t x[9]) {
#if defined(__CUDA_ARCH__)
  x[0] = w.ev[0];
  x[1] = p_add_cc(w.ev[1], w.od[0]); x[2] = p_addc_cc(w.ev[2], w.od[1]); x[3] = p_addc_cc(w.ev[3], w.od[2]);
  uint32_t m4 = p_addc_cc(w.ev[4], w.od[3]), m5 = p_addc_cc(w.ev[5], w.od[4]), m6 = p_addc_cc(w.ev[6], w.od[5]),
           m7 = p_addc_cc(w.ev[7], 0), m8 = p_addc(0, 0);
  x[4] = p_add_cc(m4, w.cnt[0]); x[5] = p_addc_cc(m5, w.cnt[1]); x[6] = p_addc_cc(m6, w.cnt[2]); x[7] = p_addc_cc(m7, w.cnt[3]);
  x[8] = p_addc(m8, w.cnt[4]);
  return;
#else
  uint32_t c;
  x[0] = w.ev[0];
  x[1] = add_co(w.ev[1], w.od[0], c);
  x[2] = add_cio(w.ev[2], w.od[1], c); x[3] = add_cio(w.ev[3], w.od[2], c); x[4] = add_cio(w.ev[4], w.od[3], c);
  x[5] = add_cio(w.ev[5], w.od[4], c); x[6] = add_cio(w.ev[6], w.od[5], c); x[7] = add_cio(w.ev[7], 0, c);
  x[8] = c;
  x[4] = add_co(x[4], w.cnt[0], c); x[5] = add_cio(x[5], w.cnt[1], c); x[6] = add_cio(x[6], w.cnt[2], c);
  x[7] = add_cio(x[7], w.cnt[3], c); x[8] = add_cio(x[8], w.cnt[4], c);
#endif
}
ZKL_HD fe wacc_reduce(const wide_acc &w) {
  uint32_t x[9]; wacc_merge(w, x); return reduce9(x);
}
// add a canonical element to the accumulator (round constants): goes through the even limbs
ZKL_HD void wacc_add_fe(wide_acc &w, const fe &a) {
#if defined(__CUDA_ARCH__)
  w.ev[0] = p_add_cc(w.ev[0], a.l[0]); w.ev[1] = p_addc_cc(w.ev[1], a.l[1]);
  w.ev[2] = p_addc_cc(w.ev[2], a.l[2]); w.ev[3] = p_addc_cc(w.ev[3], a.l[3]);
  w.cnt[0] = p_addc(w.cnt[0], 0);
#else
  uint32_t c;
  w.ev[0] = add_co(w.ev[0], a.l[0], c); w.ev[1] = add_cio(w.ev[1], a.l[1], c);
  w.ev[2] = add_cio(w.ev[2], a.l[2], c); w.ev[3] = add_cio(w.ev[3], a.l[3], c);
  w.cnt[0] += c;
#endif
}

ZKL_HD fe fe_mul(const fe &a, const fe &b) {
  wide_acc w; wacc_zero(w); wacc_mad(w, a, b); return wacc_reduce(w);
}
ZKL_HD fe fe_sqr(const fe &a) { return fe_mul(a, a); }
ZKL_HD fe fe_cube(const fe &a) { return fe_mul(fe_mul(a, a), a); }

ZKL_HD fe fe_pow(fe b, uint64_t e) {
  fe r = fe_one();
  while (e) { if (e & 1) r = fe_mul(r, b); b = fe_sqr(b); e >>= 1; }
  return r;
}
// a^(p-2); inv(0) = 0 like winter-math
ZKL_HD fe fe_inv(const fe &a) {
  // p - 2 = 0xffffffffffffffffffffd2ffffffffff
  const uint32_t e[4] = {0xFFFFFFFFu, 0xFFFFD2FFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  fe r = fe_one(), b = a;
  for (int i = 0; i < 128; i++) {
    if ((e[i >> 5] >> (i & 31)) & 1) r = fe_mul(r, b);
    b = fe_sqr(b);
  }
  return r;
}

// x mod p for an arbitrary 128-bit value (x < 2^128 < 2p): utils.rs:50-66 be_from_u128
ZKL_HD fe fe_from_u128_mod(const fe &x) {
#if defined(__CUDA_ARCH__)
  {
    uint32_t q0 = p_add_cc(x.l[0], ZKL_C0), q1 = p_addc_cc(x.l[1], ZKL_C1), q2 = p_addc_cc(x.l[2], 0), q3 = p_addc_cc(x.l[3], 0),
             c = p_addc(0, 0);
    fe r; r.l[0] = c ? q0 : x.l[0]; r.l[1] = c ? q1 : x.l[1]; r.l[2] = c ? q2 : x.l[2]; r.l[3] = c ? q3 : x.l[3];
    return r;
  }
#endif
  fe q; uint32_t c;
  q.l[0] = add_co(x.l[0], ZKL_C0, c); q.l[1] = add_cio(x.l[1], ZKL_C1, c);
  q.l[2] = add_cio(x.l[2], 0, c); q.l[3] = add_cio(x.l[3], 0, c);
  return c ? q : x;
}
// lo + hi * 2^64 mod p for canonical lo, hi: utils.rs:359-371 fold_bytes32_to_fe (after per-half reduction)
ZKL_HD fe fe_fold_pair(const fe &lo, const fe &hi) {
#if defined(__CUDA_ARCH__)
  {
    uint32_t x[9];
    x[0] = lo.l[0]; x[1] = lo.l[1];
    x[2] = p_add_cc(lo.l[2], hi.l[0]); x[3] = p_addc_cc(lo.l[3], hi.l[1]); x[4] = p_addc_cc(hi.l[2], 0); x[5] = p_addc_cc(hi.l[3], 0);
    x[6] = p_addc(0, 0); x[7] = 0; x[8] = 0;
    return reduce9(x);
  }
#endif
  uint32_t x[9], c;
  x[0] = lo.l[0]; x[1] = lo.l[1];
  x[2] = add_co(lo.l[2], hi.l[0], c); x[3] = add_cio(lo.l[3], hi.l[1], c);
  x[4] = add_cio(0, hi.l[2], c); x[5] = add_cio(0, hi.l[3], c);
  x[6] = c; x[7] = 0; x[8] = 0;
  return reduce9(x);
}
