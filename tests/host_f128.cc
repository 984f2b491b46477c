// Host build of the DEVICE limb arithmetic (zk-lisp_b200/csrc/f128.cuh compiled with g++): the
// carry primitives fall back to plain C so the limb algorithms can be checked on the CPU.
#include "../zk-lisp_b200/csrc/f128_lazy.cuh"
extern "C" {
void h_binop(int op, const fe *a, const fe *b, fe *out, long n) {
  for (long i = 0; i < n; i++) {
    switch (op) {
      case 0: out[i] = fe_add(a[i], b[i]); break;
      case 1: out[i] = fe_sub(a[i], b[i]); break;
      case 2: out[i] = fe_mul(a[i], b[i]); break;
      case 3: out[i] = fe_inv(a[i]); break;
      case 4: out[i] = fe_cube(a[i]); break;
      case 5: out[i] = fe_fold_pair(a[i], b[i]); break;
      case 6: out[i] = fe_from_u128_mod(a[i]); break;
      // lazy arithmetic (tcperm.cuh): inputs are ANY 128-bit values, outputs are 128-bit representatives
      case 7: out[i] = lazy_mul(a[i], b[i]); break;
      case 8: out[i] = lazy_add(a[i], b[i]); break;
      case 9: out[i] = lazy_canon(a[i]); break;
      case 10: out[i] = lazy_sqr(a[i]); break;
      case 11: { uint32_t x[8]; lazy_mul_wide(a[i], b[i], x); out[i] = lazy_reduce8_mac(x); } break;
    }
  }
}
// out = sum_k a[k]*b[k] + rc  (12 terms, lazy reduction) — the MDS row pattern
void h_dot12(const fe *a, const fe *b, const fe *rc, fe *out, long n) {
  for (long i = 0; i < n; i++) {
    wide_acc w; wacc_zero(w);
    for (int k = 0; k < 12; k++) wacc_mad(w, a[12 * i + k], b[12 * i + k]);
    wacc_add_fe(w, rc[i]);
    out[i] = wacc_reduce(w);
  }
}
// tensor-core epilogue: 16 byte-position accumulators + round constant -> lazy element
void h_tc_epilogue(const uint32_t *acc, const fe *rc, fe *out, long n) {
  for (long i = 0; i < n; i++) out[i] = rc ? tc_epilogue(acc + 16 * i, rc[i]) : tc_epilogue_plain(acc + 16 * i);
}
// raw 256-bit product (the cube bytes fed to the MMA)
void h_mul_wide(const fe *a, const fe *b, uint32_t *out, long n) {
  for (long i = 0; i < n; i++) lazy_mul_wide(a[i], b[i], out + 8 * i);
}
void h_sqr_wide(const fe *a, uint32_t *out, long n) {
  for (long i = 0; i < n; i++) lazy_sqr_wide(a[i], out + 8 * i);
}
uint32_t h_tc_acc_max() { return TC_ACC_MAX; }
}
