// kernels_ntt.cuh — column iNTT / coset LDE over f128 (winter-math fft::interpolate_poly,
// evaluate_poly_with_offset; DefaultTraceLde::new / CompositionPoly::new call sites at
// zk-lisp-proof-winterfell/src/prove.rs:462,502).
//
// Layout. A trace segment has n <= 8192 rows (segment_planner.rs:24 caps it at 4096), so one column
// (n * 16 B <= 128 KB) fits in the shared memory of one CTA: a transform is ONE pass over HBM —
// coalesced 128-bit loads, all log2(n) radix-2 stages in shared memory, coalesced 128-bit stores.
// The LDE of column c is stored COSET-MAJOR: lde[c][k][i] = T_c(offset * w_N^k * w_n^i) is the
// natural LDE row r = k + blowup*i. One CTA produces one (column, coset) slab with unit-stride
// stores, the constraint frame's "next" row (r + blowup) is the neighbouring element, and FRI
// pairs (r, r + N/2) sit n/2 apart inside one coset.
#pragma once
#include "kernels_hash.cuh"

// wpow[m] = w^m for m < count (w = primitive root of the LDE domain)
__global__ void k_fill_powers(fe w, fe *out, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fe(out + i, fe_pow(w, (uint64_t)i));
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

// In-place decimation-in-frequency NTT over shared memory: natural in, bit-reversed out.
// tw(j) must return w_n^j (or w_n^-j for the inverse).  wpow has N = n * tw_stride entries.
__device__ __forceinline__ void smem_ntt_dif(fe *sm, int n, int logn, const fe *__restrict__ wpow, size_t N, size_t tw_stride,
                                             bool inverse) {
  for (int s = 0; s < logn; s++) {
    int half = n >> (s + 1);
    for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
      int blk = t / half, j = t - blk * half;
      int i0 = blk * 2 * half + j, i1 = i0 + half;
      fe a = sm[i0], b = sm[i1];
      size_t e = (size_t)j * ((size_t)n / (2 * half)) * tw_stride;  // exponent of w_N
      if (inverse && e) e = N - e;
      fe w = ld_fe(wpow + e);
      sm[i0] = fe_add(a, b);
      sm[i1] = fe_mul(fe_sub(a, b), w);
    }
    __syncthreads();
  }
}

// interpolate_poly: src[c][0..n) evaluations over <w_n> (any coset: the offset is undone by the
// caller) -> dst[c][0..n) coefficients. grid = ncols. src == dst allowed.
__global__ void k_interpolate(const fe *src, fe *dst, int n, int logn, const fe *__restrict__ wpow, size_t N, fe n_inv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fe *sm = reinterpret_cast<fe *>(smem_raw);
  const fe *col = src + (size_t)blockIdx.x * n;
  fe *o = dst + (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = ld_fe(col + i);
  __syncthreads();
  smem_ntt_dif(sm, n, logn, wpow, N, N / n, true);
  for (int j = threadIdx.x; j < n; j += blockDim.x) st_fe(o + j, fe_mul(sm[bitrev(j, logn)], n_inv));
}

// coset scaling table: scale[k][j] = (offset * w_N^k)^j = opow[j] * w_N^(j k), k < blowup, j < n (built once per (n, blowup))
__global__ void k_lde_scale_table(const fe *__restrict__ opow, const fe *__restrict__ wpow, size_t N, int n, int blowup, fe *__restrict__ scale) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * blowup) return;
  size_t k = t / n, j = t % n;
  st_fe(scale + t, fe_mul(ld_fe(opow + j), ld_fe(wpow + (j * k) % N)));
}

// coset LDE: grid = (blowup, ncols). out[c][k][i] = poly_c(offset * w_N^k * w_n^i).
__global__ void k_lde(const fe *__restrict__ polys, int n, int logn, int blowup, const fe *__restrict__ wpow, size_t N,
                      const fe *__restrict__ scale, fe *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fe *sm = reinterpret_cast<fe *>(smem_raw);
  int k = blockIdx.x; size_t c = blockIdx.y;
  const fe *poly = polys + c * n;
  const fe *sc = scale + (size_t)k * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) sm[j] = fe_mul(ld_fe(poly + j), ld_fe(sc + j));
  __syncthreads();
  smem_ntt_dif(sm, n, logn, wpow, N, blowup, false);
  fe *dst = out + c * N + (size_t)k * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) st_fe(dst + i, sm[bitrev(i, logn)]);
}

// natural-order evaluation copy for the C-ABI primitive zkl_evaluate_columns: nat[c][k + blowup*i] = cm[c][k*n + i]
__global__ void k_coset_to_natural(const fe *__restrict__ cm, fe *__restrict__ nat, size_t N, size_t n, int blowup, size_t total) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  size_t c = t / N, s = t % N;
  size_t r = (s / n) + (size_t)blowup * (s % n);
  st_fe(nat + c * N + r, ld_fe(cm + t));
}

// ================================================================= n > 8192: two-pass (four-step) transforms
// A column no longer fits one CTA's shared memory, so n = n1 * n2 and with j = j1 + n1 j2, i = n2 i1 + i2:
//   y[n2 i1 + i2] = sum_j1 w_n1^(i1 j1) * ( w_n^(i2 j1) * sum_j2 x[j1 + n1 j2] w_n2^(i2 j2) )
// pass 1: T consecutive j1 per CTA, T size-n2 transforms in shared memory, twiddle, B[j1][i2] (unit stride in i2)
// pass 2: T consecutive i2 per CTA, T size-n1 transforms, y[n2 i1 + i2] — the same address set the CTA read, so
//         pass 2 runs IN PLACE on B. Every global access moves T*16-byte contiguous pieces (T >= 4: full sectors).
// Two passes over HBM instead of one; used only above 8192 rows (larger segments).
// `sign`: forward transforms use w, inverse ones w^-1 (exponent e -> Ntab - e). wtab has Ntab = n * tstride entries.

// T independent DIF transforms of length m over sm[t * m + .]; natural in, bit-reversed out
__device__ __forceinline__ void smem_ntt_dif_batch(fe *sm, int T, int m, int logm, const fe *__restrict__ wtab, size_t Ntab, size_t stride_m,
                                                   bool inverse) {
  const int per = m / 2;
  for (int s = 0; s < logm; s++) {
    int half = m >> (s + 1);
    for (int q = threadIdx.x; q < T * per; q += blockDim.x) {
      int t = q / per, b = q - t * per;
      int blk = b / half, j = b - blk * half;
      fe *row = sm + (size_t)t * m;
      int i0 = blk * 2 * half + j, i1 = i0 + half;
      fe a = row[i0], c = row[i1];
      size_t e = (size_t)j * ((size_t)m / (2 * half)) * stride_m;
      if (inverse && e) e = Ntab - e;
      row[i0] = fe_add(a, c);
      row[i1] = fe_mul(fe_sub(a, c), ld_fe(wtab + e));
    }
    __syncthreads();
  }
}
// pass 1. grid = (n1 / T, ncols, cosets). src column c: src + c*n (natural order). scale (optional): scale + z*n, multiplied in.
// dst slab (c, z): dst + (c * cosets + z) * n, written as B[j1][i2].
__global__ void k_ntt_pass1(const fe *__restrict__ src, fe *__restrict__ dst, int n1, int n2, int logn2, int T, const fe *__restrict__ wtab, size_t Ntab,
                            const fe *__restrict__ scale, bool inverse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fe *sm = reinterpret_cast<fe *>(smem_raw);
  const size_t n = (size_t)n1 * n2;
  const int j1_0 = blockIdx.x * T; const size_t c = blockIdx.y, z = blockIdx.z, cosets = gridDim.z;
  const fe *col = src + c * n;
  const fe *sc = scale ? scale + z * n : nullptr;
  for (int q = threadIdx.x; q < T * n2; q += blockDim.x) {
    int t = q % T, j2 = q / T;
    size_t j = (size_t)(j1_0 + t) + (size_t)n1 * j2;
    fe v = ld_fe(col + j);
    if (sc) v = fe_mul(v, ld_fe(sc + j));
    sm[(size_t)t * n2 + j2] = v;
  }
  __syncthreads();
  smem_ntt_dif_batch(sm, T, n2, logn2, wtab, Ntab, Ntab / n2, inverse);
  fe *out = dst + (c * cosets + z) * n;
  const size_t stride_n = Ntab / n;
  for (int q = threadIdx.x; q < T * n2; q += blockDim.x) {
    int t = q / n2, i2 = q - t * n2;
    size_t e = ((size_t)i2 * (size_t)(j1_0 + t) * stride_n) % Ntab;
    if (inverse && e) e = Ntab - e;
    st_fe(out + (size_t)(j1_0 + t) * n2 + i2, fe_mul(sm[(size_t)t * n2 + bitrev(i2, logn2)], ld_fe(wtab + e)));
  }
}
// pass 2, in place. grid = (n2 / T, nslabs). slab s: buf + s*n. Optional final factor (1/n for inverse transforms).
__global__ void k_ntt_pass2(fe *__restrict__ buf, int n1, int n2, int logn1, int T, const fe *__restrict__ wtab, size_t Ntab, bool inverse,
                            bool has_factor, fe factor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fe *sm = reinterpret_cast<fe *>(smem_raw);
  const size_t n = (size_t)n1 * n2;
  const int i2_0 = blockIdx.x * T;
  fe *slab = buf + (size_t)blockIdx.y * n;
  for (int q = threadIdx.x; q < T * n1; q += blockDim.x) {
    int t = q % T, j1 = q / T;
    sm[(size_t)t * n1 + j1] = ld_fe(slab + (size_t)j1 * n2 + i2_0 + t);
  }
  __syncthreads();
  smem_ntt_dif_batch(sm, T, n1, logn1, wtab, Ntab, Ntab / n1, inverse);
  for (int q = threadIdx.x; q < T * n1; q += blockDim.x) {
    int t = q % T, i1 = q / T;
    fe v = sm[(size_t)t * n1 + bitrev(i1, logn1)];
    if (has_factor) v = fe_mul(v, factor);
    st_fe(slab + (size_t)i1 * n2 + i2_0 + t, v);
  }
}
