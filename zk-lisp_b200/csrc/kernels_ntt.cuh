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
