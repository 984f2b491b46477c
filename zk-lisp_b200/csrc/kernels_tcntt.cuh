// kernels_tcntt.cuh — coset LDE of 4096-row columns with the radix-16 butterflies on the tensor cores.
//
// A size-4096 transform is three radix-16 stages (4096 = 16^3). One stage is 256 independent 16-point DFTs
//   out[i] = sum_j in[j] * w16^(i j),   i, j < 16,
// a product with the CONSTANT 16 x 16 matrix (w16^(ij)) over f128 — the same shape as the sponge's MDS product
// (tcperm.cuh), so it runs the same way: the 16 inputs of a DFT are 256 bytes of K, the constant byte matrix
// B[(i,u),(j,v)] = byte_u(w16^(ij) * 2^(8v) mod p) has N = 256 rows, and
//   out[i] ≡ sum_u 2^(8u) * (sum_{j,v} B[(i,u),(j,v)] * byte_v(in[j]))
// with exact s32 accumulators (256 terms < 2^16: at most 16 646 400 < TC_ACC_MAX for ANY constants). The CUDA cores
// keep the carry propagation of the 16 accumulators per output (tc_epilogue_plain) and the inter-stage twiddles:
// per element and stage ~120 instructions instead of 4 radix-2 butterflies (4 x ~1/2 product + add/sub ≈ 340).
//
// Index maps (validated against a recursive FFT in Python), j = jA + 16 jC + 256 jD, w = w_4096:
//   stage 1: instance m1 = jA + 16 jC, inputs over jD (natural order), outputs iD, twiddle w_256^(iD jC)
//   stage 2: instance m2 = jA + 16 iD, inputs over jC,                outputs iC, twiddle w^((16 iC + iD) jA)
//   stage 3: instance m3 = iB = 16 iC + iD, inputs over jA,            outputs iA -> y[256 iA + iB]
// One CTA transforms one (column, coset) at a time: two shared-memory working arrays [input index][instance]
// (ping/pong, 16 chunks of 256 x 16 B, chunk pitch 4112 B so that both scatter patterns are bank-conflict free)
// are at once the MMA A operands (K-major, no swizzle: LBO = chunk pitch, SBO = 128 B); instances 0..127 / 128..255
// are the two M = 128 tiles, their accumulators the two halves of TMEM. 512 compute threads (two per instance, 8
// outputs each) + one MMA-issuing warp.
#pragma once
#include "kernels_ntt.cuh"
#include "tcperm.cuh"

#define NTT16_N 4096
#define NTT16_CROSS_MAX_R 16                    // two-pass tensor-core transform: R blocks of NTT16_N, R in {2, 4, 8, 16}
#define NTT16_PITCH 4112                       // bytes between input chunks: 256 instances x 16 B + 16 B pad
#define NTT16_BUF_BYTES (16 * NTT16_PITCH)
#define NTT16_B_BYTES (16 * 256 * 16)
#define NTT16_THREADS (512 + 32)
#define NTT16_SMEM_BYTES (NTT16_B_BYTES + 2 * NTT16_BUF_BYTES + 256 + 128)
#define NTT16_IDESC ((2u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))

// byte images of the DFT-16 matrix for w16 (forward) and w16^-1 (inverse): entry (chunk j, row 16 i + u) = 16 bytes b:
// byte_u(w16^(± i j) * 2^(8b) mod p)
__device__ uint4 g_dft16_bytes[2][NTT16_B_BYTES / 16];
__global__ void k_build_dft16_bytes(fe w16, fe w16_inv) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * 16 * 256) return;
  const int inv = t >> 12; t &= 4095;
  int j = t >> 8, row = t & 255, i = row >> 4, u = row & 15;
  fe cur = fe_pow(inv ? w16_inv : w16, (uint64_t)(i * j));
  fe two8 = fe_from_u64(256);
  uint32_t out[4] = {0, 0, 0, 0};
  for (int b = 0; b < 16; b++) {
    uint32_t limb = u < 4 ? cur.l[0] : (u < 8 ? cur.l[1] : (u < 12 ? cur.l[2] : cur.l[3]));
    out[b >> 2] |= ((limb >> (8 * (u & 3))) & 0xFF) << (8 * (b & 3));
    cur = fe_mul(cur, two8);
  }
  g_dft16_bytes[inv][t] = make_uint4(out[0], out[1], out[2], out[3]);
}

// grid = min(ncols * blowup, #SMs), persistent.
//   forward (coset LDE):  polys: ncols x 4096 coefficients; scale[k][j] = (offset w_N^k)^j; out[c][k][i] coset-major
//   inverse (interpolate): polys: ncols x 4096 evaluations over <w_4096>, blowup = 1, scale = nullptr, every output is
//                          multiplied by `factor` = 4096^-1; in place allowed (a CTA reads its whole column before it writes)
// wpow: w_N^e, N = 4096 * blowup.
__global__ void __launch_bounds__(NTT16_THREADS, 1) k_tc_lde4096(const fe *polys, int ncols, int blowup, const fe *__restrict__ scale,
                                                                const fe *__restrict__ wpow, fe *out, bool inverse, fe factor) {
  extern __shared__ unsigned char ntt_smem[];
  const uint32_t base = (smem_u32(ntt_smem) + 127u) & ~127u;
  unsigned char *gen = ntt_smem + (base - smem_u32(ntt_smem));
  const uint32_t b_tile = base, buf0 = base + NTT16_B_BYTES, buf1 = buf0 + NTT16_BUF_BYTES, ctl = buf1 + NTT16_BUF_BYTES;
  uint32_t *ctl_gen = reinterpret_cast<uint32_t *>(gen + NTT16_B_BYTES + 2 * NTT16_BUF_BYTES);
  {
    uint4 *bdst = reinterpret_cast<uint4 *>(gen);
    for (int i = threadIdx.x; i < NTT16_B_BYTES / 16; i += NTT16_THREADS) bdst[i] = g_dft16_bytes[inverse ? 1 : 0][i];
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 64), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(ctl, 1); mbar_init(ctl + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl_gen[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_mma = warp >= 16;
  const int tile = (warp >> 3) & 1, half = (warp >> 2) & 1;          // compute warps: tile = instances 128 t .. 128 t + 127
  const int m = 128 * tile + 32 * (warp & 3) + lane;                 // DFT instance of this thread
  const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 256 * tile + 128 * half;
  const size_t N = (size_t)NTT16_N * blowup;
  const size_t s256 = N / 256, s4096 = N / NTT16_N;                  // w_256 = w_N^s256, w_4096 = w_N^s4096
  uint32_t phase = 0;
  uint32_t src = buf0, dst = buf1;      // ping/pong roles persist across problems: the next column is loaded into the buffer
                                        // whose last readers (the MMAs of stage 2 of the previous column) are known complete
  const int problems = ncols * blowup;
  for (int p = blockIdx.x; p < problems; p += gridDim.x) {
    const int c = p / blowup, k = p - c * blowup;
    if (!is_mma) {   // x[j] = a_j * (offset w_N^k)^j, natural order: chunk jD = j >> 8, instance m1 = j & 255
      const fe *poly = polys + (size_t)c * NTT16_N, *sc = scale + (size_t)k * NTT16_N;
#pragma unroll 2
      for (int q = 0; q < 8; q++) {
        int j = threadIdx.x + 512 * q;
        fe v = ld_fe(poly + j);
        if (scale) v = lazy_mul(v, ld_fe(sc + j));
        sts_u4(src + (uint32_t)(j >> 8) * NTT16_PITCH + (uint32_t)(j & 255) * 16, v.l[0], v.l[1], v.l[2], v.l[3]);
      }
    }
#pragma unroll 1
    for (int stage = 0; stage < 3; stage++) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (is_mma) {
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const uint64_t adesc = tc_smem_desc(src + t * 2048, NTT16_PITCH, 128);
            const uint64_t bdesc = tc_smem_desc(b_tile, 4096, 128);
#pragma unroll
            for (int kk = 0; kk < 8; kk++)     // one MMA = two input indices = 32 bytes of K
              tc_mma_i8(tmem_base + 256 * t, adesc + (uint64_t)((2 * kk * NTT16_PITCH) >> 4), bdesc + (uint64_t)((2 * kk * 4096) >> 4), NTT16_IDESC, kk > 0);
            tc_commit(ctl + 8 * t);
          }
        }
        __syncwarp();
      } else {
        mbar_wait(ctl + 8 * tile, phase);
        tc_fence_after();
        uint32_t acc[2][16];
        tmem_ld16(tmem_mine, acc[0]);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          tmem_wait_ld(acc[q & 1]);
          if (q < 7) tmem_ld16(tmem_mine + 16 * (q + 1), acc[(q + 1) & 1]);
          const int i = 8 * half + q;                              // output index of the 16-point DFT
          fe v = tc_epilogue_plain(acc[q & 1]);
          if (stage == 0) {          // m = jA + 16 jC, i = iD
            const int jA = m & 15, jC = m >> 4;
            { size_t e = (size_t)(i * jC) * s256; if (inverse && e) e = N - e; v = lazy_mul(v, ld_fe(wpow + e)); }
            sts_u4(dst + (uint32_t)jC * NTT16_PITCH + (uint32_t)(jA + 16 * i) * 16, v.l[0], v.l[1], v.l[2], v.l[3]);
          } else if (stage == 1) {   // m = jA + 16 iD, i = iC
            const int jA = m & 15, iB = 16 * i + (m >> 4);
            { size_t e = (size_t)(iB * jA) * s4096; if (inverse && e) e = N - e; v = lazy_mul(v, ld_fe(wpow + e)); }
            sts_u4(dst + (uint32_t)jA * NTT16_PITCH + (uint32_t)iB * 16, v.l[0], v.l[1], v.l[2], v.l[3]);
          } else {                   // m = iB, i = iA
            if (inverse) v = lazy_mul(v, factor);
            st_fe(out + ((size_t)c * blowup + k) * NTT16_N + 256 * i + m, lazy_canon(v));
          }
        }
      }
      phase ^= 1;
      uint32_t t = src; src = dst; dst = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}


// ================================================================= version 2: coset shift folded into the constant matrix
// x[j] = a_j s^j with s = offset * w_N^k; for j = jA + 16 jC + 256 jD:  s^j = s^(jA + 16 jC) * (s^256)^jD.  The second
// factor depends on the stage-1 INPUT index only, so stage 1 of coset k uses the constant matrix
//   B_k[(iD,u),(jD,v)] = byte_u((s^256)^jD * w16^(iD jD) * 2^(8v) mod p)
// on the RAW coefficients, and the first factor joins the stage-1 output twiddle: tw0_k[iD][m1] = s^(m1) * w_256^(iD jC),
// m1 = jA + 16 jC. (Shared memory holds ONE matrix, so stages 2 and 3 run with B_k as well; the factors (s^256)^jC and
// (s^256)^jA this adds are constants of m1 and are divided out of tw0_k.) One of the three modular products per element disappears, and because the MMA A operand is now the
// coefficient column itself the next column is fetched by TMA bulk copies (16 x 4 KB into the chunk layout) while the
// current one is in its last stage. For the inverse transform B = 4096^-1 * w16^(-iD jD) (the final scaling disappears too)
// and tw0[iD][m1] = w_256^(-iD jC). A CTA serves ONE coset (k = blockIdx.x % blowup): B_k is loaded once.
//
// consts layout per coset k: bmat[k][4096] uint4 (shared-memory image, as g_dft16_bytes), tw0[k][16][256] fe.
__global__ void k_build_lde16_consts(const fe *__restrict__ scale /* [blowup][scale_n] or null (no coset shift) */, size_t scale_n, const fe *__restrict__ wpow, size_t N,
                                     int blowup, bool inverse, fe factor, uint4 *__restrict__ bmat, fe *__restrict__ tw0) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= blowup * 4096) return;
  const int k = t >> 12, e = t & 4095;
  const size_t s16 = N / 16, s256 = N / 256;
  {  // B image entry (chunk j = jD, row = 16 i + u)
    const int j = e >> 8, row = e & 255, i = row >> 4, u = row & 15;
    size_t ex = (size_t)((i * j) & 15) * s16; if (inverse && ex) ex = N - ex;
    fe cur = ld_fe(wpow + ex);
    if (scale) cur = fe_mul(cur, ld_fe(scale + (size_t)k * scale_n + 256 * j));
    if (inverse) cur = fe_mul(cur, factor);
    const fe two8 = fe_from_u64(256);
    uint32_t out[4] = {0, 0, 0, 0};
    for (int b = 0; b < 16; b++) {
      uint32_t limb = u < 4 ? cur.l[0] : (u < 8 ? cur.l[1] : (u < 12 ? cur.l[2] : cur.l[3]));
      out[b >> 2] |= ((limb >> (8 * (u & 3))) & 0xFF) << (8 * (b & 3));
      cur = fe_mul(cur, two8);
    }
    bmat[(size_t)k * 4096 + e] = make_uint4(out[0], out[1], out[2], out[3]);
  }
  {  // tw0[k][iD][m1]. B_k also serves stages 2 and 3 (one matrix fits shared memory), whose input indices are jC and jA:
     // they pick up (s^256)^jC and (s^256)^jA — resp. 4096^-1 twice — which this twiddle takes out again.
    const int iD = e >> 8, m1 = e & 255, jA = m1 & 15, jC = m1 >> 4;
    size_t ex = (size_t)(iD * jC) * s256; if (inverse && ex) ex = N - ex;
    fe v = ld_fe(wpow + ex);
    if (scale) {
      v = fe_mul(v, ld_fe(scale + (size_t)k * scale_n + m1));
      v = fe_mul(v, fe_pow(fe_inv(ld_fe(scale + (size_t)k * scale_n + 256)), (uint64_t)(jA + jC)));
    }
    if (inverse) { fe f1 = fe_inv(factor); v = fe_mul(v, fe_mul(f1, f1)); }
    st_fe(tw0 + (size_t)k * 4096 + e, v);
  }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// grid >= blowup CTAs; CTA b serves coset b % blowup and every (gridDim.x / blowup)-th column of it.
__global__ void __launch_bounds__(NTT16_THREADS, 1) k_tc_lde4096_v2(const fe *polys, int ncols, int blowup, const uint4 *__restrict__ bmat,
                                                                   const fe *__restrict__ tw0, const fe *__restrict__ wpow, fe *out, bool inverse, int cross, size_t Ntab /* entries of wpow: w_Ntab^e */,
                                                                   uint32_t radices /* 4 bits per cross level: log2 R_1 (lowest nibble) .. log2 R_L; 0 = the one level `cross` */,
                                                                   bool block_major_out /* cross mode: block q's outputs stay together (k_tc_unscatter finishes) */) {
  extern __shared__ unsigned char ntt_smem[];
  const uint32_t base = (smem_u32(ntt_smem) + 127u) & ~127u;
  unsigned char *gen = ntt_smem + (base - smem_u32(ntt_smem));
  const uint32_t b_tile = base, buf0 = base + NTT16_B_BYTES, buf1 = buf0 + NTT16_BUF_BYTES, ctl = buf1 + NTT16_BUF_BYTES;
  uint32_t *ctl_gen = reinterpret_cast<uint32_t *>(gen + NTT16_B_BYTES + 2 * NTT16_BUF_BYTES);
  const int k = blockIdx.x % blowup, rank = blockIdx.x / blowup;
  const int peers = (gridDim.x - k + blowup - 1) / blowup;      // CTAs serving coset k
  {
    uint4 *bdst = reinterpret_cast<uint4 *>(gen);
    const uint4 *bsrc = bmat + (size_t)k * 4096;
    for (int i = threadIdx.x; i < NTT16_B_BYTES / 16; i += NTT16_THREADS) bdst[i] = bsrc[i];
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 64), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(ctl, 1); mbar_init(ctl + 8, 1); mbar_init(ctl + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl_gen[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_mma = warp >= 16;
  const int tile = (warp >> 3) & 1, half = (warp >> 2) & 1;
  const int m = 128 * tile + 32 * (warp & 3) + lane;
  const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 256 * tile + 128 * half;
  const size_t N = Ntab;
  const size_t s4096 = N / NTT16_N;
  const fe *tw0k = tw0 + (size_t)k * 4096;
  uint32_t phase = 0, ld_phase = 0;
  uint32_t src = buf0, dst = buf1;
  // column c as 16 chunks of 256 coefficients: chunk jD -> src + jD * PITCH (instance m1 = j & 255 at 16 B pitch)
  // cross mode (columns of R * 4096 rows, second pass; cross = R): problem q = R c + i2 reads block i2 of slab (c, k) of the first
  // pass's output and writes y[i2 + R i1] of the slab's transform.
  // Several cross levels (n = R_1 ... R_L 4096, `radices`): q = ((c R_1 + i2_1) R_2 + i2_2 ...) R_L + i2_L, `cross` = R_L; the block's
  // outputs i1 go to y[i2_1 + R_1 (i2_2 + R_2 (... + R_L i1))] of column c.
  if (!radices) radices = cross ? (uint32_t)(31 - __clz(cross)) : 0;
  int prod_log = 0;
  for (uint32_t r = radices; r; r >>= 4) prod_log += (int)(r & 15);
  auto out_base = [&](int q, size_t *stride) -> fe * {
    const size_t P = (size_t)1 << prod_log;
    size_t rem = (size_t)q & (P - 1), off = 0;
    int shift = prod_log;
    // digits from the last level (least significant in q) to the first
    int nl = 0; for (uint32_t r = radices; r; r >>= 4) nl++;
    for (int l = nl - 1; l >= 0; l--) {
      const int lg = (int)((radices >> (4 * l)) & 15);
      shift -= lg;                                        // log2 (R_1 ... R_{l})
      off += (rem & (((size_t)1 << lg) - 1)) << shift;
      rem >>= lg;
    }
    *stride = P;
    return out + (((size_t)q >> prod_log) * blowup + k) * (P * NTT16_N) + off;
  };
  auto in_block = [&](int q) -> const fe * {
    return cross ? polys + ((size_t)((q / cross) * blowup + k) * cross + (q % cross)) * NTT16_N : polys + (size_t)q * NTT16_N;
  };
  auto fetch = [&](int c, uint32_t to) {
    mbar_expect_tx(ctl + 16, NTT16_N * 16);
    const fe *poly = in_block(c);
#pragma unroll 1
    for (int jD = 0; jD < 16; jD++) bulk_g2s(to + (uint32_t)jD * NTT16_PITCH, poly + 256 * jD, 4096, ctl + 16);
  };
  if (is_mma && lane == 0 && rank < ncols) fetch(rank, src);
  for (int c = rank; c < ncols; c += peers) {
#pragma unroll 1
    for (int stage = 0; stage < 3; stage++) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (is_mma) {
        tc_fence_after();
        if (lane == 0) {
          if (stage == 0) mbar_wait(ctl + 16, ld_phase);   // the column has landed (async proxy -> async proxy: no fence)
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const uint64_t adesc = tc_smem_desc(src + t * 2048, NTT16_PITCH, 128);
            const uint64_t bdesc = tc_smem_desc(b_tile, 4096, 128);
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
              tc_mma_i8(tmem_base + 256 * t, adesc + (uint64_t)((2 * kk * NTT16_PITCH) >> 4), bdesc + (uint64_t)((2 * kk * 4096) >> 4), NTT16_IDESC, kk > 0);
            tc_commit(ctl + 8 * t);
          }
          // `dst` is idle in the last stage (results go to HBM; its last readers were the MMAs of stage 2): prefetch into it
          if (stage == 2 && c + peers < ncols) fetch(c + peers, dst);
        }
        __syncwarp();
      } else {
        mbar_wait(ctl + 8 * tile, phase);
        tc_fence_after();
        uint32_t acc[2][16];
        tmem_ld16(tmem_mine, acc[0]);
#pragma unroll 2
        for (int q = 0; q < 8; q++) {
          tmem_wait_ld(acc[q & 1]);
          if (q < 7) tmem_ld16(tmem_mine + 16 * (q + 1), acc[(q + 1) & 1]);
          const int i = 8 * half + q;
          fe v = tc_epilogue_plain(acc[q & 1]);
          if (stage == 0) {          // m = jA + 16 jC, i = iD
            const int jA = m & 15, jC = m >> 4;
            v = lazy_mul(v, ld_fe(tw0k + 256 * i + m));
            sts_u4(dst + (uint32_t)jC * NTT16_PITCH + (uint32_t)(jA + 16 * i) * 16, v.l[0], v.l[1], v.l[2], v.l[3]);
          } else if (stage == 1) {   // m = jA + 16 iD, i = iC
            const int jA = m & 15, iB = 16 * i + (m >> 4);
            { size_t e = (size_t)(iB * jA) * s4096; if (inverse && e) e = N - e; v = lazy_mul(v, ld_fe(wpow + e)); }
            sts_u4(dst + (uint32_t)jA * NTT16_PITCH + (uint32_t)iB * 16, v.l[0], v.l[1], v.l[2], v.l[3]);
          } else {                   // m = iB, i = iA
            if (cross && block_major_out) st_fe(out + ((((size_t)c >> prod_log) * blowup + k) << prod_log) * NTT16_N + ((size_t)c & (((size_t)1 << prod_log) - 1)) * NTT16_N + 256 * i + m, lazy_canon(v));
            else if (cross) { size_t stride; fe *ob = out_base(c, &stride); st_fe(ob + stride * (size_t)(256 * i + m), lazy_canon(v)); }
            else st_fe(out + ((size_t)c * blowup + k) * NTT16_N + 256 * i + m, lazy_canon(v));
          }
        }
      }
      phase ^= 1;
      uint32_t t = src; src = dst; dst = t;
    }
    ld_phase ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}


// The last step of a transform with P = R_1 ... R_L >= 32 blocks per column: slab (column, coset) holds block b's 4096 outputs
// together (tmp[b][i1], the 4096-point kernel's unit-stride stores); y[o + P i1] = tmp[b(o)][i1] where o's digits are b's in
// reverse order (o = i2_1 + R_1 (i2_2 + ...), b = ((i2_1 R_2 + i2_2) R_3 ...)). Scattering straight from the 4096-point kernel puts
// every 16-byte element 16 P bytes from its neighbour — at P = 256 the capture showed 2.6x the algorithmic DRAM reads (sector
// read-modify-write) — so the transpose goes through shared memory: 32 x 32 tiles, 512-byte runs on both sides.
__global__ void __launch_bounds__(256) k_tc_unscatter(const fe *__restrict__ tmp, fe *__restrict__ out, int prod_log, uint32_t radices) {
  __shared__ fe tile[32][33];
  const size_t P = (size_t)1 << prod_log;
  const size_t slab = blockIdx.z;
  const size_t o0 = (size_t)blockIdx.y * 32, i0 = (size_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows of 32 per pass
  const fe *src = tmp + slab * P * NTT16_N;
  fe *dst = out + slab * P * NTT16_N;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    size_t o = o0 + r, b = 0;
    for (uint32_t rd = radices; rd; rd >>= 4) { const int lg = (int)(rd & 15); b = (b << lg) | (o & (((size_t)1 << lg) - 1)); o >>= lg; }
    tile[r][tx] = ld_fe(src + b * NTT16_N + i0 + tx);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) st_fe(dst + (i0 + r) * P + o0 + tx, tile[tx][r]);
}

// ================================================================= 65536-row columns: two passes, both on the tensor cores
// n = 16 m, m = 4096, j = j1 + m j2, i = i2 + 16 i1:
//   y[i2 + 16 i1] = sum_j1 [s^j1 w_m^(j1 i1)] * w_n^(j1 i2) * sum_j2 a[j1 + m j2] (s^m)^j2 w16^(j2 i2)
// pass 1 (k_tc_cross16): ONE radix-16 stage across the 16 blocks of m coefficients with the per-coset matrix
//   C_k[(i2,u),(j2,v)] = byte_u((s^m)^j2 w16^(i2 j2) 2^(8v) mod p), then the twiddle w_n^(j1 i2); a CTA takes slices of 256
//   consecutive j1 (16 TMA bulk copies of 4 KB, double-buffered: the next slice lands while this one is multiplied) and
//   writes block i2 of the slab (c, k) with unit stride in j1;
// pass 2 (k_tc_lde4096_v2, cross = 1): the coset LDE of every block — its B_k / tw0_k carry exactly the s^j1 factor —
//   scattered to y[i2 + 16 i1].
// Inverse (interpolation): s = 1, inverse roots, 1/n folded into C.
__global__ void k_build_cross16_consts(const fe *__restrict__ scale /* [blowup][kstride] or null */, size_t kstride, const fe *__restrict__ wpow, size_t N, size_t m, int R,
                                       int blowup, bool inverse, fe factor, uint4 *__restrict__ cmat) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= blowup * 4096) return;
  const int k = t >> 12, e = t & 4095;
  const int j = e >> 8, row = e & 255, i = row >> 4, u = row & 15;
  uint32_t out[4] = {0, 0, 0, 0};
  if (i < R && j < R) {   // a radix-R stage (R | 16) is the radix-16 tile with zero rows / columns: unused inputs may hold anything
    size_t ex = (size_t)((i * j) % R) * (N / R); if (inverse && ex) ex = N - ex;
    fe cur = ld_fe(wpow + ex);
    if (scale) cur = fe_mul(cur, ld_fe(scale + (size_t)k * kstride + m * j));
    if (inverse) cur = fe_mul(cur, factor);
    const fe two8 = fe_from_u64(256);
    for (int b = 0; b < 16; b++) {
      uint32_t limb = u < 4 ? cur.l[0] : (u < 8 ? cur.l[1] : (u < 12 ? cur.l[2] : cur.l[3]));
      out[b >> 2] |= ((limb >> (8 * (u & 3))) & 0xFF) << (8 * (b & 3));
      cur = fe_mul(cur, two8);
    }
  }
  cmat[(size_t)k * 4096 + e] = make_uint4(out[0], out[1], out[2], out[3]);
}

// src: ncols columns of n = R m elements (column stride n), R in {2, 4, 8, 16}. dst: slabs (c, k) of n elements, block i2 at offset i2 * m.
// wpow: w_Ntab^e; wstep = Ntab / n (w_n = w_Ntab^wstep). grid >= blowup CTAs, CTA b serves coset b % blowup.
__global__ void __launch_bounds__(NTT16_THREADS, 1) k_tc_cross16(const fe *__restrict__ src, int ncols, int blowup, size_t m, int R, const uint4 *__restrict__ cmat,
                                                                const fe *__restrict__ wpow, size_t Ntab, size_t wstep, fe *__restrict__ dst, bool inverse,
                                                                bool src_per_coset /* src = [column][coset][n] (the output of a level above) */,
                                                                bool dst_blocks_outer /* dst = [column R + i2][coset][m]: the columns of the level below */) {
  extern __shared__ unsigned char ntt_smem[];
  const uint32_t base = (smem_u32(ntt_smem) + 127u) & ~127u;
  unsigned char *gen = ntt_smem + (base - smem_u32(ntt_smem));
  const uint32_t b_tile = base, buf0 = base + NTT16_B_BYTES, buf1 = buf0 + NTT16_BUF_BYTES, ctl = buf1 + NTT16_BUF_BYTES;
  uint32_t *ctl_gen = reinterpret_cast<uint32_t *>(gen + NTT16_B_BYTES + 2 * NTT16_BUF_BYTES);
  const int k = blockIdx.x % blowup, rank = blockIdx.x / blowup;
  const int peers = (gridDim.x - k + blowup - 1) / blowup;
  {
    uint4 *bdst = reinterpret_cast<uint4 *>(gen);
    const uint4 *bsrc = cmat + (size_t)k * 4096;
    for (int i = threadIdx.x; i < NTT16_B_BYTES / 16; i += NTT16_THREADS) bdst[i] = bsrc[i];
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 64), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(ctl, 1); mbar_init(ctl + 8, 1); mbar_init(ctl + 16, 1); mbar_init(ctl + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl_gen[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_mma = warp >= 16;
  const int tile = (warp >> 3) & 1, half = (warp >> 2) & 1;
  const int inst = 128 * tile + 32 * (warp & 3) + lane;          // j1 - j1_0
  const uint32_t tmem_mine = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 256 * tile + 128 * half;
  const size_t n = (size_t)R * m;
  const int slices = (int)(m / 256);
  const int problems = ncols * slices;
  auto fetch = [&](int p, int which) {   // slice p -> buffer `which`: chunk j2 = 256 coefficients of block j2
    const uint32_t to = which ? buf1 : buf0, bar = ctl + 16 + 8 * which;
    mbar_expect_tx(bar, (uint32_t)R * 4096);
    const fe *col = src + (src_per_coset ? ((size_t)(p / slices) * blowup + k) : (size_t)(p / slices)) * n + (size_t)(p % slices) * 256;
#pragma unroll 1
    for (int j2 = 0; j2 < R; j2++) bulk_g2s(to + (uint32_t)j2 * NTT16_PITCH, col + m * j2, 4096, bar);
  };
  if (is_mma && lane == 0 && rank < problems) fetch(rank, 0);
  int it = 0;
  for (int p = rank; p < problems; p += peers, it++) {
    const int cur = it & 1;
    tc_fence_before();
    __syncthreads();          // every epilogue of the previous slice has read its accumulators
    if (is_mma) {
      tc_fence_after();
      if (lane == 0) {
        if (p + peers < problems) fetch(p + peers, cur ^ 1);   // the other buffer's readers (MMAs of the previous slice) are complete
        mbar_wait(ctl + 16 + 8 * cur, (uint32_t)(it >> 1) & 1);
        const uint32_t a = cur ? buf1 : buf0;
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const uint64_t adesc = tc_smem_desc(a + t * 2048, NTT16_PITCH, 128);
          const uint64_t bdesc = tc_smem_desc(b_tile, 4096, 128);
#pragma unroll
          for (int kk = 0; kk < 8; kk++)
            tc_mma_i8(tmem_base + 256 * t, adesc + (uint64_t)((2 * kk * NTT16_PITCH) >> 4), bdesc + (uint64_t)((2 * kk * 4096) >> 4), NTT16_IDESC, kk > 0);
          tc_commit(ctl + 8 * t);
        }
      }
      __syncwarp();
    } else {
      mbar_wait(ctl + 8 * tile, (uint32_t)it & 1);
      tc_fence_after();
      const int c = p / slices;
      const size_t j1 = (size_t)(p % slices) * 256 + inst;
      fe *slab = dst + ((size_t)c * blowup + k) * n;
      uint32_t acc[2][16];
      tmem_ld16(tmem_mine, acc[0]);
#pragma unroll 2
      for (int q = 0; q < 8; q++) {
        tmem_wait_ld(acc[q & 1]);
        if (q < 7) tmem_ld16(tmem_mine + 16 * (q + 1), acc[(q + 1) & 1]);
        const int i2 = 8 * half + q;
        if (i2 >= R) continue;                       // warp-uniform
        fe v = tc_epilogue_plain(acc[q & 1]);
        size_t e = (j1 * (size_t)i2 * wstep) & (Ntab - 1);
        if (inverse && e) e = Ntab - e;
        if (i2) v = lazy_mul(v, ld_fe(wpow + e));
        // any 128-bit representative: the next pass feeds the bytes to the tensor cores
        if (dst_blocks_outer) st_fe(dst + (((size_t)c * R + i2) * blowup + k) * m + j1, v);
        else st_fe(slab + (size_t)i2 * m + j1, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}
