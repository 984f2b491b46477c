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
