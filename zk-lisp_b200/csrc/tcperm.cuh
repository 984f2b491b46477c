// tcperm.cuh — the 27-round sponge permutation (hasher.rs:173-190) with the dense 12x12 MDS product on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Why this is exact. A round is s' = MDS * s^3 + rc over f128. Write the (lazy, 256-bit) cube of lane k as
// bytes c_{k,v}, v = 0..31. Then (MDS * s^3)_i = sum_{k,v} T(i,k,v) * c_{k,v}  (mod p) with the constants
// T(i,k,v) = MDS[i][k] * 2^(8v) mod p. Splitting T into its 16 bytes T_u gives
//     (MDS * s^3)_i  ≡  sum_u 2^(8u) * ( sum_{k,v} T_u(i,k,v) * c_{k,v} )
// and the inner sums are ONE u8 x u8 -> s32 matrix product with a constant 192 x 384 matrix B
// (row n = 16 i + u, column 32 k + v): every term is < 2^16, 384 of them fit an s32 accumulator exactly.
// The CUDA cores keep what is not linear: the cubes (two 128x128-bit products and ONE reduction, the second
// reduction is absorbed by the matrix) and the carry propagation of the 16 accumulators of each output lane
// (tc_epilogue). Per state and round this replaces 144 field products (2304 IMAD.WIDE + carries) by a tensor-core
// tile, leaving 24 products on the integer pipe.
//
// Organisation: a CTA runs TWO independent sets of 128 sponges (one sponge per TMEM lane, two compute threads per
// sponge owning lanes 0..5 / 6..11: 2 x 256 compute threads) plus one MMA-issuing warp per set (576 threads).
//   shared memory: B (24 x 192 x 16 B = 72 KB, K-major, no swizzle) + per set A (24 x 128 x 16 B = 48 KB)
//   TMEM: 512 columns, set s uses columns [256 s, 256 s + 192): D[sponge][16 i + u]
// Per round and set (producer/consumer over mbarriers, no CTA-wide barrier): for j = 0..5 each compute warp writes its
// cubes of lanes j and 6 + j to A (st.shared, fence.proxy.async) and arrives on full[j]; the MMA warp waits on
// full[j] and issues the two tcgen05.mma (M 128, N 192, K 32) of those lanes, after j = 5 a tcgen05.commit on
// dready; compute warps wait on dready, tcgen05.ld their 6 x 16 columns and run the epilogue.
#pragma once
#include "f128_lazy.cuh"
#include "poseidon.cuh"

#define TC_KCH 24                 // 16-byte K chunks: 12 lanes x 32 cube bytes
#define TC_SPONGES 128            // per set = TMEM lanes = MMA M
#ifndef TC_LPT
#define TC_LPT 6                  // state lanes per compute thread: 6 (two threads per sponge) or 4 (three threads per sponge)
#endif
#define TC_SPLIT (12 / TC_LPT)    // compute threads per sponge
#ifndef TC_GRP
#define TC_GRP 1                  // lanes a thread cubes between two publishes to the MMA warp (divides TC_LPT)
#endif
#define TC_NGRP (TC_LPT / TC_GRP)
#define TC_SET_THREADS (TC_SPONGES * TC_SPLIT)
#define TC_SETS 2
#define TC_THREADS (TC_SETS * TC_SET_THREADS + 32 * TC_SETS)   // compute warps + one MMA warp per set
#define TC_N 192                  // MMA N = 12 output lanes x 16 bytes
#define TC_B_BYTES (TC_KCH * TC_N * 16)
#define TC_A_BYTES (TC_KCH * TC_SPONGES * 16)
#define TC_CTL_BYTES 256          // mbarriers + TMEM base address
#define TC_SMEM_BYTES (TC_B_BYTES + TC_SETS * TC_A_BYTES + TC_CTL_BYTES + 128)
#define TC_TMEM_COLS 512
#define TC_SET_COLS 256
#define TC_CTA_SPONGES (TC_SETS * TC_SPONGES)

#if defined(__CUDACC__)
// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], u8 x u8 -> s32, issued by ONE thread
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// wait for this thread's outstanding tcgen05.ld; the registers are threaded through so no use can be hoisted above it
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,n),2):((16 B, SBO), LBO)):
// bits [0,14) start >> 4, [16,30) LBO >> 4 (between the two 16-byte K chunks of one MMA), [32,46) SBO >> 4
// (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// instruction descriptor: c_format S32 (2) at [4,6), a/b format u8 (0) at [7,10)/[10,13), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
#define TC_IDESC ((2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_SPONGES >> 4) << 24))

// B matrix image (exact shared-memory layout), built once per suite by k_build_mds_bytes
__device__ uint4 g_mds_bytes[TC_B_BYTES / 16];

// one thread per (chunk c, row n): 16 bytes = byte u of MDS[i][k] * 2^(8 (16 (c & 1) + b)) mod p, b = 0..15, n = 16 i + u, k = c >> 1
__global__ void k_build_mds_bytes(const suite_consts *suite) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= TC_KCH * TC_N) return;
  int c = t / TC_N, n = t % TC_N;
  int i = n >> 4, u = n & 15, k = c >> 1;
  fe m = suite->mds[i][k];
  fe sh = fe_pow(fe_from_u64(2), (uint64_t)(128 * (c & 1)));   // 2^(8*16*(c&1))
  fe cur = fe_mul(m, sh);
  fe two8 = fe_from_u64(256);
  uint32_t out[4] = {0, 0, 0, 0};
  for (int b = 0; b < 16; b++) {
    uint32_t byte = (cur.l[u >> 2] >> (8 * (u & 3))) & 0xFF;
    out[b >> 2] |= byte << (8 * (b & 3));
    cur = fe_mul(cur, two8);
  }
  g_mds_bytes[t] = make_uint4(out[0], out[1], out[2], out[3]);
}

#endif  // __CUDACC__ (device part 1)
// Host-side check of tc_epilogue's precondition for a given MDS matrix: the largest accumulator any state can
// produce is 255 * (largest row sum of the byte matrix). Returns that bound.
static inline uint32_t tc_max_accumulator(const suite_consts &su) {
  uint32_t best = 0;
  for (int i = 0; i < 12; i++) {
    uint32_t sum[16] = {0};
    for (int k = 0; k < 12; k++) {
      fe cur = su.mds[i][k];
      for (int v = 0; v < 32; v++) {
        for (int u = 0; u < 16; u++) sum[u] += (cur.l[u >> 2] >> (8 * (u & 3))) & 0xFF;
        cur = fe_mul(cur, fe_from_u64(256));
      }
    }
    for (int u = 0; u < 16; u++) if (sum[u] > best) best = sum[u];
  }
  return best * 255u;
}
#if defined(__CUDACC__)
// ---------------------------------------------------------------- per-thread context
// Warp roles: warps 0..15 compute (set = warp / 8; sponge = 32 (warp & 3) + lane — the TMEM lane quadrant a warp may
// read is warp % 4; half = (warp >> 2) & 1), warps 16 and 17 issue the MMAs of set 0 and set 1.
// Shared-memory control block per set (64 bytes): full[j], j < TC_LPT (one arrival per compute warp, after the
// warp's cubes of lanes j, TC_LPT + j, ... are in A), then dready (tcgen05.commit of the round's last MMA).
struct tc_ctx {
  uint32_t a_addr;    // shared address of this sponge's 16-byte slot in chunk 0 of the set's A tile
  uint32_t a_tile;    // shared address of the set's A tile
  uint32_t b_tile;    // shared address of B
  uint32_t tmem;      // TMEM address: lane quadrant of this warp, first column of this thread's 6 output lanes
  uint32_t tmem_d;    // TMEM address of the set's accumulator (lane 0)
  uint32_t full;      // shared address of full[0] of the set
  uint32_t dready;    // shared address of the set's accumulator-ready barrier
  uint32_t phase;     // phase parity of the current round (all barriers of a set flip once per round)
  volatile uint32_t *flags;   // generic pointer to the set's two control words (tile-level decisions broadcast to the MMA warp)
  int set, half, sponge;
  bool is_mma;        // MMA-issuing warp (no sponge)
  bool active;        // warp-uniform: this warp's 32 sponges carry work in the current tile (half-tile mode idles every other quadrant)
};
#define TC_COMPUTE_WARPS_PER_SET (TC_SET_THREADS / 32)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// CTA prologue: copies B, allocates TMEM, initialises the barriers. `smem` = dynamic shared memory base (>= TC_SMEM_BYTES).
__device__ __forceinline__ tc_ctx tc_setup(unsigned char *smem_raw) {
  uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint32_t b_tile = base, a0 = base + TC_B_BYTES, ctl = a0 + TC_SETS * TC_A_BYTES;
  unsigned char *gen = smem_raw + (base - smem_u32(smem_raw));
  // B image -> shared
  uint4 *bdst = reinterpret_cast<uint4 *>(gen);
  for (int i = threadIdx.x; i < TC_B_BYTES / 16; i += TC_THREADS) bdst[i] = g_mds_bytes[i];
  uint32_t *ctl_gen = reinterpret_cast<uint32_t *>(gen + TC_B_BYTES + TC_SETS * TC_A_BYTES);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 128), "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < TC_SETS; s++) {
      for (int j = 0; j < TC_NGRP; j++) mbar_init(ctl + 64 * s + 8 * j, TC_COMPUTE_WARPS_PER_SET);
      mbar_init(ctl + 64 * s + 48, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async_smem();   // B was written through the generic proxy, the MMA reads it through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base = ctl_gen[32];
  tc_ctx c;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  c.is_mma = warp >= TC_SETS * TC_COMPUTE_WARPS_PER_SET;
  c.set = c.is_mma ? warp - TC_SETS * TC_COMPUTE_WARPS_PER_SET : warp / TC_COMPUTE_WARPS_PER_SET;
  c.half = (warp % TC_COMPUTE_WARPS_PER_SET) >> 2;   // which group of TC_LPT lanes this thread owns
  c.sponge = 32 * (warp & 3) + lane;
  c.a_tile = a0 + c.set * TC_A_BYTES;
  c.a_addr = c.a_tile + c.sponge * 16;
  c.b_tile = b_tile;
  c.tmem_d = tmem_base + c.set * TC_SET_COLS;
  c.tmem = c.tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + 16 * TC_LPT * c.half;
  c.full = ctl + 64 * c.set;
  c.dready = c.full + 48;
  c.phase = 0;
  c.flags = ctl_gen + 48 + 2 * c.set;
  c.active = true;
  return c;
}
// CTA epilogue (all threads)
__device__ __forceinline__ void tc_teardown(unsigned char *smem_raw) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t *ctl_gen = reinterpret_cast<const uint32_t *>(smem_raw + (base - smem_u32(smem_raw)) + TC_B_BYTES + TC_SETS * TC_A_BYTES);
    uint32_t tmem_base = ctl_gen[32];
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// MMA warp: `rounds` = 27 x (number of permutations the set's compute warps will run in this kernel). Per round the
// MMAs of lanes j, TC_LPT + j, ... (one MMA = one lane = 32 cube bytes of K) are issued as soon as all compute warps
// have published those cubes, so only the last group is exposed latency; the commit after the last group signals dready.
__device__ __forceinline__ void tc_mma_loop(tc_ctx &c, uint64_t rounds) {
  const uint64_t adesc = tc_smem_desc(c.a_tile, TC_SPONGES * 16, 128);
  const uint64_t bdesc = tc_smem_desc(c.b_tile, TC_N * 16, 128);
  const bool leader = (threadIdx.x & 31) == 0;
  uint32_t phase = c.phase;
#pragma unroll 1
  for (uint64_t r = 0; r < rounds; r++) {
#pragma unroll
    for (int q = 0; q < TC_NGRP; q++) {
      mbar_wait(c.full + 8 * q, phase);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int jj = 0; jj < TC_GRP; jj++) {
#pragma unroll
          for (int g = 0; g < TC_SPLIT; g++) {   // one MMA = one state lane = 32 cube bytes of K
            const int l = g * TC_LPT + q * TC_GRP + jj;
            tc_mma_i8(c.tmem_d, adesc + (uint64_t)((2 * l * TC_SPONGES * 16) >> 4), bdesc + (uint64_t)((2 * l * TC_N * 16) >> 4), TC_IDESC, (q | jj | g) > 0);
          }
        }
        if (q == TC_NGRP - 1) tc_commit(c.dready);
      }
      __syncwarp();
    }
    phase ^= 1;
  }
  c.phase = phase;
}

// A permutation's worth of barrier traffic for a compute warp whose 32 sponges carry no work in this tile: the set's other
// warps and its MMA warp see the arrivals they count on, nothing is cubed, stored or read back (the warp's rows of the A tile
// keep stale bytes; the accumulator rows they produce are never loaded).
__device__ __forceinline__ void tc_permute_idle(tc_ctx &c) {
  const bool leader = (threadIdx.x & 31) == 0;
#pragma unroll 1
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    __syncwarp();
    if (leader) {
#pragma unroll
      for (int q = 0; q < TC_NGRP; q++) mbar_arrive(c.full + 8 * q);
    }
    mbar_wait(c.dready, c.phase);   // one round at a time: full[] must not see two arrivals of one warp in a phase
    c.phase ^= 1;
  }
}

// The permutation, compute side. Called by all compute threads of a set together; st[j] = lazy lane TC_LPT*half + j.
// HALF (half-tile kernels only): a warp without work idles through the rounds. The full-tile instantiation carries no trace of it.
template <bool HALF = false>
__device__ __forceinline__ void tc_permute(fe (&st)[TC_LPT], tc_ctx &c) {
  if constexpr (HALF) { if (!c.active) { tc_permute_idle(c); return; } }
  const int lane0 = TC_LPT * c.half;
  const bool leader = (threadIdx.x & 31) == 0;
#pragma unroll 1
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    tc_fence_before();              // this thread's tcgen05.ld of the previous round are complete (wait::ld) and ordered
    // cubes: x^2 reduced lazily, x^2 * x left as 256 raw bits = two K chunks of this sponge's A row
#pragma unroll
    for (int j = 0; j < TC_LPT; j++) {
      fe sq = lazy_sqr(st[j]);
      uint32_t x[8]; lazy_mul_wide(sq, st[j], x);
      uint32_t a = c.a_addr + (uint32_t)(2 * (lane0 + j)) * (TC_SPONGES * 16);
      sts_u4(a, x[0], x[1], x[2], x[3]);
      sts_u4(a + TC_SPONGES * 16, x[4], x[5], x[6], x[7]);
      if ((j + 1) % TC_GRP == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (leader) mbar_arrive(c.full + 8 * (j / TC_GRP));
      }
    }
    mbar_wait(c.dready, c.phase);
    c.phase ^= 1;
    tc_fence_after();
    uint32_t acc[2][16];
    tmem_ld16(c.tmem, acc[0]);
#pragma unroll
    for (int j = 0; j < TC_LPT; j++) {
      tmem_wait_ld(acc[j & 1]);
      if (j < TC_LPT - 1) tmem_ld16(c.tmem + 16 * (j + 1), acc[(j + 1) & 1]);
      st[j] = tc_epilogue(acc[j & 1], ldc_fe(&c_suite.rc[r][lane0 + j]));
    }
  }
}
#endif
