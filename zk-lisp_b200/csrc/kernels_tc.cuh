// kernels_tc.cuh — bulk sponge kernels on the tensor-core permutation (tcperm.cuh): row hashing over the
// coset-major LDE, Merkle levels, FRI leaves, grinding and the batched primitives used by the parity tests.
// Reference semantics: zk-lisp-proof-winterfell/src/poseidon/hasher.rs:57-140 (entry points), :144-231 (sponge);
// row hashing with partitions src/agg/child.rs:1025-1045.
//
// Every kernel is persistent: grid = min(#tiles, #SMs) CTAs of 576 threads (512 compute + 2 MMA warps), one CTA per SM (the B matrix image,
// two A tiles and all 512 TMEM columns are per CTA), each looping over tiles of 256 sponges (2 sets x 128).
// Thread (set, sponge, half) owns lanes 6*half .. 6*half+5 of its sponge; the absorb of item l (state lane l,
// l < 10) is done by the thread that owns lane l.
#pragma once
#include "kernels_hash.cuh"
#include "tcperm.cuh"

// Sponge index handled by this thread in tile `tile` (sets c.active).
//   full tiles: 256 sponges, every compute warp busy — a pass costs ~119 us (two 32-sponge units per warp scheduler);
//   half tiles: 128 sponges, TMEM lane quadrant q belongs to set q & 1 and the other set's warps of that quadrant idle —
//     one unit per scheduler, ~80 us per pass (profiles/tc_pass_time_r01.log). Used by a lone prover for launches that fit one wave
//     of half tiles: the pass latency, not SM time, is what such a launch costs.
#define TC_HALF_SPONGES 128
template <bool HALF> __device__ __forceinline__ constexpr size_t tc_tile_sponges() { return HALF ? TC_HALF_SPONGES : TC_CTA_SPONGES; }
template <bool HALF>
__device__ __forceinline__ size_t tc_index(tc_ctx &c, size_t tile) {
  if constexpr (!HALF) return tile * TC_CTA_SPONGES + (size_t)c.set * TC_SPONGES + c.sponge;
  c.active = ((c.sponge >> 5) & 1) == c.set;
  return tile * TC_HALF_SPONGES + c.sponge;
}
template <bool HALF> __device__ __forceinline__ bool tc_active(const tc_ctx &c) { if constexpr (HALF) return c.active; else return true; }

// Generic sponge over a lazily produced item stream: the same contract as sponge_run (kernels_hash.cuh), but all
// 256 threads of a set run the same number of permutations (`total` must be uniform across the set).
// Returns lane 0 (canonical) in the half-0 thread of each sponge.
template <bool HALF = false, typename Src>
__device__ __forceinline__ fe tc_sponge_run(const Src &src, int total, tc_ctx &c) {
  fe st[TC_LPT];
#pragma unroll
  for (int j = 0; j < TC_LPT; j++) st[j] = fe_zero();
  if (c.half == TC_SPLIT - 1) { st[TC_LPT - 2] = ldc_fe(&c_suite.dom[0]); st[TC_LPT - 1] = ldc_fe(&c_suite.dom[1]); }
  const int lane0 = TC_LPT * c.half;
#pragma unroll 1
  for (int m = 0; m < total; m += ZKL_RATE) {
    if (tc_active<HALF>(c)) {
#pragma unroll
      for (int j = 0; j < TC_LPT; j++) {
        int l = lane0 + j;
        if (l < ZKL_RATE && m + l < total) st[j] = lazy_add(st[j], src.item(m + l));
      }
    }
    tc_permute<HALF>(st, c);
  }
  return lazy_canon(st[0]);
}

// number of tiles the persistent CTA `blockIdx.x` processes
__device__ __forceinline__ uint64_t tc_my_tiles(size_t tiles) { return blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0; }
__device__ __forceinline__ uint64_t tc_perms(int total_items) { return (uint64_t)((total_items + ZKL_RATE - 1) / ZKL_RATE); }

// ---------------------------------------------------------------- kernels
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_permute_batch(fe *states, size_t count) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (count + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t i = tc_index<HALF>(c, tile), ic = i < count ? i : count - 1;
    fe st[TC_LPT];
#pragma unroll
    for (int j = 0; j < TC_LPT; j++) st[j] = ld_fe(states + 12 * ic + TC_LPT * c.half + j);
    tc_permute<HALF>(st, c);
    if (tc_active<HALF>(c) && i < count) {
#pragma unroll
      for (int j = 0; j < TC_LPT; j++) st_fe(states + 12 * i + TC_LPT * c.half + j, lazy_canon(st[j]));
    }
  }
  tc_teardown(tc_smem);
}

// Row hashing over a coset-major LDE — same contract as k_hash_rows: sponge t = (s, q) hashes columns
// [q*psize, min(W,(q+1)*psize)) of storage row s (natural row r = s/n + blowup*(s%n)) and writes out[q*N + r].
// Sponges of one tile may straddle two partitions only if their lengths agree, so tiles are formed per partition:
// tile index = q * tiles_per_part + tile_in_part.
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_hash_rows(const fe *__restrict__ lde, size_t N, size_t n, int blowup, int W, int psize,
                                                                int nchunks, fe *__restrict__ out, size_t cstride) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles_per_part = (N + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  size_t tiles = tiles_per_part * (size_t)nchunks;
  if (c.is_mma) {
    uint64_t perms = 0;
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) perms += tc_perms(1 + (min(psize, W - (int)(tile / tiles_per_part) * psize) + 1) / 2);
    tc_mma_loop(c, perms * ZKL_ROUNDS);
  } else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    int q = (int)(tile / tiles_per_part);
    size_t s = tc_index<HALF>(c, tile % tiles_per_part), sc = s < N ? s : N - 1;
    int first = q * psize, len = min(psize, W - first);
    elems_src src{lde + (size_t)first * cstride + sc, cstride, len};
    fe d = tc_sponge_run<HALF>(src, 1 + (len + 1) / 2, c);
    if (tc_active<HALF>(c) && s < N && c.half == 0) {
      size_t r = (s / n) + (size_t)blowup * (s % n);
      st_fe(out + (size_t)q * N + r, d);
    }
  }
  tc_teardown(tc_smem);
}
// leaf = merge_many(partition digests)
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_merge_many_rows(const fe *__restrict__ parts, size_t stride, size_t N, int nchunks, fe *__restrict__ out) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (N + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * tc_perms(1 + nchunks) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t r = tc_index<HALF>(c, tile), rc = r < N ? r : N - 1;
    many_src src{parts + rc, stride};
    fe d = tc_sponge_run<HALF>(src, 1 + nchunks, c);
    if (tc_active<HALF>(c) && r < N && c.half == 0) st_fe(out + r, d);
  }
  tc_teardown(tc_smem);
}
// one Merkle level: out[i] = merge(in[2i], in[2i+1])
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_merkle_level(const fe *__restrict__ in, fe *__restrict__ out, size_t count) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (count + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t i = tc_index<HALF>(c, tile), ic = i < count ? i : count - 1;
    merge_src src{fe_zero(), fe_zero()};
    if (c.half == 0) { src.a = ld_fe(in + 2 * ic); src.b = ld_fe(in + 2 * ic + 1); }   // items 1, 2 live in lanes 1, 2 (TC_LPT >= 3)
    fe d = tc_sponge_run<HALF>(src, 3, c);
    if (tc_active<HALF>(c) && i < count && c.half == 0) st_fe(out + i, d);
  }
  tc_teardown(tc_smem);
}
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_hash_elements_batch(const fe *elems, int len, size_t count, fe *out) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (count + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * tc_perms(1 + (len + 1) / 2) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t i = tc_index<HALF>(c, tile), ic = i < count ? i : count - 1;
    elems_src src{elems + ic * (size_t)len, 1, len};
    fe d = tc_sponge_run<HALF>(src, 1 + (len + 1) / 2, c);
    if (tc_active<HALF>(c) && i < count && c.half == 0) st_fe(out + i, d);
  }
  tc_teardown(tc_smem);
}
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_merge_with_int_batch(fe seed, const uint64_t *vals, size_t count, fe *out) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (count + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t i = tc_index<HALF>(c, tile), ic = i < count ? i : count - 1;
    mwi_src src{seed, vals[ic]};
    fe d = tc_sponge_run<HALF>(src, 3, c);
    if (tc_active<HALF>(c) && i < count && c.half == 0) st_fe(out + i, d);
  }
  tc_teardown(tc_smem);
}

// coin draws in bulk (same contract as k_coin_draw, kernels_stark.cuh): out[i] = merge_with_int(*seed, counter_base + 1 + i). The
// composition coefficients of a large segment are millions of draws (one per assertion).
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_coin_draw(const fe *__restrict__ seed, uint64_t counter_base, size_t count, fe *__restrict__ out) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t tiles = (count + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * ZKL_ROUNDS);
  else {
    const fe sd = ld_fe(seed);
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      size_t i = tc_index<HALF>(c, tile), ic = i < count ? i : count - 1;
      mwi_src src{sd, counter_base + 1 + (uint64_t)ic};
      fe d = tc_sponge_run<HALF>(src, 3, c);
      if (tc_active<HALF>(c) && i < count && c.half == 0) st_fe(out + i, d);
    }
  }
  tc_teardown(tc_smem);
}

// FRI layer leaves (same contract as k_fri_leaves, kernels_stark.cuh): leaf[k + B*i'] = hash_elements([ev[k][i'], ev[k][i' + nd/2]])
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_fri_leaves(const fe *__restrict__ ev, int nd, int B, fe *__restrict__ leaves) {
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  size_t half_rows = (size_t)nd / 2, total = half_rows * B;
  size_t tiles = (total + tc_tile_sponges<HALF>() - 1) / tc_tile_sponges<HALF>();
  if (c.is_mma) tc_mma_loop(c, tc_my_tiles(tiles) * ZKL_ROUNDS);
  else for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    size_t g = tc_index<HALF>(c, tile), t = g < total ? g : total - 1;
    size_t k = t / half_rows, ip = t % half_rows;
    elems_src src{ev + k * nd + ip, half_rows, 2};
    fe d = tc_sponge_run<HALF>(src, 2, c);
    if (tc_active<HALF>(c) && g < total && c.half == 0) st_fe(leaves + (k + (size_t)B * ip), d);
  }
  tc_teardown(tc_smem);
}
// grinding (same contract as k_grind): atomicMin over nonces in [nonce_base, nonce_base + count) whose digest has
// >= bits trailing zeros in its first 8 little-endian bytes (agg/fs.rs:202-207). ONE launch covers the whole search
// window: a CTA walks its tiles in increasing nonce order and stops at the first tile that starts above the best
// nonce found so far, so at most one wave of tiles is wasted. The stop decision is data dependent, so one thread per
// set takes it and broadcasts it through shared memory to the set's compute warps AND its MMA warp (named barrier).
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_grind(const fe *__restrict__ seed, uint64_t nonce_base, uint32_t count, uint32_t bits,
                                                            unsigned long long *result) {
  if (*(volatile unsigned long long *)result < nonce_base) return;
  extern __shared__ unsigned char tc_smem[];
  tc_ctx c = tc_setup(tc_smem);
  const fe sd = ld_fe(seed);
  size_t tiles = ((size_t)count + TC_CTA_SPONGES - 1) / TC_CTA_SPONGES;
  const bool decider = !c.is_mma && c.half == 0 && c.sponge == 0;
  int it = 0;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, it ^= 1) {
    if (decider) c.flags[it] = *(volatile unsigned long long *)result < nonce_base + tile * TC_CTA_SPONGES ? 1u : 0u;
    named_bar_sync(1 + c.set, TC_SET_THREADS + 32);
    if (c.flags[it]) break;
    if (c.is_mma) { tc_mma_loop(c, ZKL_ROUNDS); continue; }
    size_t g = tc_index<false>(c, tile);
    uint64_t nonce = nonce_base + (g < count ? g : count - 1);
    mwi_src src{sd, nonce};
    fe d = tc_sponge_run<false>(src, 3, c);
    if (g < count && c.half == 0) {
      uint64_t head = ((uint64_t)d.l[1] << 32) | d.l[0];
      uint32_t tz = head == 0 ? 64 : (uint32_t)__ffsll((long long)head) - 1;
      if (tz >= bits) atomicMin(result, (unsigned long long)nonce);
    }
  }
  tc_teardown(tc_smem);
}
