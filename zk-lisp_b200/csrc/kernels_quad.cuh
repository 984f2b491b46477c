// kernels_quad.cuh — the lowest-latency sponge: 64 threads (two warps) run ONE sponge.
// Reference semantics: zk-lisp-proof-winterfell/src/poseidon/hasher.rs:144-231 (sponge), :173-190 (permutation).
//
// Why. A segment proof walks ~250 permutations that depend on each other (16 + 16 Merkle levels of the two
// commitment trees, 5..15 levels for each of the 11 FRI layer trees, two coin permutations per layer, the 22 permutations
// of the OOD-frame digest): with one stream per segment the proof's wall time is that chain times the latency of ONE
// permutation. The 16-thread cooperative sponge (kernels_hash.cuh) spends ~2 800 clk per round, most of it in the 192
// dependent-by-pipe IMAD.WIDE of a thread's MDS row (5.2 clk each per warp, profiles/probe_pipes2_r01.log).
// Here thread (l, h), l < 12, h < 4, owns a QUARTER of row l: 3 of its 12 products; the four partial sums are reduced,
// exchanged with two xor-shuffles and added, so every thread of a lane holds the new lane value and cubes it again
// (redundantly — SIMT makes that free). Per round and warp: 26 + 48 + 6 wide multiplies instead of 224.
//
// Thread t of a group: lane l = t >> 2 (t >= 48 idles along on clamped indices), quarter h = t & 3; lanes 0..7 live in
// the group's first warp, 8..11 in the second. Cubes are published through a double-buffered shared-memory slot, one
// 64-thread named barrier per round. The state is canonical between rounds (reduce9 + fe_add), the cube is lazy
// (f128_lazy.cuh: any 128-bit representative) — the unreduced row accumulator takes either.
#pragma once
#include "f128_lazy.cuh"
#include "kernels_hash.cuh"

#define ZKL_QUAD_GROUP 64
#define ZKL_QUAD_THREADS 256   // 4 sponges per CTA (named barriers 1..4)
#define ZKL_QUAD_MAX_GROUPS 8  // the fused kernels (k_tree_top, k_reseed_draw) run up to 8 groups in one CTA

struct quad_ctx {
  int l, h;        // state lane (clamped to 11), quarter
  bool writer;     // publishes the lane's cube (h == 0 and a real lane)
  int bar;         // named barrier of the group
  fe m[3];         // MDS[l][3h .. 3h+2]
  fe *xch;         // the group's 2 x 12 exchange slots
};
__device__ __forceinline__ void quad_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ quad_ctx quad_setup(fe *xch_all) {
  quad_ctx q;
  const int t = threadIdx.x & (ZKL_QUAD_GROUP - 1), g = threadIdx.x / ZKL_QUAD_GROUP;
  const int l = t >> 2;
  q.l = l < 12 ? l : 11;
  q.h = t & 3;
  q.writer = l < 12 && q.h == 0;
  q.bar = 1 + g;
  q.xch = xch_all + g * 24;
#pragma unroll
  for (int k = 0; k < 3; k++) q.m[k] = ld_fe(&g_suite.mds[q.l][3 * q.h + k]);
  return q;
}

// `parity`: which exchange buffer round 0 uses. ZKL_ROUNDS is odd, so a sponge that runs several permutations must flip it per
// permutation: otherwise the last round of one permutation and the first round of the next share a buffer with no barrier
// between the slow warp's reads and the fast warp's writes (a real race with two warps per sponge: it corrupted a few leaf
// digests per thousand rows under load — the proofs still verified, the bytes differed).
__device__ __forceinline__ void quad_permute(fe &st, const quad_ctx &q, int parity = 0) {
#pragma unroll 1
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    fe *buf = q.xch + ((r + parity) & 1) * 12;
    fe c = lazy_mul(lazy_sqr(st), st);
    if (q.writer) sts_fe(buf + q.l, c);
    quad_bar(q.bar);
    wide_acc w; wacc_zero(w);
#pragma unroll
    for (int k = 0; k < 3; k++) wacc_mad(w, lds_fe(buf + 3 * q.h + k), q.m[k]);
    if (q.h == 0) wacc_add_fe(w, ld_fe(&g_suite.rc[r][q.l]));
    fe part = wacc_reduce(w);
#pragma unroll
    for (int x = 1; x <= 2; x <<= 1) {
      fe o;
#pragma unroll
      for (int i = 0; i < 4; i++) o.l[i] = __shfl_xor_sync(0xFFFFFFFFu, part.l[i], x);
      part = fe_add(part, o);
    }
    st = part;
  }
}

// Same contract as coop_sponge_run; the digest is returned in the group's FIRST warp (threads 0..31 of the group).
template <typename Src>
__device__ __forceinline__ fe quad_sponge_run(const Src &src, int total) {
  __shared__ fe s_qxch[ZKL_QUAD_MAX_GROUPS * 24];
  const quad_ctx q = quad_setup(s_qxch);
  fe st = fe_zero();
  if (q.l == 10) st = ld_fe(&g_suite.dom[0]);
  if (q.l == 11) st = ld_fe(&g_suite.dom[1]);
  int parity = 0;
#pragma unroll 1
  for (int m = 0; m < total; m += ZKL_RATE, parity ^= 1) {
    if (q.l < ZKL_RATE && m + q.l < total) st = fe_add(st, src.item(m + q.l));
    quad_permute(st, q, parity);
  }
  fe out;
#pragma unroll
  for (int i = 0; i < 4; i++) out.l[i] = __shfl_sync(0xFFFFFFFFu, st.l[i], 0);
  return out;
}
__device__ __forceinline__ fe quad_merge(const fe &a, const fe &b) { merge_src s{a, b}; return quad_sponge_run(s, 3); }
__device__ __forceinline__ fe quad_merge_with_int(const fe &seed, uint64_t v) { mwi_src s{seed, v}; return quad_sponge_run(s, 3); }
__device__ __forceinline__ fe quad_hash_elements(const fe *base, size_t stride, int len) {
  elems_src s{base, stride, len};
  return quad_sponge_run(s, 1 + (len + 1) / 2);
}
__device__ __forceinline__ bool quad_leader() { return (threadIdx.x & (ZKL_QUAD_GROUP - 1)) == 0; }
__device__ __forceinline__ size_t quad_index() { return ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_QUAD_GROUP; }

// Group-size generic helpers for the fused kernels: G = 64 (this file) or G = 16 (coop_sponge_run, kernels_hash.cuh — 4x fewer
// warp instructions per permutation: the choice when many segment streams share the GPU and SM time, not latency, is scarce).
template <int G, typename Src>
__device__ __forceinline__ fe grp_sponge_run(const Src &src, int total) {
  if constexpr (G == ZKL_QUAD_GROUP) return quad_sponge_run(src, total);
  else return coop_sponge_run(src, total);
}
template <int G> __device__ __forceinline__ fe grp_merge(const fe &a, const fe &b) { merge_src s{a, b}; return grp_sponge_run<G>(s, 3); }
template <int G> __device__ __forceinline__ fe grp_merge_with_int(const fe &seed, uint64_t v) { mwi_src s{seed, v}; return grp_sponge_run<G>(s, 3); }
#define ZKL_GRP_MAX_GROUPS 8   // groups per CTA of the fused kernels (both sizes: 8 x 16 = ZKL_COOP_THREADS, 8 x 64 = 512)

// ---------------------------------------------------------------- kernels (one group per sponge; out-of-range groups run on clamped inputs)
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_merkle_level(const fe *__restrict__ in, fe *__restrict__ out, size_t count) {
  size_t g = quad_index(), gi = g < count ? g : count - 1;
  fe d = quad_merge(ld_fe(in + 2 * gi), ld_fe(in + 2 * gi + 1));
  if (g < count && quad_leader()) st_fe(out + g, d);
}
// The top of a Merkle tree in ONE launch: levels count, count/2, ..., 1 (count <= ZKL_GRP_MAX_GROUPS), single CTA of
// G * count threads. in: 2*count digests; nodes as in dev_merkle (nodes[lvl + i] = node i of the level with lvl nodes).
template <int G>
__global__ void __launch_bounds__(ZKL_GRP_MAX_GROUPS * G) k_tree_top(const fe *__restrict__ in, fe *__restrict__ nodes, int count) {
  const int g = threadIdx.x / G;
  const fe *src = in;
  for (int lvl = count; lvl >= 1; lvl >>= 1) {
    if (g < lvl) {   // whole groups skip: the barriers inside the sponge are per group
      fe d = grp_merge<G>(ld_fe(src + 2 * g), ld_fe(src + 2 * g + 1));
      if ((threadIdx.x & (G - 1)) == 0) st_fe(nodes + lvl + g, d);
    }
    __syncthreads();   // orders the level's global stores before the next level's loads (same CTA)
    src = nodes + lvl;
  }
}
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_merge_many_rows(const fe *__restrict__ parts, size_t stride, size_t rows, int nchunks, fe *__restrict__ out) {
  size_t g = quad_index(), r = g < rows ? g : rows - 1;
  many_src s{parts + r, stride};
  fe d = quad_sponge_run(s, 1 + nchunks);
  if (g < rows && quad_leader()) st_fe(out + r, d);
}
// coset-major row hashing, one group per (row, chunk) — same contract as k_hash_rows
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_hash_rows(const fe *__restrict__ lde, size_t N, size_t n, int blowup, int W, int psize,
                                                                     int nchunks, fe *__restrict__ out, size_t cstride) {
  size_t total = N * (size_t)nchunks;
  size_t g = quad_index(), t = g < total ? g : total - 1;
  size_t s = t % N; int q = (int)(t / N);
  size_t r = (s / n) + (size_t)blowup * (s % n);
  int first = q * psize, len = min(psize, W - first);
  fe d = quad_hash_elements(lde + (size_t)first * cstride + s, cstride, len);
  if (g < total && quad_leader()) st_fe(out + (size_t)q * N + r, d);
}
// FRI layer leaves — same contract as k_fri_leaves (kernels_stark.cuh)
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_fri_leaves(const fe *__restrict__ ev, int nd, int B, fe *__restrict__ leaves) {
  size_t half_rows = (size_t)nd / 2, total = half_rows * B;
  size_t g = quad_index(), t = g < total ? g : total - 1;
  size_t k = t / half_rows, ip = t % half_rows;
  fe d = quad_hash_elements(ev + k * nd + ip, half_rows, 2);
  if (g < total && quad_leader()) st_fe(leaves + (k + (size_t)B * ip), d);
}

// ---------------------------------------------------------------- Fiat-Shamir coin (winter-crypto DefaultRandomCoin [WF-mem])
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_coin_init(const fe *__restrict__ elems, int count, fe *seed) {
  fe d = quad_hash_elements(elems, 1, count);
  if (threadIdx.x == 0) st_fe(seed, d);
}
// seed = merge(seed, digest) [digest also copied to *copy_to when given], then out[i] = merge_with_int(seed, 1 + i), i < draws:
// a reseed followed by the draws that depend on it, in one launch. One CTA of G * (1..8) threads; the draws are dealt to its groups.
template <int G>
__global__ void __launch_bounds__(ZKL_GRP_MAX_GROUPS * G) k_reseed_draw(fe *seed, const fe *__restrict__ digest, fe *copy_to, int draws, fe *__restrict__ out) {
  __shared__ fe s_seed;
  const int groups = blockDim.x / G, g = threadIdx.x / G;
  const fe dg = ld_fe(digest);
  if (g == 0) {
    fe d = grp_merge<G>(ld_fe(seed), dg);
    if (threadIdx.x == 0) { s_seed = d; st_fe(seed, d); if (copy_to) st_fe(copy_to, dg); }
  }
  __syncthreads();
  for (int base = 0; base < draws; base += groups) {   // uniform trip count; whole groups without a draw skip the sponge
    int i = base + g;
    if (i < draws) {
      fe d = grp_merge_with_int<G>(s_seed, (uint64_t)i + 1);
      if ((threadIdx.x & (G - 1)) == 0) st_fe(out + i, d);
    }
  }
}
// Query positions in three short launches instead of one CTA walking the 64 draws in turn (agg/fs.rs:223-228):
//   k_seed_with_nonce: seed' = merge_with_int(seed, nonce);  k_*_coin_draw: value_i = next() in parallel;
//   k_positions_finish: mask, sort, de-duplicate.
template <int G>
__global__ void __launch_bounds__(G) k_seed_with_nonce(const fe *__restrict__ seed, const unsigned long long *__restrict__ nonce, fe *__restrict__ out_seed) {
  fe d = grp_merge_with_int<G>(ld_fe(seed), (uint64_t)*nonce);
  if (threadIdx.x == 0) st_fe(out_seed, d);
}
// positions[0..*count) = sorted, de-duplicated (u64_le(draw_i[..8]) & mask), i < num_queries <= 255; one warp
__global__ void k_positions_finish(const fe *__restrict__ draws, int num_queries, uint32_t mask, uint32_t *positions, uint32_t *count) {
  __shared__ uint32_t s_vals[256];
  for (int i = threadIdx.x; i < num_queries; i += blockDim.x) s_vals[i] = ld_fe(draws + i).l[0] & mask;  // N <= 2^32
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < num_queries; i++) {  // insertion sort
      uint32_t v = s_vals[i]; int j = i - 1;
      while (j >= 0 && s_vals[j] > v) { s_vals[j + 1] = s_vals[j]; j--; }
      s_vals[j + 1] = v;
    }
    uint32_t c = 0;
    for (int i = 0; i < num_queries; i++) if (i == 0 || s_vals[i] != s_vals[i - 1]) positions[c++] = s_vals[i];
    *count = c;
  }
}
// out[i] = merge_with_int(seed, counter_base + 1 + i), i < count (many draws: composition / DEEP coefficients)
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_coin_draw(const fe *__restrict__ seed, uint64_t counter_base, int count, fe *__restrict__ out) {
  int g = (int)quad_index(), gi = g < count ? g : count - 1;
  fe d = quad_merge_with_int(ld_fe(seed), counter_base + 1 + (uint64_t)gi);
  if (g < count && quad_leader()) st_fe(out + g, d);
}
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_hash_elements_single(const fe *__restrict__ elems, int count, fe *out) {
  fe d = quad_hash_elements(elems, 1, count);
  if (threadIdx.x == 0) st_fe(out, d);
}
