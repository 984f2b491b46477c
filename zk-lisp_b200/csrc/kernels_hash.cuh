// kernels_hash.cuh — row hashing, Merkle levels, coin primitives (device code).
// Reference semantics: zk-lisp-proof-winterfell/src/poseidon/hasher.rs:57-140 (entry points),
// :144-231 (sponge); row hashing with partitions src/agg/child.rs:1025-1045.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------- generic sponge over a lazily produced item stream
// `Src::item(m)` returns absorbed element m (m = 0 is the domain label). Every block of up to 10
// absorbs is followed by exactly one permutation (hasher.rs:193-201 + the final `lane != 0` permute,
// hasher.rs:226-228), i.e. ceil(total/10) permutations.
template <typename Src>
__device__ __forceinline__ fe sponge_run(const Src &src, int total) {
  __shared__ fe s_state[12 * ZKL_HASH_THREADS];
  fe *st = s_state + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 10; i++) sts_fe(st + i * ZKL_HASH_THREADS, fe_zero());
  sts_fe(st + 10 * ZKL_HASH_THREADS, ldc_fe(&c_suite.dom[0]));
  sts_fe(st + 11 * ZKL_HASH_THREADS, ldc_fe(&c_suite.dom[1]));
  int m = 0;
#pragma unroll 1
  while (m < total) {
    int take = min(ZKL_RATE, total - m);
#pragma unroll 1
    for (int l = 0; l < take; l++) {
      fe *slot = st + l * ZKL_HASH_THREADS;
      sts_fe(slot, fe_add(lds_fe(slot), src.item(m + l)));
    }
    permute(st);
    m += ZKL_RATE;
  }
  return lds_fe(st);
}

__device__ __forceinline__ fe ld_fe(const fe *p) {
  uint4 v = *reinterpret_cast<const uint4 *>(p);
  fe r; r.l[0] = v.x; r.l[1] = v.y; r.l[2] = v.z; r.l[3] = v.w; return r;
}
__device__ __forceinline__ void st_fe(fe *p, const fe &a) {
  *reinterpret_cast<uint4 *>(p) = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
}

// hash_elements over `len` canonical elements at base[(first + j) * stride]  (hasher.rs:126-139):
// item m>=1 folds elements 2(m-1), 2(m-1)+1 as e0 + e1*2^64 (utils.rs:359-371); odd tail alone.
struct elems_src {
  const fe *base; size_t stride; int len;
  __device__ __forceinline__ fe item(int m) const {
    if (m == 0) return c_labels.elements;
    int j = 2 * (m - 1);
    fe e0 = ld_fe(base + (size_t)j * stride);
    if (j + 1 < len) return fe_fold_pair(e0, ld_fe(base + (size_t)(j + 1) * stride));
    return e0;
  }
};
__device__ __forceinline__ fe hash_elements_dev(const fe *base, size_t stride, int len) {
  elems_src s{base, stride, len};
  return sponge_run(s, 1 + (len + 1) / 2);
}

// digests are carried as one element (upper 16 digest bytes are zero, utils.rs:375-381), so a digest
// chunk folds to the element itself.
struct merge_src {
  fe a, b;
  __device__ __forceinline__ fe item(int m) const { return m == 0 ? c_labels.merge : (m == 1 ? a : b); }
};
__device__ __forceinline__ fe merge_dev(const fe &a, const fe &b) {  // hasher.rs:72-85
  merge_src s{a, b};
  return sponge_run(s, 3);
}
struct mwi_src {
  fe seed; uint64_t v;
  __device__ __forceinline__ fe item(int m) const {
    return m == 0 ? c_labels.merge_with_int : (m == 1 ? seed : fe_from_u64(v));
  }
};
__device__ __forceinline__ fe merge_with_int_dev(const fe &seed, uint64_t v) {  // hasher.rs:107-120
  mwi_src s{seed, v};
  return sponge_run(s, 3);
}
struct many_src {
  const fe *d; size_t stride;
  __device__ __forceinline__ fe item(int m) const { return m == 0 ? c_labels.merge_many : ld_fe(d + (size_t)(m - 1) * stride); }
};
__device__ __forceinline__ fe merge_many_dev(const fe *d, size_t stride, int count) {  // hasher.rs:87-105
  many_src s{d, stride};
  return sponge_run(s, 1 + count);
}

// ---------------------------------------------------------------- kernels
// Row hashing over a coset-major LDE: column c lives at lde + c*N, storage index s = k*n + i holds
// the natural LDE row r = k + blowup*i. Thread (s, q) hashes columns [q*psize, min(W,(q+1)*psize))
// of row r and writes digest to out[q*N + r]. `cstride` = distance between columns (N for a whole matrix; the column-sharded prover
// hashes the rows of a coset range of a larger matrix: lde offset to the first coset, N = rows in the range, cstride = the whole N).
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_hash_rows(const fe *__restrict__ lde, size_t N, size_t n, int blowup, int W, int psize,
                                                   int nchunks, fe *__restrict__ out, size_t cstride) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * (size_t)nchunks) return;
  size_t s = t % N; int q = (int)(t / N);
  size_t r = (s / n) + (size_t)blowup * (s % n);
  int first = q * psize, len = min(psize, W - first);
  fe d = hash_elements_dev(lde + (size_t)first * cstride + s, cstride, len);
  st_fe(out + (size_t)q * N + r, d);
}
// leaf = merge_many(partition digests) (agg/child.rs:1040-1044)
// parts[q * stride + r], r < rows
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_merge_many_rows(const fe *__restrict__ parts, size_t stride, size_t rows, int nchunks, fe *__restrict__ out) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  st_fe(out + r, merge_many_dev(parts + r, stride, nchunks));
}
// one Merkle level: out[i] = merge(in[2i], in[2i+1]), i < count
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_merkle_level(const fe *__restrict__ in, fe *__restrict__ out, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fe(out + i, merge_dev(ld_fe(in + 2 * i), ld_fe(in + 2 * i + 1)));
}
// batched primitives for the parity tests
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_permute_batch(fe *states, size_t count) {
  __shared__ fe s_state[12 * ZKL_HASH_THREADS];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  fe *st = s_state + threadIdx.x;
#pragma unroll
  for (int l = 0; l < 12; l++) sts_fe(st + l * ZKL_HASH_THREADS, ld_fe(states + 12 * i + l));
  permute(st);
#pragma unroll
  for (int l = 0; l < 12; l++) st_fe(states + 12 * i + l, lds_fe(st + l * ZKL_HASH_THREADS));
}
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_hash_elements_batch(const fe *elems, int len, size_t count, fe *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fe(out + i, hash_elements_dev(elems + i * (size_t)len, 1, len));
}
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_merge_with_int_batch(fe seed, const uint64_t *vals, size_t count, fe *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fe(out + i, merge_with_int_dev(seed, vals[i]));
}
// one seed per value: out[i] = merge_with_int(seeds[i], vals[i]) — the coin draws of many child proofs in lockstep (replay)
__global__ void __launch_bounds__(ZKL_HASH_THREADS) k_merge_with_int_pairs(const fe *__restrict__ seeds, const uint64_t *__restrict__ vals, size_t count,
                                                                         fe *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  st_fe(out + i, merge_with_int_dev(ld_fe(seeds + i), vals[i]));
}
// generic byte-chunk sponge: `chunks` = pre-folded chunk elements (label excluded), label selects the entry point
__global__ void k_sponge_single(const fe *chunks, int count, int label_kind, fe *out) {
  if (threadIdx.x || blockIdx.x) return;
  struct src_t {
    const fe *c; int kind;
    __device__ __forceinline__ fe item(int m) const {
      if (m == 0) return kind == 0 ? c_labels.bytes : (kind == 1 ? c_labels.merge_many : c_labels.elements);
      return ld_fe(c + m - 1);
    }
  } s{chunks, label_kind};
  st_fe(out, sponge_run(s, 1 + count));
}

// ================================================================= cooperative (low-latency) sponge
// A 16-thread group runs ONE sponge: thread l < 12 owns lane l. Per round each thread cubes its lane,
// publishes it through a double-buffered shared-memory slot, and accumulates its own MDS row. One
// permutation costs ~650 instructions per thread per round instead of ~6800 for the thread-per-state
// kernel, i.e. ~10x lower latency — used wherever the number of independent sponges cannot fill the GPU
// (Merkle levels, FRI layer trees, the Fiat-Shamir coin, grinding tails, the OOD frame digest).
// The suite is read from a global-memory copy (per-thread MDS rows diverge, which __constant__ serialises).
#define ZKL_COOP_GROUP 16
#define ZKL_COOP_THREADS 128
__device__ suite_consts g_suite;  // global-memory mirror of c_suite (written by zkl_suite_load)

// `parity`: exchange buffer of round 0; flipped per permutation by multi-permutation sponges (ZKL_ROUNDS is odd: the last round
// of one permutation and the first of the next must not share a buffer, there is no barrier between them)
__device__ __forceinline__ void coop_permute(fe &st, int l, fe *xch /* this group's 2 x 12 slots */, uint32_t gmask, int parity = 0) {
  const bool active = l < 12;
#pragma unroll 1
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    fe *buf = xch + ((r + parity) & 1) * 12;
    if (active) sts_fe(buf + l, fe_cube(st));
    __syncwarp(gmask);  // group-scoped: the two groups of a warp may run different numbers of permutations
    if (active) {
      wide_acc w; wacc_zero(w);
#pragma unroll
      for (int k = 0; k < 12; k++) wacc_mad(w, lds_fe(buf + k), ld_fe(&g_suite.mds[l][k]));
      wacc_add_fe(w, ld_fe(&g_suite.rc[r][l]));
      st = wacc_reduce(w);
    }
  }
}
// returns the digest element in every thread of the group
template <typename Src>
__device__ __forceinline__ fe coop_sponge_run(const Src &src, int total) {
  __shared__ fe s_xch[(ZKL_COOP_THREADS / ZKL_COOP_GROUP) * 24];
  const int l = threadIdx.x & (ZKL_COOP_GROUP - 1);
  fe *xch = s_xch + (threadIdx.x / ZKL_COOP_GROUP) * 24;
  const uint32_t gmask = 0xFFFFu << ((threadIdx.x & 16));
  fe st = fe_zero();
  if (l == 10) st = ld_fe(&g_suite.dom[0]);
  if (l == 11) st = ld_fe(&g_suite.dom[1]);
  int parity = 0;
#pragma unroll 1
  for (int m = 0; m < total; m += ZKL_RATE, parity ^= 1) {
    if (l < ZKL_RATE && m + l < total) st = fe_add(st, src.item(m + l));
    coop_permute(st, l, xch, gmask, parity);
  }
  fe out;
  const int src_lane = (threadIdx.x & 31) & ~(ZKL_COOP_GROUP - 1);
#pragma unroll
  for (int i = 0; i < 4; i++) out.l[i] = __shfl_sync(gmask, st.l[i], src_lane);
  return out;
}
__device__ __forceinline__ fe coop_merge(const fe &a, const fe &b) { merge_src s{a, b}; return coop_sponge_run(s, 3); }
__device__ __forceinline__ fe coop_merge_with_int(const fe &seed, uint64_t v) { mwi_src s{seed, v}; return coop_sponge_run(s, 3); }
__device__ __forceinline__ fe coop_hash_elements(const fe *base, size_t stride, int len) {
  elems_src s{base, stride, len};
  return coop_sponge_run(s, 1 + (len + 1) / 2);
}

// group index / count helpers: a kernel launched with ZKL_COOP_THREADS threads handles 8 sponges per block.
// Out-of-range groups still run (all threads must reach __syncwarp / shfl) on clamped inputs, then skip the store.
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_coop_merkle_level(const fe *__restrict__ in, fe *__restrict__ out, size_t count) {
  size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  size_t gi = g < count ? g : count - 1;
  fe d = coop_merge(ld_fe(in + 2 * gi), ld_fe(in + 2 * gi + 1));
  if (g < count && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) st_fe(out + g, d);
}
// coset-major row hashing, one group per (row, chunk) — same contract as k_hash_rows
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_coop_hash_rows(const fe *__restrict__ lde, size_t N, size_t n, int blowup, int W, int psize,
                                                                     int nchunks, fe *__restrict__ out, size_t cstride) {
  size_t total = N * (size_t)nchunks;
  size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  size_t t = g < total ? g : total - 1;
  size_t s = t % N; int q = (int)(t / N);
  size_t r = (s / n) + (size_t)blowup * (s % n);
  int first = q * psize, len = min(psize, W - first);
  fe d = coop_hash_elements(lde + (size_t)first * cstride + s, cstride, len);
  if (g < total && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) st_fe(out + (size_t)q * N + r, d);
}
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_coop_merge_many_rows(const fe *__restrict__ parts, size_t stride, size_t rows, int nchunks, fe *__restrict__ out) {
  size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  size_t r = g < rows ? g : rows - 1;
  many_src s{parts + r, stride};
  fe d = coop_sponge_run(s, 1 + nchunks);
  if (g < rows && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) st_fe(out + r, d);
}
