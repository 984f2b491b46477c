// poseidon.cuh — zk-lisp's 12-lane, 27-round Poseidon-style sponge on the device.
// Reference: zk-lisp-proof-winterfell/src/poseidon/hasher.rs:144-231 (sponge), :173-190 (permutation:
// x^3 on all 12 lanes, then MDS*s + rc[r]), src/utils.rs:359-381 (byte folds, digest layout).
//
// One thread owns one sponge state (48 registers). The commitment/coin suite (suite_id = [0;32],
// hasher.rs:23) lives in __constant__ memory so every IMAD takes its MDS limb straight from the
// constant bank with a warp-uniform address. Each MDS row is accumulated UNREDUCED (12 products
// + the round constant in a 260-bit even/odd accumulator) and reduced once: 168 wide products
// and 36 reductions per round instead of 168 modmuls.
#pragma once
#include "f128.cuh"

#define ZKL_ROUNDS 27
#define ZKL_RATE 10

struct suite_consts {
  fe dom[2];
  fe mds[12][12];
  fe rc[ZKL_ROUNDS][12];
};
// domain-label elements fold_bytes32_to_fe(label padded to 32 bytes), hasher.rs:155-160
struct label_consts {
  fe elements;        // "winter/hash/elements"            hasher.rs:133
  fe merge;           // "zkl/winter/hash/merge"           hasher.rs:79
  fe merge_many;      // "zkl/winter/hash/merge_many"      hasher.rs:99
  fe merge_with_int;  // "zkl/winter/hash/merge_with_int"  hasher.rs:114
  fe bytes;           // "zkl/winter/hash/bytes"           hasher.rs:65
};

#if defined(__CUDACC__)
// single translation unit: the constants are defined here
__constant__ suite_consts c_suite;
__constant__ label_consts c_labels;

// Sponge state lives in shared memory, lane-major: lane l of thread t is st[l * ZKL_HASH_THREADS + t]
// (consecutive threads -> consecutive 16-byte slots: conflict-free 128-bit LDS/STS). Keeping the
// state out of registers lets the MDS loop run over output lanes without unrolling (dynamic lane
// index), which keeps the round body ~400 instructions (inside the L0/L1 instruction caches) and the
// register count near 100 instead of 255 with spills.
#ifndef ZKL_HASH_THREADS
#define ZKL_HASH_THREADS 64
#endif

__device__ __forceinline__ fe lds_fe(const fe *p) {
  uint4 v = *reinterpret_cast<const uint4 *>(p);
  fe r; r.l[0] = v.x; r.l[1] = v.y; r.l[2] = v.z; r.l[3] = v.w; return r;
}
__device__ __forceinline__ void sts_fe(fe *p, const fe &a) {
  *reinterpret_cast<uint4 *>(p) = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
}
__device__ __forceinline__ fe ldc_fe(const fe *p) {  // constant-bank load (warp-uniform address)
  uint4 v = *reinterpret_cast<const uint4 *>(p);
  fe r; r.l[0] = v.x; r.l[1] = v.y; r.l[2] = v.z; r.l[3] = v.w; return r;
}

// hasher.rs:173-190: x^3 on all lanes, then MDS * s + rc[r], 27 rounds. `st` points at this
// thread's lane-0 slot; lanes are ZKL_HASH_THREADS elements apart.
__device__ __noinline__ void permute(fe *st) {
#pragma unroll 1
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    fe c[12];
#pragma unroll
    for (int i = 0; i < 12; i++) c[i] = fe_cube(lds_fe(st + i * ZKL_HASH_THREADS));
#pragma unroll 1
    for (int i = 0; i < 12; i += 2) {  // two output lanes per iteration: two independent carry chains in flight
      wide_acc w0, w1; wacc_zero(w0); wacc_zero(w1);
#pragma unroll
      for (int k = 0; k < 12; k++) {
        wacc_mad(w0, c[k], ldc_fe(&c_suite.mds[i][k]));
        wacc_mad(w1, c[k], ldc_fe(&c_suite.mds[i + 1][k]));
      }
      wacc_add_fe(w0, ldc_fe(&c_suite.rc[r][i]));
      wacc_add_fe(w1, ldc_fe(&c_suite.rc[r][i + 1]));
      sts_fe(st + i * ZKL_HASH_THREADS, wacc_reduce(w0));
      sts_fe(st + (i + 1) * ZKL_HASH_THREADS, wacc_reduce(w1));
    }
  }
}
#endif
