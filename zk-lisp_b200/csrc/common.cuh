// common.cuh — context, error handling and launch helpers shared by the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/zkl_b200.h"
#include "f128.cuh"
#include "poseidon.cuh"

#define ZKL_CHECK(ctx, expr)                                                                         \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +      \
                       std::to_string(__LINE__) + ")");                                              \
      return 1;                                                                                      \
    }                                                                                                \
  } while (0)
#define ZKL_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r) return _r;         \
  } while (0)

// per-(N) device tables
struct domain_tables {
  size_t N = 0;
  fe *wpow = nullptr;  // wpow[m] = w_N^m, m in [0, N)
};

struct zkl_ctx {
  int device = 0;
  int num_sms = 148;
  bool tc_half_tiles = true;               // a lone prover's tensor-core sponge launches use 128-sponge tiles when they fit one wave
  int tc_grid_cap = 148;                   // persistent tensor-core sponge kernels launch at most this many CTAs
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;            // private stream-ordered memory pool
  std::string err;
  uint64_t launches = 0;
  bool suite_loaded = false;
  // sponge count from which the bulk kernels are used instead of the 16-thread cooperative ones
  long bulk_threshold = 0;   // 0 = automatic (bulk_min in zkl_b200.cu)
  // bulk hashing path: 0 = tensor-core MDS (tcperm.cuh, the default), 1 = integer-pipe thread-per-sponge kernels
  int hash_path = 0;
  int coop_path = 0;   // small sponge batches: 0 = automatic (by the number of contexts sharing the device), 1 = 16-thread groups
                       // (kernels_hash.cuh, least SM time), 2 = 64-thread groups (kernels_quad.cuh, lowest latency)
  int ntt_path = 0;    // 4096-row coset LDE: 0 = radix-16 stages on the tensor cores, coset shift folded into the matrix (kernels_tcntt.cuh v2),
                       // 1 = radix-2 in shared memory, 2 = tensor cores with a separate coset-shift product (v1)
  bool tc_ok = true;   // the loaded MDS constants satisfy the tensor-core epilogue's accumulator bound
  std::map<size_t, domain_tables> tables;  // by N
  std::map<size_t, fe *> opow;             // GENERATOR^j tables by n
  std::map<size_t, std::pair<void *, void *>> lde16;   // tensor-core LDE constants (byte matrices, twiddles) by (blowup, direction)
  std::map<size_t, fe *> periodic;         // periodic-column tables by (n << 8 | ce)
  std::vector<cudaEvent_t> ev;             // stage timing events
  int wait_mode = 0;                       // host waits: 0 = automatic, 1 = spin, 2 = block, 3 = poll + yield (ZKL_WAIT_MODE; api_stark.inl: ctx_wait)
  cudaEvent_t wait_ev = nullptr;           // cudaEventBlockingSync event of the blocking wait
  void *stage_h = nullptr; size_t stage_cap = 0, stage_used = 0;   // pinned host staging for per-call parameter uploads (api_stark.inl: upload_air)
  std::vector<uint32_t> sort_a, sort_b, sort_cnt;  // scratch of the assertion ordering
  size_t shard_min_leaves = (size_t)1 << 15;       // column-sharded prover: smaller trees / FRI layers are built on every rank (zkl_set_shard_min_leaves)
  void set_error(const std::string &e) { err = e; }
};

static inline int ceil_div(size_t a, size_t b) { return (int)((a + b - 1) / b); }
static inline int ilog2_sz(size_t n) { int k = 0; while (((size_t)1 << k) < n) k++; return k; }

// device scratch with RAII. Allocation is STREAM-ORDERED (cudaMallocAsync / cudaFreeAsync on the calling
// context's stream, pool release threshold = unlimited), so concurrent segment streams never hit the
// device-wide synchronisation of cudaMalloc/cudaFree.
static thread_local cudaStream_t tl_stream = nullptr;
static thread_local cudaMemPool_t tl_pool = nullptr;
static thread_local bool tl_stream_set = false;
template <typename T>
struct dev_buf {
  T *p = nullptr; size_t n = 0; cudaStream_t s = nullptr; bool async = false;
  dev_buf() {}
  dev_buf(const dev_buf &) = delete;
  dev_buf &operator=(const dev_buf &) = delete;
  dev_buf(dev_buf &&o) noexcept : p(o.p), n(o.n), s(o.s), async(o.async) { o.p = nullptr; }
  ~dev_buf() { release(); }
  void release() {
    if (!p) return;
    if (async) cudaFreeAsync(p, s); else cudaFree(p);
    p = nullptr;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (tl_stream_set) {
      // each context allocates from its OWN pool: blocks freed on one stream are never handed to another stream,
      // so the allocator never inserts cross-stream dependencies between concurrently proving segments
      async = true; s = tl_stream;
      return tl_pool ? cudaMallocFromPoolAsync((void **)&p, count * sizeof(T) + 16, tl_pool, s)
                     : cudaMallocAsync((void **)&p, count * sizeof(T) + 16, s);
    }
    async = false;
    return cudaMalloc((void **)&p, count * sizeof(T) + 16);
  }
};

// host-side u128 helpers for CONSTANT derivation only (domain labels, roots of unity seeds)
typedef unsigned __int128 hu128;
static const hu128 H_P = ((hu128)0xffffffffffffffffULL << 64) | 0xffffd30000000001ULL;
static inline fe h_to_fe(hu128 x) {
  fe r; r.l[0] = (uint32_t)x; r.l[1] = (uint32_t)(x >> 32); r.l[2] = (uint32_t)(x >> 64); r.l[3] = (uint32_t)(x >> 96); return r;
}
static inline fe fe_from_abi(const zkl_fe &x) {
  fe r; r.l[0] = (uint32_t)x.lo; r.l[1] = (uint32_t)(x.lo >> 32); r.l[2] = (uint32_t)x.hi; r.l[3] = (uint32_t)(x.hi >> 32); return r;
}
