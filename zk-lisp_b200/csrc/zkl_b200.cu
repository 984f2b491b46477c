// zkl_b200.cu — C ABI (include/zkl_b200.h) over the sm_100a kernels. Single translation unit so the
// __constant__ hasher suite is shared by every kernel without relocatable device code.
// No CPU fallback: every entry point launches device work; a missing device fails at ctx creation.
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>

#include "kernels_air.cuh"
#include "kernels_hash.cuh"
#include "kernels_ntt.cuh"
#include "kernels_quad.cuh"
#include "kernels_stark.cuh"
#include "kernels_tc.cuh"
#include "kernels_tcntt.cuh"

static thread_local std::string g_create_error;   // zkl_last_error(NULL) reports the calling thread's failed zkl_ctx_create
static std::mutex g_suite_mutex;
struct device_suite { suite_consts consts; bool loaded = false; int users = 0; };
static device_suite g_device_suite[64];

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                 \
  do {                                                                              \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
    (ctx)->launches++;                                                              \
    ZKL_CHECK(ctx, cudaGetLastError());                                             \
  } while (0)
// tensor-core sponge kernels: persistent, one 576-thread CTA per SM, `sponges` = number of sponges in the launch. Tiles of 256
// sponges; a lone prover (use_quad) takes half tiles of 128 when the launch then still fits one wave — the pass is shorter
// (kernels_tc.cuh: tc_index). The mode is a template parameter: the full-tile kernels carry no trace of it.
#define LAUNCH_TC(ctx, kernel, sponges, ...)                                                                                        \
  do {                                                                                                                              \
    const size_t cap_ = (size_t)(ctx)->tc_grid_cap;                                                                                 \
    if (tc_half_tiles(ctx, (size_t)(sponges)))                                                                                      \
      LAUNCH(ctx, kernel<true>, std::min<size_t>(((size_t)(sponges) + TC_HALF_SPONGES - 1) / TC_HALF_SPONGES, cap_), TC_THREADS, TC_SMEM_BYTES, __VA_ARGS__); \
    else                                                                                                                            \
      LAUNCH(ctx, kernel<false>, std::min<size_t>(((size_t)(sponges) + TC_CTA_SPONGES - 1) / TC_CTA_SPONGES, cap_), TC_THREADS, TC_SMEM_BYTES, __VA_ARGS__); \
  } while (0)
// the grinding kernel walks 256-nonce tiles in increasing order (early exit per tile)
#define LAUNCH_TC_TILES(ctx, kernel, sponges, ...)                                                                  \
  LAUNCH(ctx, kernel, std::min<size_t>(((size_t)(sponges) + TC_CTA_SPONGES - 1) / TC_CTA_SPONGES, (size_t)(ctx)->tc_grid_cap), TC_THREADS, \
         TC_SMEM_BYTES, __VA_ARGS__)

// low-latency sponge kernels: one 64-thread group per sponge (kernels_quad.cuh) or the older 16-thread groups (coop_path 1)
#define LAUNCH_QUAD(ctx, kernel, sponges, ...) LAUNCH(ctx, kernel, ceil_div((size_t)(sponges) * ZKL_QUAD_GROUP, ZKL_QUAD_THREADS), ZKL_QUAD_THREADS, 0, __VA_ARGS__)
#define LAUNCH_COOP(ctx, kernel, sponges, ...) LAUNCH(ctx, kernel, ceil_div((size_t)(sponges) * ZKL_COOP_GROUP, ZKL_COOP_THREADS), ZKL_COOP_THREADS, 0, __VA_ARGS__)

// Which low-latency sponge a context uses (coop_path 0 = decide by how many contexts share the device: a lone prover is
// latency-bound and takes the 64-thread groups, a pool of provers is SM-time-bound and takes the 16-thread groups).
static std::atomic<int> g_ctx_per_device[64];
static std::atomic<int> g_ctx_total{0};
static int g_host_cores = 1;   // cores this process may run on (sched_getaffinity), read at context creation
static inline bool use_quad(const zkl_ctx *ctx) {
  if (ctx->coop_path == 1) return false;
  if (ctx->coop_path == 2) return true;
  return g_ctx_per_device[ctx->device & 63].load(std::memory_order_relaxed) < 4;
}

// sponge count from which the bulk (tensor-core) kernels take over from the small-batch kernels: one dependent permutation
// step costs ~117 us on the tensor-core kernel at any size, 29 / 43 / 78 us on the 64-thread groups at <= 512 / 1024 / 2048
// sponges and 35 / 37 / 45 us on the 16-thread groups (profiles/perm_latency_r01.log)
static inline int tc_half_tiles(const zkl_ctx *ctx, size_t sponges) {
  return ctx->tc_half_tiles && use_quad(ctx) && sponges <= (size_t)TC_HALF_SPONGES * (size_t)ctx->tc_grid_cap;
}
static inline size_t bulk_min(const zkl_ctx *ctx) {
  if (ctx->bulk_threshold > 0) return (size_t)ctx->bulk_threshold;   // set explicitly (zkl_set_bulk_threshold)
  return use_quad(ctx) ? 4096 : 1024;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static fe host_root_of_unity(int logn);
static fe host_label(const char *s) {  // hasher.rs:155-160: label zero-padded to 32 bytes, folded
  uint8_t b[32]; memset(b, 0, 32);
  size_t l = strlen(s); if (l > 32) l = 32;
  memcpy(b, s, l);
  fe lo, hi; memcpy(lo.l, b, 16); memcpy(hi.l, b + 16, 16);
  return fe_fold_pair(fe_from_u128_mod(lo), fe_from_u128_mod(hi));
}

extern "C" int zkl_ctx_create(int device, zkl_ctx **out) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return 1;
  }
  if (device < 0 || device >= count) { g_create_error = "device index out of range"; return 1; }
  // poseidon/hasher.rs:235-241: the reference reads the commitment hasher's round count from ZKL_POSEIDON_HASHER_ROUNDS
  // (positive integers only, anything else means 27). The kernels are built for 27 rounds: refuse to run under another value
  // rather than emit commitments the reference process would not reproduce.
  if (const char *hr = getenv("ZKL_POSEIDON_HASHER_ROUNDS")) {
    char *end = nullptr; unsigned long long v = strtoull(hr, &end, 10);
    const bool parses = *hr && end && *end == 0 && hr[0] != '-' && hr[0] != '+' && hr[0] != ' ';
    if (parses && v > 0 && v != ZKL_ROUNDS) {
      g_create_error = std::string("ZKL_POSEIDON_HASHER_ROUNDS=") + hr + " is set: this library implements the 27-round commitment hasher only";
      return 1;
    }
  }
  zkl_ctx *c = new zkl_ctx();
  c->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete c; return 1; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete c; return 1; }
  c->num_sms = prop.multiProcessorCount;
  c->tc_grid_cap = c->num_sms;
  if (const char *e = getenv("ZKL_TC_GRID_CAP")) { int v = atoi(e); if (v > 0 && v <= c->num_sms) c->tc_grid_cap = v; }   // experiment knob
  if (const char *e = getenv("ZKL_COOP_PATH")) { int v = atoi(e); if (v >= 0 && v <= 2) c->coop_path = v; }   // experiment knobs (A/B runs of bench.py)
  if (const char *e = getenv("ZKL_TC_HALF_TILES")) c->tc_half_tiles = atoi(e) != 0;
  if (const char *e = getenv("ZKL_WAIT_MODE")) { int v = atoi(e); if (v >= 0 && v <= 3) c->wait_mode = v; }
  if (const char *e = getenv("ZKL_SHARD_MIN_LEAVES")) { long long v = atoll(e); if (v >= 0) c->shard_min_leaves = (size_t)v; }
  if (const char *e = getenv("ZKL_NTT_PATH")) { int v = atoi(e); if (v >= 0 && v <= 2) c->ntt_path = v; }
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e); delete c; return 1;
  }
  {  // private pool per context, freed blocks are kept (no trimming at synchronisation points)
    cudaMemPoolProps props; memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&c->pool, &props) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else {
      c->pool = nullptr; (void)cudaGetLastError();
    }
  }
  // opt in to large dynamic shared memory for the single-pass column transforms
  cudaFuncSetAttribute(k_interpolate, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(k_lde, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(k_tc_lde4096, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT16_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_lde4096_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT16_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_cross16, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT16_SMEM_BYTES);
  cudaFuncSetAttribute(k_ntt_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(k_tc_permute_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_permute_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_hash_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_hash_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merge_many_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merge_many_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merkle_level<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merkle_level<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_hash_elements_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_hash_elements_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merge_with_int_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_merge_with_int_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_fri_leaves<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_fri_leaves<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_grind, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_coin_draw<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaFuncSetAttribute(k_tc_coin_draw<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  label_consts lc;
  lc.elements = host_label("winter/hash/elements");
  lc.merge = host_label("zkl/winter/hash/merge");
  lc.merge_many = host_label("zkl/winter/hash/merge_many");
  lc.merge_with_int = host_label("zkl/winter/hash/merge_with_int");
  lc.bytes = host_label("zkl/winter/hash/bytes");
  {
    std::lock_guard<std::mutex> g(g_suite_mutex);
    if ((e = cudaMemcpyToSymbol(c_labels, &lc, sizeof lc)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete c; return 1; }
    // byte image of the DFT-16 matrix for the tensor-core LDE (kernels_tcntt.cuh); w16 = primitive 16th root of unity
    k_build_dft16_bytes<<<32, 256, 0, c->stream>>>(host_root_of_unity(4), fe_inv(host_root_of_unity(4)));
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete c; return 1; }
  }
  g_ctx_per_device[device & 63]++;
  g_ctx_total++;
  { cpu_set_t set; CPU_ZERO(&set); g_host_cores = sched_getaffinity(0, sizeof set, &set) == 0 ? std::max(CPU_COUNT(&set), 1) : 1; }
  *out = c;
  return 0;
}

extern "C" void zkl_ctx_destroy(zkl_ctx *ctx) {
  if (!ctx) return;
  g_ctx_per_device[ctx->device & 63]--;
  g_ctx_total--;
  if (ctx->suite_loaded) { std::lock_guard<std::mutex> g(g_suite_mutex); g_device_suite[ctx->device & 63].users--; }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
  if (ctx->wait_ev) cudaEventDestroy(ctx->wait_ev);
  if (tl_stream == ctx->stream) { tl_stream = nullptr; tl_pool = nullptr; tl_stream_set = false; }
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->tables) { cudaFree(kv.second.wpow); }
  for (auto &kv : ctx->opow) cudaFree(kv.second);
  for (auto &kv : ctx->periodic) cudaFree(kv.second);
  for (auto &kv : ctx->lde16) { cudaFree(kv.second.first); if (kv.second.second) cudaFree(kv.second.second); }
  for (auto e : ctx->ev) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
}
extern "C" const char *zkl_last_error(const zkl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" uint64_t zkl_launch_count(const zkl_ctx *ctx) { return ctx->launches; }

extern "C" int zkl_suite_load(zkl_ctx *ctx, const zkl_suite *s) {
  ZKL_CHECK(ctx, cudaSetDevice(ctx->device));
  suite_consts h;
  for (int i = 0; i < 2; i++) h.dom[i] = fe_from_abi(s->dom[i]);
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) h.mds[i][j] = fe_from_abi(s->mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 12; j++) h.rc[r][j] = fe_from_abi(s->rc[r][j]);
  std::lock_guard<std::mutex> g(g_suite_mutex);
  // tc_epilogue packs accumulator pairs into 32 bits: verify the bound for these constants, else use the integer pipe
  ctx->tc_ok = tc_max_accumulator(h) <= TC_ACC_MAX;
  if (!ctx->tc_ok) ctx->hash_path = 1;
  // The suite lives in device-global symbols shared by every context of a device (the commitment hasher has ONE suite,
  // suite_id = [0;32], hasher.rs:23). The same constants again: nothing is rewritten under kernels other contexts may have
  // in flight. Different constants while other contexts hold the first suite: refused.
  device_suite &ds = g_device_suite[ctx->device & 63];
  if (ds.loaded && memcmp(&ds.consts, &h, sizeof h) == 0) { ds.users += ctx->suite_loaded ? 0 : 1; ctx->suite_loaded = true; return 0; }
  if (ds.loaded && ds.users > (ctx->suite_loaded ? 1 : 0)) {
    ctx->set_error("another context on this device holds a different hasher suite: the suite is per device (include/zkl_b200.h)");
    return 1;
  }
  ZKL_CHECK(ctx, cudaDeviceSynchronize());   // no kernel of this (only) user may still read the old constants
  ZKL_CHECK(ctx, cudaMemcpyToSymbol(c_suite, &h, sizeof h));
  ZKL_CHECK(ctx, cudaMemcpyToSymbol(g_suite, &h, sizeof h));
  {  // byte image of the MDS matrix for the tensor-core permutation (tcperm.cuh)
    suite_consts *gs; ZKL_CHECK(ctx, cudaGetSymbolAddress((void **)&gs, g_suite));
    LAUNCH(ctx, k_build_mds_bytes, ceil_div(TC_KCH * TC_N, 128), 128, 0, gs);
    ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ds.consts = h; ds.loaded = true; ds.users = 1;
  ctx->suite_loaded = true;
  return 0;
}
// every entry point binds the calling thread to the context's device and stream (stream-ordered allocator)
static int enter(zkl_ctx *ctx) {
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->set_error("cudaSetDevice failed"); return 1; }
  tl_stream = ctx->stream; tl_pool = ctx->pool; tl_stream_set = true;
  return 0;
}
static int need_suite(zkl_ctx *ctx) {
  if (!ctx->suite_loaded) { ctx->set_error("zkl_suite_load has not been called on this context"); return 1; }
  return enter(ctx);
}

// w_N^m table and offset powers -------------------------------------------------------------------
static fe host_root_of_unity(int logn) {  // G^(2^(40-logn)), G = 3^((p-1)/2^40)
  // (p-1)/2^40 = 2^88 - 45 ; 3^(2^88-45) computed as 3^(2^88) * inv(3^45)
  fe g = fe_from_u64(3), a = g;
  for (int i = 0; i < 88; i++) a = fe_sqr(a);
  fe G = fe_mul(a, fe_inv(fe_pow(g, 45)));
  for (int i = logn; i < 40; i++) G = fe_sqr(G);
  return G;
}
static int get_tables(zkl_ctx *ctx, size_t N, domain_tables **out) {
  auto it = ctx->tables.find(N);
  if (it == ctx->tables.end()) {
    domain_tables t; t.N = N;
    ZKL_CHECK(ctx, cudaMalloc((void **)&t.wpow, N * sizeof(fe)));
    fe w = host_root_of_unity(ilog2_sz(N));
    LAUNCH(ctx, k_fill_powers, ceil_div(N, 256), 256, 0, w, t.wpow, N);
    it = ctx->tables.emplace(N, t).first;
  }
  *out = &it->second;
  return 0;
}
// opow[j] = offset^j, j < n, for offset = GENERATOR (the only offset Winterfell uses)
static int get_offset_powers(zkl_ctx *ctx, size_t n, fe **out) {
  auto it = ctx->opow.find(n);
  if (it == ctx->opow.end()) {
    fe *p;
    ZKL_CHECK(ctx, cudaMalloc((void **)&p, n * sizeof(fe)));
    LAUNCH(ctx, k_fill_powers, ceil_div(n, 256), 256, 0, fe_from_u64(3), p, n);
    it = ctx->opow.emplace(n, p).first;
  }
  *out = it->second;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// device-side building blocks (all stream-ordered, no host sync)
// ------------------------------------------------------------------------------------------------
static const int NTT_THREADS = 512;

// split of a large transform: n = n1 * n2, T columns per CTA (T * max(n1, n2) elements = 128 KB of shared memory)
struct ntt_split { int n1, n2, T; };
static ntt_split split_for(size_t n) {
  int logn = ilog2_sz(n);
  ntt_split s; s.n2 = 1 << ((logn + 1) / 2); s.n1 = (int)(n / s.n2);
  s.T = std::min(8192 / std::max(s.n1, s.n2), std::min(s.n1, s.n2));
  return s;
}
static const size_t NTT_SINGLE_PASS_MAX = 8192, NTT_MAX = (size_t)1 << 24;

// constants of the v2 tensor-core transform (kernels_tcntt.cuh): per-coset byte matrices and stage-1 twiddles.
// `scale` (optional) = [blowup][scale_n] coset-shift powers; `variant` tells cache entries with different factors apart.
static int get_lde16_consts(zkl_ctx *ctx, size_t blowup, bool inverse, const fe *scale, size_t scale_n, const domain_tables *t, fe factor, int variant,
                            const uint4 **bmat, const fe **tw0) {
  const size_t key = (t->N << 10) | (blowup << 3) | ((size_t)variant << 1) | (inverse ? 1 : 0);
  auto it = ctx->lde16.find(key);
  if (it == ctx->lde16.end()) {
    void *b, *w;
    ZKL_CHECK(ctx, cudaMalloc(&b, blowup * 4096 * sizeof(uint4)));
    ZKL_CHECK(ctx, cudaMalloc(&w, blowup * 4096 * sizeof(fe)));
    LAUNCH(ctx, k_build_lde16_consts, ceil_div(blowup * 4096, 256), 256, 0, scale, scale_n, t->wpow, t->N, (int)blowup, inverse, factor, (uint4 *)b, (fe *)w);
    it = ctx->lde16.emplace(key, std::make_pair(b, w)).first;
  }
  *bmat = (const uint4 *)it->second.first; *tw0 = (const fe *)it->second.second;
  return 0;
}
// per-coset matrices of the cross stage of a 65536-point transform (k_tc_cross16)
static int get_cross16_consts(zkl_ctx *ctx, size_t blowup, bool inverse, const fe *scale, size_t kstride, size_t m, int R, const domain_tables *t, fe factor,
                              bool top, const uint4 **cmat) {
  // one entry per (domain, blowup, direction, block size m, radix, top level or not: the top level of an inverse transform carries 1/n)
  const size_t key = ((size_t)1 << 63) | (t->N << 24) | ((size_t)ilog2_sz(m) << 16) | ((size_t)R << 10) | (blowup << 2) | (top ? 2 : 0) | (inverse ? 1 : 0);
  auto it = ctx->lde16.find(key);
  if (it == ctx->lde16.end()) {
    void *b;
    ZKL_CHECK(ctx, cudaMalloc(&b, blowup * 4096 * sizeof(uint4)));
    LAUNCH(ctx, k_build_cross16_consts, ceil_div(blowup * 4096, 256), 256, 0, scale, kstride, t->wpow, t->N, m, R, (int)blowup, inverse, factor, (uint4 *)b);
    it = ctx->lde16.emplace(key, std::make_pair(b, (void *)nullptr)).first;
  }
  *cmat = (const uint4 *)it->second.first;
  return 0;
}
static unsigned lde16_grid(const zkl_ctx *ctx, size_t ncols, size_t blowup) {
  return (unsigned)std::min<size_t>(ncols * blowup, std::max<size_t>((size_t)ctx->num_sms, blowup));
}

// n = R_1 ... R_L * 4096 with R_i in {2, 4, 8, 16}: L cross stages (k_tc_cross16) and the 4096-point transforms, all on the tensor
// cores and L + 1 passes over HBM. 16 384 rows and up (8192 still fit one CTA's shared memory and stay on the single-pass kernel).
// The odd bits go to the TOP radix, the levels below are radix 16.
struct tc_plan { int L = 0; int R[4] = {0, 0, 0, 0}; uint32_t radices = 0; };
static bool tc_plan_for(size_t n, tc_plan &p) {
  if (n < 4 * (size_t)NTT16_N || (n & (n - 1))) return false;
  int e = ilog2_sz(n / NTT16_N);
  p = tc_plan();
  if (e % 4) { p.R[p.L++] = 1 << (e % 4); e -= e % 4; }
  for (; e > 0; e -= 4) { if (p.L == 4) return false; p.R[p.L++] = 16; }
  for (int l = 0; l < p.L; l++) p.radices |= (uint32_t)ilog2_sz((size_t)p.R[l]) << (4 * l);
  return true;
}
static int cross_radix(size_t n) { tc_plan p; return tc_plan_for(n, p) ? p.R[0] : 0; }   // non-zero: the tensor-core multi-pass transform applies

// src: ncols columns of n elements; out: [column][coset][n] (blowup = 1 for the inverse transform). `scale` = [blowup][n] coset-shift
// powers or null; `factor` = 1/n for the inverse transform (folded into the top level's matrices).
// A coset range k0 .. k0 + nk - 1 (nk = 0: all) comes out compact, [column][nk][n]: the kernels index their per-coset constants by the
// coset they serve, so a range is the same launch with the constants offset to k0 and nk in place of the blowup.
static int dev_tc_multilevel(zkl_ctx *ctx, const fe *src, fe *out, size_t ncols, size_t n, size_t blowup, bool inverse, const fe *scale, const domain_tables *t,
                             fe factor, size_t k0 = 0, size_t nk = 0) {
  if (!nk) nk = blowup;
  tc_plan pl; if (!tc_plan_for(n, pl)) { ctx->set_error("no tensor-core plan for this transform size"); return 1; }
  const uint4 *cmat[4], *bmat; const fe *tw0;
  size_t m = n;
  for (int l = 0; l < pl.L; l++) {
    m /= (size_t)pl.R[l];
    ZKL_TRY(get_cross16_consts(ctx, blowup, inverse, scale, n, m, pl.R[l], t, l == 0 ? factor : fe_one(), l == 0, &cmat[l]));
  }
  ZKL_TRY(get_lde16_consts(ctx, blowup, inverse, scale, n, t, inverse ? fe_one() : fe_zero(), inverse ? 1 : 0, &bmat, &tw0));
  for (int l = 0; l < pl.L; l++) cmat[l] += k0 * 4096;
  bmat += k0 * 4096; tw0 += k0 * 4096;
  blowup = nk;   // from here on: the cosets of the call
  const size_t chunk = std::min(ncols, std::max<size_t>(1, ((size_t)1 << 31) / (blowup * n * sizeof(fe))));   // columns per round: <= 2 GB per scratch buffer
  dev_buf<fe> sa, sb; ZKL_CHECK(ctx, sa.alloc(chunk * blowup * n));
  if (pl.L > 1) ZKL_CHECK(ctx, sb.alloc(chunk * blowup * n));
  for (size_t c0 = 0; c0 < ncols; c0 += chunk) {
    const size_t nc = std::min(chunk, ncols - c0);
    const fe *in = src + c0 * n; fe *to = sa.p;
    size_t cols = nc, len = n;   // level l: `cols` columns of `len` elements
    for (int l = 0; l < pl.L; l++) {
      const size_t sub = len / (size_t)pl.R[l];
      LAUNCH(ctx, k_tc_cross16, lde16_grid(ctx, cols * 16, blowup), NTT16_THREADS, NTT16_SMEM_BYTES, in, (int)cols, (int)blowup, sub, pl.R[l], cmat[l], t->wpow, t->N,
             t->N / len, to, inverse, l > 0, l + 1 < pl.L);
      in = to; to = (to == sa.p) ? sb.p : sa.p;
      if (l + 1 < pl.L) { cols *= (size_t)pl.R[l]; len = sub; }
    }
    // the last pass scatters to y[o + P i1]; from P = 32 blocks per column on it keeps blocks together and a tiled transpose finishes
    int prod_log = 0; for (int l = 0; l < pl.L; l++) prod_log += ilog2_sz((size_t)pl.R[l]);
    const bool unscatter = pl.L > 1 && prod_log >= 5;
    LAUNCH(ctx, k_tc_lde4096_v2, lde16_grid(ctx, cols * (size_t)pl.R[pl.L - 1], blowup), NTT16_THREADS, NTT16_SMEM_BYTES, in, (int)(cols * (size_t)pl.R[pl.L - 1]), (int)blowup, bmat,
           tw0, t->wpow, unscatter ? to : out + c0 * blowup * n, inverse, pl.R[pl.L - 1], t->N, pl.radices, unscatter);
    if (unscatter)
      LAUNCH(ctx, k_tc_unscatter, dim3(NTT16_N / 32, (unsigned)(((size_t)1 << prod_log) / 32), (unsigned)(nc * blowup)), 256, 0, to, out + c0 * blowup * n, prod_log, pl.radices);
  }
  return 0;
}

// cols (ncols x n, evaluations over <w_n>) -> coefficients, in place or out of place
static int dev_interpolate(zkl_ctx *ctx, const fe *src, fe *dst, size_t ncols, size_t n) {
  if (n > NTT_MAX || (n & (n - 1)) || n < 2) { ctx->set_error("trace length must be a power of two in [2, 2^24]"); return 1; }
  domain_tables *t; ZKL_TRY(get_tables(ctx, n, &t));
  fe n_inv = fe_inv(fe_from_u64(n));
  if (n == NTT16_N && ctx->ntt_path == 0) {   // inverse radix-16 stages on the tensor cores, 1/n folded into the matrix (kernels_tcntt.cuh v2)
    const uint4 *bmat; const fe *tw0;
    ZKL_TRY(get_lde16_consts(ctx, 1, true, nullptr, 0, t, n_inv, 0, &bmat, &tw0));
    LAUNCH(ctx, k_tc_lde4096_v2, lde16_grid(ctx, ncols, 1), NTT16_THREADS, NTT16_SMEM_BYTES, src, (int)ncols, 1, bmat, tw0, t->wpow, dst, true, 0, t->N, 0u, false);
    return 0;
  }
  if (ctx->ntt_path == 0 && cross_radix(n))   // 16 384 rows and up: cross stages + 4096-point transforms, all on the tensor cores
    return dev_tc_multilevel(ctx, src, dst, ncols, n, 1, true, nullptr, t, n_inv);
  if (n == NTT16_N && ctx->ntt_path == 2) {   // v1: separate 1/n product
    LAUNCH(ctx, k_tc_lde4096, (unsigned)std::min<size_t>(ncols, (size_t)ctx->num_sms), NTT16_THREADS, NTT16_SMEM_BYTES, src, (int)ncols, 1, (const fe *)nullptr,
           t->wpow, dst, true, n_inv);
    return 0;
  }
  if (n <= NTT_SINGLE_PASS_MAX) {
    LAUNCH(ctx, k_interpolate, (unsigned)ncols, NTT_THREADS, n * sizeof(fe), src, dst, (int)n, ilog2_sz(n), t->wpow, n, n_inv);
    return 0;
  }
  // two passes over HBM (kernels_ntt.cuh); pass 1 goes through a scratch buffer when the call is in place
  ntt_split sp = split_for(n);
  dev_buf<fe> tmp; fe *mid = dst;
  if (src == dst) { ZKL_CHECK(ctx, tmp.alloc(ncols * n)); mid = tmp.p; }
  size_t smem = (size_t)sp.T * std::max(sp.n1, sp.n2) * sizeof(fe);
  LAUNCH(ctx, k_ntt_pass1, dim3(sp.n1 / sp.T, (unsigned)ncols, 1), NTT_THREADS, smem, src, mid, sp.n1, sp.n2, ilog2_sz(sp.n2), sp.T, t->wpow, n, (const fe *)nullptr, true);
  LAUNCH(ctx, k_ntt_pass2, dim3(sp.n2 / sp.T, (unsigned)ncols), NTT_THREADS, smem, mid, sp.n1, sp.n2, ilog2_sz(sp.n1), sp.T, t->wpow, n, true, true, n_inv);
  if (mid != dst) ZKL_CHECK(ctx, cudaMemcpyAsync(dst, mid, ncols * n * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
// polys (ncols x n) -> coset-major LDE (ncols x blowup x n) over GENERATOR * <w_N>
static int dev_lde(zkl_ctx *ctx, const fe *polys, fe *out, size_t ncols, size_t n, size_t blowup, size_t k0 = 0, size_t nk = 0) {
  if (!nk) nk = blowup;
  if (n > NTT_MAX || (n & (n - 1)) || n < 2) { ctx->set_error("trace length must be a power of two in [2, 2^24]"); return 1; }
  size_t N = n * blowup;
  domain_tables *t; ZKL_TRY(get_tables(ctx, N, &t));
  // scale[k][j] = (GENERATOR * w_N^k)^j, cached per (n, blowup)
  const size_t key = n | (blowup << 32) | ((size_t)1 << 61);
  auto it = ctx->opow.find(key);
  if (it == ctx->opow.end()) {
    fe *opow; ZKL_TRY(get_offset_powers(ctx, n, &opow));
    fe *p; ZKL_CHECK(ctx, cudaMalloc((void **)&p, N * sizeof(fe)));
    LAUNCH(ctx, k_lde_scale_table, ceil_div(N, 256), 256, 0, opow, t->wpow, N, (int)n, (int)blowup, p);
    it = ctx->opow.emplace(key, p).first;
  }
  if (nk != blowup && !(ctx->ntt_path == 0 && cross_radix(n))) {   // a coset range (dev_lde_cosets) on the radix-2 two-pass path, compact output
    ntt_split sp = split_for(n);
    size_t smem = (size_t)sp.T * std::max(sp.n1, sp.n2) * sizeof(fe);
    LAUNCH(ctx, k_ntt_pass1, dim3(sp.n1 / sp.T, (unsigned)ncols, (unsigned)nk), NTT_THREADS, smem, polys, out, sp.n1, sp.n2, ilog2_sz(sp.n2), sp.T, t->wpow, N,
           it->second + k0 * n, false);
    LAUNCH(ctx, k_ntt_pass2, dim3(sp.n2 / sp.T, (unsigned)(ncols * nk)), NTT_THREADS, smem, out, sp.n1, sp.n2, ilog2_sz(sp.n1), sp.T, t->wpow, N, false, false, fe_zero());
    return 0;
  }
  if (n == NTT16_N && ctx->ntt_path == 0) {   // radix-16 stages on the tensor cores, coset shift folded into the matrix (kernels_tcntt.cuh v2)
    const uint4 *bmat; const fe *tw0;
    ZKL_TRY(get_lde16_consts(ctx, blowup, false, it->second, n, t, fe_zero(), 0, &bmat, &tw0));
    LAUNCH(ctx, k_tc_lde4096_v2, lde16_grid(ctx, ncols, blowup), NTT16_THREADS, NTT16_SMEM_BYTES, polys, (int)ncols, (int)blowup, bmat, tw0, t->wpow, out, false, 0, t->N, 0u, false);
    return 0;
  }
  if (ctx->ntt_path == 0 && cross_radix(n))   // 16 384 rows and up: cross stages + coset LDE of the 4096-row blocks, all on the tensor cores
    return dev_tc_multilevel(ctx, polys, out, ncols, n, blowup, false, it->second, t, fe_zero(), k0, nk);
  if (n == NTT16_N && ctx->ntt_path == 2) {   // v1: separate coset-shift product
    LAUNCH(ctx, k_tc_lde4096, (unsigned)std::min<size_t>(ncols * blowup, (size_t)ctx->num_sms), NTT16_THREADS, NTT16_SMEM_BYTES, polys, (int)ncols, (int)blowup,
           it->second, t->wpow, out, false, fe_zero());
    return 0;
  }
  if (n <= NTT_SINGLE_PASS_MAX) {
    dim3 grid((unsigned)blowup, (unsigned)ncols);
    LAUNCH(ctx, k_lde, grid, NTT_THREADS, n * sizeof(fe), polys, (int)n, ilog2_sz(n), (int)blowup, t->wpow, N, it->second, out);
    return 0;
  }
  ntt_split sp = split_for(n);
  size_t smem = (size_t)sp.T * std::max(sp.n1, sp.n2) * sizeof(fe);
  LAUNCH(ctx, k_ntt_pass1, dim3(sp.n1 / sp.T, (unsigned)ncols, (unsigned)blowup), NTT_THREADS, smem, polys, out, sp.n1, sp.n2, ilog2_sz(sp.n2), sp.T, t->wpow, N, it->second, false);
  LAUNCH(ctx, k_ntt_pass2, dim3(sp.n2 / sp.T, (unsigned)(ncols * blowup)), NTT_THREADS, smem, out, sp.n1, sp.n2, ilog2_sz(sp.n1), sp.T, t->wpow, N, false, false, fe_zero());
  return 0;
}
// the cosets k0 .. k0 + nk - 1 of the same LDE, compact: out[c][j][i] = LDE[c][k0 + j][i] (column-sharded prover: a rank extends what
// every rank holds as coefficients only over the cosets it owns). Natively for the two-pass sizes — where it matters —, through the
// whole LDE and a strided copy otherwise.
static int dev_lde_cosets(zkl_ctx *ctx, const fe *polys, fe *out, size_t ncols, size_t n, size_t blowup, size_t k0, size_t nk) {
  if (k0 + nk > blowup || !nk) { ctx->set_error("coset range outside the LDE domain"); return 1; }
  if (nk == blowup) return dev_lde(ctx, polys, out, ncols, n, blowup);
  const bool native = n > NTT_SINGLE_PASS_MAX;   // tensor-core multi-pass or radix-2 two-pass transforms: both take a coset range
  if (native) return dev_lde(ctx, polys, out, ncols, n, blowup, k0, nk);
  dev_buf<fe> full; ZKL_CHECK(ctx, full.alloc(ncols * n * blowup));
  ZKL_TRY(dev_lde(ctx, polys, full.p, ncols, n, blowup));
  ZKL_CHECK(ctx, cudaMemcpy2DAsync(out, nk * n * sizeof(fe), full.p + k0 * n, blowup * n * sizeof(fe), nk * n * sizeof(fe), ncols, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
// leaves[r] for the coset-major matrix `lde` (ncols x N); partition size per PartitionOptions
static size_t partition_size_of(const zkl_options *o, size_t cols, size_t ext = 1) {  // [WF-mem] PartitionOptions::partition_size::<E>
  if (o->num_partitions == 1 && o->hash_rate == 0) return cols;
  size_t ps = (cols + o->num_partitions - 1) / o->num_partitions;
  return std::max<size_t>(ps, o->hash_rate / ext);
}
// partition digests of a coset-major matrix: out[q * N + r] = hash_elements(row r, columns [q*ps, min(ncols, (q+1)*ps)))
static int dev_hash_partitions(zkl_ctx *ctx, const fe *lde, size_t ncols, size_t n, size_t blowup, size_t ps, fe *out, size_t cstride = 0) {
  if (!cstride) cstride = n * blowup;
  size_t N = n * blowup, nchunks = (ncols + ps - 1) / ps;
  // bulk (tensor-core or thread-per-sponge) kernel when the sponges fill SMs, the 16-thread cooperative one otherwise
  size_t sponges = N * nchunks;
  bool bulk = sponges >= bulk_min(ctx);
  if (bulk && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_hash_rows, sponges, lde, N, n, (int)blowup, (int)ncols, (int)ps, (int)nchunks, out, cstride);
  else if (bulk) LAUNCH(ctx, k_hash_rows, ceil_div(sponges, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, lde, N, n, (int)blowup, (int)ncols, (int)ps, (int)nchunks, out, cstride);
  else if (use_quad(ctx)) LAUNCH_QUAD(ctx, k_quad_hash_rows, sponges, lde, N, n, (int)blowup, (int)ncols, (int)ps, (int)nchunks, out, cstride);
  else LAUNCH_COOP(ctx, k_coop_hash_rows, sponges, lde, N, n, (int)blowup, (int)ncols, (int)ps, (int)nchunks, out, cstride);
  return 0;
}
// leaves[r] = merge_many(parts[0][r0 + r], ..., parts[nchunks-1][r0 + r]), r < rows (agg/child.rs:1040-1044); parts[q] at parts + q*stride
static int dev_merge_partitions(zkl_ctx *ctx, const fe *parts, size_t stride, size_t rows, size_t nchunks, fe *leaves) {
  if (rows >= bulk_min(ctx) && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_merge_many_rows, rows, parts, stride, rows, (int)nchunks, leaves);
  else if (rows >= bulk_min(ctx)) LAUNCH(ctx, k_merge_many_rows, ceil_div(rows, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, parts, stride, rows, (int)nchunks, leaves);
  else if (use_quad(ctx)) LAUNCH_QUAD(ctx, k_quad_merge_many_rows, rows, parts, stride, rows, (int)nchunks, leaves);
  else LAUNCH_COOP(ctx, k_coop_merge_many_rows, rows, parts, stride, rows, (int)nchunks, leaves);
  return 0;
}
static int dev_hash_rows(zkl_ctx *ctx, const fe *lde, size_t ncols, size_t n, size_t blowup, const zkl_options *o, fe *leaves,
                         dev_buf<fe> &scratch, size_t ext = 1, size_t cstride = 0) {
  // rows of E elements (ext base columns each): PartitionOptions::partition_size::<E> counts E columns, minimum hash_rate / E [WF-mem]
  size_t N = n * blowup, ps = ext * partition_size_of(o, ncols / ext, ext);
  size_t nchunks = (ncols + ps - 1) / ps;
  if (nchunks == 1) return dev_hash_partitions(ctx, lde, ncols, n, blowup, ps, leaves, cstride);
  ZKL_CHECK(ctx, scratch.alloc(N * nchunks));
  ZKL_TRY(dev_hash_partitions(ctx, lde, ncols, n, blowup, ps, scratch.p, cstride));
  return dev_merge_partitions(ctx, scratch.p, N, N, nchunks, leaves);
}
// MerkleTree::new: nodes[count + i] = leaves[i] conceptually; we store nodes[1..count) and keep leaves
// separately. nodes has `count` entries (nodes[0] unused), nodes[1] = root.
static int dev_merkle(zkl_ctx *ctx, const fe *leaves, fe *nodes, size_t count) {
  if (count < 2 || (count & (count - 1))) { ctx->set_error("Merkle leaf count must be a power of two >= 2"); return 1; }
  const fe *in = leaves;
  for (size_t lvl = count / 2; lvl >= 1; lvl >>= 1) {
    if (lvl >= bulk_min(ctx) && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_merkle_level, lvl, in, nodes + lvl, lvl);
    else if (lvl >= bulk_min(ctx)) LAUNCH(ctx, k_merkle_level, ceil_div(lvl, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, in, nodes + lvl, lvl);
    else if (lvl <= ZKL_GRP_MAX_GROUPS) {   // the remaining levels in one launch
      if (use_quad(ctx)) LAUNCH(ctx, k_tree_top<ZKL_QUAD_GROUP>, 1, (unsigned)(lvl * ZKL_QUAD_GROUP), 0, in, nodes, (int)lvl);
      else LAUNCH(ctx, k_tree_top<ZKL_COOP_GROUP>, 1, (unsigned)(lvl * ZKL_COOP_GROUP), 0, in, nodes, (int)lvl);
      break;
    }
    else if (use_quad(ctx)) LAUNCH_QUAD(ctx, k_quad_merkle_level, lvl, in, nodes + lvl, lvl);
    else LAUNCH_COOP(ctx, k_coop_merkle_level, lvl, in, nodes + lvl, lvl);
    in = nodes + lvl;
  }
  return 0;
}

static void digest_from_fe_host(const fe &x, uint8_t out[32]) { memset(out, 0, 32); memcpy(out, x.l, 16); }

// ------------------------------------------------------------------------------------------------
// batched hasher primitives
// ------------------------------------------------------------------------------------------------
extern "C" int zkl_permute_batch(zkl_ctx *ctx, zkl_fe *states, size_t count) {
  ZKL_TRY(need_suite(ctx));
  dev_buf<fe> d; ZKL_CHECK(ctx, d.alloc(count * 12));
  ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, states, count * 12 * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_permute_batch, count, d.p, count);
  else LAUNCH(ctx, k_permute_batch, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, d.p, count);
  ZKL_CHECK(ctx, cudaMemcpyAsync(states, d.p, count * 12 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
static int download_digests(zkl_ctx *ctx, const fe *dev, size_t count, uint8_t *out) {
  std::vector<fe> h(count);
  ZKL_CHECK(ctx, cudaMemcpyAsync(h.data(), dev, count * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < count; i++) digest_from_fe_host(h[i], out + 32 * i);
  return 0;
}
extern "C" int zkl_hash_elements_batch(zkl_ctx *ctx, const zkl_fe *elems, size_t len, size_t count, uint8_t *out) {
  ZKL_TRY(need_suite(ctx));
  dev_buf<fe> d, o; ZKL_CHECK(ctx, d.alloc(std::max<size_t>(count * len, 1))); ZKL_CHECK(ctx, o.alloc(count));
  if (len) ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, elems, count * len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_hash_elements_batch, count, d.p, (int)len, count, o.p);
  else LAUNCH(ctx, k_hash_elements_batch, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, d.p, (int)len, count, o.p);
  return download_digests(ctx, o.p, count, out);
}
// a 32-byte digest as the element its sponge chunk folds to (utils.rs:359-371): lo16 mod p + (hi16 mod p)*2^64
static fe fold_digest_host(const uint8_t *b) {
  fe lo, hi; memcpy(lo.l, b, 16); memcpy(hi.l, b + 16, 16);
  return fe_fold_pair(fe_from_u128_mod(lo), fe_from_u128_mod(hi));
}
extern "C" int zkl_merge_batch(zkl_ctx *ctx, const uint8_t *pairs, size_t count, uint8_t *out) {
  ZKL_TRY(need_suite(ctx));
  std::vector<fe> h(2 * count);
  for (size_t i = 0; i < 2 * count; i++) h[i] = fold_digest_host(pairs + 32 * i);
  dev_buf<fe> d, o; ZKL_CHECK(ctx, d.alloc(2 * count)); ZKL_CHECK(ctx, o.alloc(count));
  ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, h.data(), 2 * count * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_merkle_level, count, d.p, o.p, count);
  else LAUNCH(ctx, k_merkle_level, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, d.p, o.p, count);
  return download_digests(ctx, o.p, count, out);
}
// Batched Merkle authentication (include/zkl_b200.h): the merge schedule is data independent (it follows from the query
// positions), so the host sends it once and the levels run back to back on the stream.
extern "C" int zkl_merge_plan(zkl_ctx *ctx, const uint8_t *known, size_t num_known, const uint32_t *left, const uint32_t *right,
                              const uint32_t *level_counts, size_t num_levels, const uint32_t *want, size_t num_want, uint8_t *out) {
  ZKL_TRY(need_suite(ctx));
  if (num_known == 0 || num_want == 0) { ctx->set_error("merge plan needs known digests and requested outputs"); return 1; }
  size_t total = 0, avail = num_known;
  for (size_t l = 0; l < num_levels; l++) {
    if (level_counts[l] == 0) { ctx->set_error("merge plan level without merges"); return 1; }
    for (size_t i = 0; i < level_counts[l]; i++)
      if (left[total + i] >= avail || right[total + i] >= avail) { ctx->set_error("merge plan refers to a digest that does not exist yet"); return 1; }
    total += level_counts[l]; avail += level_counts[l];
  }
  for (size_t i = 0; i < num_want; i++) if (want[i] >= avail) { ctx->set_error("merge plan output index out of range"); return 1; }
  if (avail > 0xFFFFFFFFull) { ctx->set_error("merge plan too large"); return 1; }
  std::vector<fe> h(num_known);
  for (size_t i = 0; i < num_known; i++) h[i] = fold_digest_host(known + 32 * i);
  std::vector<uint32_t> idx(2 * total + num_want);            // (left, right) interleaved: a gather yields the pairs a Merkle level kernel reads
  for (size_t i = 0; i < total; i++) { idx[2 * i] = left[i]; idx[2 * i + 1] = right[i]; }
  for (size_t i = 0; i < num_want; i++) idx[2 * total + i] = want[i];
  size_t widest = 0;
  for (size_t l = 0; l < num_levels; l++) widest = std::max<size_t>(widest, level_counts[l]);
  dev_buf<fe> table, pairs, res; dev_buf<uint32_t> didx;
  ZKL_CHECK(ctx, table.alloc(avail)); ZKL_CHECK(ctx, pairs.alloc(std::max<size_t>(2 * widest, 1))); ZKL_CHECK(ctx, res.alloc(num_want));
  ZKL_CHECK(ctx, didx.alloc(idx.size()));
  ZKL_CHECK(ctx, cudaMemcpyAsync(table.p, h.data(), num_known * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(didx.p, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  size_t off = 0, next = num_known;
  for (size_t l = 0; l < num_levels; l++) {
    const size_t cnt = level_counts[l];
    LAUNCH(ctx, k_gather_fe, ceil_div(2 * cnt, (size_t)256), 256, 0, table.p, didx.p + 2 * off, (int)(2 * cnt), pairs.p);
    if (cnt >= bulk_min(ctx) && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_merkle_level, cnt, pairs.p, table.p + next, cnt);
    else if (cnt >= bulk_min(ctx)) LAUNCH(ctx, k_merkle_level, ceil_div(cnt, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, pairs.p, table.p + next, cnt);
    else if (use_quad(ctx)) LAUNCH_QUAD(ctx, k_quad_merkle_level, cnt, pairs.p, table.p + next, cnt);
    else LAUNCH_COOP(ctx, k_coop_merkle_level, cnt, pairs.p, table.p + next, cnt);
    off += cnt; next += cnt;
  }
  LAUNCH(ctx, k_gather_fe, ceil_div(num_want, (size_t)256), 256, 0, table.p, didx.p + 2 * total, (int)num_want, res.p);
  return download_digests(ctx, res.p, num_want, out);
}
extern "C" int zkl_merge_with_int_batch(zkl_ctx *ctx, const uint8_t seed[32], const uint64_t *values, size_t count, uint8_t *out) {
  ZKL_TRY(need_suite(ctx));
  dev_buf<uint64_t> v; dev_buf<fe> o; ZKL_CHECK(ctx, v.alloc(count)); ZKL_CHECK(ctx, o.alloc(count));
  ZKL_CHECK(ctx, cudaMemcpyAsync(v.p, values, count * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_merge_with_int_batch, count, fold_digest_host(seed), v.p, count, o.p);
  else LAUNCH(ctx, k_merge_with_int_batch, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, fold_digest_host(seed), v.p, count, o.p);
  return download_digests(ctx, o.p, count, out);
}
// merge_with_int with one seed per value (the coins of many child proofs advanced in lockstep, replay.py)
extern "C" int zkl_merge_with_int_pairs(zkl_ctx *ctx, const uint8_t *seeds, const uint64_t *values, size_t count, uint8_t *out) {
  ZKL_TRY(need_suite(ctx));
  if (count == 0) return 0;
  std::vector<fe> h(count);
  for (size_t i = 0; i < count; i++) h[i] = fold_digest_host(seeds + 32 * i);
  dev_buf<fe> s, o; dev_buf<uint64_t> v;
  ZKL_CHECK(ctx, s.alloc(count)); ZKL_CHECK(ctx, o.alloc(count)); ZKL_CHECK(ctx, v.alloc(count));
  ZKL_CHECK(ctx, cudaMemcpyAsync(s.p, h.data(), count * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(v.p, values, count * 8, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_merge_with_int_pairs, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, s.p, v.p, count, o.p);
  return download_digests(ctx, o.p, count, out);
}
static int sponge_chunks(zkl_ctx *ctx, const uint8_t *data, size_t len, int kind, uint8_t out[32]) {
  size_t nch = (len + 31) / 32;
  std::vector<fe> h(std::max<size_t>(nch, 1));
  for (size_t i = 0; i < nch; i++) {  // hasher.rs:205-224: 32-byte chunks, zero padded, folded
    uint8_t c[32]; memset(c, 0, 32);
    memcpy(c, data + 32 * i, std::min<size_t>(32, len - 32 * i));
    h[i] = fold_digest_host(c);
  }
  dev_buf<fe> d, o; ZKL_CHECK(ctx, d.alloc(h.size())); ZKL_CHECK(ctx, o.alloc(1));
  ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_sponge_single, 1, 32, 0, d.p, (int)nch, kind, o.p);
  return download_digests(ctx, o.p, 1, out);
}
extern "C" int zkl_hash_bytes(zkl_ctx *ctx, const uint8_t *data, size_t len, uint8_t out[32]) {
  ZKL_TRY(need_suite(ctx));
  return sponge_chunks(ctx, data, len, 0, out);
}
extern "C" int zkl_merge_many(zkl_ctx *ctx, const uint8_t *digests, size_t count, uint8_t out[32]) {
  ZKL_TRY(need_suite(ctx));
  if (count == 0) { memset(out, 0, 32); return 0; }  // hasher.rs:88-90
  return sponge_chunks(ctx, digests, count * 32, 1, out);
}
extern "C" int zkl_merkle_build(zkl_ctx *ctx, const uint8_t *leaves, size_t count, uint8_t out_root[32], uint8_t *out_nodes) {
  ZKL_TRY(need_suite(ctx));
  std::vector<fe> h(count);
  for (size_t i = 0; i < count; i++) h[i] = fold_digest_host(leaves + 32 * i);
  dev_buf<fe> l, nd; ZKL_CHECK(ctx, l.alloc(count)); ZKL_CHECK(ctx, nd.alloc(count));
  ZKL_CHECK(ctx, cudaMemcpyAsync(l.p, h.data(), count * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemsetAsync(nd.p, 0, sizeof(fe), ctx->stream));
  ZKL_TRY(dev_merkle(ctx, l.p, nd.p, count));
  std::vector<uint8_t> nodes(32 * count);
  ZKL_TRY(download_digests(ctx, nd.p, count, nodes.data()));
  memcpy(out_root, nodes.data() + 32, 32);
  if (out_nodes) memcpy(out_nodes, nodes.data(), 32 * count);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NTT primitives
// ------------------------------------------------------------------------------------------------
extern "C" int zkl_interpolate_columns(zkl_ctx *ctx, zkl_fe *cols, size_t ncols, size_t n) {
  ZKL_TRY(enter(ctx));
  dev_buf<fe> d; ZKL_CHECK(ctx, d.alloc(ncols * n));
  ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, cols, ncols * n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_TRY(dev_interpolate(ctx, d.p, d.p, ncols, n));
  ZKL_CHECK(ctx, cudaMemcpyAsync(cols, d.p, ncols * n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int zkl_evaluate_columns(zkl_ctx *ctx, const zkl_fe *polys, size_t ncols, size_t n, size_t blowup, zkl_fe offset,
                                    zkl_fe *out) {
  ZKL_TRY(enter(ctx));
  if (!(offset.lo == 3 && offset.hi == 0)) { ctx->set_error("only the Winterfell domain offset GENERATOR = 3 is supported"); return 1; }
  size_t N = n * blowup;
  dev_buf<fe> p, cm, nat; ZKL_CHECK(ctx, p.alloc(ncols * n)); ZKL_CHECK(ctx, cm.alloc(ncols * N)); ZKL_CHECK(ctx, nat.alloc(ncols * N));
  ZKL_CHECK(ctx, cudaMemcpyAsync(p.p, polys, ncols * n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_TRY(dev_lde(ctx, p.p, cm.p, ncols, n, blowup));
  LAUNCH(ctx, k_coset_to_natural, ceil_div(ncols * N, 256), 256, 0, cm.p, nat.p, N, n, (int)blowup, ncols * N);
  ZKL_CHECK(ctx, cudaMemcpyAsync(out, nat.p, ncols * N * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

#include "host_air.inl"
#include "comm_local.inl"
#include "api_stark.inl"
#include "api_replay.inl"
