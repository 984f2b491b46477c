// tc_probe.cu — stand-alone check + timing of the tensor-core permutation (zk-lisp_b200/csrc/tcperm.cuh)
// against a host permutation written with the canonical field routines, and against the integer-pipe kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../zk-lisp_b200/csrc/kernels_tc.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static uint64_t sm_state = 0x5A4B4C;
static uint64_t splitmix() {
  uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
static fe rand_fe() {
  fe r; uint64_t a = splitmix(), b = splitmix();
  r.l[0] = (uint32_t)a; r.l[1] = (uint32_t)(a >> 32); r.l[2] = (uint32_t)b; r.l[3] = (uint32_t)(b >> 32);
  return fe_from_u128_mod(r);
}
static void host_permute(fe *s, const suite_consts &su) {
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    fe c[12], o[12];
    for (int i = 0; i < 12; i++) c[i] = fe_cube(s[i]);
    for (int i = 0; i < 12; i++) {
      fe acc = su.rc[r][i];
      for (int k = 0; k < 12; k++) acc = fe_add(acc, fe_mul(su.mds[i][k], c[k]));
      o[i] = acc;
    }
    memcpy(s, o, sizeof o);
  }
}
template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms / reps;
}

int main(int argc, char **argv) {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  suite_consts su;
  for (int i = 0; i < 2; i++) su.dom[i] = rand_fe();
  for (int i = 0; i < 12; i++) for (int k = 0; k < 12; k++) su.mds[i][k] = rand_fe();
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int i = 0; i < 12; i++) su.rc[r][i] = rand_fe();
  label_consts lc; lc.elements = rand_fe(); lc.merge = rand_fe(); lc.merge_many = rand_fe(); lc.merge_with_int = rand_fe(); lc.bytes = rand_fe();
  CK(cudaMemcpyToSymbol(c_suite, &su, sizeof su)); CK(cudaMemcpyToSymbol(g_suite, &su, sizeof su));
  CK(cudaMemcpyToSymbol(c_labels, &lc, sizeof lc));
  suite_consts *gs; CK(cudaGetSymbolAddress((void **)&gs, g_suite));
  k_build_mds_bytes<<<(TC_KCH * TC_N + 127) / 128, 128>>>(gs);
  CK(cudaDeviceSynchronize());
  CK(cudaFuncSetAttribute(k_tc_permute_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_tc_hash_elements_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_tc_merkle_level<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_tc_hash_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  int fails = 0;
  printf("largest possible MMA accumulator %u (bound %u)\n", tc_max_accumulator(su), TC_ACC_MAX);
  if (tc_max_accumulator(su) > TC_ACC_MAX) { printf("random MDS exceeds the accumulator bound\n"); return 3; }

  {  // 1. permutation vs host
    const size_t count = 1000;
    std::vector<fe> h(count * 12), ref;
    for (auto &x : h) x = rand_fe();
    h[0] = fe_zero(); for (int i = 0; i < 4; i++) h[1].l[i] = (i == 1 ? 0xFFFFD300u : (i == 0 ? 0u : 0xFFFFFFFFu));  // p - 1
    ref = h;
    for (size_t i = 0; i < count; i++) host_permute(&ref[12 * i], su);
    fe *d; CK(cudaMalloc(&d, count * 12 * sizeof(fe)));
    CK(cudaMemcpy(d, h.data(), count * 12 * sizeof(fe), cudaMemcpyHostToDevice));
    k_tc_permute_batch<false><<<2, TC_THREADS, TC_SMEM_BYTES>>>(d, count);
    CK(cudaDeviceSynchronize());
    std::vector<fe> got(count * 12);
    CK(cudaMemcpy(got.data(), d, count * 12 * sizeof(fe), cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < count * 12; i++) if (!fe_eq(got[i], ref[i])) { if (bad < 4) printf("  mismatch state %zu lane %zu: got %08x%08x%08x%08x want %08x%08x%08x%08x\n", i / 12, i % 12, got[i].l[3], got[i].l[2], got[i].l[1], got[i].l[0], ref[i].l[3], ref[i].l[2], ref[i].l[1], ref[i].l[0]); bad++; }
    printf("permute_batch vs host: %zu / %zu lanes differ\n", bad, count * 12);
    fails += bad != 0;
    CK(cudaFree(d));
  }
  {  // 2. hash_elements and merkle level vs the integer-pipe kernels
    for (int len : {1, 7, 18, 19, 40, 204, 219}) {
      const size_t count = 777;
      std::vector<fe> h(count * len); for (auto &x : h) x = rand_fe();
      fe *d, *o1, *o2; CK(cudaMalloc(&d, h.size() * sizeof(fe))); CK(cudaMalloc(&o1, count * sizeof(fe))); CK(cudaMalloc(&o2, count * sizeof(fe)));
      CK(cudaMemcpy(d, h.data(), h.size() * sizeof(fe), cudaMemcpyHostToDevice));
      k_hash_elements_batch<<<(count + ZKL_HASH_THREADS - 1) / ZKL_HASH_THREADS, ZKL_HASH_THREADS>>>(d, len, count, o1);
      k_tc_hash_elements_batch<false><<<3, TC_THREADS, TC_SMEM_BYTES>>>(d, len, count, o2);
      CK(cudaDeviceSynchronize());
      std::vector<fe> a(count), b(count);
      CK(cudaMemcpy(a.data(), o1, count * sizeof(fe), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), o2, count * sizeof(fe), cudaMemcpyDeviceToHost));
      size_t bad = 0; for (size_t i = 0; i < count; i++) bad += !fe_eq(a[i], b[i]);
      printf("hash_elements len %3d: %zu / %zu differ\n", len, bad, count);
      fails += bad != 0;
      CK(cudaFree(d)); CK(cudaFree(o1)); CK(cudaFree(o2));
    }
    const size_t count = 4096 + 5;
    std::vector<fe> h(2 * count); for (auto &x : h) x = rand_fe();
    fe *d, *o1, *o2; CK(cudaMalloc(&d, h.size() * sizeof(fe))); CK(cudaMalloc(&o1, count * sizeof(fe))); CK(cudaMalloc(&o2, count * sizeof(fe)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(fe), cudaMemcpyHostToDevice));
    k_merkle_level<<<(count + ZKL_HASH_THREADS - 1) / ZKL_HASH_THREADS, ZKL_HASH_THREADS>>>(d, o1, count);
    k_tc_merkle_level<false><<<sms, TC_THREADS, TC_SMEM_BYTES>>>(d, o2, count);
    CK(cudaDeviceSynchronize());
    std::vector<fe> a(count), b(count);
    CK(cudaMemcpy(a.data(), o1, count * sizeof(fe), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), o2, count * sizeof(fe), cudaMemcpyDeviceToHost));
    size_t bad = 0; for (size_t i = 0; i < count; i++) bad += !fe_eq(a[i], b[i]);
    printf("merkle level: %zu / %zu differ\n", bad, count);
    fails += bad != 0;
    CK(cudaFree(d)); CK(cudaFree(o1)); CK(cudaFree(o2));
  }
  {  // 3. timing
    for (size_t count : {(size_t)1 << 16, (size_t)1 << 18, (size_t)1 << 20}) {
      fe *d; CK(cudaMalloc(&d, count * 12 * sizeof(fe))); CK(cudaMemset(d, 0x11, count * 12 * sizeof(fe)));
      size_t tiles = (count + TC_CTA_SPONGES - 1) / TC_CTA_SPONGES;
      int grid = (int)(tiles < (size_t)sms ? tiles : sms);
      float t_old = time_ms([&] { k_permute_batch<<<(count + ZKL_HASH_THREADS - 1) / ZKL_HASH_THREADS, ZKL_HASH_THREADS>>>(d, count); }, 3);
      float t_tc = time_ms([&] { k_tc_permute_batch<false><<<grid, TC_THREADS, TC_SMEM_BYTES>>>(d, count); }, 3);
      printf("permute %8zu states: integer pipe %8.3f ms (%.1f M/s)   tensor core %8.3f ms (%.1f M/s)   x%.2f\n", count, t_old, count / t_old / 1e3, t_tc,
             count / t_tc / 1e3, t_old / t_tc);
      CK(cudaFree(d));
    }
  }
  CK(cudaDeviceSynchronize());
  printf(fails ? "TC PROBE FAILED\n" : "TC PROBE OK\n");
  return fails != 0;
}
