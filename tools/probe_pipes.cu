// probe_pipes.cu — instruction-throughput probes for the design of the tensor-core MDS path.
// Measures, per SM and per clock: IMMA (mma.sync m16n8k32 u8*u8->s32), IMAD.WIDE.U32, IMAD (32-bit),
// IADD3, SHF, and IMMA co-issued with ALU/IMAD work. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void k_imma(int *out, uint32_t seed) {
  int c[NACC][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = seed * (threadIdx.x + i);
  b[0] = seed + threadIdx.x; b[1] = seed ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < NACC; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[j][i] = 0;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) imma(c[j], a, b);
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IMAD.WIDE chains: 8 independent 64-bit accumulators
__global__ void k_imadw(uint64_t *out, uint32_t seed) {
  uint64_t acc[8];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = j;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(x), "r"(y));
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad(uint32_t *out, uint32_t seed) {
  uint32_t acc[8];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = j;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(x), "r"(y));
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t *out, uint32_t seed) {
  uint32_t acc[8];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = j;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %0, %0, %2;" : "+r"(acc[j]) : "r"(x), "r"(y));
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shf(uint32_t *out, uint32_t seed) {
  uint32_t acc[8];
  uint32_t x = seed + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = j + seed;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) asm volatile("shf.l.clamp.b32 %0, %0, %1, 7;" : "+r"(acc[j]) : "r"(x));
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// IMMA + ALU + IMAD.WIDE mix: per iteration 4 IMMA, NALU add/addc pairs, NMW IMAD.WIDE
template <int NALU, int NMW>
__global__ void k_mix(int *out, uint32_t seed) {
  int c[4][4];
  uint32_t a[4], b[2];
  uint32_t acc[8]; uint64_t w[8];
  uint32_t x = seed + threadIdx.x, y = seed * 3 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = seed * (threadIdx.x + i);
  b[0] = seed + threadIdx.x; b[1] = seed ^ threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[j][i] = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) { acc[j] = j; w[j] = j; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 4; j++) imma(c[j], a, b);
#pragma unroll
    for (int j = 0; j < NALU; j++) asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %0, %0, %2;" : "+r"(acc[j & 7]) : "r"(x), "r"(y));
#pragma unroll
    for (int j = 0; j < NMW; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j & 7]) : "r"(x), "r"(y));
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j] + (int)w[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  printf("device %s, %d SMs, clock %.3f GHz\n", p.name, sms, ghz);
  void *buf; cudaMalloc(&buf, 64 << 20);
  const int threads = 256; const int blocks = sms * 8;  // 2048 threads / SM = 16 warps per SMSP
  double warps = (double)blocks * threads / 32;
  auto report = [&](const char *name, float ms, double instr_per_warp, double unit_per_instr, const char *unit) {
    double clk = ms * 1e-3 * ghz * 1e9;
    double per_sm_clk = warps * instr_per_warp / sms / clk;  // warp-instructions per SM per clk
    printf("%-28s %8.3f ms  %.3f warp-instr/clk/SM  => %.1f %s/clk/SM\n", name, ms, per_sm_clk, per_sm_clk * unit_per_instr, unit);
  };
  float ms;
  ms = time_ms([&] { k_imma<4><<<blocks, threads>>>((int *)buf, 12345u); });
  report("IMMA.16832.u8 (4 acc)", ms, 4.0 * ITERS, 16 * 8 * 32, "MAC");
  ms = time_ms([&] { k_imma<8><<<blocks, threads>>>((int *)buf, 12345u); });
  report("IMMA.16832.u8 (8 acc)", ms, 8.0 * ITERS, 16 * 8 * 32, "MAC");
  ms = time_ms([&] { k_imma<8><<<sms * 4, 128>>>((int *)buf, 12345u); });
  {  // 4 warps per SMSP
    double w2 = (double)sms * 4 * 128 / 32; double clk = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %8.3f ms  => %.1f MAC/clk/SM\n", "IMMA (8 acc, 4 warps/SMSP)", ms, w2 * 8.0 * ITERS / sms / clk * 4096);
  }
  ms = time_ms([&] { k_imadw<<<blocks, threads>>>((uint64_t *)buf, 12345u); });
  report("IMAD.WIDE.U32", ms, 8.0 * ITERS, 32, "lane-op");
  ms = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t *)buf, 12345u); });
  report("IMAD (lo)", ms, 8.0 * ITERS, 32, "lane-op");
  ms = time_ms([&] { k_iadd3<<<blocks, threads>>>((uint32_t *)buf, 12345u); });
  report("IADD3+IADD3.X", ms, 16.0 * ITERS, 32, "lane-op");
  ms = time_ms([&] { k_shf<<<blocks, threads>>>((uint32_t *)buf, 12345u); });
  report("SHF", ms, 8.0 * ITERS, 32, "lane-op");
  ms = time_ms([&] { k_mix<0, 0><<<blocks, threads>>>((int *)buf, 12345u); });
  report("mix 4 IMMA only", ms, 4.0 * ITERS, 4096, "MAC");
  ms = time_ms([&] { k_mix<16, 0><<<blocks, threads>>>((int *)buf, 12345u); });
  report("mix 4 IMMA + 32 ALU", ms, 4.0 * ITERS, 4096, "MAC");
  ms = time_ms([&] { k_mix<16, 8><<<blocks, threads>>>((int *)buf, 12345u); });
  report("mix 4 IMMA + 32 ALU + 8 MW", ms, 4.0 * ITERS, 4096, "MAC");
  ms = time_ms([&] { k_mix<32, 16><<<blocks, threads>>>((int *)buf, 12345u); });
  report("mix 4 IMMA + 64 ALU + 16 MW", ms, 4.0 * ITERS, 4096, "MAC");
  ms = time_ms([&] { k_mix<32, 0><<<blocks, threads>>>((int *)buf, 12345u); });
  report("mix 4 IMMA + 64 ALU", ms, 4.0 * ITERS, 4096, "MAC");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
