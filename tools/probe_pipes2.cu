// probe_pipes2.cu — per-opcode issue rates on sm_100a with data-dependent operands (nothing hoistable):
// IMAD.WIDE.U32 (plain / with carry chain), IMAD, IADD3, IADD3.X chains, SHF, LOP3, SEL, LEA, PRMT.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define NCH 8

#define KERNEL(name, body)                                                             \
  __global__ void __launch_bounds__(256) name(uint32_t *out, uint32_t seed) {          \
    uint32_t a[NCH], b[NCH];                                                           \
    uint64_t w[NCH];                                                                   \
    _Pragma("unroll") for (int j = 0; j < NCH; j++) { a[j] = seed * (threadIdx.x + j + 1); b[j] = seed ^ (threadIdx.x * 7 + j); w[j] = a[j]; } \
    for (int it = 0; it < ITERS; it++) {                                               \
      _Pragma("unroll") for (int j = 0; j < NCH; j++) { body }                         \
    }                                                                                  \
    uint32_t s = 0;                                                                    \
    _Pragma("unroll") for (int j = 0; j < NCH; j++) s += a[j] + b[j] + (uint32_t)w[j] + (uint32_t)(w[j] >> 32); \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                    \
  }

KERNEL(k_imadw, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(b[j]));)
KERNEL(k_imadw_cc, asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(b[j]) : "r"(a[j]), "r"(b[j]));)
KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_iadd3x, asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.cc.u32 %1, %1, %0;\n\taddc.cc.u32 %0, %0, %1;\n\taddc.u32 %1, %1, %0;" : "+r"(a[j]), "+r"(b[j]));)
KERNEL(k_shf, asm volatile("shf.l.clamp.b32 %0, %0, %1, 7;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_sel, asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %0, p;}" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_mix_wa, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(b[j])); asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %1, %1, %0;" : "+r"(a[j]), "+r"(b[j]));)
KERNEL(k_mix_wa4, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(b[j])); asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.cc.u32 %1, %1, %0;\n\taddc.cc.u32 %0, %0, %1;\n\taddc.u32 %1, %1, %0;" : "+r"(a[j]), "+r"(b[j]));)

template <typename K>
static void run(const char *name, K kern, double instr_per_body, uint32_t *buf, int sms, double ghz, int warps_per_sm) {
  int threads = 256, blocks = sms * warps_per_sm / 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(buf, 12345u); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 3; i++) kern<<<blocks, threads>>>(buf, 12345u + i);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  double clk = ms * 1e-3 * ghz * 1e9;
  double warp_instr = (double)blocks * threads / 32 * ITERS * NCH * instr_per_body;
  double per_smsp = warp_instr / sms / 4 / clk;
  printf("%-34s warps/SM %2d  %7.3f ms  %.3f warp-instr/clk/SMSP  (%.2f clk per warp-instr)\n", name, warps_per_sm, ms, per_smsp, 1.0 / per_smsp);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  uint32_t *buf; cudaMalloc(&buf, 64 << 20);
  for (int wps : {64, 16}) {
    run("IMAD.WIDE.U32 (64-bit acc)", k_imadw, 1, buf, sms, ghz, wps);
    run("IMAD.WIDE lo.cc + hi (pair)", k_imadw_cc, 1, buf, sms, ghz, wps);
    run("IMAD (32-bit)", k_imad, 1, buf, sms, ghz, wps);
    run("IADD3", k_iadd3, 1, buf, sms, ghz, wps);
    run("IADD3 + 3x IADD3.X chain", k_iadd3x, 4, buf, sms, ghz, wps);
    run("SHF", k_shf, 1, buf, sms, ghz, wps);
    run("LOP3", k_lop3, 1, buf, sms, ghz, wps);
    run("ISETP+SEL", k_sel, 2, buf, sms, ghz, wps);
    run("PRMT", k_prmt, 1, buf, sms, ghz, wps);
    run("mix: 1 IMAD.WIDE + 2 IADD3", k_mix_wa, 3, buf, sms, ghz, wps);
    run("mix: 1 IMAD.WIDE + 4 IADD3", k_mix_wa4, 5, buf, sms, ghz, wps);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
