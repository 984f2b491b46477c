// probe_pipes3.cu — round 2 probes for the sponge permutation's wide multiplies (VERDICT r1 task 4):
//   (i)  the FP64 pipe as a second multiplier: DFMA issue rate alone and CO-ISSUED with IMAD.WIDE.U32 / IADD3, plus the
//        conversions a limb product on doubles needs (I2F.F64, the 2^52 magic-add extraction);
//   (ii) a 32x32->64 product as IMAD (lo) + IMAD.HI instead of one IMAD.WIDE.U32;
//   (iii) IMAD.WIDE with more independent chains per thread (latency vs issue).
// Every body works on data-dependent operands (nothing hoistable). Output: clk per warp instruction per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define NCH 8

#define KERNEL(name, body)                                                             \
  __global__ void __launch_bounds__(256) name(uint32_t *out, uint32_t seed) {          \
    uint32_t a[NCH], b[NCH];                                                           \
    uint64_t w[NCH];                                                                   \
    double d[NCH], e[NCH];                                                             \
    _Pragma("unroll") for (int j = 0; j < NCH; j++) { a[j] = seed * (threadIdx.x + j + 1); b[j] = seed ^ (threadIdx.x * 7 + j); w[j] = a[j]; d[j] = (double)(a[j] & 0xFFFFF); e[j] = 1.0 + 1e-9 * (b[j] & 0xFF); } \
    for (int it = 0; it < ITERS; it++) {                                               \
      _Pragma("unroll") for (int j = 0; j < NCH; j++) { body }                         \
    }                                                                                  \
    uint32_t s = 0;                                                                    \
    _Pragma("unroll") for (int j = 0; j < NCH; j++) s += a[j] + b[j] + (uint32_t)w[j] + (uint32_t)(w[j] >> 32) + (uint32_t)__double2loint(d[j]) + (uint32_t)__double2hiint(e[j]); \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                    \
  }

#define IMADW asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(b[j]));
#define DFMA asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[j]) : "d"(e[j]));
#define DFMA_RZ asm volatile("fma.rz.f64 %0, %0, %1, %1;" : "+d"(d[j]) : "d"(e[j]));
#define IADD asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %1, %1, %0;" : "+r"(a[j]), "+r"(b[j]));

KERNEL(k_imadw, IMADW)
KERNEL(k_dfma, DFMA)
KERNEL(k_dfma_rz, DFMA_RZ)
KERNEL(k_dadd, asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(e[j]));)
KERNEL(k_dfma_imadw, DFMA IMADW)
KERNEL(k_dfma2_imadw, DFMA DFMA_RZ IMADW)
KERNEL(k_dfma_iadd2, DFMA IADD)
KERNEL(k_dfma_imadw_iadd2, DFMA IMADW IADD)
KERNEL(k_i2d, asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[j]) : "r"(a[j])); a[j] += (uint32_t)__double2loint(d[j]);)
KERNEL(k_d2i_magic, asm volatile("add.rn.f64 %0, %0, 0d4338000000000000;" : "+d"(d[j])); a[j] ^= (uint32_t)__double2loint(d[j]); d[j] = __hiloint2double(0x43300000, a[j] & 0xFFFFF);)
KERNEL(k_mul_lo_hi, asm volatile("{.reg .u32 l, h; mul.lo.u32 l, %0, %1; mul.hi.u32 h, %0, %1; add.u32 %0, l, h;}" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_mad_lo_hi_cc, asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(b[j]) : "r"(a[j] | 1), "r"(b[j] | 3));)
KERNEL(k_imadw_x, asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %2; add.u64 %0, %0, t;}" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(b[j]));)
KERNEL(k_mul24, asm volatile("mul24.lo.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_imad16, asm volatile("{.reg .u32 x, y; and.b32 x, %0, 0xFFFF; and.b32 y, %1, 0xFFFF; mad.lo.u32 %0, x, y, %1;}" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_dp4a, asm volatile("dp4a.u32.u32 %0, %0, %1, %1;" : "+r"(a[j]) : "r"(b[j]));)
KERNEL(k_dp4a_imadw, asm volatile("dp4a.u32.u32 %0, %0, %1, %1;" : "+r"(a[j]) : "r"(b[j])); IMADW)
KERNEL(k_fmul_imadw, { float f = __uint_as_float(a[j]); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f)); a[j] = __float_as_uint(f); } IMADW)

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
  double bodies = (double)blocks * threads / 32 * ITERS * NCH;
  double clk_per_body = clk * sms * 4 / bodies;
  printf("%-44s warps/SM %2d  %7.3f ms  %6.2f clk per body per SMSP  (%.2f clk per warp-instr, %g instr/body)\n", name, warps_per_sm, ms, clk_per_body,
         clk_per_body / instr_per_body, instr_per_body);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  printf("%s, %d SMs, %.3f GHz\n", p.name, sms, ghz);
  uint32_t *buf; cudaMalloc(&buf, 64 << 20);
  for (int wps : {64, 16}) {
    run("IMAD.WIDE.U32 (64-bit acc)", k_imadw, 1, buf, sms, ghz, wps);
    run("mul.wide + add.u64", k_imadw_x, 1, buf, sms, ghz, wps);
    run("DFMA.RN", k_dfma, 1, buf, sms, ghz, wps);
    run("DFMA.RZ", k_dfma_rz, 1, buf, sms, ghz, wps);
    run("DADD", k_dadd, 1, buf, sms, ghz, wps);
    run("co-issue: 1 DFMA + 1 IMAD.WIDE", k_dfma_imadw, 2, buf, sms, ghz, wps);
    run("co-issue: 2 DFMA + 1 IMAD.WIDE", k_dfma2_imadw, 3, buf, sms, ghz, wps);
    run("co-issue: 1 DFMA + 2 IADD3", k_dfma_iadd2, 3, buf, sms, ghz, wps);
    run("co-issue: 1 DFMA + 1 IMAD.WIDE + 2 IADD3", k_dfma_imadw_iadd2, 4, buf, sms, ghz, wps);
    run("I2F.F64.U32 (+ extract)", k_i2d, 1, buf, sms, ghz, wps);
    run("magic-add double->int + rebuild", k_d2i_magic, 1, buf, sms, ghz, wps);
    run("mul.lo + mul.hi (+add)", k_mul_lo_hi, 3, buf, sms, ghz, wps);
    run("mad.lo.cc + madc.hi (pair)", k_mad_lo_hi_cc, 2, buf, sms, ghz, wps);
    run("mul24.lo", k_mul24, 1, buf, sms, ghz, wps);
    run("16x16 via 2 LOP + IMAD", k_imad16, 3, buf, sms, ghz, wps);
    run("DP4A.U32", k_dp4a, 1, buf, sms, ghz, wps);
    run("co-issue: 1 DP4A + 1 IMAD.WIDE", k_dp4a_imadw, 2, buf, sms, ghz, wps);
    run("co-issue: 1 FFMA + 1 IMAD.WIDE", k_fmul_imadw, 2, buf, sms, ghz, wps);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
