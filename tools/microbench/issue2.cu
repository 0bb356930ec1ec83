// issue2.cu -- issue cost (SMSP cycles per warp-instruction) of the sweep's
// instruction classes, alone and mixed; every op is inline asm volatile on
// loop-carried registers so that nothing is folded away.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue2 issue2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 1024;
constexpr int kChains = 8;

#define FFMA2(a, m, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(c))
#define FFMA(a, m, c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(m), "f"(c))
#define PRMT(v, s) asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(v) : "r"(s))
#define I2FP(f, v) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v))
#define LOP3(v, s) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(v) : "r"(s))
#define IADD(v, s) asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(s))
#define VMX3(v, a, b) do { v = __vimax3_s16x2(v, a, b); asm volatile("" : "+r"(v)); } while (0)
#define FMNMX(f, g) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(g))

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, uint32_t seed, float fs) {
  unsigned long long a2[kChains];
  float a1[kChains], f1[kChains];
  uint32_t v[kChains], u[kChains];
  const unsigned long long m2 = ((unsigned long long) __float_as_uint(0.999f) << 32) | __float_as_uint(0.998f);
  const unsigned long long c2 = ((unsigned long long) __float_as_uint(0.001f) << 32) | __float_as_uint(0.002f);
  for (int i = 0; i < kChains; ++i) { a2[i] = threadIdx.x + i; a1[i] = i + fs; f1[i] = fs; v[i] = seed * (i + 1 + threadIdx.x); u[i] = seed + i; }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 10 || MODE == 12 || MODE == 13) FFMA2(a2[i], m2, c2);
      if (MODE == 1 || MODE == 5 || MODE == 7 || MODE == 8 || MODE == 9) PRMT(v[i], seed);
      if (MODE == 7) PRMT(u[i], seed);
      if (MODE == 2 || MODE == 6 || MODE == 9) I2FP(f1[i], u[i]);
      if (MODE == 3 || MODE == 12) VMX3(v[i], u[i], seed);
      if (MODE == 4) LOP3(v[i], seed);
      if (MODE == 11) IADD(v[i], seed);
      if (MODE == 8 || MODE == 10 || MODE == 14) FFMA(a1[i], fs, fs);
      if (MODE == 13 || MODE == 15) FMNMX(f1[i], a1[i]);
      if (MODE == 14) LOP3(v[i], seed);
    }
  }
  float s = 0;
  for (int i = 0; i < kChains; ++i) s += __uint_as_float((uint32_t) a2[i]) + a1[i] + f1[i] + __uint_as_float(v[i] ^ u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int inst, float* out, int sms, double ghz) {
  const int ctas = sms * 4, threads = 256;     // 8 warps per SMSP
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  k<MODE><<<ctas, threads>>>(out, 12345u, 0.5f); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0)); k<MODE><<<ctas, threads>>>(out, 12345u, 0.5f); CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1)); float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double groups_per_smsp = 8.0 * kIters * kChains;      // warps per SMSP x groups per warp
  const double clk = best * 1e-3 * ghz * 1e9;
  printf("%-28s %8.3f ms  %6.2f SMSP-cycles per group (%d instr)\n", name, best, clk / groups_per_smsp, inst);
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, %.3f GHz nominal (cycle counts assume it)\n", p.name, sms, ghz);
  float* out; CHECK(cudaMalloc(&out, (size_t) sms * 4 * 256 * sizeof(float)));
  run<0>("ffma2", 1, out, sms, ghz);
  run<1>("prmt", 1, out, sms, ghz);
  run<2>("i2fp", 1, out, sms, ghz);
  run<3>("vimnmx3.s16x2", 1, out, sms, ghz);
  run<4>("lop3", 1, out, sms, ghz);
  run<11>("iadd", 1, out, sms, ghz);
  run<15>("fmnmx", 1, out, sms, ghz);
  run<5>("ffma2 + prmt", 2, out, sms, ghz);
  run<6>("ffma2 + i2fp", 2, out, sms, ghz);
  run<7>("ffma2 + 2 prmt", 3, out, sms, ghz);
  run<12>("ffma2 + vimnmx3", 2, out, sms, ghz);
  run<13>("ffma2 + fmnmx", 2, out, sms, ghz);
  run<8>("ffma + prmt", 2, out, sms, ghz);
  run<14>("ffma + lop3", 2, out, sms, ghz);
  run<9>("prmt + i2fp", 2, out, sms, ghz);
  run<10>("ffma2 + ffma", 2, out, sms, ghz);
  return 0;
}
