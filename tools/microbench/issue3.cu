// issue3.cu -- what one 16-bit stereo frame of the packed sweep costs in SMSP
// cycles, by how the frame word is turned into two floats.  A "frame group" is
// 11 FFMA2 (the K filter, energy and mode sums of both channels) + the peak
// VIMNMX3 + one conversion variant; every op is inline asm volatile on
// loop-carried registers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue3 issue3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 512;
constexpr int kChains = 4;

#define FFMA2(a, m, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(c))
#define PRMT_LO(d, v) asm volatile("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(d) : "r"(v))
#define PRMT_HI(d, v) asm volatile("prmt.b32 %0, %1, %1, 0xBB32;" : "=r"(d) : "r"(v))
#define I2FP(f, v) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v))
#define SGXT(d, v) asm volatile("bfe.s32 %0, %1, 0, 16;" : "=r"(d) : "r"(v))
#define SHRS(d, v) asm volatile("shr.s32 %0, %1, 16;" : "=r"(d) : "r"(v))
#define SHL16(d, v) asm volatile("shl.b32 %0, %1, 16;" : "=r"(d) : "r"(v))
#define ANDHI(d, v) asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(d) : "r"(v))
#define XORB(d, v) asm volatile("xor.b32 %0, %1, 0x80008000;" : "=r"(d) : "r"(v))
#define MAGIC_LO(d, v) asm volatile("prmt.b32 %0, %1, 0x4b000000, 0x7610;" : "=r"(d) : "r"(v))
#define MAGIC_HI(d, v) asm volatile("prmt.b32 %0, %1, 0x4b000000, 0x7632;" : "=r"(d) : "r"(v))
#define CVT16_LO(f, v) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, lo;}" : "=f"(f) : "r"(v))
#define CVT16_HI(f, v) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi;}" : "=f"(f) : "r"(v))
#define VMX3(v, a, b) do { v = __vimax3_s16x2(v, a, b); asm volatile("" : "+r"(v)); } while (0)

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return ((unsigned long long) __float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// CONV: 0 none, 1 prmt+i2fp (the kernel's), 2 sgxt / shr.s32 + i2fp, 3 shl / and + i2fp (both
// channels scaled by 65536), 4 cvt.f32.s16 on both halves (XU), 5 xor + magic prmt + fadd2,
// 6 prmt+i2fp low half, cvt.f32.s16 high half.   NF = FFMA2 per frame, PEAK = with VIMNMX3
template <int CONV, int NF, int PEAK>
__global__ void __launch_bounds__(128) k(float* out, uint32_t seed, float fs) {
  unsigned long long a2[kChains][4];
  uint32_t w[kChains], mx[kChains];
  const unsigned long long m2 = pack(0.998f, 0.999f), c2 = pack(0.002f, 0.001f);
  const unsigned long long magic = pack(-8421376.0f, -8421376.0f), one = pack(1.0f, 1.0f);
  for (int i = 0; i < kChains; ++i) {
    for (int j = 0; j < 4; ++j) a2[i][j] = threadIdx.x + i + j;
    w[i] = seed * (i + 1 + threadIdx.x); mx[i] = 0;
  }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      float xl = 0.0f, xr = 0.0f;
      uint32_t t0, t1;
      if (CONV == 1) { PRMT_LO(t0, w[i]); PRMT_HI(t1, w[i]); I2FP(xl, t0); I2FP(xr, t1); }
      if (CONV == 2) { SGXT(t0, w[i]); SHRS(t1, w[i]); I2FP(xl, t0); I2FP(xr, t1); }
      if (CONV == 3) { SHL16(t0, w[i]); ANDHI(t1, w[i]); I2FP(xl, t0); I2FP(xr, t1); }
      if (CONV == 4) { CVT16_LO(xl, w[i]); CVT16_HI(xr, w[i]); }
      if (CONV == 6) { PRMT_LO(t0, w[i]); I2FP(xl, t0); CVT16_HI(xr, w[i]); }
      unsigned long long x2 = pack(xl, xr);
      if (CONV == 5) {
        uint32_t b; XORB(b, w[i]); MAGIC_LO(t0, b); MAGIC_HI(t1, b);
        x2 = ((unsigned long long) t1 << 32) | t0;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x2) : "l"(one), "l"(magic));
      }
      if (PEAK) VMX3(mx[i], w[i], seed);
      // the sample enters the first chain; NF FFMA2 in four dependency chains
      if (NF > 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i][0]) : "l"(m2), "l"(x2));
#pragma unroll
      for (int j = 1; j < NF; ++j) FFMA2(a2[i][j & 3], m2, c2);
      // next frame word depends on the filter state (nothing can be hoisted)
      w[i] += (uint32_t) a2[i][0] & 0x10001u;
    }
  }
  float s = 0;
  for (int i = 0; i < kChains; ++i)
    for (int j = 0; j < 4; ++j) s += __uint_as_float((uint32_t) a2[i][j]) + __uint_as_float(w[i] ^ mx[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CONV, int NF, int PEAK>
static void run(const char* name, float* out, int sms, double ghz) {
  const int ctas = sms * 4, threads = 128;     // 4 warps per SMSP, as in the sweep
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  k<CONV, NF, PEAK><<<ctas, threads>>>(out, 12345u, 0.5f); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0)); k<CONV, NF, PEAK><<<ctas, threads>>>(out, 12345u, 0.5f); CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1)); float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double frames_per_smsp = 4.0 * kIters * kChains;      // warps per SMSP x frames per warp
  const double clk = best * 1e-3 * ghz * 1e9;
  printf("%-58s %8.3f ms  %6.2f SMSP-cycles per warp-frame\n", name, best, clk / frames_per_smsp);
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, %.3f GHz nominal (cycle counts assume it); 4 warps per SMSP\n", p.name, sms, ghz);
  float* out; CHECK(cudaMalloc(&out, (size_t) sms * 4 * 128 * sizeof(float)));
  run<0, 11, 0>("11 ffma2", out, sms, ghz);
  run<0, 11, 1>("11 ffma2 + vimnmx3", out, sms, ghz);
  run<1, 11, 1>("11 ffma2 + vimnmx3 + 2 prmt + 2 i2fp (kernel)", out, sms, ghz);
  run<2, 11, 1>("11 ffma2 + vimnmx3 + sgxt + shr.s32 + 2 i2fp", out, sms, ghz);
  run<3, 11, 1>("11 ffma2 + vimnmx3 + shl + and + 2 i2fp (x65536)", out, sms, ghz);
  run<4, 11, 1>("11 ffma2 + vimnmx3 + 2 cvt.f32.s16 (XU)", out, sms, ghz);
  run<6, 11, 1>("11 ffma2 + vimnmx3 + prmt + i2fp + cvt.f32.s16", out, sms, ghz);
  run<5, 11, 1>("11 ffma2 + vimnmx3 + xor + 2 magic prmt + fadd2", out, sms, ghz);
  run<1, 0, 0>("2 prmt + 2 i2fp alone", out, sms, ghz);
  run<2, 0, 0>("sgxt + shr.s32 + 2 i2fp alone", out, sms, ghz);
  run<3, 0, 0>("shl + and + 2 i2fp alone", out, sms, ghz);
  run<4, 0, 0>("2 cvt.f32.s16 alone", out, sms, ghz);
  return 0;
}
