// issue.cu -- does a packed FFMA2 cost one issue slot or two?  Mixes of FMA-pipe
// and ALU-pipe instructions, reported as warp-instructions per clock per SM
// (issue limit: 4) and FP32 lane-ops per clock per SM (FMA pipe limit: 128).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue issue.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 2048;
constexpr int kChains = 6;

struct Consts { float c[8]; };

// MODE: 0 = FFMA2 only, 1 = FFMA2 + 1 ALU (LOP3), 2 = FFMA2 + 2 ALU, 3 = 2 FFMA + 1 ALU,
//       4 = 2 FFMA only, 5 = FFMA2 (UR broadcast) + PRMT + I2FP, 6 = ALU only (LOP3),
//       7 = I2FP only, 8 = VIMNMX3.S16x2 only, 9 = FMNMX3 only, 10 = FFMA2 + FMNMX3
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(float2* out, const __grid_constant__ Consts K, uint32_t seed) {
  float2 acc[kChains];
  uint32_t v[kChains], u[kChains];
  float f[kChains];
  for (int i = 0; i < kChains; ++i) {
    acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    v[i] = seed * (threadIdx.x + i + 1); u[i] = seed + i; f[i] = i;
  }
  const float2 b = make_float2(K.c[2], K.c[3]);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 10) acc[i] = __ffma2_rn(acc[i], make_float2(K.c[0], K.c[0]), b);
      if (MODE == 3 || MODE == 4) { acc[i].x = fmaf(acc[i].x, K.c[0], b.x); acc[i].y = fmaf(acc[i].y, K.c[1], b.y); }
      if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 6) v[i] = (v[i] & u[i]) ^ seed;
      if (MODE == 2) u[i] = (u[i] | v[i]) ^ seed;
      if (MODE == 5) {
        int s; asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(s) : "r"(v[i]));
        const float x = (float) s;
        acc[i] = __ffma2_rn(acc[i], make_float2(K.c[0], K.c[0]), make_float2(x, x));
        v[i] += 0x10001u;
      }
      if (MODE == 7) { f[i] += 1.0f; v[i] = __float_as_uint((float) (int) v[i]); }
      if (MODE == 8) v[i] = __vimax3_s16x2(v[i], u[i], seed);
      if (MODE == 9 || MODE == 10) f[i] = fmaxf(fmaxf(f[i], fabsf(__uint_as_float(u[i]))), fabsf(__uint_as_float(seed)));
    }
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < kChains; ++i) { s.x += acc[i].x + __uint_as_float(v[i]) + f[i]; s.y += acc[i].y + __uint_as_float(u[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double inst_per_chain, double fp_per_chain, float2* out, int sms, double ghz) {
  Consts K; for (int i = 0; i < 8; ++i) K.c[i] = 0.99f - 0.01f * i;
  const int ctas = sms * 4, threads = 256;
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  k_mix<MODE><<<ctas, threads>>>(out, K, 12345u); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0)); k_mix<MODE><<<ctas, threads>>>(out, K, 12345u); CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1)); float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double warps = (double) ctas * threads / 32, n = (double) kIters * kChains;
  const double clk = best * 1e-3 * ghz * 1e9;
  printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM  %7.1f fp32 lane-ops/clk/SM\n", name, best,
         warps * n * inst_per_chain / clk / sms, warps * n * fp_per_chain * 32 / clk / sms);
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, %.3f GHz (nominal; numbers assume this clock)\n", p.name, sms, ghz);
  float2* out; CHECK(cudaMalloc(&out, (size_t) sms * 4 * 256 * sizeof(float2)));
  run<0>("ffma2", 1, 2, out, sms, ghz);
  run<4>("2 ffma", 2, 2, out, sms, ghz);
  run<1>("ffma2 + lop3", 2, 2, out, sms, ghz);
  run<2>("ffma2 + 2 lop3", 3, 2, out, sms, ghz);
  run<3>("2 ffma + lop3", 3, 2, out, sms, ghz);
  run<6>("lop3", 1, 0, out, sms, ghz);
  run<5>("ffma2 + prmt + i2fp + iadd", 4, 2, out, sms, ghz);
  run<7>("i2fp + fadd", 2, 1, out, sms, ghz);
  run<8>("vimnmx3.s16x2", 1, 0, out, sms, ghz);
  run<9>("fmnmx3", 1, 0, out, sms, ghz);
  run<10>("ffma2 + fmnmx3", 2, 2, out, sms, ghz);
  return 0;
}
