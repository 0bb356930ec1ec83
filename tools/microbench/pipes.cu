// pipes.cu -- instruction-throughput microbenchmarks that shape the sweep
// kernel (run on the B200 via gpurun; results summarised in profiles/).
//   ffma_reg   : FFMA with three register sources
//   ffma_const : FFMA with a constant-bank coefficient (FIR tap form)
//   ffma2      : packed FFMA2 (two FP32 FMAs per lane per instruction)
//   dfma       : FP64 FMA
//   fmnmx      : FMNMX with |x| (peak tracking)
//   i2f        : int16 -> float conversion
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;
__constant__ float c_coef[16] = {0.99f, 0.98f, 0.97f, 0.96f, 0.95f, 0.94f, 0.93f, 0.92f,
                                 0.91f, 0.9f, 0.89f, 0.88f, 0.87f, 0.86f, 0.85f, 0.84f};

__global__ void k_ffma_reg(float* out, float a, float b) {
  float acc[kChains];
  for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FIR-like: acc += x[t] * const_coef[t]; x values in registers, rotated
__global__ void k_ffma_const(float* out, float seed) {
  float x[8], acc[kChains];
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  for (int i = 0; i < kChains; ++i) acc[i] = 0.f;
  for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) acc[i] = fmaf(x[(i + t) & 7], c_coef[t], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = acc[i] * 1e-3f;   // 8 extra ops per 64
  }
  float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float2* out, float2 a, float2 b) {
  float2 acc[kChains];
  for (int i = 0; i < kChains; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < kChains; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FIR-like with coefficient pairs held in registers
__global__ void k_ffma2_fir(float2* out, float seed) {
  float2 x[8], acc[kChains], c[8];
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(seed + threadIdx.x * 1e-3f + i, i); c[i] = make_float2(c_coef[i], c_coef[i]); }
  for (int i = 0; i < kChains; ++i) acc[i] = make_float2(0, 0);
  for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) acc[i] = __ffma2_rn(x[(i + t) & 7], c[t], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(acc[i].x * 1e-3f, acc[i].y * 1e-3f);
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < kChains; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, double a, double b) {
  double acc[kChains];
  for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fmnmx(float* out, float a) {
  float acc[kChains];
  for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = fmaxf(fabsf(acc[i] - a), a);
  }
  float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: 2 FFMA2 + 1 FMNMX per group (issue-slot sharing between pipes)
__global__ void k_mix(float2* out, float2 a, float2 b) {
  float2 acc[kChains]; float m[kChains];
  for (int i = 0; i < kChains; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); m[i] = 0; }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      acc[i] = __ffma2_rn(acc[i], a, b);
      m[i] = fmaxf(m[i], fabsf(acc[i].x));
    }
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < kChains; ++i) { s.x += acc[i].x + m[i]; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_i2f(float* out, const int* in) {
  int v[kChains]; float acc[kChains];
  for (int i = 0; i < kChains; ++i) { v[i] = in[threadIdx.x + i]; acc[i] = 0; }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) { acc[i] += (float) (short) (v[i] + it); }
  }
  float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  launch(); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CHECK(cudaEventRecord(e0)); launch(); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, threads = 512, ctas = sms * 4;
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
  void* out; CHECK(cudaMalloc(&out, (size_t) ctas * threads * 16));
  int* in; CHECK(cudaMalloc(&in, 4096 * 4)); CHECK(cudaMemset(in, 1, 4096 * 4));
  const double lanes = (double) ctas * threads * kIters * kChains;
  auto rep = [&](const char* name, double ms, double ops_per_lane_iter) {
    const double ops = lanes * ops_per_lane_iter;
    printf("%-12s %8.3f ms  %8.2f Tops/s  %7.1f lane-ops/clk/SM @1.965GHz\n", name, ms,
           ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9);
  };
  rep("ffma_reg", time_ms([&] { k_ffma_reg<<<ctas, threads>>>((float*) out, 0.999f, 0.001f); }), 1);
  rep("ffma_const", time_ms([&] { k_ffma_const<<<ctas, threads>>>((float*) out, 0.5f); }), 1);
  rep("ffma2", time_ms([&] { k_ffma2<<<ctas, threads>>>((float2*) out, make_float2(0.999f, 0.998f), make_float2(0.001f, 0.002f)); }), 2);
  rep("ffma2_fir", time_ms([&] { k_ffma2_fir<<<ctas, threads>>>((float2*) out, 0.5f); }), 2);
  rep("dfma", time_ms([&] { k_dfma<<<ctas, threads>>>((double*) out, 0.999, 0.001); }), 1);
  rep("fmnmx", time_ms([&] { k_fmnmx<<<ctas, threads>>>((float*) out, 0.5f); }), 2);
  rep("mix2+1", time_ms([&] { k_mix<<<ctas, threads>>>((float2*) out, make_float2(0.999f, 0.998f), make_float2(0.001f, 0.002f)); }), 3);
  rep("i2f.s16", time_ms([&] { k_i2f<<<ctas, threads>>>((float*) out, in); }), 1);
  return 0;
}
