// core.cu -- compute-ceiling prototypes of the sweep's inner loop, fed from
// shared memory only (no HBM traffic): how many SM lane-clocks one PCM sample
// costs when a lane owns one channel (scalar FFMA) or both channels of a
// stereo frame (packed FFMA2).  Run on the B200 via gpurun; results are
// summarised in profiles/.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Coef {
  float c, ne2, np1, np2, q1, q2;
  float lre[12], lim[12];
  float rr, ri;
  int reps;
};

constexpr int kRowFrames = 96;           // frames per lane row in shared memory (S16 stereo: 384 B)
constexpr int kRowBytes = kRowFrames * 4 + 16;   // odd number of 16-byte units

__device__ __forceinline__ int sext_lo(uint32_t w) { int r; asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w)); return r; }
__device__ __forceinline__ int sext_hi(uint32_t w) { int r; asm("prmt.b32 %0, %1, 0, 0xBB32;" : "=r"(r) : "r"(w)); return r; }

// ---- scalar: one lane = one channel of a row (two lanes share a row)
template <bool XI, bool PEAK>
__global__ void __launch_bounds__(128) k_scalar(float* out, const __grid_constant__ Coef k) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  unsigned char* wsm = sm + wic * 16 * kRowBytes;
  for (int i = lane; i < 16 * kRowBytes / 4; i += 32) ((uint32_t*) wsm)[i] = (i * 2654435761u) >> 3;
  __syncwarp();
  const uint4* row = (const uint4*) (wsm + (lane >> 1) * kRowBytes);
  const uint32_t sel = (lane & 1) ? 0xBB32u : 0x9910u;
  float d1 = 0, w1 = 0, w2 = 0, v1 = 0, v2 = 0, yr = 0, yi = 0, sp = 0;
  double e0 = 0;
  for (int rep = 0; rep < k.reps; ++rep) {
#pragma unroll 1
    for (int it = 0; it < kRowFrames / 12; ++it) {
      float x[12];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const uint4 v = row[it * 3 + u];
        int a, b, c, d;
        asm("prmt.b32 %0, %1, 0, %2;" : "=r"(a) : "r"(v.x), "r"(sel));
        asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(v.y), "r"(sel));
        asm("prmt.b32 %0, %1, 0, %2;" : "=r"(c) : "r"(v.z), "r"(sel));
        asm("prmt.b32 %0, %1, 0, %2;" : "=r"(d) : "r"(v.w), "r"(sel));
        x[4 * u] = (float) a; x[4 * u + 1] = (float) b; x[4 * u + 2] = (float) c; x[4 * u + 3] = (float) d;
      }
      float e = 0, sr = 0, si = 0;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float t = fmaf(k.ne2, w2, x[i]);
        const float d = fmaf(k.c, d1, t);
        const float w = w1 + d;
        const float yh = d - d1;
        const float u = fmaf(k.np2, v2, yh);
        const float v = fmaf(k.np1, v1, u);
        const float y = fmaf(k.q2, v2, fmaf(k.q1, v1, v));
        w2 = w1; w1 = w; d1 = d; v2 = v1; v1 = v;
        e = fmaf(y, y, e);
        if (XI) { sr = fmaf(y, k.lre[i], sr); si = fmaf(y, k.lim[i], si); }
      }
      e0 += (double) e;
      if (XI) {
        const float nr = fmaf(yr, k.rr, fmaf(-yi, k.ri, sr));
        const float ni = fmaf(yr, k.ri, fmaf(yi, k.rr, si));
        yr = nr; yi = ni;
      }
      if (PEAK) {
#pragma unroll
        for (int i = 0; i < 12; i += 2) sp = fmaxf(sp, fmaxf(fabsf(x[i]), fabsf(x[i + 1])));
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float) e0 + yr + yi + sp;
}

// ---- packed: one lane = both channels of a row
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

template <bool XI, bool PEAK>
__global__ void __launch_bounds__(128) k_packed(float* out, const __grid_constant__ Coef k) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  unsigned char* wsm = sm + wic * 32 * kRowBytes;
  for (int i = lane; i < 32 * kRowBytes / 4; i += 32) ((uint32_t*) wsm)[i] = (i * 2654435761u) >> 3;
  __syncwarp();
  const uint4* row = (const uint4*) (wsm + lane * kRowBytes);
  float2 d1 = bc(0), w1 = bc(0), w2 = bc(0), v1 = bc(0), v2 = bc(0), yr = bc(0), yi = bc(0);
  uint32_t mx = 0x80008000u, mn = 0x7fff7fffu;
  double e0a = 0, e0b = 0;
  for (int rep = 0; rep < k.reps; ++rep) {
#pragma unroll 1
    for (int it = 0; it < kRowFrames / 12; ++it) {
      uint32_t wd[12];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const uint4 v = row[it * 3 + u];
        wd[4 * u] = v.x; wd[4 * u + 1] = v.y; wd[4 * u + 2] = v.z; wd[4 * u + 3] = v.w;
      }
      float2 e = bc(0), sr = bc(0), si = bc(0);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float2 x = make_float2((float) sext_lo(wd[i]), (float) sext_hi(wd[i]));
        const float2 t = ffma2(bc(k.ne2), w2, x);
        const float2 d = ffma2(bc(k.c), d1, t);
        const float2 w = __fadd2_rn(w1, d);
        const float2 yh = __fadd2_rn(d, make_float2(-d1.x, -d1.y));
        const float2 u = ffma2(bc(k.np2), v2, yh);
        const float2 v = ffma2(bc(k.np1), v1, u);
        const float2 y = ffma2(bc(k.q2), v2, ffma2(bc(k.q1), v1, v));
        w2 = w1; w1 = w; d1 = d; v2 = v1; v1 = v;
        e = ffma2(y, y, e);
        if (XI) { sr = ffma2(y, bc(k.lre[i]), sr); si = ffma2(y, bc(k.lim[i]), si); }
      }
      e0a += (double) e.x; e0b += (double) e.y;
      if (XI) {
        const float2 nr = ffma2(yr, bc(k.rr), ffma2(yi, bc(-k.ri), sr));
        const float2 ni = ffma2(yr, bc(k.ri), ffma2(yi, bc(k.rr), si));
        yr = nr; yi = ni;
      }
      if (PEAK) {
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          mx = __vimax3_s16x2(mx, wd[i], wd[i + 1]);
          mn = __vimin3_s16x2(mn, wd[i], wd[i + 1]);
        }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float) (e0a + e0b) + yr.x + yi.y + (float) (mx ^ mn);
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
  const int sms = p.multiProcessorCount;
  float* out; CHECK(cudaMalloc(&out, (size_t) sms * 16 * 128 * 4));
  Coef k;
  k.c = 0.9891f; k.ne2 = -2.9e-5f; k.np1 = 1.6906f; k.np2 = -0.7325f; k.q1 = -1.75f; k.q2 = 0.78f;
  for (int i = 0; i < 12; ++i) { k.lre[i] = 0.99f - 0.005f * i; k.lim[i] = 0.005f * i; }
  k.rr = 1.05f; k.ri = -0.06f; k.reps = 200;
  printf("device %s, %d SMs\n", p.name, sms);
  auto rep = [&](const char* name, int cps, double ms, double samples_per_lane) {
    const double lanes = (double) sms * cps * 128;
    const double samples = lanes * samples_per_lane * k.reps;
    const double clk = ms * 1e-3 * 1.965e9 * sms * 128;
    printf("%-28s ctas/SM=%d  %8.3f ms  %8.1f Gsamples/s  %6.2f lane-clk/sample\n", name, cps, ms,
           samples / ms / 1e6, clk / samples);
  };
#define RUN_S(XI, PK, CPS) { \
    CHECK(cudaFuncSetAttribute(k_scalar<XI, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * kRowBytes)); \
    rep("scalar xi=" #XI " peak=" #PK, CPS, time_ms([&] { k_scalar<XI, PK><<<sms * CPS, 128, 4 * 16 * kRowBytes>>>(out, k); }), kRowFrames); }
#define RUN_P(XI, PK, CPS) { \
    CHECK(cudaFuncSetAttribute(k_packed<XI, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32 * kRowBytes)); \
    rep("packed xi=" #XI " peak=" #PK, CPS, time_ms([&] { k_packed<XI, PK><<<sms * CPS, 128, 4 * 32 * kRowBytes>>>(out, k); }), 2 * kRowFrames); }
  RUN_S(true, true, 2) RUN_S(true, true, 4) RUN_S(true, true, 8) RUN_S(true, true, 12)
  RUN_S(false, true, 8) RUN_S(true, false, 8) RUN_S(false, false, 8)
  RUN_P(true, true, 1) RUN_P(true, true, 2) RUN_P(true, true, 3) RUN_P(true, true, 4)
  RUN_P(false, true, 4) RUN_P(true, false, 4) RUN_P(false, false, 4)
  return 0;
}
