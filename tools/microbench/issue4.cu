// issue4.cu -- does the FP64 pipe (and the legacy tensor path) run NEXT TO the packed FP32 stream?
// The 16-bit sweep is bound by FFMA2 dispatch (issue3.cu); its mode sums (2 of 11 FFMA2 per frame)
// can be formed on the RAW samples instead (adjoint of the filter, see DESIGN), and raw 16-bit
// samples turn into doubles without a conversion instruction.  A "frame group" here is NF FFMA2 +
// the peak VIMNMX3 + shl / and + 2 I2FP + one variant of FP64 work; every op is inline asm
// volatile on loop-carried registers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue4 issue4.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 512;
constexpr int kChains = 4;

#define FFMA2(a, m, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(m), "l"(c))
#define I2FP(f, v) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v))
#define SHL16(d, v) asm volatile("shl.b32 %0, %1, 16;" : "=r"(d) : "r"(v))
#define ANDHI(d, v) asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(d) : "r"(v))
#define VMX3(v, a, b) do { v = __vimax3_s16x2(v, a, b); asm volatile("" : "+r"(v)); } while (0)
#define DFMA(a, x, c) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a) : "d"(x), "d"(c))
#define DADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define I2D(d, v) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(v))
#define F2D(d, f) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f))
#define MKD(d, lo, hi) asm volatile("mov.b64 %0, {%1, %2};" : "=d"(d) : "r"(lo), "r"(hi))
// biased forms: flip the sign bit while moving the sample to the top of the word
#define LEAB(d, v) asm volatile("{.reg .b32 t; shl.b32 t, %1, 16; xor.b32 %0, t, 0x80000000;}" : "=r"(d) : "r"(v))
#define ANDX(d, v) asm volatile("lop3.b32 %0, %1, 0xffff0000, 0x80000000, 0x6a;" : "=r"(d) : "r"(v))

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return ((unsigned long long) __float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// D64: 0 none, 1 4 DFMA on loop-carried doubles (no conversion), 2 2 I2F.F64.S32 + 4 DFMA,
// 3 2 F2F.F64.F32 + 4 DFMA, 4 magic pair (register pair, biased words) + 2 DADD + 4 DFMA,
// 5 magic pair + 4 DFMA (bias removed later), 6 4 DFMA + 2 DADD (no conversion)
// 7 one mma.sync m16n8k32 s8 per two frames
template <int D64, int NF, int BASE>
__global__ void __launch_bounds__(128) k(float* out, uint32_t seed, double cz) {
  unsigned long long a2[kChains][4];
  uint32_t w[kChains], mx[kChains];
  double zr[kChains][2], zi[kChains][2];
  int acc[kChains][4];
  const unsigned long long m2 = pack(0.998f, 0.999f), c2 = pack(0.002f, 0.001f);
  const double cr = cz, ci = cz * 0.5, magic = -4503601774854144.0;
  for (int i = 0; i < kChains; ++i) {
    for (int j = 0; j < 4; ++j) { a2[i][j] = threadIdx.x + i + j; acc[i][j] = 0; }
    w[i] = seed * (i + 1 + threadIdx.x); mx[i] = 0;
    zr[i][0] = zr[i][1] = zi[i][0] = zi[i][1] = threadIdx.x;
  }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      float xl = 0.0f, xr = 0.0f;
      uint32_t t0 = 0, t1 = 0;
      if (BASE) {
        if (D64 == 4 || D64 == 5) { LEAB(t0, w[i]); ANDX(t1, w[i]); }
        else { SHL16(t0, w[i]); ANDHI(t1, w[i]); }
        I2FP(xl, t0); I2FP(xr, t1);
        VMX3(mx[i], w[i], seed);
      }
      unsigned long long x2 = pack(xl, xr);
      if (D64 == 1 || D64 == 6) {
        double dl = zr[i][0], dr = zr[i][1];
        if (D64 == 6) { DADD(dl, zi[i][0], magic); DADD(dr, zi[i][1], magic); }
        DFMA(zr[i][0], dl, cr); DFMA(zi[i][0], dl, ci); DFMA(zr[i][1], dr, cr); DFMA(zi[i][1], dr, ci);
      }
      if (D64 == 2) {
        double dl, dr; I2D(dl, t0); I2D(dr, t1);
        DFMA(zr[i][0], dl, cr); DFMA(zi[i][0], dl, ci); DFMA(zr[i][1], dr, cr); DFMA(zi[i][1], dr, ci);
      }
      if (D64 == 3) {
        double dl, dr; F2D(dl, xl); F2D(dr, xr);
        DFMA(zr[i][0], dl, cr); DFMA(zi[i][0], dl, ci); DFMA(zr[i][1], dr, cr); DFMA(zi[i][1], dr, ci);
      }
      if (D64 == 4 || D64 == 5) {
        double dl, dr; MKD(dl, t0, 0x43300000u); MKD(dr, t1, 0x43300000u);
        if (D64 == 4) { DADD(dl, dl, magic); DADD(dr, dr, magic); }
        DFMA(zr[i][0], dl, cr); DFMA(zi[i][0], dl, ci); DFMA(zr[i][1], dr, cr); DFMA(zi[i][1], dr, ci);
      }
      if (D64 == 7 && (it & 1)) {
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3])
                     : "r"(w[i]), "r"(t0), "r"(t1), "r"(seed), "r"(seed), "r"(mx[i]));
      }
      if (NF > 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i][0]) : "l"(m2), "l"(x2));
#pragma unroll
      for (int j = 1; j < NF; ++j) FFMA2(a2[i][j & 3], m2, c2);
      w[i] += (uint32_t) a2[i][0] & 0x10001u;
    }
  }
  float s = 0;
  for (int i = 0; i < kChains; ++i)
    for (int j = 0; j < 4; ++j)
      s += __uint_as_float((uint32_t) a2[i][j]) + __uint_as_float(w[i] ^ mx[i]) +
           (float) (zr[i][j & 1] + zi[i][j & 1]) + (float) acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int D64, int NF, int BASE>
static void run(const char* name, float* out, int sms, double ghz) {
  const int ctas = sms * 4, threads = 128;     // 4 warps per SMSP, as in the sweep
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  k<D64, NF, BASE><<<ctas, threads>>>(out, 12345u, 0.999); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0)); k<D64, NF, BASE><<<ctas, threads>>>(out, 12345u, 0.999); CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1)); float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double frames_per_smsp = 4.0 * kIters * kChains;
  const double clk = best * 1e-3 * ghz * 1e9;
  printf("%-64s %8.3f ms  %6.2f SMSP-cycles per warp-frame\n", name, best, clk / frames_per_smsp);
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, %.3f GHz nominal (cycle counts assume it); 4 warps per SMSP\n", p.name, sms, ghz);
  float* out; CHECK(cudaMalloc(&out, (size_t) sms * 4 * 128 * sizeof(float)));
  run<0, 11, 1>("11 ffma2 + peak + shl/and + 2 i2fp (today)", out, sms, ghz);
  run<0, 9, 1>("9 ffma2 + peak + shl/and + 2 i2fp (lean)", out, sms, ghz);
  run<1, 9, 1>("9 ffma2 + base + 4 dfma", out, sms, ghz);
  run<6, 9, 1>("9 ffma2 + base + 2 dadd + 4 dfma", out, sms, ghz);
  run<2, 9, 1>("9 ffma2 + base + 2 i2f.f64.s32 + 4 dfma", out, sms, ghz);
  run<3, 9, 1>("9 ffma2 + base + 2 f2f.f64.f32 + 4 dfma", out, sms, ghz);
  run<4, 9, 1>("9 ffma2 + base(biased) + magic pair + 2 dadd + 4 dfma", out, sms, ghz);
  run<5, 9, 1>("9 ffma2 + base(biased) + magic pair + 4 dfma", out, sms, ghz);
  run<7, 9, 1>("9 ffma2 + base + 1/2 mma.m16n8k32.s8", out, sms, ghz);
  run<1, 0, 0>("4 dfma alone", out, sms, ghz);
  run<6, 0, 0>("2 dadd + 4 dfma alone", out, sms, ghz);
  run<2, 0, 1>("base + 2 i2f.f64.s32 + 4 dfma, no ffma2", out, sms, ghz);
  run<3, 0, 1>("base + 2 f2f.f64.f32 + 4 dfma, no ffma2", out, sms, ghz);
  run<7, 0, 1>("base + 1/2 mma.m16n8k32.s8, no ffma2", out, sms, ghz);
  return 0;
}
