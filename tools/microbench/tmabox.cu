// tmabox.cu -- how fast does a B200 stream HBM through 2-D tensor copies when every
// lane of a warp owns its own row stream (the access pattern of the run sweep,
// loudgain_b200/csrc/lg_run.cu)?  A "track" of 1 GiB is viewed as rows of `pitch`
// bytes; a warp owns 32 consecutive rows and pulls `piece` bytes of each per stage,
// as `nbox` tensor copies of 32 / nbox rows, `ring` stages deep, `warps` warps per
// persistent CTA (one CTA per SM), no compute.  Also: one cp.async.bulk per row.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmabox tmabox.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mbar_init(uint32_t m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t m, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t d, const void* s, uint32_t b, uint32_t m) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(s), "r"(b), "r"(m) : "memory"); }
__device__ __forceinline__ void tma2d(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t m, uint32_t p) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 4000;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(m), "r"(p) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
  return p != 0;
}

struct Params {
  const unsigned char* src;
  const CUtensorMap* maps;     // [nbox] (same view, box of 32 / nbox rows)
  uint32_t pitch, piece, nbox, ring, nstages, nitems, mode;   // mode 0: tensor boxes, 1: bulk copy per row
  unsigned long long* sink;
};

__global__ void __launch_bounds__(512, 1) k_box(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char sm_all[];
  const uint32_t wic = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t stage_bytes = 32 * P.piece;
  const uint32_t warp_bytes = (P.ring * stage_bytes + 64 + 127) & ~127u;
  const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sm_all) + wic * warp_bytes;
  const uint32_t mb = sa + P.ring * stage_bytes;
  if (lane == 0) for (uint32_t i = 0; i < P.ring; ++i) mbar_init(mb + 8 * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  unsigned long long acc = 0;
  for (uint32_t item = blockIdx.x * nw + wic; item < P.nitems; item += gridDim.x * nw) {
    auto issue = [&](uint32_t s, uint32_t slot) {
      const uint32_t bar = mb + 8 * slot, dst = sa + slot * stage_bytes;
      if (P.mode == 0) {
        if (elect_one()) {
          mbar_expect(bar, stage_bytes);
          const uint32_t rows = 32 / P.nbox;
          for (uint32_t b = 0; b < P.nbox; ++b)
            tma2d(dst + b * rows * P.piece, P.maps + b * 0 + (P.nbox == 1 ? 0 : P.nbox == 2 ? 1 : 2), (int) (s * P.piece / 4),
                  (int) (item * 32 + b * rows), bar);
        }
      } else {
        if (lane == 0) mbar_expect(bar, stage_bytes);
        __syncwarp();
        bulk(dst + lane * P.piece, P.src + (size_t) (item * 32 + lane) * P.pitch + (size_t) s * P.piece, P.piece, bar);
      }
    };
    for (uint32_t i = 0; i + 1 < P.ring && i < P.nstages; ++i) issue(i, i);
    uint32_t slot = 0;
    for (uint32_t s = 0; s < P.nstages; ++s) {
      mbar_wait(mb + 8 * slot, (phase >> slot) & 1u);
      phase ^= 1u << slot;
      __syncwarp();
      const uint32_t ps = s + P.ring - 1;
      if (ps < P.nstages) issue(ps, slot == 0 ? P.ring - 1 : slot - 1);
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa + slot * stage_bytes + lane * P.piece));
      acc += v;
      if (++slot == P.ring) slot = 0;
    }
    __syncwarp();
  }
  if (acc == 0x1234567) P.sink[0] = acc;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const size_t total = (size_t) 1 << 30;
  unsigned char* src; CHECK(cudaMalloc(&src, total + 4096)); CHECK(cudaMemset(src, 1, total + 4096));
  unsigned long long* sink; CHECK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  PFN_cuTensorMapEncodeTiled encode = (PFN_cuTensorMapEncodeTiled) fn;
  CUtensorMap* dmaps; CHECK(cudaMalloc(&dmaps, 3 * sizeof(CUtensorMap)));
  CHECK(cudaFuncSetAttribute(k_box, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  printf("%-64s %8s %10s\n", "pattern", "ms", "GB/s");
  struct Cfg { uint32_t pitch, piece, nbox, ring, warps, mode, misalign; };
  const Cfg cfgs[] = {
      // pitch 14112 = 3528 stereo S16 frames (16-byte aligned rows only)
      {14112, 208, 1, 3, 8, 0, 0},  {14112, 208, 1, 3, 10, 0, 0}, {14112, 208, 2, 3, 10, 0, 0}, {14112, 208, 4, 3, 10, 0, 0},
      {14112, 208, 1, 2, 16, 0, 0}, {14112, 208, 4, 2, 16, 0, 0}, {14112, 192, 1, 3, 10, 0, 0}, {14112, 192, 4, 3, 10, 0, 0},
      {14112, 400, 1, 2, 8, 0, 0},  {14112, 400, 2, 2, 8, 0, 0},  {14112, 400, 4, 2, 8, 0, 0},  {14112, 384, 1, 2, 8, 0, 0},
      {14112, 400, 1, 3, 5, 0, 0},  {14112, 784, 1, 2, 4, 0, 0},  {14112, 784, 4, 2, 4, 0, 0},
      // rows on 128-byte boundaries: pitch 14208 = 111 lines
      {14208, 384, 1, 2, 8, 0, 0},  {14208, 256, 1, 3, 9, 0, 0},  {14208, 256, 1, 2, 14, 0, 0}, {14208, 128, 1, 3, 16, 0, 0},
      {14208, 384, 1, 2, 8, 0, 64}, {14208, 208, 1, 3, 10, 0, 0},
      // one bulk copy per row
      {14112, 208, 1, 3, 10, 1, 0}, {14112, 400, 1, 2, 8, 1, 0},  {14208, 384, 1, 2, 8, 1, 0},  {14112, 208, 1, 2, 16, 1, 0},
  };
  for (const Cfg& c : cfgs) {
    const uint32_t nrows = (uint32_t) (total / c.pitch) / 32 * 32;
    const unsigned char* base = src + c.misalign;
    CUtensorMap maps[3];
    const uint32_t rowsv[3] = {32, 16, 8};
    for (int i = 0; i < 3; ++i) {
      const cuuint64_t dims[2] = {c.pitch / 4, nrows};
      const cuuint64_t strides[1] = {c.pitch};
      const cuuint32_t box[2] = {c.piece / 4, rowsv[i]};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult rc = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*) base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int) rc); return 1; }
    }
    CHECK(cudaMemcpy(dmaps, maps, sizeof maps, cudaMemcpyHostToDevice));
    Params P;
    P.src = base; P.maps = dmaps; P.pitch = c.pitch; P.piece = c.piece; P.nbox = c.nbox; P.ring = c.ring;
    P.nstages = c.pitch / c.piece; P.nitems = nrows / 32; P.mode = c.mode; P.sink = sink;
    const size_t smem = (size_t) c.warps * ((c.ring * 32 * c.piece + 64 + 127) & ~127u);
    if (smem > 226 * 1024) { printf("skip (smem %zu)\n", smem); continue; }
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      CHECK(cudaEventRecord(e0));
      k_box<<<sms, c.warps * 32, smem>>>(P);
      CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1)); CHECK(cudaGetLastError());
      float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    char name[96];
    snprintf(name, sizeof name, "%s pitch=%u piece=%u boxes=%u ring=%u warps=%u off=%u", c.mode ? "bulk/row" : "tensor2d",
             c.pitch, c.piece, c.nbox, c.ring, c.warps, c.misalign);
    printf("%-64s %8.3f %10.1f\n", name, best, (double) P.nitems * 32 * P.nstages * c.piece / best / 1e6);
  }
  return 0;
}
