// rowstream.cu -- how fast can the B200 stream HBM when every warp pulls
// `rows` separate row streams, `piece` bytes per row and step, rows `stride`
// bytes apart (the access pattern of a sweep whose lanes own long time chunks)?
// Compared with one contiguous stream per warp.  TMA bulk copies into shared
// memory, 3-deep ring, no compute.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mbar_init(uint32_t m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t m, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t d, const void* s, uint32_t b, uint32_t m) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(s), "r"(b), "r"(m) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t m, uint32_t p) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(m), "r"(p) : "memory");
}

// Each warp owns `rows` rows of `row_len` bytes, consecutive rows `row_len` apart (so a warp's
// region is contiguous: rows * row_len), and reads them piece by piece, all rows in lock step.
__global__ void k_rows(const unsigned char* src, uint32_t rows, uint32_t piece, uint32_t row_len,
                       uint32_t nwarps, unsigned long long* sink) {
  extern __shared__ __align__(16) unsigned char sm_all[];
  const uint32_t wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wic;
  if (warp >= nwarps) return;
  const uint32_t stage_bytes = rows * piece;
  unsigned char* sm = sm_all + wic * (3 * stage_bytes + 64);
  const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sm);
  const uint32_t mb = sa + 3 * stage_bytes;
  if (lane == 0) for (int i = 0; i < 3; ++i) mbar_init(mb + 8 * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const unsigned char* base = src + (size_t) warp * rows * row_len + (size_t) lane * row_len;
  const uint32_t nst = row_len / piece;
  auto pre = [&](uint32_t s) {
    const uint32_t b = s % 3;
    if (lane == 0) mbar_expect(mb + 8 * b, stage_bytes);
    __syncwarp();
    if (lane < rows) bulk(sa + b * stage_bytes + lane * piece, base + (size_t) s * piece, piece, mb + 8 * b);
  };
  pre(0); if (nst > 1) pre(1);
  unsigned long long acc = 0;
  for (uint32_t s = 0; s < nst; ++s) {
    mbar_wait(mb + 8 * (s % 3), (s / 3) & 1);
    __syncwarp();
    if (s + 2 < nst) pre(s + 2);
    acc += *reinterpret_cast<const unsigned long long*>(sm + (s % 3) * stage_bytes + (lane * 8) % stage_bytes);
  }
  if (acc == 0x1234567) sink[0] = acc;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  const size_t total = (size_t) 1 << 30;           // 1 GiB
  unsigned char* src; CHECK(cudaMalloc(&src, total)); CHECK(cudaMemset(src, 1, total));
  unsigned long long* sink; CHECK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  printf("%-28s %8s %10s\n", "pattern", "ms", "GB/s");
  const uint32_t row_len = 3584;                   // ~ one 882-frame stereo S16 chunk
  struct Cfg { uint32_t rows, piece, wpb; } cfgs[] = {
      {16, 96 + 16, 4}, {16, 192 + 32, 4}, {16, 448, 4}, {16, 896, 4}, {16, 1792, 2},
      {8, 448, 4}, {8, 896, 4}, {8, 1792, 4}, {4, 1792, 4}, {4, 3584, 4}, {1, 3584, 8}};
  for (auto c : cfgs) {
    // make piece divide row_len
    uint32_t piece = c.piece; while (row_len % piece) piece -= 16;
    const uint32_t nwarps = (uint32_t) (total / ((size_t) c.rows * row_len));
    const uint32_t blocks = (nwarps + c.wpb - 1) / c.wpb;
    const size_t smem = (size_t) c.wpb * (3 * c.rows * piece + 64);
    CHECK(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      CHECK(cudaEventRecord(e0));
      k_rows<<<blocks, c.wpb * 32, smem>>>(src, c.rows, piece, row_len, nwarps, sink);
      CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1)); CHECK(cudaGetLastError());
      float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    char name[64]; snprintf(name, sizeof name, "rows=%u piece=%u wpb=%u", c.rows, piece, c.wpb);
    printf("%-28s %8.3f %10.1f\n", name, best, (double) nwarps * c.rows * row_len / best / 1e6);
  }
  return 0;
}
