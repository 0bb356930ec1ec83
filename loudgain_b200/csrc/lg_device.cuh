// lg_device.cuh -- small device-side helpers shared by the sweep kernels
// (lg_kernels.cu, lg_pair.cu): cp.async staging, sample extraction, register
// pinning.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lg {

// LG_CPASYNC_L2 (tuning): L2 prefetch size hint of the streaming copies, e.g. .L2::128B
#ifndef LG_CPASYNC_L2
#define LG_CPASYNC_L2 ""
#endif
// cache operator of the streaming copies: cg = L2 only, ca = also L1
#ifndef LG_CPASYNC_OP
#define LG_CPASYNC_OP "cg"
#endif
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async." LG_CPASYNC_OP ".shared.global" LG_CPASYNC_L2 " [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- 2-D TMA tile loads (global -> shared, completion on an mbarrier)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// One lane of the (converged) warp: the form the compiler recognises as "a single
// thread issues this", so a tensor copy behind it needs no loop over active lanes.
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 4000;\n"   // suspend up to ~4 us per try
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// box of the tensor map `tmap` (a CUtensorMap in global memory) at element
// coordinates (x, y) -> shared memory at dst (128-byte aligned)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ int sext_half(uint32_t w, uint32_t sel) {
  int r;
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "r"(sel));
  return r;
}

// Shared-memory loads by 32-bit shared address (a pinned generic pointer would
// compile to generic LD instructions).
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Makes a per-lane invariant opaque to the compiler, so that it is kept in a
// register instead of being recomputed (from S2R / parameter loads) inside
// the streaming loop.
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
template <class T>
__device__ __forceinline__ T* pin(T* v) { asm volatile("" : "+l"(v)); return v; }

}  // namespace lg
