// lg_kernels.cu -- sm_100a kernels of the loudness path and their launchers.
//
//   sweep_kernel    fused K-weight + chunk energy + sample peak + true peak
//                   (what ebur128_add_frames_short does per call in the
//                   reference path, /root/reference/src/scan.c:448)
//   fixslot_kernel  FP64 state carry + energy correction + channel weighting: 100 ms slot energies
//   block_kernel    400 ms gating blocks and 3 s short-term blocks
//   query_kernel    gated integrated loudness and loudness range over a set
//                   of tracks (ebur128_loudness_global[_multiple],
//                   ebur128_loudness_range[_multiple]: scan.c:294,297,383,388)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "lg_common.h"
#include "lg_device.cuh"
#include "lg_kernels.h"
#include "lg_post.cuh"
#include "lg_sweep.cuh"

namespace lg {

// ------------------------------------------------------------------ sweep
//
// One warp = up to 32/min(C,32) consecutive chunks of one track; one lane =
// one (chunk, channel).  The warp is autonomous (no CTA barrier):
//
//  * it streams its rows HBM -> shared memory with 16-byte cp.async copies laid
//    out so that consecutive lanes fetch consecutive units of one row
//    (coalesced), kRing stages deep; rows start on 16-byte boundaries of the
//    track and the row stride is an odd number of units, so 128-bit shared
//    loads of different rows do not collide;
//  * every lane filters its channel (FP32) and accumulates energy, the
//    correction cross terms and the sample peak; the interior of a chunk runs
//    as straight-line code over 24 frames, the few warm-up / edge iterations
//    go through a rolled loop;
//  * true peak is NOT evaluated here: the lane only records max |x| of each
//    12-frame iteration (16-bit code, one coalesced 128-byte store per warp
//    and pair).  truepeak_kernel evaluates the polyphase FIR afterwards, only
//    on the iterations whose bound ||c||_1 * max|x| exceeds the channel's
//    final sample peak.

// How a lane finds its samples in a staged row.
enum RowLayout : int {
  ROW_S16_STEREO = 0,   // frame = one 32-bit word (L | R << 16)
  ROW_S16_MONO = 1,     // two frames per word
  ROW_S16_ANY = 2,      // 16-bit loads, stride = frame bytes
  ROW_F32_MONO = 3,
  ROW_F32_STEREO = 4,
  ROW_F32_ANY = 5,
};

// Raw samples of one iteration of the lane's channel, from shared memory.
// rowp -> first frame of the iteration in the lane's row (16-byte aligned for
// the vector layouts: iterations are 12 frames).
template <int LAYOUT>
__device__ __forceinline__ void smem_load_iter(const unsigned char* rowp, uint32_t fb, uint32_t ch,
                                               float* x) {
  if (LAYOUT == ROW_S16_STEREO) {
    // One PRMT sign-extends the lane's half of a frame word (selector nibble
    // bit 3 = replicate the sign of that byte).
    const uint4* p = reinterpret_cast<const uint4*>(rowp);
    const uint32_t sel = ch ? 0xBB32u : 0x9910u;
#pragma unroll
    for (int u = 0; u < kIter / 4; ++u) {
      const uint4 v = p[u];
      x[4 * u + 0] = (float) sext_half(v.x, sel);
      x[4 * u + 1] = (float) sext_half(v.y, sel);
      x[4 * u + 2] = (float) sext_half(v.z, sel);
      x[4 * u + 3] = (float) sext_half(v.w, sel);
    }
  } else if (LAYOUT == ROW_S16_MONO) {
    // 12 frames = 24 bytes: an 8-byte-aligned window of three 64-bit loads
    const uint2* p = reinterpret_cast<const uint2*>(rowp);
#pragma unroll
    for (int u = 0; u < kIter / 4; ++u) {
      const uint2 v = p[u];
      x[4 * u + 0] = (float) sext_half(v.x, 0x9910u);
      x[4 * u + 1] = (float) sext_half(v.x, 0xBB32u);
      x[4 * u + 2] = (float) sext_half(v.y, 0x9910u);
      x[4 * u + 3] = (float) sext_half(v.y, 0xBB32u);
    }
  } else if (LAYOUT == ROW_S16_ANY) {
    const unsigned char* q = rowp + ch * 2u;
#pragma unroll
    for (int i = 0; i < kIter; ++i) x[i] = (float) (int) *reinterpret_cast<const short*>(q + i * fb);
  } else if (LAYOUT == ROW_F32_MONO) {
    const float4* p = reinterpret_cast<const float4*>(rowp);
#pragma unroll
    for (int u = 0; u < kIter / 4; ++u) {
      const float4 v = p[u];
      x[4 * u] = v.x; x[4 * u + 1] = v.y; x[4 * u + 2] = v.z; x[4 * u + 3] = v.w;
    }
  } else if (LAYOUT == ROW_F32_STEREO) {
    const float4* p = reinterpret_cast<const float4*>(rowp);
#pragma unroll
    for (int u = 0; u < kIter / 2; ++u) {
      const float4 v = p[u];
      x[2 * u] = ch ? v.y : v.x;
      x[2 * u + 1] = ch ? v.w : v.z;
    }
  } else {
    const unsigned char* q = rowp + ch * 4u;
#pragma unroll
    for (int i = 0; i < kIter; ++i) x[i] = *reinterpret_cast<const float*>(q + i * fb);
  }
}

#ifndef LG_SWEEP_MINBLOCKS
#define LG_SWEEP_MINBLOCKS 8
#endif

// Lanes per chunk row when the layout fixes it (0 = run-time, P.lpc).
template <int LAYOUT>
__host__ __device__ constexpr uint32_t layout_lpc() {
  return (LAYOUT == ROW_S16_STEREO || LAYOUT == ROW_F32_STEREO) ? 2u
       : (LAYOUT == ROW_S16_MONO || LAYOUT == ROW_F32_MONO) ? 1u : 0u;
}

// KCOPY = 16-byte copies a lane issues per stage (0 = run-time count, for
// more than 32 channels).
template <int LAYOUT, bool TP, int KCOPY>
__global__ void __launch_bounds__(kSweepThreads, LG_SWEEP_MINBLOCKS)
sweep_kernel(const __grid_constant__ SweepParams P) {
  extern __shared__ __align__(16) unsigned char smem_all[];
  const uint32_t wic = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wic;
  if (warp >= P.nwarps) return;                // whole warps leave; no CTA barrier below
  unsigned char* sm = smem_all + wic * P.warp_smem;

  const WarpWork ww = P.work[warp];
  const Track& tr = P.tracks[ww.track];
  const int W = P.W, L = P.L;
  const uint32_t C = P.channels, fb = P.fb;
  const long long frames = (long long) tr.frames;
  const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm);

  constexpr uint32_t LPC = layout_lpc<LAYOUT>();
  const uint32_t lpc = LPC ? LPC : P.lpc;
  const uint32_t slot = lane / lpc;
  const uint32_t chl = lane - slot * lpc;      // channel within the warp's group
  const uint32_t ch = ww.ch_base + chl;
  const uint32_t chunk = ww.first_chunk + slot;
  const bool compute = slot < P.cpw && ch < C;
  const bool active = compute && chunk < tr.nchunks;

  // ---- lane state
  const LaneGeom geo = lane_geometry(frames, L, W, P.aq, chunk);
  LaneCtx c;
  lane_init(c, W, L, geo);
  const unsigned char* my_row = pin(sm + slot * P.row_stride);

  // ---- staging.  Per stage every row receives one contiguous piece of
  // kStageFrames frames, moved with 16-byte cp.async copies.  The lpc lanes of
  // a row copy its units interleaved (lane chl takes units chl, chl + lpc, ...),
  // so consecutive lanes fetch consecutive 16-byte units, and a lane only ever
  // needs its own row's address.  Warps at a track boundary use the
  // zero-filling form.
  const uint32_t ustride = lpc << 4;                            // bytes between a lane's copies
  const long long row_byte0 = geo.a * (long long) fb + (chl << 4);   // first copy, may be negative
  const long long track_bytes = frames * (long long) fb;
  const bool interior = ww.interior != 0;
  const bool copier = slot < P.cpw;
  const uint32_t ncopy = KCOPY ? (uint32_t) KCOPY : P.kcopies;
  const uint32_t dst_row =
      pin((uint32_t) __cvta_generic_to_shared(sm) + slot * P.row_stride + (chl << 4));
  const unsigned char* src = pcm + row_byte0;                   // advanced stage by stage
  uint32_t pf_off = 0;                                          // ring offset of the next prefetch

  auto prefetch = [&]() {
    const uint32_t dst = dst_row + pf_off;
    if (copier) {
      if (interior) {
        if (KCOPY) {
#pragma unroll
          for (int k = 0; k < (KCOPY ? KCOPY : 1); ++k) cp_async16(dst + k * ustride, src + k * ustride);
        } else {
          for (uint32_t k = 0; k < ncopy; ++k)
            if (chl + k * lpc < P.units) cp_async16(dst + k * ustride, src + k * ustride);
        }
      } else {
#pragma unroll 1
        for (uint32_t k = 0; k < ncopy; ++k) {
          if (chl + k * lpc >= P.units) break;
          const long long g = (src - pcm) + (long long) (k * ustride);
          long long ok = g < 0 ? 0 : track_bytes - g;
          ok = ok < 0 ? 0 : (ok > 16 ? 16 : ok);
          cp_async16_zfill(dst + k * ustride, pcm + (ok ? g : 0), (uint32_t) ok);
        }
      }
    }
    src += P.stage_row_bytes;
    pf_off += P.stage_bytes;
    if (pf_off == P.ring_bytes) pf_off = 0;
  };

  uint32_t* mrec = pin(P.mrec + (size_t) warp * P.npairs * 32u + lane);   // advanced pair by pair

  const uint32_t niters = (uint32_t) P.niters;
  const uint32_t npairs = P.npairs;
  const uint32_t nstages = (npairs + kPairsPerStage - 1) / kPairsPerStage;
  // Pairs whose two iterations are both "fast" for every lane of the warp.
  const int lfast = ww.lmin_valid < L ? ww.lmin_valid : L;
  const uint32_t fast_lo = (uint32_t) ((W + P.aq - 1 + kPairFrames - 1) / kPairFrames);
  const uint32_t fast_hi = (uint32_t) ((W + lfast) / kPairFrames);     // exclusive
#pragma unroll
  for (int i = 0; i < kRing - 1; ++i) {
    if ((uint32_t) i < nstages) prefetch();
    cp_async_commit();
  }

  uint32_t cs_off = 0;                                          // ring offset of the stage being consumed
  uint32_t pair = 0;
  for (uint32_t s = 0; s < nstages; ++s) {
    cp_async_wait<kRing - 2>();
    __syncwarp();                // everyone's data has landed; the previous stage is consumed
    if (s + kRing - 1 < nstages) prefetch();
    cp_async_commit();
    const unsigned char* sbuf = my_row + cs_off;
    cs_off += P.stage_bytes;
    if (cs_off == P.ring_bytes) cs_off = 0;
#pragma unroll 1
    for (uint32_t pr = 0; pr < (uint32_t) kPairsPerStage; ++pr, ++pair) {
      if (pair >= npairs) break;
      const unsigned char* buf = sbuf + pr * kPairFrames * fb;
      float m0 = 0.0f, m1 = 0.0f;
      if (pair >= fast_lo && pair < fast_hi) {
        // ---- interior of the chunk: straight-line code, loads first
        if (compute) {
          float x0[kIter], x1[kIter];
          smem_load_iter<LAYOUT>(buf, fb, ch, x0);
          smem_load_iter<LAYOUT>(buf + kIter * fb, fb, ch, x1);
          const int f0 = (int) (pair * kPairFrames);
          m0 = iter_fast(c, P, x0, f0);
          m1 = iter_fast(c, P, x1, f0 + kIter);
        }
      } else if (compute) {
        // ---- chunk edges: warm-up and masked iterations, one at a time
#pragma unroll 1
        for (uint32_t it = 0; it < 2u; ++it) {
          const uint32_t iter = pair * 2u + it;
          if (iter >= niters) break;
          const int f0 = (int) (iter * kIter);
          float x[kIter];
          smem_load_iter<LAYOUT>(buf + it * kIter * fb, fb, ch, x);
          const float m = f0 + kIter <= W ? iter_warm(c, P, x, iter == 0) : iter_masked(c, P, x, f0);
          if (it == 0) m0 = m; else m1 = m;
        }
      }
      if (TP) { *mrec = peak_code(m0) | (peak_code(m1) << 16); mrec += 32; }
    }
  }
  cp_async_wait<0>();

  if (active) {
    ChunkRec v;
    v.e0 = c.e0; v.yr = c.yr; v.yi = c.yi;
    v.pd = c.pd; v.pw = c.pw; v.qd = c.qd; v.qw = c.qw;
    P.recs[tr.rec_base + (uint64_t) chunk * C + ch] = v;
  }
  // Sample peak: non-negative floats order like their bit patterns.  Reduce
  // over the lanes of the warp that hold the same channel, one atomic each.
  const unsigned peers = __match_any_sync(0xffffffffu, compute ? chl : 0xffffu);
  const uint32_t spb = __reduce_max_sync(peers, active ? __float_as_uint(c.sp) : 0u);
  if (compute && lane == (uint32_t) (__ffs(peers) - 1))
    atomicMax(P.peaks + 2 * (tr.peak_base + ch), spb);
}

template <int LAYOUT, bool TP, int KCOPY>
static cudaError_t launch_sweep_k(const SweepParams& p, cudaStream_t stream) {
  const uint32_t wpb = kSweepThreads / 32;
  const uint32_t blocks = (p.nwarps + wpb - 1) / wpb;
  const size_t smem = (size_t) p.warp_smem * wpb;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<LAYOUT, TP, KCOPY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sweep_kernel<LAYOUT, TP, KCOPY>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sweep_kernel<LAYOUT, TP, KCOPY><<<blocks, kSweepThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// Copies per lane and stage: units / lpc = 3 per pair for 16-bit and 6 per
// pair for float frames as long as a row's channels fit one warp.
template <int LAYOUT, bool TP>
static cudaError_t launch_sweep_t(const SweepParams& p, cudaStream_t stream) {
  constexpr bool f32 = LAYOUT >= ROW_F32_MONO;
  constexpr int K = (f32 ? 6 : 3) * kPairsPerStage;
  if (p.kcopies == (uint32_t) K && p.units == p.kcopies * p.lpc)
    return launch_sweep_k<LAYOUT, TP, K>(p, stream);
  return launch_sweep_k<LAYOUT, TP, 0>(p, stream);
}

template <int LAYOUT>
static cudaError_t launch_sweep_l(const SweepParams& p, int tpf, cudaStream_t stream) {
  return tpf ? launch_sweep_t<LAYOUT, true>(p, stream) : launch_sweep_t<LAYOUT, false>(p, stream);
}

cudaError_t launch_sweep(const SweepParams& p, uint32_t format, int tpf, cudaStream_t stream) {
  if (p.nwarps == 0) return cudaSuccess;
  if (format == FMT_S16) {
    if (p.channels == 2) return launch_sweep_l<ROW_S16_STEREO>(p, tpf, stream);
    if (p.channels == 1) return launch_sweep_l<ROW_S16_MONO>(p, tpf, stream);
    return launch_sweep_l<ROW_S16_ANY>(p, tpf, stream);
  }
  if (p.channels == 2) return launch_sweep_l<ROW_F32_STEREO>(p, tpf, stream);
  if (p.channels == 1) return launch_sweep_l<ROW_F32_MONO>(p, tpf, stream);
  return launch_sweep_l<ROW_F32_ANY>(p, tpf, stream);
}

// -------------------------------------------------------------- true peak
//
// Second pass over the iteration maxima the sweep left behind.  The true peak
// is a maximum, and a polyphase output cannot exceed ||c||_1 * max|x| over
// its taps' window; the channel's true peak is at least its sample peak
// (which is final by now).  So only iterations whose bound exceeds the
// channel's current peak can matter -- a few percent of the audio on
// programme material.  CTAs scan tiles of the code array, collect those
// iterations in shared memory and evaluate them densely, one iteration per
// thread, re-reading their 24 (36) frames from the PCM (L2 / HBM).  The
// result is identical to evaluating every frame.

constexpr int kTpThreads = 128;
constexpr int kTpWarps = kTpThreads / 32;
constexpr int kTpBatch = 4;                             // pairs a lane has in flight
constexpr int kTpCap = 32 + 2 * kTpBatch * 32;          // per-warp queue: remainder + one batch

template <int FMT>
__device__ __forceinline__ float pcm_sample(const unsigned char* p) {
  if (FMT == FMT_S16) return (float) (int) *reinterpret_cast<const short*>(p);
  return *reinterpret_cast<const float*>(p);
}

// A sweep lane's place in its track.
struct TpLane {
  uint32_t track, ch;
  long long a;       // track frame of lane-local frame 0
  int f_lo, f_end;   // lane-local frames the true-peak pass owns: [f_lo, f_end)
  long long lead_in; // track frames of leading context (no true-peak output inside)
  bool ok;
};

__device__ __forceinline__ TpLane tp_locate(const SweepParams& P, uint32_t w, uint32_t lane) {
  const WarpWork ww = P.work[w];
  const Track& tr = P.tracks[ww.track];
  TpLane r;
  r.track = ww.track;
  const uint32_t slot = lane / P.lpc;
  r.ch = ww.ch_base + (lane - slot * P.lpc);
  const uint32_t chunk = ww.first_chunk + slot;
  r.ok = slot < P.cpw && r.ch < P.channels && chunk < tr.nchunks;
  const LaneGeom g = lane_geometry((long long) tr.frames, P.L, P.W, P.aq, chunk);
  r.a = g.a;
  r.lead_in = (long long) tr.lead_in;
  r.f_lo = P.W + g.o;
  const long long left = (long long) tr.frames - g.a;      // frames of the track from local 0
  const int f_hi = r.f_lo + P.L;
  r.f_end = left < (long long) f_hi ? (left < 0 ? 0 : (int) left) : f_hi;
  return r;
}

// One queued iteration: (track, channel, track frame of its first frame).
__device__ __forceinline__ uint4 tp_entry(uint32_t track, uint32_t ch, long long t0) {
  return make_uint4(track, ch, (uint32_t) (unsigned long long) t0,
                    (uint32_t) ((unsigned long long) t0 >> 32));
}

template <int FMT, int TPF>
__device__ __forceinline__ void tp_evaluate(const SweepParams& P, const uint4 cd) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  const Track& tr = P.tracks[cd.x];
  const long long frames = (long long) tr.frames;
  const long long t0 = (long long) ((unsigned long long) cd.z | ((unsigned long long) cd.w << 32));
  const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm) +
                             cd.y * (FMT == FMT_S16 ? 2u : 4u);
  float win[NT + kIter];
  float m;
  if (t0 >= NT && t0 + kIter <= frames) {
    const unsigned char* q = pcm + (t0 - NT) * (long long) P.fb;
#pragma unroll
    for (int k = 0; k < NT + kIter; ++k) win[k] = pcm_sample<FMT>(q + (uint32_t) k * P.fb);
    m = tp_window_valid<TPF>(win, kIter);
  } else {
#pragma unroll
    for (int k = 0; k < NT + kIter; ++k) {
      const long long t = t0 - NT + k;
      win[k] = (t >= 0 && t < frames) ? pcm_sample<FMT>(pcm + t * (long long) P.fb) : 0.0f;
    }
    const long long left = frames - t0;
    m = tp_window_valid<TPF>(win, left > kIter ? kIter : (int) left);
  }
  uint32_t* cell = P.peaks + 2 * (tr.peak_base + cd.y) + 1;
  if (__float_as_uint(m) > __ldcg(cell)) atomicMax(cell, __float_as_uint(m));
}

// Warps are autonomous.  A work item is kTpSegment consecutive pairs of one
// sweep warp; scanning lane = sweep lane, so the lane's geometry and its
// channel's current peak are loaded once per item and the previous pair's
// codes stay in a register.  Iterations that can still matter go into the
// warp's queue and are evaluated 32 at a time, one per lane.
template <int FMT, int TPF>
__global__ void __launch_bounds__(kTpThreads, 5)
truepeak_kernel(const __grid_constant__ SweepParams P, const uint32_t seg_pairs) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  __shared__ uint4 queue_all[kTpWarps][kTpCap];
  const uint32_t npairs = P.npairs;
  const uint32_t lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
  uint4* queue = queue_all[wic];
  const uint32_t nseg = (npairs + seg_pairs - 1) / seg_pairs;
  const uint64_t nitems = (uint64_t) P.nwarps * nseg;
  const uint64_t nscan = (uint64_t) gridDim.x * kTpWarps;
  uint32_t qn = 0;                       // warp-uniform

  for (uint64_t item = (uint64_t) blockIdx.x * kTpWarps + wic; item < nitems; item += nscan) {
    const uint32_t w = (uint32_t) (item / nseg);
    const uint32_t p_begin = (uint32_t) (item - (uint64_t) w * nseg) * seg_pairs;
    const uint32_t p_end = p_begin + seg_pairs < npairs ? p_begin + seg_pairs : npairs;
    const TpLane me = tp_locate(P, w, lane);
    const uint32_t* codes = P.mrec + ((size_t) w * npairs) * 32u + lane;
    uint32_t* cell = P.peaks + 2 * ((me.ok ? P.tracks[me.track].peak_base : 0) + (me.ok ? me.ch : 0));
    const uint2 pk = __ldcg(reinterpret_cast<const uint2*>(cell));
    float floor_ = __uint_as_float(pk.x > pk.y ? pk.x : pk.y);
    // Integer form of the bound test: codes order like the values they stand
    // for, so "bound * value(code) > floor" can only hold for codes above the
    // (truncated) code of floor / bound; the margin keeps the pre-test a
    // superset of the exact test, which is repeated on the survivors.
    uint32_t code_gate = me.ok ? __float_as_uint(floor_ / P.tp_bound * 0.999f) >> 16 : 0xffffffffu;
    uint32_t prev = p_begin ? __ldcs(codes + (size_t) (p_begin - 1) * 32u) : 0u;
    uint32_t nxt[kTpBatch];
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j)
      nxt[j] = p_begin + j < p_end ? __ldcs(codes + (size_t) (p_begin + j) * 32u) : 0u;
    for (uint32_t p0 = p_begin; p0 < p_end; p0 += kTpBatch) {
      uint32_t cw[kTpBatch];
      uint32_t top = prev >> 16;
#pragma unroll
      for (int j = 0; j < kTpBatch; ++j) {
        cw[j] = nxt[j];
        const uint32_t pn = p0 + kTpBatch + j;
        nxt[j] = pn < p_end ? __ldcs(codes + (size_t) pn * 32u) : 0u;
        const uint32_t hi = cw[j] >> 16, lo = cw[j] & 0xffffu;
        top = max(top, max(hi, lo));
      }
      if (NT > kIter) top = max(top, prev & 0xffffu);
      if (__any_sync(0xffffffffu, top > code_gate)) {
#pragma unroll
        for (int j = 0; j < kTpBatch; ++j) {
          const uint32_t code = cw[j];
          // codes of iterations 2p-2 .. 2p+1
          const uint32_t c0 = prev & 0xffffu, c1 = prev >> 16, c2 = code & 0xffffu, c3 = code >> 16;
          prev = code;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            uint32_t cm = it ? (c3 > c2 ? c3 : c2) : (c2 > c1 ? c2 : c1);
            if (NT > kIter) { const uint32_t cb = it ? c1 : c0; cm = cm > cb ? cm : cb; }
            const uint32_t iter = (p0 + j) * 2u + it;
            const int f0 = (int) iter * kIter;
            const bool hit = me.ok && p0 + j < p_end && P.tp_bound * peak_code_value(cm) > floor_ &&
                             f0 + kIter > me.f_lo && f0 < me.f_end && me.a + f0 + kIter > me.lead_in;
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
              if (hit)
                queue[qn + __popc(mask & ((1u << lane) - 1u))] = tp_entry(me.track, me.ch, me.a + f0);
              qn += __popc(mask);
            }
          }
        }
        if (qn >= 32u) {
          __syncwarp();
          while (qn >= 32u) {
            qn -= 32u;
            tp_evaluate<FMT, TPF>(P, queue[qn + lane]);
          }
          __syncwarp();
          floor_ = fmaxf(floor_, __uint_as_float(__ldcg(cell + 1)));
          code_gate = me.ok ? __float_as_uint(floor_ / P.tp_bound * 0.999f) >> 16 : 0xffffffffu;
        }
      } else {
        prev = cw[kTpBatch - 1];
      }
    }
  }
  __syncwarp();
  if (lane < qn) tp_evaluate<FMT, TPF>(P, queue[lane]);
}

template <int FMT, int TPF>
static cudaError_t launch_truepeak_t(const SweepParams& p, uint32_t sms, cudaStream_t stream) {
  // Persistent grid: every resident warp scans work items of `seg` pairs of
  // one sweep warp.  Items are as long as possible (the per-item set-up is a
  // chain of dependent loads) while leaving about four per scanning warp.
  static int per_sm = 0;
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, truepeak_kernel<FMT, TPF>, kTpThreads, 0) !=
            cudaSuccess || per_sm < 1)
      per_sm = 4;
  }
  const uint64_t nscan = (uint64_t) sms * per_sm * kTpWarps;
  uint64_t nseg = (4 * nscan + p.nwarps - 1) / p.nwarps;
  const uint64_t max_seg = (p.npairs + kTpBatch - 1) / kTpBatch;
  if (nseg > max_seg) nseg = max_seg;
  if (nseg < 1) nseg = 1;
  uint32_t seg = (uint32_t) ((p.npairs + nseg - 1) / nseg);
  seg = (seg + kTpBatch - 1) / kTpBatch * kTpBatch;
  const uint64_t nitems = (uint64_t) p.nwarps * ((p.npairs + seg - 1) / seg);
  const uint64_t ctas = (nitems + kTpWarps - 1) / kTpWarps;
  const uint64_t want = (uint64_t) sms * per_sm;
  truepeak_kernel<FMT, TPF><<<(unsigned) (ctas < want ? ctas : want), kTpThreads, 0, stream>>>(p, seg);
  return cudaGetLastError();
}

cudaError_t launch_truepeak(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                            cudaStream_t stream, cudaEvent_t hold) {
  if (p.nwarps == 0 || tpf == 0) return cudaSuccess;
  if (hold) {
    const cudaError_t e = cudaStreamWaitEvent(stream, hold, 0);
    if (e != cudaSuccess) return e;
  }
  if (format == FMT_S16)
    return tpf == 4 ? launch_truepeak_t<FMT_S16, 4>(p, sms, stream)
                    : launch_truepeak_t<FMT_S16, 2>(p, sms, stream);
  return tpf == 4 ? launch_truepeak_t<FMT_F32, 4>(p, sms, stream)
                  : launch_truepeak_t<FMT_F32, 2>(p, sms, stream);
}

// --------------------------------------------------------- post-processing

// Index of the last track whose `base` field is <= idx (tracks are laid out
// in increasing base order; empty tracks share a base with their successor,
// so step past them).
template <class GetBase>
__device__ uint32_t find_track(const Track* tracks, uint32_t ntracks, uint64_t idx, GetBase base) {
  uint32_t lo = 0, hi = ntracks;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (base(tracks[mid]) <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

// The same with the tracks' base offsets staged in shared memory by the CTA
// (one coalesced pass instead of a chain of dependent global loads per thread);
// batches with more than kTrackCache tracks search the global table.
constexpr uint32_t kTrackCache = 1024;

template <class GetBase>
__device__ uint32_t find_track_cta(const Track* tracks, uint32_t ntracks, uint64_t idx, bool valid,
                                   uint64_t* s_base, GetBase base) {
  if (ntracks > kTrackCache) return valid ? find_track(tracks, ntracks, idx, base) : 0u;
  for (uint32_t i = threadIdx.x; i < ntracks; i += blockDim.x) s_base[i] = base(tracks[i]);
  __syncthreads();
  uint32_t lo = 0, hi = ntracks;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s_base[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

// FP64 state carry + energy correction + channel weighting, one thread per 100 ms slot
// (lg_post.cuh: slot_energy_fused).  (Staging a CTA's stretch of chunk records in shared
// memory first made it slower, 47 -> 61 us next to the true-peak evaluation:
// profiles/r02_tuning.txt E.)
constexpr int kFixThreads = 128;

__global__ void __launch_bounds__(kFixThreads, 5)
fixslot_kernel(const Track* __restrict__ tracks, uint32_t ntracks, const CoefSet* __restrict__ coefs,
               const ChunkRec* __restrict__ recs, uint64_t total_slots, double* __restrict__ eslot,
               const cplx* __restrict__ xi_table) {
  __shared__ uint64_t s_base[kTrackCache];
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ti = find_track_cta(tracks, ntracks, s, s < total_slots, s_base,
                                     [](const Track& t) { return t.slot_base; });
  if (s >= total_slots) return;
  const Track& tr = tracks[ti];
  eslot[s] = slot_energy_fused(tr, coefs[tr.coef], recs, (uint32_t) (s - tr.slot_base), xi_table,
                               31 - __clz((int) tr.aq));
}

// EBUR128_MODE_HISTOGRAM: libebur128 then does not store block energies but counts
// them in 1000 bins of 0.1 LU from -70 to +30 LUFS, and every later sum, gate and
// percentile works on the bins' centre energies.  That is the same as replacing each
// block energy by the centre of its bin at the moment it is stored (a block below
// the first edge -- the absolute gate -- is dropped: 0 never passes a gate), which
// is what the block kernels do for tracks with Track::flags bit 0.  tab = 1001 bin
// edges, then 1000 centre energies (hist_table(): the reference's own expressions,
// evaluated on the host); the bin is the reference's binary search result.
constexpr int kHistBins = 1000;

__device__ double hist_quantise(double e, const double* __restrict__ tab) {
  if (!(e >= tab[0])) return 0.0;
  int i = (int) floor((10.0 * log10(e) + 0.691 + 70.0) * 10.0);
  i = i < 0 ? 0 : (i > kHistBins - 1 ? kHistBins - 1 : i);
  while (i > 0 && e < tab[i]) --i;
  while (i < kHistBins - 1 && e >= tab[i + 1]) ++i;
  return tab[kHistBins + 1 + i];
}

__global__ void __launch_bounds__(256)
block_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
             const CoefSet* __restrict__ coefs, const double* __restrict__ eslot,
             uint64_t total_blocks, uint64_t total_st, double* __restrict__ zblock,
             double* __restrict__ zst, const double* __restrict__ hist_tab) {
  __shared__ uint64_t s_bbase[kTrackCache], s_sbase[kTrackCache];
  const bool cached = ntracks <= kTrackCache;
  if (cached) {
    for (uint32_t k = threadIdx.x; k < ntracks; k += blockDim.x) {
      s_bbase[k] = tracks[k].block_base;
      s_sbase[k] = tracks[k].st_base;
    }
    __syncthreads();
  }
  auto locate = [&](const uint64_t* sb, uint64_t idx, auto base) {
    if (!cached) return find_track(tracks, ntracks, idx, base);
    uint32_t lo = 0, hi = ntracks;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sb[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
  };
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total_blocks) {
    const uint32_t ti = locate(s_bbase, i, [](const Track& t) { return t.block_base; });
    const Track& tr = tracks[ti];
    const double e = gating_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (i - tr.block_base));
    zblock[i] = (tr.flags & 1u) ? hist_quantise(e, hist_tab) : e;
  } else if (i < total_blocks + total_st) {
    const uint64_t j = i - total_blocks;
    const uint32_t ti = locate(s_sbase, j, [](const Track& t) { return t.st_base; });
    const Track& tr = tracks[ti];
    const double e = shortterm_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (j - tr.st_base));
    zst[j] = (tr.flags & 1u) ? hist_quantise(e, hist_tab) : e;
  }
}

// Blocks of one stream from its complete slot list (time segments measured
// separately, e.g. on several GPUs, and concatenated: lgb_slots_query).
__global__ void __launch_bounds__(256)
stream_block_kernel(const double* __restrict__ eslot, int s100, uint64_t nblocks, uint64_t nst,
                    double* __restrict__ zblock, double* __restrict__ zst,
                    const double* __restrict__ hist_tab) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblocks) {
    const double e = gating_block(eslot, s100, i);
    zblock[i] = hist_tab ? hist_quantise(e, hist_tab) : e;
  } else if (i < nblocks + nst) {
    const double e = shortterm_block(eslot, s100, i - nblocks);
    zst[i - nblocks] = hist_tab ? hist_quantise(e, hist_tab) : e;
  }
}

cudaError_t launch_stream_blocks(const double* eslot, int s100, uint64_t nblocks, uint64_t nst,
                                 double* zblock, double* zst, cudaStream_t stream, const double* hist_tab) {
  if (nblocks + nst == 0) return cudaSuccess;
  const unsigned blocks = (unsigned) ((nblocks + nst + 255) / 256);
  stream_block_kernel<<<blocks, 256, 0, stream>>>(eslot, s100, nblocks, nst, zblock, zst, hist_tab);
  return cudaGetLastError();
}

// The histogram table of the current device (created on first use, never freed).
const double* hist_table() {
  static std::mutex mu;
  static const double* tabs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (tabs[dev]) return tabs[dev];
  std::vector<double> h(2 * kHistBins + 1);
  // libebur128 1.2.x ebur128.c: histogram_energy_boundaries / histogram_energies
  h[0] = pow(10.0, (-70.0 + 0.691) / 10.0);
  for (int i = 1; i <= kHistBins; ++i) h[i] = pow(10.0, ((double) i / 10.0 - 70.0 + 0.691) / 10.0);
  for (int i = 0; i < kHistBins; ++i) h[kHistBins + 1 + i] = pow(10.0, ((double) i / 10.0 - 69.95 + 0.691) / 10.0);
  double* d = nullptr;
  if (cudaMalloc((void**) &d, h.size() * sizeof(double)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  tabs[dev] = d;
  return d;
}

// ------------------------------------------------------------- reductions

constexpr int kQueryThreads = 1024;     // the most a query CTA may have (select_two needs 256 or more)
// What the query kernels are launched with.  A 1024-thread CTA needs a whole SM's register
// file (64 per thread), so it waits for any SM that still holds CTAs of the true-peak
// evaluation; smaller CTAs slip in next to them but have fewer loads in flight.  512 was
// the best of 256 / 512 / 1024 on the album step (profiles/r02_tuning.txt E).
constexpr int kQueryLaunchDefault = 512;
static int query_launch_threads() {
  static const int n = [] {
    const char* e = getenv("LOUDGAIN_B200_QUERY_THREADS");      // tuning: 256, 512 or 1024
    const int v = e ? atoi(e) : 0;
    return (v == 256 || v == 512 || v == 1024) ? v : kQueryLaunchDefault;
  }();
  return n;
}

struct SumCount {
  double s;
  unsigned long long n;
};

// Deterministic CTA-wide (sum, count): fixed shuffle tree, then warp 0.
__device__ SumCount block_sum_count(double s, unsigned long long n, SumCount* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    n += __shfl_down_sync(0xffffffffu, n, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { scratch[wid].s = s; scratch[wid].n = n; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? scratch[lane].s : 0.0;
    n = lane < nw ? scratch[lane].n : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_down_sync(0xffffffffu, s, o);
      n += __shfl_down_sync(0xffffffffu, n, o);
    }
    if (lane == 0) { scratch[0].s = s; scratch[0].n = n; }
  }
  __syncthreads();
  const SumCount r = scratch[0];
  __syncthreads();
  return r;
}

// A query's member block lists, flattened: element g of the concatenation of
// all members' lists.  Up to kQueryCache member descriptors (and their prefix
// offsets) are staged in shared memory so that every thread's loads are
// independent; larger queries walk the member list.
constexpr int kQueryCache = 256;
constexpr int kStCache = 4096;          // short-term energies kept in shared memory

struct QueryView {
  const BlockList* lists;     // global
  const uint32_t* mem;        // global member indices of this query
  uint32_t count;
  bool cached;
  const BlockList* s_lists;   // shared copies (cached)
  const uint32_t* s_zoff;     // [count + 1]
  const uint32_t* s_stoff;    // [count + 1]
  // how the elements are dealt out: this thread takes first, first + stride, ...
  // (one CTA: threadIdx.x / blockDim.x; a cluster sharing the gating blocks:
  // rank * blockDim.x + threadIdx.x / cluster size * blockDim.x)
  uint32_t first, stride;
};

// f(e, g) for every gating block energy (ST = false) or short-term energy
// (ST = true).  Cached queries work in batches of kQueryBatch elements per
// thread: all searches, then all loads, then the callbacks, so that a thread
// has kQueryBatch independent loads in flight instead of one.
constexpr int kQueryBatch = 8;

template <bool ST, class F>
__device__ __forceinline__ void for_each_energy(const QueryView& v, F f) {
  if (v.cached) {
    const uint32_t* off = ST ? v.s_stoff : v.s_zoff;
    const uint32_t total = off[v.count];
    // a thread's elements come in increasing order, so its member index only
    // ever moves forward: no search per element
    uint32_t m = 0;
    for (uint32_t g0 = v.first; g0 < total; g0 += v.stride * kQueryBatch) {
      const double* src[kQueryBatch];
      double e[kQueryBatch];
#pragma unroll
      for (int u = 0; u < kQueryBatch; ++u) {
        const uint32_t g = g0 + u * v.stride;
        const uint32_t gc = g < total ? g : total - 1;
        while (off[m + 1] <= gc) ++m;            // off[count] = total > gc: stops in range
        src[u] = (ST ? v.s_lists[m].st : v.s_lists[m].z) + (gc - off[m]);
      }
#pragma unroll
      for (int u = 0; u < kQueryBatch; ++u) e[u] = *src[u];
#pragma unroll
      for (int u = 0; u < kQueryBatch; ++u) {
        const uint32_t g = g0 + u * v.stride;
        if (g < total) f(e[u], g);
      }
    }
  } else {
    uint32_t base = 0;
    for (uint32_t m = 0; m < v.count; ++m) {
      const BlockList bl = v.lists[v.mem[m]];
      const double* p = ST ? bl.st : bl.z;
      const uint32_t n = ST ? bl.nst : bl.nz;
      for (uint32_t i = v.first; i < n; i += v.stride) f(p[i], base + i);
      base += n;
    }
  }
}

// The k-th smallest (0-based) values, for two ranks at once, among the
// short-term energies >= floor_e.  Positive doubles order like their 64-bit
// patterns: 8 passes of 8-bit radix select, one 256-bin histogram per rank and
// pass, bin search by a block-wide scan.  The values come from shared memory
// (s_st, n_st of them) when they fit, else from the member lists.
struct SelectState {
  unsigned long long prefix[2];
  unsigned long long k[2];
};

__device__ void select_two(const QueryView& v, const double* s_st, uint32_t n_st, bool st_cached,
                           double floor_e, unsigned long long k_lo, unsigned long long k_hi,
                           unsigned int* hist /* [2][256] */, unsigned int* wsum /* [2][8] */,
                           SelectState* st, double* out_lo, double* out_hi) {
  if (threadIdx.x == 0) { st->prefix[0] = st->prefix[1] = 0; st->k[0] = k_lo; st->k[1] = k_hi; }
  unsigned long long mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    auto count = [&](double e, uint32_t) {
      if (!(e >= floor_e)) return;
      const unsigned long long bits = (unsigned long long) __double_as_longlong(e);
      const unsigned int digit = (unsigned int) ((bits >> shift) & 0xffull);
      if ((bits & mask) == p0) atomicAdd(&hist[digit], 1u);
      if ((bits & mask) == p1) atomicAdd(&hist[256 + digit], 1u);
    };
    if (st_cached) {
      for (uint32_t i = threadIdx.x; i < n_st; i += blockDim.x) count(s_st[i], i);
    } else {
      for_each_energy<true>(v, count);
    }
    __syncthreads();
    // a CTA of 512 threads or more: threads 0..255 own the bins of rank 0, 256..511 those
    // of rank 1; a 256-thread CTA scans the two histograms one after the other
    const int nscan = blockDim.x >= 512 ? 1 : 2;
    for (int sc = 0; sc < nscan; ++sc) {
      const bool owner = threadIdx.x < (nscan == 1 ? 512 : 256);
      const int which = nscan == 1 ? (threadIdx.x >> 8) & 1 : sc, bin = threadIdx.x & 255;
      const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 7;
      const unsigned int cnt = owner ? hist[which * 256 + bin] : 0u;
      unsigned int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (owner && lane == 31) wsum[which * 8 + w] = inc;
      __syncthreads();
      unsigned int base = 0;
      for (int j = 0; j < w; ++j) base += wsum[which * 8 + j];
      const unsigned long long excl = (unsigned long long) base + inc - cnt;
      const unsigned long long kk = st->k[which];
      __syncthreads();
      if (owner && cnt && excl <= kk && kk < excl + cnt) {
        st->k[which] = kk - excl;
        st->prefix[which] |= (unsigned long long) bin << shift;
      }
    }
    mask |= 0xffull << shift;
    __syncthreads();
  }
  *out_lo = __longlong_as_double((long long) st->prefix[0]);
  *out_hi = __longlong_as_double((long long) st->prefix[1]);
}

// One query per thread-block CLUSTER.  The gating blocks (the long lists: 10 per
// second of audio) are dealt out over the cluster's CTAs, whose partial sums
// meet in rank 0's shared memory (DSMEM); the short-term energies (1 per
// second) and the final result are rank 0's alone.  Cluster size 1 is the
// plain one-CTA query.
constexpr int kMaxQueryCluster = 8;

__device__ SumCount cluster_sum_count(cooperative_groups::cluster_group& cluster, SumCount mine,
                                      SumCount* xch /* [kMaxQueryCluster], this round's buffer */) {
  const unsigned R = cluster.num_blocks();
  if (R == 1) return mine;
  if (threadIdx.x == 0) cluster.map_shared_rank(xch, 0)[cluster.block_rank()] = mine;
  cluster.sync();
  const SumCount* all = cluster.map_shared_rank(xch, 0);
  SumCount t{0.0, 0ull};
  for (unsigned r = 0; r < R; ++r) { t.s += all[r].s; t.n += all[r].n; }   // fixed order
  return t;
}

// PART 0: the whole query; 1: integrated loudness only; 2: loudness range only (the two
// halves read different lists and write different fields of the result, so a step runs
// them side by side on two streams: lg_batch.cu).
template <int PART>
__global__ void __launch_bounds__(kQueryThreads)
query_kernel(const BlockList* __restrict__ lists, const Query* __restrict__ queries,
             const uint32_t* __restrict__ members, double abs_gate,
             QueryResult* __restrict__ results) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned R = cluster.num_blocks(), rk = cluster.block_rank();
  const unsigned qi = blockIdx.x / R;
  if (R > 1) cluster.sync();      // every CTA of the cluster runs before its shared memory is touched
  __shared__ SumCount scratch[32];
  __shared__ SumCount xch[2][kMaxQueryCluster];
  __shared__ unsigned int hist[512];
  __shared__ unsigned int wsum[16];
  __shared__ SelectState sel;
  __shared__ BlockList s_lists[kQueryCache];
  __shared__ uint32_t s_zoff[kQueryCache + 1], s_stoff[kQueryCache + 1];
  __shared__ double s_st[kStCache];
  const Query q = queries[qi];
  QueryView v;
  v.lists = lists; v.mem = members + q.first; v.count = q.count;
  v.cached = q.count <= (uint32_t) kQueryCache;
  v.s_lists = s_lists; v.s_zoff = s_zoff; v.s_stoff = s_stoff;
  if (v.cached) {
    for (uint32_t m = threadIdx.x; m < q.count; m += blockDim.x) s_lists[m] = lists[v.mem[m]];
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t az = 0, ast = 0;
      for (uint32_t m = 0; m < q.count; ++m) {
        s_zoff[m] = az; s_stoff[m] = ast;
        az += s_lists[m].nz; ast += s_lists[m].nst;
      }
      s_zoff[q.count] = az; s_stoff[q.count] = ast;
    }
    __syncthreads();
  }
  QueryResult res;
  res.loudness = -HUGE_VAL; res.range = 0.0; res.rel_thr = 0.0;
  res.sum1 = res.sum2 = 0.0; res.n1 = res.n2 = res.nst = 0;

  // ---- integrated loudness: absolute gate, then relative gate at -10 LU;
  //      the blocks are shared out over the cluster
  v.first = rk * blockDim.x + threadIdx.x;
  v.stride = R * blockDim.x;
  double s = 0.0;
  unsigned long long n = 0;
  SumCount a{0.0, 0ull};
  if (PART != 2) {
  for_each_energy<false>(v, [&](double e, uint32_t) { if (e >= abs_gate) { s += e; ++n; } });
  a = cluster_sum_count(cluster, block_sum_count(s, n, scratch), xch[0]);
  res.sum1 = a.s; res.n1 = a.n;
  SumCount b{0.0, 0ull};
  if (a.n) {                                   // cluster-uniform
    const double thr = a.s / (double) a.n * 0.1;
    res.rel_thr = thr;
    s = 0.0; n = 0;
    for_each_energy<false>(v, [&](double e, uint32_t) { if (e >= abs_gate && e >= thr) { s += e; ++n; } });
    b = cluster_sum_count(cluster, block_sum_count(s, n, scratch), xch[1]);
    res.sum2 = b.s; res.n2 = b.n;
    if (b.n) res.loudness = energy_to_lufs(b.s / (double) b.n);
  }
  if (R > 1) {
    cluster.sync();                            // everyone has read rank 0's exchange buffers
    if (rk != 0) return;
  }
  }
  if (PART == 1) {
    if (threadIdx.x == 0) {
      QueryResult& o = results[qi];
      o.loudness = res.loudness; o.rel_thr = res.rel_thr;
      o.sum1 = res.sum1; o.sum2 = res.sum2; o.n1 = res.n1; o.n2 = res.n2;
    }
    return;
  }

  // ---- loudness range (rank 0): -20 LU relative gate on short-term energies,
  //      then the 10th / 95th percentile by rank.  The energies are staged in
  //      shared memory during the first pass when they fit.
  v.first = threadIdx.x;
  v.stride = blockDim.x;
  const bool st_cached = v.cached && s_stoff[q.count] <= (uint32_t) kStCache;
  const uint32_t n_st = st_cached ? s_stoff[q.count] : 0u;
  s = 0.0; n = 0;
  for_each_energy<true>(v, [&](double e, uint32_t g) {
    if (st_cached) s_st[g] = e;
    if (e >= abs_gate) { s += e; ++n; }
  });
  a = block_sum_count(s, n, scratch);          // (its barriers publish s_st)
  res.nst = a.n;
  if (a.n) {
    double floor_e = a.s / (double) a.n * 0.01;
    if (floor_e < abs_gate) floor_e = abs_gate;
    n = 0;
    if (st_cached) {
      for (uint32_t i = threadIdx.x; i < n_st; i += blockDim.x) if (s_st[i] >= floor_e) ++n;
    } else {
      for_each_energy<true>(v, [&](double e, uint32_t) { if (e >= floor_e) ++n; });
    }
    const SumCount c = block_sum_count(0.0, n, scratch);
    if (c.n) {
      const unsigned long long k_hi = (unsigned long long) ((double) (c.n - 1) * 0.95 + 0.5);
      const unsigned long long k_lo = (unsigned long long) ((double) (c.n - 1) * 0.1 + 0.5);
      double lo, hi;
      select_two(v, s_st, n_st, st_cached, floor_e, k_lo, k_hi, hist, wsum, &sel, &lo, &hi);
      res.range = energy_to_lufs(hi) - energy_to_lufs(lo);
    }
  }
  if (threadIdx.x == 0) {
    if (PART == 2) { results[qi].range = res.range; results[qi].nst = res.nst; }
    else results[qi] = res;
  }
}

// ---------------------------------------------------------------------------
// Album queries over tracks that are sharded across GPUs
// (ebur128_loudness_global_multiple / _range_multiple of scan.c:383-391 when the
// album's tracks were measured on different ranks).
//
// The gating is distributed, not the block lists: every rank reduces ITS OWN
// gating blocks and only (sum, count) pairs travel -- written by the kernels
// themselves into every peer's exchange region over NVLink (peer-mapped memory,
// lg_common.h: XchgParams), ordered by system-scope release / acquire on one
// flag per rank and phase.  Four launches per step, each of which waits only
// for the peers' PREVIOUS launch, so nothing depends on how any GPU schedules
// its CTAs:
//   xchg_publish_kernel   sums behind the absolute gate (gating and short-term
//                         blocks) + the rank's short-term energies -> all peers
//   xchg_gate_kernel      waits for phase 0 of every rank; relative threshold
//                         from the rank-ordered totals; sums behind it -> all peers
//   xchg_range_kernel     (on another stream, next to the gating) waits for phase 0;
//                         range percentiles over the union of the short-term
//                         energies every rank received
//   xchg_finish_kernel    waits for phase 1; loudness from the totals
// Totals are added in rank order on every rank: all ranks get the same bits.
// The short-term lists are one value per second of audio, a tenth of the gating
// blocks, which never leave their GPU.
struct ViewSmem {
  BlockList lists[kQueryCache];
  uint32_t zoff[kQueryCache + 1], stoff[kQueryCache + 1];
};

__device__ void load_view(QueryView& v, ViewSmem& sm, const BlockList* lists, const uint32_t* members,
                          const Query q) {
  v.lists = lists; v.mem = members + q.first; v.count = q.count;
  v.cached = q.count <= (uint32_t) kQueryCache;
  v.s_lists = sm.lists; v.s_zoff = sm.zoff; v.s_stoff = sm.stoff;
  v.first = threadIdx.x; v.stride = blockDim.x;
  if (v.cached) {
    for (uint32_t m = threadIdx.x; m < q.count; m += blockDim.x) sm.lists[m] = lists[v.mem[m]];
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t az = 0, ast = 0;
      for (uint32_t m = 0; m < q.count; ++m) {
        sm.zoff[m] = az; sm.stoff[m] = ast;
        az += sm.lists[m].nz; ast += sm.lists[m].nst;
      }
      sm.zoff[q.count] = az; sm.stoff[q.count] = ast;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ XchgHdr* xchg_hdr(const XchgParams& X, uint32_t target, uint32_t parity, uint32_t r,
                                             uint32_t a) {
  return reinterpret_cast<XchgHdr*>(X.peer[target] + xchg_hdr_off(X.world, X.nalbums, parity, r, a));
}
__device__ __forceinline__ double* xchg_st(const XchgParams& X, uint32_t target, uint32_t parity, uint32_t r) {
  return reinterpret_cast<double*>(X.peer[target] + xchg_st_off(X.world, X.nalbums, X.st_cap, parity, r));
}
__device__ __forceinline__ unsigned long long* xchg_flag(const XchgParams& X, uint32_t target, uint32_t phase,
                                                         uint32_t r) {
  return reinterpret_cast<unsigned long long*>(X.peer[target]) + phase * kMaxWorld + r;
}

// Every thread has fenced its own peer stores; the CTA that finishes last tells
// every rank that this rank's `phase` of `step` is complete.
__device__ void xchg_signal(const XchgParams& X, uint32_t phase, unsigned long long step, uint32_t nctas,
                            bool wrote) {
  if (wrote) __threadfence_system();           // only threads that stored into a peer's region
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(X.ctl + 1 + phase, 1ull);
    if (done == nctas - 1) {
      X.ctl[1 + phase] = 0ull;
      __threadfence_system();
      for (uint32_t p = 0; p < X.world; ++p) st_release_sys(xchg_flag(X, p, phase, X.rank), step);
    }
  }
}

// Waits until every rank has completed `phase` of `step` (flags only grow).  A
// peer that never arrives (a rank that died) is given up on after kXchgTimeoutNs:
// the time-out is counted in ctl[4] and reported by the fetch, the GPU does not hang.
constexpr unsigned long long kXchgTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ void xchg_wait(const XchgParams& X, uint32_t phase, unsigned long long step) {
  if (threadIdx.x < X.world) {
    const unsigned long long* f = xchg_flag(X, X.rank, phase, threadIdx.x);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < step) {
      if (global_ns() - t0 > kXchgTimeoutNs) { atomicAdd(X.ctl + 4, 1ull); break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
}

// (A thread-block cluster per album shares the album's local blocks out, like query_kernel.)
__global__ void __launch_bounds__(kQueryThreads)
xchg_publish_kernel(const BlockList* __restrict__ lists, const Query* __restrict__ queries,
                    const uint32_t* __restrict__ members, double abs_gate, const __grid_constant__ XchgParams X) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned R = cluster.num_blocks(), rk = cluster.block_rank();
  if (R > 1) cluster.sync();      // every CTA of the cluster runs before its shared memory is touched
  __shared__ SumCount scratch[32];
  __shared__ SumCount xch[2][kMaxQueryCluster];
  __shared__ ViewSmem vs;
  const uint32_t a = blockIdx.x / R;
  const unsigned long long step = __ldcg(X.ctl);
  const uint32_t parity = (uint32_t) (step & 1ull);
  QueryView v;
  load_view(v, vs, lists, members, queries[X.first_query + a]);
  v.first = rk * blockDim.x + threadIdx.x;
  v.stride = R * blockDim.x;
  double s = 0.0;
  unsigned long long n = 0;
  for_each_energy<false>(v, [&](double e, uint32_t) { if (e >= abs_gate) { s += e; ++n; } });
  const SumCount za = cluster_sum_count(cluster, block_sum_count(s, n, scratch), xch[0]);
  const uint32_t off = X.st_off[a], cnt = X.st_off[a + 1] - off;
  s = 0.0; n = 0;
  bool wrote = false;
  for_each_energy<true>(v, [&](double e, uint32_t g) {
    for (uint32_t p = 0; p < X.world; ++p) xchg_st(X, p, parity, X.rank)[off + g] = e;
    wrote = true;
    if (e >= abs_gate) { s += e; ++n; }
  });
  const SumCount sa = cluster_sum_count(cluster, block_sum_count(s, n, scratch), xch[1]);
  if (rk == 0 && threadIdx.x == 0) {
    for (uint32_t p = 0; p < X.world; ++p) {
      XchgHdr* h = xchg_hdr(X, p, parity, X.rank, a);
      h->s1 = za.s; h->n1 = za.n; h->sst = sa.s; h->nst = sa.n;
      h->st_off = off; h->st_cnt = cnt;
    }
    wrote = true;
  }
  xchg_signal(X, 0, step, gridDim.x, wrote);
  if (R > 1) cluster.sync();      // rank 0's exchange buffers have been read by everyone
}

struct XchgTotals {
  double s1, sst, s2;
  unsigned long long n1, nst, n2;
};

// Totals of album `a` over all ranks.  Thread r fetches rank r's header (four 16-byte loads,
// all ranks at once), thread 0 adds them up in rank order: the same bits on every rank.
// CTA-collective; the headers stay in s_hdr for the caller.
__device__ void xchg_totals(const XchgParams& X, uint32_t parity, uint32_t a, XchgHdr* s_hdr /* [kMaxWorld] */,
                            XchgTotals* out /* shared */) {
  if (threadIdx.x < X.world) {
    const uint4* src = reinterpret_cast<const uint4*>(xchg_hdr(X, X.rank, parity, threadIdx.x, a));
    uint4* dst = reinterpret_cast<uint4*>(s_hdr + threadIdx.x);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = __ldcg(src + k);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    XchgTotals t{0.0, 0.0, 0.0, 0ull, 0ull, 0ull};
    for (uint32_t r = 0; r < X.world; ++r) {
      t.s1 += s_hdr[r].s1; t.n1 += s_hdr[r].n1;
      t.sst += s_hdr[r].sst; t.nst += s_hdr[r].nst;
      t.s2 += s_hdr[r].s2; t.n2 += s_hdr[r].n2;
    }
    *out = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kQueryThreads)
xchg_gate_kernel(const BlockList* __restrict__ lists, const Query* __restrict__ queries,
                 const uint32_t* __restrict__ members, double abs_gate, const __grid_constant__ XchgParams X) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned R = cluster.num_blocks(), rk = cluster.block_rank();
  if (R > 1) cluster.sync();
  __shared__ SumCount scratch[32];
  __shared__ SumCount xch[kMaxQueryCluster];
  __shared__ ViewSmem vs;
  __shared__ XchgTotals tot;
  __shared__ __align__(16) XchgHdr s_hdr[kMaxWorld];
  const uint32_t a = blockIdx.x / R;
  const unsigned long long step = __ldcg(X.ctl);
  const uint32_t parity = (uint32_t) (step & 1ull);
  QueryView v;
  load_view(v, vs, lists, members, queries[X.first_query + a]);
  v.first = rk * blockDim.x + threadIdx.x;
  v.stride = R * blockDim.x;
  xchg_wait(X, 0, step);
  xchg_totals(X, parity, a, s_hdr, &tot);
  double s = 0.0;
  unsigned long long n = 0;
  if (tot.n1) {
    const double thr = tot.s1 / (double) tot.n1 * 0.1;
    for_each_energy<false>(v, [&](double e, uint32_t) { if (e >= abs_gate && e >= thr) { s += e; ++n; } });
  }
  const SumCount b = cluster_sum_count(cluster, block_sum_count(s, n, scratch), xch);
  if (rk == 0 && threadIdx.x == 0) {
    for (uint32_t p = 0; p < X.world; ++p) {
      XchgHdr* h = xchg_hdr(X, p, parity, X.rank, a);
      h->s2 = b.s; h->n2 = b.n;
    }
  }
  xchg_signal(X, 1, step, gridDim.x, rk == 0 && threadIdx.x == 0);
  if (R > 1) cluster.sync();
}

// The range of an album needs nothing of the relative gate: only what phase 0 brought
// (every rank's short-term energies and their sums behind the absolute gate).  It is
// selected by its own launch, next to the gating.  The union of the short-term energies is
// staged in (dynamic) shared memory when it fits.
__global__ void __launch_bounds__(kQueryThreads)
xchg_range_kernel(double abs_gate, QueryResult* __restrict__ results, uint32_t st_smem_cap,
                  const __grid_constant__ XchgParams X) {
  extern __shared__ double s_stx[];
  __shared__ SumCount scratch[32];
  __shared__ ViewSmem vs;
  __shared__ XchgTotals tot;
  __shared__ unsigned int hist[512];
  __shared__ unsigned int wsum[16];
  __shared__ SelectState sel;
  const uint32_t a = blockIdx.x;
  const unsigned long long step = __ldcg(X.ctl);
  const uint32_t parity = (uint32_t) (step & 1ull);
  __shared__ __align__(16) XchgHdr s_hdr[kMaxWorld];
  xchg_wait(X, 0, step);
  xchg_totals(X, parity, a, s_hdr, &tot);
  if (threadIdx.x == 0) {
    uint32_t ast = 0;
    for (uint32_t r = 0; r < X.world; ++r) {
      const uint32_t cnt = s_hdr[r].st_cnt;
      vs.lists[r] = BlockList{nullptr, xchg_st(X, X.rank, parity, r) + s_hdr[r].st_off, 0u, cnt};
      vs.zoff[r] = 0; vs.stoff[r] = ast;
      ast += cnt;
    }
    vs.zoff[X.world] = 0; vs.stoff[X.world] = ast;
  }
  __syncthreads();
  QueryView v;
  v.lists = nullptr; v.mem = nullptr; v.count = X.world; v.cached = true;
  v.s_lists = vs.lists; v.s_zoff = vs.zoff; v.s_stoff = vs.stoff;
  v.first = threadIdx.x; v.stride = blockDim.x;
  const uint32_t n_st = vs.stoff[X.world];
  const bool st_cached = n_st <= st_smem_cap;
  double range = 0.0;
  if (tot.nst) {
    double floor_e = tot.sst / (double) tot.nst * 0.01;
    if (floor_e < abs_gate) floor_e = abs_gate;
    unsigned long long n = 0;
    for_each_energy<true>(v, [&](double e, uint32_t g) {
      if (st_cached) s_stx[g] = e;
      if (e >= floor_e) ++n;
    });
    const SumCount c = block_sum_count(0.0, n, scratch);      // (its barriers publish s_stx)
    if (c.n) {
      const unsigned long long k_hi = (unsigned long long) ((double) (c.n - 1) * 0.95 + 0.5);
      const unsigned long long k_lo = (unsigned long long) ((double) (c.n - 1) * 0.1 + 0.5);
      double lo, hi;
      select_two(v, s_stx, n_st, st_cached, floor_e, k_lo, k_hi, hist, wsum, &sel, &lo, &hi);
      range = energy_to_lufs(hi) - energy_to_lufs(lo);
    }
  }
  if (threadIdx.x == 0) {
    results[X.first_query + a].range = range;
    results[X.first_query + a].nst = tot.nst;
  }
}

__global__ void __launch_bounds__(64)
xchg_finish_kernel(QueryResult* __restrict__ results, const __grid_constant__ XchgParams X) {
  const uint32_t a = blockIdx.x;
  const unsigned long long step = __ldcg(X.ctl);
  const uint32_t parity = (uint32_t) (step & 1ull);
  __shared__ __align__(16) XchgHdr s_hdr[kMaxWorld];
  __shared__ XchgTotals tot;
  xchg_wait(X, 1, step);
  xchg_totals(X, parity, a, s_hdr, &tot);
  if (threadIdx.x == 0) {
    QueryResult& o = results[X.first_query + a];
    o.loudness = -HUGE_VAL; o.rel_thr = 0.0;
    o.sum1 = tot.s1; o.n1 = tot.n1; o.sum2 = tot.s2; o.n2 = tot.n2;
    if (tot.n1) {
      o.rel_thr = tot.s1 / (double) tot.n1 * 0.1;
      if (tot.n2) o.loudness = energy_to_lufs(tot.s2 / (double) tot.n2);
    }
  }
}

// The step is over on this rank: the step counter moves on.  Its own launch, behind BOTH chains of
// the exchange (gate -> finish, and the albums' ranges on the stream next to it): every kernel
// of a step reads the counter when it starts, and as long as the finish kernel advanced it, a
// range kernel that started late -- behind the finish kernel of its own step: it runs on another
// stream, needs most of an SM's shared memory, and at low priority it can wait for that -- read
// the NEXT step's number, waited for flags nobody was going to raise before this step was
// fetched, and the exchange timed out (seen once at two and once at eight GPUs).
__global__ void xchg_advance_kernel(const __grid_constant__ XchgParams X) {
  if (threadIdx.x == 0 && blockIdx.x == 0) X.ctl[0] = X.ctl[0] + 1ull;
}

static cudaError_t launch_clustered(const void* kernel, uint32_t nalbums, uint32_t cluster, cudaStream_t stream,
                                    void** args) {
  if (cluster < 1) cluster = 1;
  if (cluster > (uint32_t) kMaxQueryCluster) cluster = kMaxQueryCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nalbums * cluster);
  cfg.blockDim = dim3(query_launch_threads());
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

cudaError_t launch_exchange_publish(const BlockList* lists, const Query* queries, const uint32_t* members,
                                    double abs_gate, const XchgParams& x, uint32_t cluster, cudaStream_t stream) {
  if (!x.nalbums) return cudaSuccess;
  void* args[] = {(void*) &lists, (void*) &queries, (void*) &members, (void*) &abs_gate, (void*) &x};
  return launch_clustered((const void*) xchg_publish_kernel, x.nalbums, cluster, stream, args);
}

cudaError_t launch_exchange_gate(const BlockList* lists, const Query* queries, const uint32_t* members,
                                 double abs_gate, const XchgParams& x, uint32_t cluster, cudaStream_t stream) {
  if (!x.nalbums) return cudaSuccess;
  void* args[] = {(void*) &lists, (void*) &queries, (void*) &members, (void*) &abs_gate, (void*) &x};
  return launch_clustered((const void*) xchg_gate_kernel, x.nalbums, cluster, stream, args);
}

cudaError_t launch_exchange_range(double abs_gate, QueryResult* results, const XchgParams& x,
                                  uint32_t st_smem_doubles, cudaStream_t stream) {
  if (!x.nalbums) return cudaSuccess;
  static size_t smem_limit = 0;
  const size_t smem = (size_t) st_smem_doubles * sizeof(double);
  if (smem > 32u * 1024u && smem > smem_limit) {       // (the kernel also has ~12 KB of static shared memory)
    const cudaError_t e = cudaFuncSetAttribute(xchg_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int) smem);
    if (e != cudaSuccess) return e;
    smem_limit = smem;
  }
  xchg_range_kernel<<<x.nalbums, query_launch_threads(), smem, stream>>>(abs_gate, results, st_smem_doubles, x);
  return cudaGetLastError();
}

cudaError_t launch_exchange_finish(QueryResult* results, const XchgParams& x, cudaStream_t stream) {
  if (!x.nalbums) return cudaSuccess;
  xchg_finish_kernel<<<x.nalbums, 64, 0, stream>>>(results, x);
  return cudaGetLastError();
}

cudaError_t launch_exchange_advance(const XchgParams& x, cudaStream_t stream) {
  if (!x.nalbums) return cudaSuccess;
  xchg_advance_kernel<<<1, 32, 0, stream>>>(x);
  return cudaGetLastError();
}

cudaError_t launch_post(const DeviceTables& t, const PostSizes& z, cudaStream_t stream, cudaEvent_t fixed) {
  if (z.total_slots) {
    const unsigned blocks = (unsigned) ((z.total_slots + kFixThreads - 1) / kFixThreads);
    fixslot_kernel<<<blocks, kFixThreads, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.recs, z.total_slots, t.eslot,
                                                      t.xi_table);
  }
  if (fixed) {
    const cudaError_t e = cudaEventRecord(fixed, stream);
    if (e != cudaSuccess) return e;
  }
  if (z.total_blocks + z.total_st) {
    const unsigned blocks = (unsigned) ((z.total_blocks + z.total_st + 255) / 256);
    block_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.eslot, z.total_blocks,
                                            z.total_st, t.zblock, t.zst, t.hist_tab);
  }
  return cudaGetLastError();
}

// ---- peaks of a frame range (ebur128_prev_sample_peak / _prev_true_peak) ------
// max |x| and max |polyphase output| over track frames [first, first + count) of
// every channel: what libebur128 reports for the frames of the LAST add_frames
// call (the interpolator's history being the frames before them, zero before
// the start of the audio).  Exhaustive evaluation with the sweep's own tap order
// (lg_sweep.cuh: tp_frame); a thread keeps one channel and strides over frames.
template <int FMT, int TPF>
__global__ void __launch_bounds__(256)
range_peak_kernel(const void* __restrict__ pcm, uint32_t C, uint64_t first, uint64_t count,
                  uint32_t* __restrict__ out) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  const uint64_t n = count * C;
  // stride: a multiple of C, so that a thread's channel never changes
  const uint64_t stride = ((uint64_t) gridDim.x * blockDim.x + C - 1) / C * C;
  uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint32_t c = (uint32_t) (idx % C);
  auto sample = [&](long long f) -> float {
    if (f < 0) return 0.0f;
    if (FMT == FMT_S16) return (float) reinterpret_cast<const short*>(pcm)[(uint64_t) f * C + c];
    return reinterpret_cast<const float*>(pcm)[(uint64_t) f * C + c];
  };
  float sp = 0.0f, tp = 0.0f;
  for (; idx < n; idx += stride) {
    const long long t = (long long) (first + idx / C);
    float win[NT > 0 ? NT : 1];
    if (NT > 0) {
#pragma unroll
      for (int q = 0; q < NT; ++q) win[q] = sample(t - (NT - 1) + q);
      sp = fmaxf(sp, fabsf(win[NT - 1]));
      tp = fmaxf(tp, tp_frame<TPF>(win, NT - 1));
    } else {
      sp = fmaxf(sp, fabsf(sample(t)));
    }
  }
  // non-negative floats order like their bit patterns
  atomicMax(out + 2 * c, __float_as_uint(sp));
  atomicMax(out + 2 * c + 1, __float_as_uint(tp));
}

cudaError_t launch_range_peaks(const void* pcm, uint32_t format, uint32_t channels, uint64_t first,
                               uint64_t count, int tpf, uint32_t* out, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * channels * sizeof(uint32_t), stream);
  if (e != cudaSuccess || !count) return e;
  const uint64_t n = count * channels;
  const uint32_t blocks = (uint32_t) ((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
#define LG_RANGE(F, T) range_peak_kernel<F, T><<<blocks, 256, 0, stream>>>(pcm, channels, first, count, out)
  if (format == FMT_S16) {
    if (tpf == 4) LG_RANGE(FMT_S16, 4); else if (tpf == 2) LG_RANGE(FMT_S16, 2); else LG_RANGE(FMT_S16, 0);
  } else {
    if (tpf == 4) LG_RANGE(FMT_F32, 4); else if (tpf == 2) LG_RANGE(FMT_F32, 2); else LG_RANGE(FMT_F32, 0);
  }
#undef LG_RANGE
  return cudaGetLastError();
}

// 16-bit samples -> float in full-scale units (x / 32768: what ebur128_add_frames_short
// does to every sample), for a state that is then fed float frames.
__global__ void __launch_bounds__(256)
widen_s16_kernel(const short* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    out[i] = (float) in[i] * (1.0f / 32768.0f);
}

cudaError_t launch_widen_s16(const void* in, void* out, size_t n, cudaStream_t stream) {
  if (!n) return cudaSuccess;
  const size_t blocks = (n + 255) / 256;
  widen_s16_kernel<<<(unsigned) (blocks < 148u * 16u ? blocks : 148u * 16u), 256, 0, stream>>>(
      (const short*) in, (float*) out, n);
  return cudaGetLastError();
}

uint32_t query_cluster_size(uint64_t max_gating_blocks) {
  // one more CTA (256 threads) per 4 k gating blocks of the largest query (about seven minutes of audio)
  static const uint64_t per_cta = [] {
    const char* e = getenv("LOUDGAIN_B200_QUERY_BLOCKS_PER_CTA");   // tuning
    const long long v = e ? atoll(e) : 0;
    return (uint64_t) (v > 0 ? v : 4096);
  }();
  uint64_t r = max_gating_blocks / per_cta;
  return (uint32_t) (r < 1 ? 1 : (r > (uint64_t) kMaxQueryCluster ? kMaxQueryCluster : r));
}

cudaError_t launch_queries(const BlockList* lists, const Query* queries, const uint32_t* members,
                           uint32_t nqueries, double abs_gate, QueryResult* results,
                           cudaStream_t stream, uint32_t cluster, int part) {
  if (!nqueries) return cudaSuccess;
  if (part == 2) cluster = 1;                  // the range is one CTA's work
  if (cluster < 1) cluster = 1;
  if (cluster > (uint32_t) kMaxQueryCluster) cluster = kMaxQueryCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nqueries * cluster);
  cfg.blockDim = dim3(query_launch_threads());
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  if (part == 1) return cudaLaunchKernelEx(&cfg, query_kernel<1>, lists, queries, members, abs_gate, results);
  if (part == 2) return cudaLaunchKernelEx(&cfg, query_kernel<2>, lists, queries, members, abs_gate, results);
  return cudaLaunchKernelEx(&cfg, query_kernel<0>, lists, queries, members, abs_gate, results);
}

}  // namespace lg
