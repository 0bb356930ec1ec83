// lg_kernels.cu -- sm_100a kernels of the loudness path and their launchers.
//
//   sweep_kernel    fused K-weight + chunk energy + sample peak + true peak
//                   (what ebur128_add_frames_short does per call in the
//                   reference path, /root/reference/src/scan.c:448)
//   fixup_kernel    FP64 state carry + energy correction per chunk
//   slot_kernel     100 ms slot energies (channel-weighted)
//   block_kernel    400 ms gating blocks and 3 s short-term blocks
//   query_kernel    gated integrated loudness and loudness range over a set
//                   of tracks (ebur128_loudness_global[_multiple],
//                   ebur128_loudness_range[_multiple]: scan.c:294,297,383,388)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lg_common.h"
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
//    (coalesced), kRing stages of 24 frames deep; rows start on 16-byte
//    boundaries of the track and the row stride is an odd number of units, so
//    128-bit shared loads of different rows do not collide;
//  * every lane filters its channel (FP32) and accumulates energy and the
//    correction cross terms;
//  * true peak: a 12-frame window is evaluated only if ||c||_1 * max|x| over
//    the window exceeds what the channel's peak is already known to reach
//    (own maximum so far, the warp's, and the track-wide value other warps
//    have published).  Such candidate windows are copied into a per-warp
//    queue and evaluated 32 at a time, one window per lane, so the FIR runs
//    dense instead of divergent.  The maximum is identical to evaluating
//    every window (see lg_sweep.cuh).

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
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

__device__ __forceinline__ int sext_half(uint32_t w, uint32_t sel) {
  int r;
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "r"(sel));
  return r;
}

// Raw samples of one iteration of the lane's channel, from shared memory.
// rowp -> first frame of the iteration in the lane's row.
template <int FMT>
__device__ __forceinline__ void smem_load_iter(const unsigned char* rowp, uint32_t fb, bool stereo,
                                               uint32_t ch, float* x) {
  if (FMT == FMT_S16 && stereo) {
    // A frame is one 32-bit word (L | R << 16).  One PRMT sign-extends the
    // lane's half (selector nibble bit 3 = replicate the sign of that byte).
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
  } else if (FMT == FMT_S16) {
    const unsigned char* q = rowp + ch * 2u;
#pragma unroll
    for (int i = 0; i < kIter; ++i) x[i] = (float) *reinterpret_cast<const short*>(q + i * fb);
  } else {
    const unsigned char* q = rowp + ch * 4u;
#pragma unroll
    for (int i = 0; i < kIter; ++i) x[i] = *reinterpret_cast<const float*>(q + i * fb);
  }
}

#ifndef LG_SWEEP_MINBLOCKS
#define LG_SWEEP_MINBLOCKS 4
#endif

constexpr uint32_t kQueue = 32;     // candidate windows a warp can hold (one dense round)

template <int TPF>
__host__ __device__ constexpr uint32_t queue_entry_bytes() {
  // window floats + one meta word, rounded up to an odd number of 16-byte units
  return ((((TpTraits<TPF>::kTaps + kIter) * 4u + 4u + 15u) >> 4) | 1u) << 4;
}

// Evaluates the n <= 32 queued windows, one per lane.
template <int TPF>
__device__ __forceinline__ void flush_round(const unsigned char* queue, uint32_t n, uint32_t lane,
                                            uint32_t* tpq) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  constexpr uint32_t EB = queue_entry_bytes<TPF>();
  __syncwarp();                                   // entries written by other lanes are visible
  if (lane < n) {
    const unsigned char* e = queue + lane * EB;
    float win[NT + kIter];
    const float4* p = reinterpret_cast<const float4*>(e);
#pragma unroll
    for (int i = 0; i < (NT + kIter) / 4; ++i) {
      const float4 v = p[i];
      win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
    }
    const uint32_t slot = *reinterpret_cast<const uint32_t*>(e + (NT + kIter) * 4);
    atomicMax(tpq + slot, __float_as_uint(tp_window<TPF>(win)));
  }
  __syncwarp();
}

// Peaks of a track-end iteration, frame by frame.  Rare (last warp of a
// track), so it is kept out of line to keep the hot loop small; everything
// goes in and out by value so that the lane state stays in registers.
template <int TPF>
struct SlowPeakArgs {
  float win[TpTraits<TPF>::kTaps + kIter];
  int f0, f_lo, f_tp;
  float sp, tp;
};

template <int TPF>
__device__ __noinline__ float2 peaks_masked_slow(const SlowPeakArgs<TPF> a) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  float sp = a.sp, tp = a.tp;
#pragma unroll 1
  for (int i = 0; i < kIter; ++i) {
    const int f = a.f0 + i;
    if (f >= a.f_lo && f < a.f_tp) {
      sp = fmaxf(sp, fabsf(a.win[NT + i]));
      if (NT > 0) {
        // same arithmetic as tp_frame, with a run-time window position
        float m = 0.0f;
        if (TPF == 4) {
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            float acc = 0.0f;
#pragma unroll
            for (int t = 0; t < 12; ++t) acc = fmaf(a.win[NT + i - t], kTp4f[p][t], acc);
            m = fmaxf(m, fabsf(acc));
          }
        } else {
          float acc = 0.0f;
#pragma unroll
          for (int t = 0; t < 24; ++t) acc = fmaf(a.win[NT + i - t], kTp2f[0][t], acc);
          m = fabsf(acc);
        }
        tp = fmaxf(tp, m);
      }
    }
  }
  return make_float2(sp, tp);
}

template <int FMT, int TPF, int KMAX>
__global__ void __launch_bounds__(kSweepThreads, LG_SWEEP_MINBLOCKS)
sweep_kernel(const __grid_constant__ SweepParams P) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const uint32_t wic = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wic;
  if (warp >= P.nwarps) return;                // whole warps leave; no CTA barrier below
  unsigned char* sm = smem_all + wic * P.warp_smem;
  const uint32_t sm_addr = (uint32_t) __cvta_generic_to_shared(sm);

  const WarpWork ww = P.work[warp];
  const Track& tr = P.tracks[ww.track];
  const int W = P.W, L = P.L;
  const uint32_t C = P.channels, fb = P.fb;
  const long long frames = (long long) tr.frames;
  const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm);

  const uint32_t slot = lane / P.lpc;
  const uint32_t chl = lane - slot * P.lpc;    // channel within the warp's group
  const uint32_t ch = ww.ch_base + chl;
  const uint32_t chunk = ww.first_chunk + slot;
  const bool compute = slot < P.cpw && ch < C;
  const bool active = compute && chunk < tr.nchunks;
  const bool stereo = C == 2;

  // ---- staging.  Per stage and row one contiguous piece of kStageFrames
  // frames, moved with 16-byte cp.async copies: copy k of a lane moves unit
  // (idx % units) of row (idx / units), idx = lane + 32 k, so consecutive
  // lanes fetch consecutive units of one row.  Warps at a track boundary use
  // the zero-filling form.  (TMA bulk copies were tried: UBLKCP is issued from
  // uniform registers, i.e. one row at a time, and the rows are too short for
  // that to pay -- see DESIGN.md.)
  const LaneGeom g0 = lane_geometry(frames, L, W, P.aq, ww.first_chunk);
  const long long warp_byte0 = g0.a * (long long) fb;          // may be negative
  const long long track_bytes = frames * (long long) fb;
  const bool interior = ww.interior != 0;
  int32_t soff[KMAX];      // source byte offset at stage 0, relative to warp_byte0
  uint32_t doff[KMAX];     // destination byte offset inside a stage buffer
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const uint32_t idx = lane + 32u * k;
    const uint32_t row = idx / P.units;
    const uint32_t unit = idx - row * P.units;
    const LaneGeom gr = lane_geometry(frames, L, W, P.aq, ww.first_chunk + row);
    soff[k] = (int32_t) ((gr.a - g0.a) * (long long) fb) + (int32_t) (unit << 4);
    doff[k] = idx < P.ncopies ? row * P.row_stride + (unit << 4) : 0xffffffffu;
  }

  auto prefetch = [&](uint32_t stage) {
    const uint32_t dst0 = sm_addr + (stage % kRing) * P.stage_bytes;
    const long long adv = (long long) stage * P.stage_row_bytes;
    if (interior) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (doff[k] != 0xffffffffu) cp_async16(dst0 + doff[k], pcm + (warp_byte0 + adv + soff[k]));
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (doff[k] == 0xffffffffu) continue;
        const long long g = warp_byte0 + adv + soff[k];
        long long ok = g < 0 ? 0 : track_bytes - g;
        ok = ok < 0 ? 0 : (ok > 16 ? 16 : ok);
        cp_async16_zfill(dst0 + doff[k], pcm + (ok ? g : 0), (uint32_t) ok);
      }
    }
  };
  auto stage_wait = [&](uint32_t) {
    cp_async_wait<kRing - 2>();
    __syncwarp();                // everyone's data has landed; the previous stage is consumed
  };

  // ---- candidate queue and per-channel true-peak cells of this warp
  constexpr uint32_t EB = queue_entry_bytes<TPF>();
  unsigned char* queue = sm + P.ring_bytes;
  uint32_t* tpq = reinterpret_cast<uint32_t*>(queue + (NT > 0 ? kQueue * EB : 0));
  tpq[lane] = 0u;
  uint32_t q_count = 0;
  // Frames at or beyond this lane-local index may lie past the end of the
  // track for some lane of the warp: true peak is then masked frame by frame.
  const LaneGeom glast = lane_geometry(frames, L, W, P.aq, ww.first_chunk + P.cpw - 1);
  const long long tp_safe_ll = frames - glast.a;
  const int tp_safe = tp_safe_ll > 0x3fffffff ? 0x3fffffff : (int) tp_safe_ll;

  // ---- lane state
  const LaneGeom geo = lane_geometry(frames, L, W, P.aq, chunk);
  LaneCtx<TPF> c;
  lane_init(c, W, L, geo);
  const unsigned char* my_row = sm + slot * P.row_stride;
  uint32_t* my_peak = P.peaks + 2 * (tr.peak_base + (compute ? ch : 0u));
  const unsigned peers = __match_any_sync(0xffffffffu, compute ? chl : 0xffffu);
  const bool leader = compute && lane == (uint32_t) (__ffs(peers) - 1);
  float thr = 0.0f;        // the channel's peak is known to reach at least this (raw units)
  float published = 0.0f;
  uint2 seen = make_uint2(0u, 0u);   // what other warps had published, as of the last poll

  // Queues the windows (hist, x) of the lanes whose flag is set, for the two
  // iterations of a pair; evaluates the queue first if it cannot take them.
  // The only place the queue is written and flushed from (the loop is kept
  // rolled so that the evaluation code exists once).
  auto enqueue2 = [&](bool cand0, const float* hist0, const float* xa, bool cand1,
                      const float* hist1, const float* xb) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const bool cand = half ? cand1 : cand0;
      const unsigned mask = __ballot_sync(0xffffffffu, cand);
      if (mask == 0u) continue;
      const uint32_t npush = __popc(mask);
      if (q_count + npush > kQueue) {
        flush_round<TPF>(queue, q_count, lane, tpq);
        q_count = 0;
        thr = fmaxf(thr, __uint_as_float(tpq[chl]));
      }
      if (cand) {
        const uint32_t pos = q_count + __popc(mask & ((1u << lane) - 1u));
        float4* e = reinterpret_cast<float4*>(queue + pos * EB);
#pragma unroll
        for (int i = 0; i < NT / 4; ++i)
          e[i] = half ? make_float4(hist1[4 * i], hist1[4 * i + 1], hist1[4 * i + 2], hist1[4 * i + 3])
                      : make_float4(hist0[4 * i], hist0[4 * i + 1], hist0[4 * i + 2], hist0[4 * i + 3]);
#pragma unroll
        for (int i = 0; i < kIter / 4; ++i)
          e[NT / 4 + i] = half ? make_float4(xb[4 * i], xb[4 * i + 1], xb[4 * i + 2], xb[4 * i + 3])
                               : make_float4(xa[4 * i], xa[4 * i + 1], xa[4 * i + 2], xa[4 * i + 3]);
        *reinterpret_cast<uint32_t*>(e + (NT + kIter) / 4) = chl;
      }
      q_count += npush;
    }
  };

  const uint32_t niters = (uint32_t) P.niters;
  const uint32_t npairs = (niters + 1u) / 2u;
  const uint32_t nstages = (npairs + kPairsPerStage - 1) / kPairsPerStage;
  // Pairs whose two iterations are both "fast" for every lane of the warp.
  const int lfast = ww.lmin_valid < L ? ww.lmin_valid : L;
  const uint32_t fast_lo = (uint32_t) ((W + P.aq - 1 + kPairFrames - 1) / kPairFrames);
  const uint32_t fast_hi = (uint32_t) ((W + lfast) / kPairFrames);     // exclusive
#pragma unroll
  for (int i = 0; i < kRing - 1; ++i) {
    if ((uint32_t) i < nstages) prefetch(i);
    cp_async_commit();
  }

  for (uint32_t s = 0; s < nstages; ++s) {
    stage_wait(s);
    if (s + kRing - 1 < nstages) prefetch(s + kRing - 1);
    cp_async_commit();
    // What other warps have published for this channel.  Polled sparingly (the
    // cells are hot) and only folded into the bound at the next poll, so the
    // L2 round trip is off the critical path.
    if (NT > 0 && compute && (s & 3u) == 0u) {
      thr = fmaxf(thr, __uint_as_float(seen.x > seen.y ? seen.x : seen.y));
      seen = __ldcg(reinterpret_cast<const uint2*>(my_peak));
    }
    const unsigned char* sbuf = my_row + (s % kRing) * P.stage_bytes;
#pragma unroll 1
    for (uint32_t pr = 0; pr < (uint32_t) kPairsPerStage; ++pr) {
      const uint32_t pair = s * kPairsPerStage + pr;
      if (pair >= npairs) break;
      const unsigned char* buf = sbuf + pr * kPairFrames * fb;
      // Both iterations of the pair leave their samples in x0 / x1 and their
      // candidate flags in c0 / c1; the queue is fed from one place below.
      float x0[kIter], x1[kIter];
      bool c0 = false, c1 = false;
      bool second = true;               // the pair has a second iteration
      if (pair >= fast_lo && pair < fast_hi && tp_safe >= (int) ((pair + 1) * kPairFrames)) {
        // ---- both iterations fast: straight-line code, loads first
        const int f0 = (int) (pair * kPairFrames);
        if (compute) {
          smem_load_iter<FMT>(buf, fb, stereo, ch, x0);
          smem_load_iter<FMT>(buf + kIter * fb, fb, stereo, ch, x1);
#if defined(LG_ABLATE_COMPUTE)
          const float m0 = max_abs12(x0), m1 = max_abs12(x1);
          c.sp = fmaxf(c.sp, fmaxf(m0, m1));
#else
          const float m0 = iter_fast_energy<TPF>(c, P, x0, f0);
          const float m1 = iter_fast_energy<TPF>(c, P, x1, f0 + kIter);
#endif
#if !defined(LG_ABLATE_TP)
          if (NT > 0) {
            const float floor_ = fmaxf(thr, c.sp);
            c0 = P.tp_bound * fmaxf(c.mprev, m0) > floor_;
            c1 = P.tp_bound * fmaxf(m0, m1) > floor_;
            c.mprev = m1;
          }
#endif
        }
      } else {
        // ---- chunk edges: warm-up and masked iterations, one at a time
        second = pair * 2u + 1u < niters;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          if (it == 1 && !second) break;
          float* x = it == 0 ? x0 : x1;
          const int f0 = (int) (pair * kPairFrames) + it * kIter;
          const int kind = iter_kind(f0, W, P.aq, L, ww.lmin_valid);
          bool cand = false;
          if (compute) {
            smem_load_iter<FMT>(buf + it * kIter * fb, fb, stereo, ch, x);
            if (kind == ITER_WARM) {
#pragma unroll
              for (int i = 0; i < kIter; ++i) (void) k_step(c.st, x[i], P);
              c.mprev = max_abs12(x);
              c.pd = c.st.d1; c.pw = c.st.w2;
            } else {
              float m;
              bool safe = true;
              if (kind == ITER_FAST) {
                m = iter_fast_energy<TPF>(c, P, x, f0);
              } else {
                iter_masked_energy<TPF>(c, P, x, f0);
                m = max_abs12(x);
                safe = f0 + kIter <= tp_safe;
                if (safe) c.sp = fmaxf(c.sp, m);
                else {                                       // track end: frame by frame
                  SlowPeakArgs<TPF> a;
                  // history of this iteration: c.hist, shifted by x0 for the second one
#pragma unroll
                  for (int i = 0; i < NT; ++i)
                    a.win[i] = it == 0 ? c.hist[i]
                                       : (i < NT - kIter ? c.hist[i + kIter] : x0[i - (NT - kIter)]);
#pragma unroll
                  for (int i = 0; i < kIter; ++i) a.win[NT + i] = x[i];
                  a.f0 = f0; a.f_lo = c.f_lo; a.f_tp = c.f_tp; a.sp = c.sp; a.tp = c.tp;
                  const float2 r = peaks_masked_slow<TPF>(a);
                  c.sp = r.x; c.tp = r.y;
                }
              }
              if (NT > 0) {
                cand = safe && P.tp_bound * fmaxf(c.mprev, m) > fmaxf(thr, c.sp);
                c.mprev = m;
              }
            }
          }
          if (it == 0) c0 = cand; else c1 = cand;
        }
        if (!second) {
#pragma unroll
          for (int i = 0; i < kIter; ++i) x1[i] = 0.0f;
        }
      }
      if (NT > 0) {
        // history of the second iteration = last NT frames of (hist, x0)
        float h1[NT > 0 ? NT : 1];
#pragma unroll
        for (int i = 0; i < NT - kIter; ++i) h1[i] = c.hist[i + kIter];
#pragma unroll
        for (int i = 0; i < kIter; ++i) h1[NT - kIter + i] = x0[i];
        enqueue2(c0, c.hist, x0, c1, h1, x1);
        if (compute) {
          if (second) {
#pragma unroll
            for (int i = 0; i < NT - kIter; ++i) c.hist[i] = h1[i + kIter];
#pragma unroll
            for (int i = 0; i < kIter; ++i) c.hist[NT - kIter + i] = x1[i];
          } else {
#pragma unroll
            for (int i = 0; i < NT; ++i) c.hist[i] = h1[i];
          }
        }
      }
    }
    if (NT > 0 && compute && (s & 3u) == 3u) {
      // publish what this warp knows, if it is news
      const float theirs = __uint_as_float(seen.x > seen.y ? seen.x : seen.y);
      const float mine = fmaxf(fmaxf(c.sp, c.tp), __uint_as_float(tpq[chl]));
      thr = fmaxf(thr, mine);
      if (mine > published && mine > theirs) {
        // goes into the true-peak cell: the reported true peak is the max of
        // both cells anyway (ebur128_true_peak folds the sample peak in)
        atomicMax(my_peak + 1, __float_as_uint(mine));
        published = mine;
      }
    }
  }
  cp_async_wait<0>();
  if (NT > 0 && q_count) flush_round<TPF>(queue, q_count, lane, tpq);
  __syncwarp();

  if (active) {
    ChunkRec v;
    v.e0 = c.e0; v.yr = c.yr; v.yi = c.yi;
    v.pd = c.pd; v.pw = c.pw; v.qd = c.qd; v.qw = c.qw;
    P.recs[tr.rec_base + (uint64_t) chunk * C + ch] = v;
  }
  // Peaks: non-negative floats order like their bit patterns.  Reduce over
  // the lanes of the warp that hold the same channel, one atomic each.
  uint32_t spb = __reduce_max_sync(peers, active ? __float_as_uint(c.sp) : 0u);
  uint32_t tpb = __reduce_max_sync(peers, active ? __float_as_uint(c.tp) : 0u);
  if (leader) {
    if (NT > 0) { const uint32_t q = tpq[chl]; tpb = tpb > q ? tpb : q; }
    atomicMax(my_peak + 0, spb);
    atomicMax(my_peak + 1, tpb);
  }
}

template <int FMT, int TPF, int KMAX>
static cudaError_t launch_sweep_k(const SweepParams& p, cudaStream_t stream) {
  const uint32_t wpb = kSweepThreads / 32;
  const uint32_t blocks = (p.nwarps + wpb - 1) / wpb;
  const size_t smem = (size_t) p.warp_smem * wpb;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<FMT, TPF, KMAX>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sweep_kernel<FMT, TPF, KMAX>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sweep_kernel<FMT, TPF, KMAX><<<blocks, kSweepThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int FMT, int TPF>
static cudaError_t launch_sweep_t(const SweepParams& p, uint32_t kmax, cudaStream_t stream) {
  if (kmax <= 3) return launch_sweep_k<FMT, TPF, 3>(p, stream);
  if (kmax <= 6) return launch_sweep_k<FMT, TPF, 6>(p, stream);
  return launch_sweep_k<FMT, TPF, 12>(p, stream);
}

cudaError_t launch_sweep(const SweepParams& p, uint32_t format, int tpf, uint32_t kmax,
                         cudaStream_t stream) {
  if (p.nwarps == 0) return cudaSuccess;
  if (format == FMT_S16) {
    if (tpf == 4) return launch_sweep_t<FMT_S16, 4>(p, kmax, stream);
    if (tpf == 2) return launch_sweep_t<FMT_S16, 2>(p, kmax, stream);
    return launch_sweep_t<FMT_S16, 0>(p, kmax, stream);
  }
  if (tpf == 4) return launch_sweep_t<FMT_F32, 4>(p, kmax, stream);
  if (tpf == 2) return launch_sweep_t<FMT_F32, 2>(p, kmax, stream);
  return launch_sweep_t<FMT_F32, 0>(p, kmax, stream);
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

__global__ void __launch_bounds__(256)
fixup_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
             const CoefSet* __restrict__ coefs, const ChunkRec* __restrict__ recs,
             uint64_t total_recs, double* __restrict__ echunk) {
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total_recs) return;
  const uint32_t ti = find_track(tracks, ntracks, r, [](const Track& t) { return t.rec_base; });
  const Track& tr = tracks[ti];
  const CoefSet& cs = coefs[tr.coef];
  const uint64_t local = r - tr.rec_base;
  const uint64_t chunk = local / tr.channels;
  const uint32_t ch = (uint32_t) (local - chunk * tr.channels);
  if (chunk >= (uint64_t) tr.nslots * cs.k) return;   // tail chunks carry peaks only
  const LaneGeom geo = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, (long long) chunk);
  echunk[r] = chunk_true_energy(cs, recs + tr.rec_base + ch, tr.channels, (long long) chunk,
                                geo.o, 31 - __clz((int) tr.aq));
}

__global__ void __launch_bounds__(256)
slot_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
            const CoefSet* __restrict__ coefs, const double* __restrict__ echunk,
            uint64_t total_slots, double* __restrict__ eslot) {
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total_slots) return;
  const uint32_t ti = find_track(tracks, ntracks, s, [](const Track& t) { return t.slot_base; });
  const Track& tr = tracks[ti];
  eslot[s] = slot_energy(tr, coefs[tr.coef], echunk, (uint32_t) (s - tr.slot_base));
}

__global__ void __launch_bounds__(256)
block_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
             const CoefSet* __restrict__ coefs, const double* __restrict__ eslot,
             uint64_t total_blocks, uint64_t total_st, double* __restrict__ zblock,
             double* __restrict__ zst) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total_blocks) {
    const uint32_t ti = find_track(tracks, ntracks, i, [](const Track& t) { return t.block_base; });
    const Track& tr = tracks[ti];
    zblock[i] = gating_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (i - tr.block_base));
  } else if (i < total_blocks + total_st) {
    const uint64_t j = i - total_blocks;
    const uint32_t ti = find_track(tracks, ntracks, j, [](const Track& t) { return t.st_base; });
    const Track& tr = tracks[ti];
    zst[j] = shortterm_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (j - tr.st_base));
  }
}

// ------------------------------------------------------------- reductions

constexpr int kQueryThreads = 512;

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

// The k-th smallest (0-based) values, for two ranks at once, among the member
// short-term energies >= floor_e.  Positive doubles order like their 64-bit
// patterns: 8 passes of 8-bit radix select over the (L2-resident) values,
// one 256-bin histogram per rank and pass, bin search by a block-wide scan.
struct SelectState {
  unsigned long long prefix[2];
  unsigned long long k[2];
};

__device__ void select_two(const BlockList* lists, const uint32_t* members, uint32_t count,
                           double floor_e, unsigned long long k_lo, unsigned long long k_hi,
                           unsigned int* hist /* [2][256] */, unsigned int* wsum /* [2][8] */,
                           SelectState* st, double* out_lo, double* out_hi) {
  if (threadIdx.x == 0) { st->prefix[0] = st->prefix[1] = 0; st->k[0] = k_lo; st->k[1] = k_hi; }
  unsigned long long mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    for (uint32_t m = 0; m < count; ++m) {
      const BlockList bl = lists[members[m]];
      for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x) {
        const double e = bl.st[i];
        if (!(e >= floor_e)) continue;
        const unsigned long long bits = (unsigned long long) __double_as_longlong(e);
        const unsigned int digit = (unsigned int) ((bits >> shift) & 0xffull);
        if ((bits & mask) == p0) atomicAdd(&hist[digit], 1u);
        if ((bits & mask) == p1) atomicAdd(&hist[256 + digit], 1u);
      }
    }
    __syncthreads();
    // threads 0..255 own the bins of rank 0, 256..511 those of rank 1
    const int which = threadIdx.x >> 8, bin = threadIdx.x & 255;
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 7;
    const unsigned int cnt = hist[threadIdx.x];
    unsigned int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[which * 8 + w] = inc;
    __syncthreads();
    unsigned int base = 0;
    for (int j = 0; j < w; ++j) base += wsum[which * 8 + j];
    const unsigned long long excl = (unsigned long long) base + inc - cnt;
    const unsigned long long kk = st->k[which];
    __syncthreads();
    if (cnt && excl <= kk && kk < excl + cnt) {
      st->k[which] = kk - excl;
      st->prefix[which] |= (unsigned long long) bin << shift;
    }
    mask |= 0xffull << shift;
    __syncthreads();
  }
  *out_lo = __longlong_as_double((long long) st->prefix[0]);
  *out_hi = __longlong_as_double((long long) st->prefix[1]);
}

__global__ void __launch_bounds__(kQueryThreads)
query_kernel(const BlockList* __restrict__ lists, const Query* __restrict__ queries,
             const uint32_t* __restrict__ members, double abs_gate,
             QueryResult* __restrict__ results) {
  __shared__ SumCount scratch[32];
  __shared__ unsigned int hist[512];
  __shared__ unsigned int wsum[16];
  __shared__ SelectState sel;
  const Query q = queries[blockIdx.x];
  const uint32_t* mem = members + q.first;
  QueryResult res;
  res.loudness = -HUGE_VAL; res.range = 0.0; res.rel_thr = 0.0;
  res.sum1 = res.sum2 = 0.0; res.n1 = res.n2 = res.nst = 0;

  // ---- integrated loudness: absolute gate, then relative gate at -10 LU
  double s = 0.0;
  unsigned long long n = 0;
  for (uint32_t m = 0; m < q.count; ++m) {
    const BlockList bl = lists[mem[m]];
    for (uint32_t i = threadIdx.x; i < bl.nz; i += blockDim.x) {
      const double e = bl.z[i];
      if (e >= abs_gate) { s += e; ++n; }
    }
  }
  SumCount a = block_sum_count(s, n, scratch);
  res.sum1 = a.s; res.n1 = a.n;
  if (a.n) {
    const double thr = a.s / (double) a.n * 0.1;
    res.rel_thr = thr;
    s = 0.0; n = 0;
    for (uint32_t m = 0; m < q.count; ++m) {
      const BlockList bl = lists[mem[m]];
      for (uint32_t i = threadIdx.x; i < bl.nz; i += blockDim.x) {
        const double e = bl.z[i];
        if (e >= abs_gate && e >= thr) { s += e; ++n; }
      }
    }
    SumCount b = block_sum_count(s, n, scratch);
    res.sum2 = b.s; res.n2 = b.n;
    if (b.n) res.loudness = energy_to_lufs(b.s / (double) b.n);
  }

  // ---- loudness range: -20 LU relative gate on short-term energies, then the
  //      10th / 95th percentile by rank
  s = 0.0; n = 0;
  for (uint32_t m = 0; m < q.count; ++m) {
    const BlockList bl = lists[mem[m]];
    for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x) {
      const double e = bl.st[i];
      if (e >= abs_gate) { s += e; ++n; }
    }
  }
  a = block_sum_count(s, n, scratch);
  res.nst = a.n;
  if (a.n) {
    double floor_e = a.s / (double) a.n * 0.01;
    if (floor_e < abs_gate) floor_e = abs_gate;
    s = 0.0; n = 0;
    for (uint32_t m = 0; m < q.count; ++m) {
      const BlockList bl = lists[mem[m]];
      for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x)
        if (bl.st[i] >= floor_e) ++n;
    }
    const SumCount c = block_sum_count(0.0, n, scratch);
    if (c.n) {
      const unsigned long long k_hi = (unsigned long long) ((double) (c.n - 1) * 0.95 + 0.5);
      const unsigned long long k_lo = (unsigned long long) ((double) (c.n - 1) * 0.1 + 0.5);
      double lo, hi;
      select_two(lists, mem, q.count, floor_e, k_lo, k_hi, hist, wsum, &sel, &lo, &hi);
      res.range = energy_to_lufs(hi) - energy_to_lufs(lo);
    }
  }
  if (threadIdx.x == 0) results[blockIdx.x] = res;
}

cudaError_t launch_post(const DeviceTables& t, const PostSizes& z, cudaStream_t stream) {
  if (z.total_recs) {
    const unsigned blocks = (unsigned) ((z.total_recs + 255) / 256);
    fixup_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.recs, z.total_recs,
                                            t.echunk);
  }
  if (z.total_slots) {
    const unsigned blocks = (unsigned) ((z.total_slots + 255) / 256);
    slot_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.echunk, z.total_slots,
                                           t.eslot);
  }
  if (z.total_blocks + z.total_st) {
    const unsigned blocks = (unsigned) ((z.total_blocks + z.total_st + 255) / 256);
    block_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.eslot, z.total_blocks,
                                            z.total_st, t.zblock, t.zst);
  }
  return cudaGetLastError();
}

cudaError_t launch_queries(const BlockList* lists, const Query* queries, const uint32_t* members,
                           uint32_t nqueries, double abs_gate, QueryResult* results,
                           cudaStream_t stream) {
  if (!nqueries) return cudaSuccess;
  query_kernel<<<nqueries, kQueryThreads, 0, stream>>>(lists, queries, members, abs_gate, results);
  return cudaGetLastError();
}

}  // namespace lg
