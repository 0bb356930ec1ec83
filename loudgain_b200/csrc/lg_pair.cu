// lg_pair.cu -- the packed sweep and its true-peak pass, for tracks with an
// even channel count (stereo above all: what /root/reference/src/scan.c:448
// feeds through ebur128_add_frames_short for almost every file).
//
// Same algorithm and the same per-channel FP32 operation sequence as
// sweep_kernel (lg_kernels.cu, lg_sweep.cuh) -- the results are bit-identical --
// but one lane owns one chunk and one channel PAIR (2j, 2j+1) and runs both
// channels through Blackwell's packed FP32 instructions (FFMA2 / FADD2: two
// independent IEEE FMAs per issue slot, the broadcast filter constant coming
// from a uniform register).  The scalar kernel is bound by instruction issue
// (ncu: 84 % of the issue slots busy at 41 % of HBM bandwidth); packing halves
// the issue slots of the 11 FMA-pipe operations per sample.
//
// Stereo rows reach shared memory as 2-D TMA tiles (cp.async.bulk.tensor.2d,
// one box per chunk class and stage, issued by one elected lane, completion on
// an mbarrier); track ends and other channel counts use 16-byte cp.async.
//
// A 16-bit stereo frame is one 32-bit word, so the sample peak of both
// channels is tracked on the raw words with packed 16-bit integer min/max
// (VIMNMX3.S16x2, two frames per instruction).
//
// True peak: the sweep leaves max |x| per (lane, iteration pair, channel) as a
// 16-bit code; tp_scan_pair_kernel + tp_eval_pair_kernel evaluate the polyphase
// FIR only on the pairs whose bound ||c||_1 * max|x| exceeds the channel's
// final sample peak.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lg_common.h"
#include "lg_device.cuh"
#include "lg_kernels.h"
#include "lg_packed.cuh"
#include "lg_sweep.cuh"

namespace lg {

// How a lane finds the frames of its channel pair in a staged row.
enum PairLayout : int {
  PAIR_S16_STEREO = 0,   // frame = one 32-bit word, 128-bit loads of 4 frames
  PAIR_F32_STEREO = 1,   // frame = 8 bytes, 128-bit loads of 2 frames
  PAIR_S16_EVEN = 2,     // 32-bit load per frame, stride = frame bytes
  PAIR_F32_EVEN = 3,     // 64-bit load per frame
};

template <int LAYOUT>
__host__ __device__ constexpr int pair_format() {
  return (LAYOUT == PAIR_S16_STEREO || LAYOUT == PAIR_S16_EVEN) ? (int) FMT_S16 : (int) FMT_F32;
}

// Everything one lane carries through its chunk: .x = channel 2j, .y = 2j+1.
struct PairCtx {
  float2 xp, d1, w1, w2, v1, v2; // filter state (lg_sweep.cuh: KState)
  float2 yr, yi;                 // running mode sums
  double e0x, e0y;
  float2 pd, pw, qd, qw;         // state snapshots
  int f_lo, f_hi;
};

// lg_sweep.cuh: iter_fast.
__device__ __forceinline__ void iter_fast2(PairCtx& c, const SweepParams& k, const float2* x, int f0) {
  float2 e = bc2(0.0f), sr = bc2(0.0f), si = bc2(0.0f);
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const float2 y = k_step2(c, x[i], k);
    e = __ffma2_rn(y, y, e);
    sr = __ffma2_rn(y, bc2(k.lam_re[i]), sr);
    si = __ffma2_rn(y, bc2(k.lam_im[i]), si);
  }
  c.e0x += (double) e.x;
  c.e0y += (double) e.y;
  // mode_accumulate: fma(-yi, rot_im, sr) == fma(yi, -rot_im, sr)
  const float2 nr = __ffma2_rn(c.yr, bc2(k.rot_re), __ffma2_rn(c.yi, bc2(-k.rot_im), sr));
  const float2 ni = __ffma2_rn(c.yr, bc2(k.rot_im), __ffma2_rn(c.yi, bc2(k.rot_re), si));
  c.yr = nr;
  c.yi = ni;
  if (f0 + kIter == c.f_hi) { c.qd = c.d1; c.qw = c.w2; }
}

// lg_sweep.cuh: iter_warm (filter state only).
__device__ __forceinline__ void iter_warm2(PairCtx& c, const SweepParams& k, const float2* x, bool first) {
  if (first) c.xp = x[0];      // lg_sweep.cuh: lane_start
#pragma unroll
  for (int i = 0; i < kIter; ++i) (void) k_step2(c, x[i], k);
  c.pd = c.d1; c.pw = c.w2;
}

// lg_sweep.cuh: iter_masked, both channels at once (the lane's two channels
// share the energy range).  Frames outside [f_lo, f_hi) only advance the
// filter; the snapshots are taken where the scalar code takes them, and the
// accumulation is the same FMA sequence, so the sums are bit-identical.
__device__ __forceinline__ void iter_masked2(PairCtx& c, const SweepParams& k, const float2* x, int f0) {
  float2 e = bc2(0.0f), sr = bc2(0.0f), si = bc2(0.0f);
  const int ilo = c.f_lo - f0, ihi = c.f_hi - f0;
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    if (i == ilo) { c.pd = c.d1; c.pw = c.w2; }
    const float2 y = k_step2(c, x[i], k);
    if (i >= ilo && i < ihi) {
      e = __ffma2_rn(y, y, e);
      sr = __ffma2_rn(y, bc2(k.lam_re[i]), sr);
      si = __ffma2_rn(y, bc2(k.lam_im[i]), si);
    }
    if (i + 1 == ihi) { c.qd = c.d1; c.qw = c.w2; }
  }
  c.e0x += (double) e.x;
  c.e0y += (double) e.y;
  const float2 nr = __ffma2_rn(c.yr, bc2(k.rot_re), __ffma2_rn(c.yi, bc2(-k.rot_im), sr));
  const float2 ni = __ffma2_rn(c.yr, bc2(k.rot_im), __ffma2_rn(c.yi, bc2(k.rot_re), si));
  c.yr = nr;
  c.yi = ni;
}

// Per-pair maxima of the lane's two channels.  16-bit input: packed signed
// max / min of the raw frame words; float input: two running |x| maxima.
// One 16-bit stereo frame word -> (left, right) as floats: PRMT sign-extends a
// half, I2FP converts (both ALU pipe; the XU-pipe I2F.S16 was measured no faster).
__device__ __forceinline__ float2 frame_s16(uint32_t w) {
  return make_float2((float) sext_half(w, 0x9910u), (float) sext_half(w, 0xBB32u));
}

template <int FMT> struct PairPeak;
template <> struct PairPeak<FMT_S16> {
  uint32_t mx, mn;               // packed s16x2
  __device__ __forceinline__ void reset() { mx = 0u; mn = 0u; }
  __device__ __forceinline__ void words(uint32_t a, uint32_t b) {
    mx = __vimax3_s16x2(mx, a, b);
    mn = __vimin3_s16x2(mn, a, b);
  }
  // max |x| per channel: 0 .. 32768.  mx >= 0 >= mn per half; the negation
  // is done on zero-extended halves so that |-32768| = 32768 survives (a
  // 16-bit negate would wrap it).
  __device__ __forceinline__ void get(int& ax, int& ay) const {
    const uint32_t nlo = (0u - (mn & 0xffffu)) & 0xffffu, nhi = (0u - (mn >> 16)) & 0xffffu;
    ax = (int) max(mx & 0xffffu, nlo);
    ay = (int) max(mx >> 16, nhi);
  }
};
template <> struct PairPeak<FMT_F32> {
  float px, py;
  __device__ __forceinline__ void reset() { px = 0.0f; py = 0.0f; }
  __device__ __forceinline__ void frames(const float2 a, const float2 b) {
    px = fmaxf(px, fmaxf(fabsf(a.x), fabsf(b.x)));
    py = fmaxf(py, fmaxf(fabsf(a.y), fabsf(b.y)));
  }
};

// Loads one iteration (kIter frames) of the lane's channel pair from the
// staged row and folds the raw samples into the pair's maxima.
template <int LAYOUT>
__device__ __forceinline__ void load_iter2(const uint32_t rowp, uint32_t fb, uint32_t chl,
                                           float2* x, PairPeak<pair_format<LAYOUT>()>& pk) {
  if constexpr (LAYOUT == PAIR_S16_STEREO) {
#pragma unroll
    for (int u = 0; u < kIter / 4; ++u) {
      const uint4 v = lds128(rowp + 16u * u);
      pk.words(v.x, v.y);
      pk.words(v.z, v.w);
      x[4 * u + 0] = frame_s16(v.x);
      x[4 * u + 1] = frame_s16(v.y);
      x[4 * u + 2] = frame_s16(v.z);
      x[4 * u + 3] = frame_s16(v.w);
    }
  } else if constexpr (LAYOUT == PAIR_F32_STEREO) {
#pragma unroll
    for (int u = 0; u < kIter / 2; ++u) {
      const uint4 v = lds128(rowp + 16u * u);
      x[2 * u] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
      x[2 * u + 1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
      pk.frames(x[2 * u], x[2 * u + 1]);
    }
  } else if constexpr (LAYOUT == PAIR_S16_EVEN) {
    const uint32_t q = rowp + chl * 4u;
#pragma unroll
    for (int i = 0; i < kIter; i += 2) {
      const uint32_t a = lds32(q + i * fb);
      const uint32_t b = lds32(q + (i + 1) * fb);
      pk.words(a, b);
      x[i] = frame_s16(a);
      x[i + 1] = frame_s16(b);
    }
  } else {
    const uint32_t q = rowp + chl * 8u;
#pragma unroll
    for (int i = 0; i < kIter; i += 2) {
      const uint2 a = lds64(q + i * fb), b = lds64(q + (i + 1) * fb);
      x[i] = make_float2(__uint_as_float(a.x), __uint_as_float(a.y));
      x[i + 1] = make_float2(__uint_as_float(b.x), __uint_as_float(b.y));
      pk.frames(x[i], x[i + 1]);
    }
  }
}

// CTA size and resident warps per SM of the packed sweep.  Warps are autonomous,
// so small CTAs only make the distribution of warps over the SMs finer.
#ifndef LG_PAIR_THREADS
#define LG_PAIR_THREADS 32
#endif
#ifndef LG_PAIR_WARPS_PER_SM
#define LG_PAIR_WARPS_PER_SM 16
#endif
constexpr int kPairThreads = LG_PAIR_THREADS;
constexpr int kPairMinBlocks = LG_PAIR_WARPS_PER_SM / (LG_PAIR_THREADS / 32);

template <int LAYOUT, bool TP>
__global__ void __launch_bounds__(kPairThreads, kPairMinBlocks)
sweep_pair_kernel(const __grid_constant__ SweepParams P) {
  constexpr int FMT = pair_format<LAYOUT>();
  constexpr int kPPS = pair_pps((uint32_t) FMT);
  constexpr bool STEREO = LAYOUT == PAIR_S16_STEREO || LAYOUT == PAIR_F32_STEREO;
  extern __shared__ __align__(128) unsigned char smem_all[];
  const uint32_t wic = threadIdx.x >> 5;
  const uint32_t lane = pin(threadIdx.x & 31u);   // kept in a register: S2R per stage and pair otherwise
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wic;
  if (warp >= P.nwarps) return;                // whole warps leave; no CTA barrier below
  unsigned char* sm = smem_all + wic * P.warp_smem;

#ifdef LG_PAIR_TRACE     // tuning: per-warp start / end times and SM at the tail of the candidate queue
  unsigned long long trace_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
  const WarpWork ww = P.work[warp];
  const Track& tr = P.tracks[ww.track];
  const int W = P.W, L = P.L;
  const uint32_t C = P.channels, fb = P.fb;
  const long long frames = (long long) tr.frames;
  const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm);

  const uint32_t lpc = STEREO ? 1u : P.lpc;    // lanes per chunk = channel pairs
  const uint32_t slot = lane / lpc;
  const uint32_t chl = lane - slot * lpc;      // channel pair within the chunk
  const uint32_t chunk = ww.first_chunk + slot;
  const bool compute = slot < P.cpw;
  const bool active = compute && chunk < tr.nchunks;

  // ---- lane state
  const LaneGeom geo = lane_geometry(frames, L, W, P.aq, chunk);
  PairCtx c;
  c.xp = c.d1 = c.w1 = c.w2 = c.v1 = c.v2 = bc2(0.0f);
  c.yr = c.yi = bc2(0.0f);
  c.e0x = c.e0y = 0.0;
  c.pd = c.pw = c.qd = c.qw = bc2(0.0f);
  c.f_lo = W + geo.o;
  c.f_hi = c.f_lo + L;
  // 2-D TMA staging (stereo, WarpWork::interior == 2): the warp's chunks of class
  // r = chunk mod m are consecutive rows of that class's tensor; one box per
  // class and stage lands them as rows r * (32 / m) ... of the stage buffer.
  const bool tma = STEREO && ww.interior == 2;
  const uint32_t tm = STEREO ? P.tma_m : 1u;
  const uint32_t my_slot = tma ? (lane % tm) * (32u / tm) + lane / tm : slot;
  const uint32_t my_row = pin((uint32_t) __cvta_generic_to_shared(sm) + my_slot * P.row_stride);

  // ---- staging.  Per stage every row receives one contiguous piece of
  // kPPS * 24 frames, moved with 16-byte cp.async copies.
  //
  // Stereo, interior warps (every byte they stage lies inside the track): kLPR
  // adjacent lanes copy kLPR adjacent units of ONE row per instruction, so an
  // instruction touches 32 / kLPR rows with 16 * kLPR contiguous bytes each
  // (fewer L1 tag look-ups and L2 sector requests than one row per lane); the
  // rows are taken in kLPR groups, lane l serving rows g * (32 / kLPR) + l / kLPR.
  //
  // Other layouts, and warps at a track boundary (zero-filling form): the lpc
  // lanes of a row copy its units interleaved.
  constexpr uint32_t kUnits = STEREO ? (uint32_t) (kPPS * kPairFrames * (FMT == FMT_S16 ? 4 : 8) / 16) : 0u;
  constexpr uint32_t kLPR = STEREO ? (kUnits % 4u == 0u ? 4u : 2u) : 1u;
  constexpr uint32_t kRowsPerCopy = 32u / kLPR;
  constexpr uint32_t kRowStride = ((kUnits | 1u) << 4);
  const uint32_t ustride = lpc << 4;
  const long long row_byte0 = geo.a * (long long) fb + (chl << 4);
  const long long track_bytes = frames * (long long) fb;
  const bool interior = ww.interior != 0;
  [[maybe_unused]] const bool coop = STEREO && interior;
  const uint32_t ncopy = P.kcopies;
  const uint32_t sm_base = (uint32_t) __cvta_generic_to_shared(sm);
  const uint32_t dst_row = pin(sm_base + slot * P.row_stride + (chl << 4));
  const unsigned char* src = pcm + row_byte0;          // own row (boundary warps, other layouts)
  // cooperative mapping: this lane's unit column and its row in every group
  const uint32_t q = lane & (kLPR - 1u), rb = lane / kLPR;
  const uint32_t dst_coop = pin(sm_base + rb * kRowStride + (q << 4));
  const unsigned char* src_g[kLPR];
#pragma unroll
  for (uint32_t g = 0; g < kLPR; ++g) {
    const long long a_row = __shfl_sync(0xffffffffu, geo.a, (int) (g * kRowsPerCopy + rb));
    src_g[g] = pcm + a_row * (long long) fb + (q << 4);
  }
  uint32_t pf_off = 0;
  // TMA bookkeeping (lane 0 issues): one mbarrier per ring stage behind the ring
  const uint32_t bar0 = sm_base + P.ring_bytes;
  uint32_t pf_buf = 0;
  int pf_x = 0;                                  // first word of the next stage within a row
  int tma_row[kTmaMaxM];                         // first row of the warp in each class's tensor
  const unsigned char* tma_map = nullptr;
  if (tma) {
    tma_map = reinterpret_cast<const unsigned char*>(P.tmaps) + (size_t) ww.track * kTmaMaxM * 128u;
#pragma unroll
    for (int r = 0; r < kTmaMaxM; ++r)
      tma_row[r] = (int) ((ww.first_chunk + (uint32_t) r) / tm) - (int) ((P.tma_shift >> r) & 1u);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kPairRing; ++i) mbar_init(bar0 + 8u * i, 1u);
      mbar_fence_init();
    }
    __syncwarp();
  }

  auto prefetch = [&]() {
    if (tma) {
      if (elect_one()) {
        const uint32_t bar = bar0 + 8u * pf_buf;
        mbar_arrive_expect_tx(bar, P.stage_bytes);
        const uint32_t dst = sm_base + pf_off;
        const uint32_t class_bytes = P.tma_class_bytes;    // (32 / tm) rows: no division per stage
#pragma unroll
        for (int r = 0; r < kTmaMaxM; ++r)
          if ((uint32_t) r < tm)
            tma_load_2d(dst + (uint32_t) r * class_bytes, tma_map + r * 128, pf_x, tma_row[r], bar);
      }
      pf_x += (int) (P.stage_row_bytes >> 2);
      pf_off += P.stage_bytes;
      if (++pf_buf == (uint32_t) kPairRing) { pf_buf = 0; pf_off = 0; }
      return;
    }
#ifdef LG_PAIR_NOLOAD   // ablation: no HBM traffic, compute on whatever is in shared memory
    if (false) {
#else
    if (coop) {
      const uint32_t dst = dst_coop + pf_off;
#pragma unroll
      for (uint32_t g = 0; g < kLPR; ++g) {
#pragma unroll
        for (uint32_t j = 0; j < (STEREO ? kUnits / kLPR : 1u); ++j)
          cp_async16(dst + g * kRowsPerCopy * kRowStride + j * (kLPR << 4), src_g[g] + j * (kLPR << 4));
        src_g[g] += P.stage_row_bytes;
      }
    } else if (compute) {
#endif
      const uint32_t dst = dst_row + pf_off;
#pragma unroll 1
      for (uint32_t k = 0; k < ncopy; ++k) {
        if (chl + k * lpc >= P.units) break;
        const long long g = (src - pcm) + (long long) (k * ustride);
        long long ok = g < 0 ? 0 : track_bytes - g;
        ok = ok < 0 ? 0 : (ok > 16 ? 16 : ok);
        cp_async16_zfill(dst + k * ustride, pcm + (ok ? g : 0), (uint32_t) ok);
      }
    }
    src += P.stage_row_bytes;
    pf_off += P.stage_bytes;
    if (pf_off == P.ring_bytes) pf_off = 0;
  };

  uint32_t* mrec = pin(P.mrec + (size_t) warp * P.npairs * 32u + lane);   // advanced pair by pair

  const uint32_t niters = (uint32_t) P.niters;
  const uint32_t npairs = P.npairs;
  const uint32_t nstages = (npairs + kPPS - 1) / kPPS;
  // Pairs whose two iterations are both "fast" for every lane of the warp.
  const int lfast = ww.lmin_valid < L ? ww.lmin_valid : L;
  const uint32_t fast_lo = (uint32_t) ((W + P.aq - 1 + kPairFrames - 1) / kPairFrames);
  const uint32_t fast_hi = (uint32_t) ((W + lfast) / kPairFrames);     // exclusive
  const uint32_t warm_hi = (uint32_t) (W / kPairFrames);               // pairs entirely in the warm-up
#pragma unroll
  for (int i = 0; i < kPairRing - 1; ++i) {
    if ((uint32_t) i < nstages) prefetch();
    cp_async_commit();
  }

  // lane-wide sample peak, raw units: integer for 16-bit input
  int spx_i = 0, spy_i = 0;
  float spx_f = 0.0f, spy_f = 0.0f;

  uint32_t cs_off = 0, cs_buf = 0, cs_parity = 0;
  uint32_t pair = 0;
  for (uint32_t s = 0; s < nstages; ++s) {
    if (tma) {
      mbar_wait(bar0 + 8u * cs_buf, cs_parity);
      if (++cs_buf == (uint32_t) kPairRing) { cs_buf = 0; cs_parity ^= 1u; }
    } else {
      cp_async_wait<kPairRing - 2>();
    }
    __syncwarp();                // everyone's data has landed; the previous stage is consumed
    if (s + kPairRing - 1 < nstages) prefetch();
    cp_async_commit();
    const uint32_t sbuf = my_row + cs_off;
    cs_off += P.stage_bytes;
    if (cs_off == P.ring_bytes) cs_off = 0;
#pragma unroll 1
    for (uint32_t pr = 0; pr < (uint32_t) kPPS; ++pr, ++pair) {
      if (pair >= npairs) break;
      const uint32_t buf = sbuf + pr * kPairFrames * fb;
      PairPeak<FMT> pk;
      pk.reset();
#ifdef LG_PAIR_NOCOMP   // ablation: staging only
      if (false) {
#else
      if (compute) {
#endif
        if (pair >= fast_lo && pair < fast_hi) {
          // ---- interior of the chunk: straight-line packed code
          float2 x0[kIter], x1[kIter];
          load_iter2<LAYOUT>(buf, fb, chl, x0, pk);
          load_iter2<LAYOUT>(buf + kIter * fb, fb, chl, x1, pk);
          const int f0 = (int) (pair * kPairFrames);
          iter_fast2(c, P, x0, f0);
          iter_fast2(c, P, x1, f0 + kIter);
        } else if (pair < warm_hi) {
          float2 x0[kIter], x1[kIter];
          load_iter2<LAYOUT>(buf, fb, chl, x0, pk);
          load_iter2<LAYOUT>(buf + kIter * fb, fb, chl, x1, pk);
          iter_warm2(c, P, x0, pair == 0);
          iter_warm2(c, P, x1, false);
        } else {
          // ---- chunk edges: one iteration at a time, each as its kind
          // (lg_common.h: iter_kind) -- of the two iterations of an edge
          // pair only the one that holds the chunk boundary is masked
#pragma unroll 1
          for (uint32_t it = 0; it < 2u; ++it) {
            const uint32_t iter = pair * 2u + it;
            if (iter >= niters) break;
            const int f0 = (int) (iter * kIter);
            float2 x[kIter];
            load_iter2<LAYOUT>(buf + it * kIter * fb, fb, chl, x, pk);
            const int kind = iter_kind(f0, W, P.aq, L, ww.lmin_valid);
            if (kind == ITER_WARM) iter_warm2(c, P, x, iter == 0);
            else if (kind == ITER_FAST) iter_fast2(c, P, x, f0);
            else iter_masked2(c, P, x, f0);
          }
        }
      }
      uint32_t cx, cy;
      if constexpr (FMT == FMT_S16) {
        int ax, ay;
        pk.get(ax, ay);
        spx_i = max(spx_i, ax); spy_i = max(spy_i, ay);
        cx = (uint32_t) ax; cy = (uint32_t) ay;
      } else {
        spx_f = fmaxf(spx_f, pk.px); spy_f = fmaxf(spy_f, pk.py);
        cx = peak_code(pk.px); cy = peak_code(pk.py);
      }
      if (TP) { *mrec = cx | (cy << 16); mrec += 32; }
    }
  }
  cp_async_wait<0>();

  const uint32_t ch = 2u * chl;
  if (active) {
    ChunkRec v;
    v.e0 = c.e0x; v.yr = c.yr.x; v.yi = c.yi.x;
    v.pd = c.pd.x; v.pw = c.pw.x; v.qd = c.qd.x; v.qw = c.qw.x;
    ChunkRec* out = P.recs + tr.rec_base + (uint64_t) chunk * C + ch;
    out[0] = v;
    v.e0 = c.e0y; v.yr = c.yr.y; v.yi = c.yi.y;
    v.pd = c.pd.y; v.pw = c.pw.y; v.qd = c.qd.y; v.qw = c.qw.y;
    out[1] = v;
  }
#ifdef LG_PAIR_TRACE
  if (lane == 0) {
    unsigned long long t1; uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    uint64_t* tb = P.tp_queue + 2ull * P.nwarps * P.npairs * 32ull - 2ull * P.nwarps;
    tb[2ull * warp] = trace_t0;
    tb[2ull * warp + 1] = (t1 - trace_t0) | ((uint64_t) smid << 48);
  }
#endif
  // Sample peak: non-negative floats order like their bit patterns.  Reduce
  // over the lanes of the warp that hold the same channel pair.
  const float spx = FMT == FMT_S16 ? (float) spx_i : spx_f;
  const float spy = FMT == FMT_S16 ? (float) spy_i : spy_f;
  const unsigned peers = __match_any_sync(0xffffffffu, compute ? chl : 0xffffu);
  const uint32_t bx = __reduce_max_sync(peers, active ? __float_as_uint(spx) : 0u);
  const uint32_t by = __reduce_max_sync(peers, active ? __float_as_uint(spy) : 0u);
  if (compute && lane == (uint32_t) (__ffs(peers) - 1)) {
    atomicMax(P.peaks + 2 * (tr.peak_base + ch), bx);
    atomicMax(P.peaks + 2 * (tr.peak_base + ch + 1), by);
  }
}

template <int LAYOUT, bool TP>
static cudaError_t launch_pair_k(const SweepParams& p, cudaStream_t stream) {
  const uint32_t wpb = kPairThreads / 32;
  const uint32_t blocks = (p.nwarps + wpb - 1) / wpb;
  size_t smem = (size_t) p.warp_smem * wpb;
  // Resident CTAs per SM can be capped by padding the dynamic shared memory
  // (p.ctas_per_sm, lg_plan.h: the planner rounds the launch to whole waves).
  if (p.ctas_per_sm) {
    const size_t cap = ((size_t) 233472 / p.ctas_per_sm - 1024) & ~(size_t) 127;
    if (cap > smem) smem = cap;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sweep_pair_kernel<LAYOUT, TP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sweep_pair_kernel<LAYOUT, TP>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sweep_pair_kernel<LAYOUT, TP><<<blocks, kPairThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int LAYOUT>
static cudaError_t launch_pair_l(const SweepParams& p, int tpf, cudaStream_t stream) {
  return tpf ? launch_pair_k<LAYOUT, true>(p, stream) : launch_pair_k<LAYOUT, false>(p, stream);
}

cudaError_t launch_sweep_pair(const SweepParams& p, uint32_t format, int tpf, cudaStream_t stream) {
  if (p.nwarps == 0) return cudaSuccess;
  if (format == FMT_S16)
    return p.channels == 2 ? launch_pair_l<PAIR_S16_STEREO>(p, tpf, stream)
                           : launch_pair_l<PAIR_S16_EVEN>(p, tpf, stream);
  return p.channels == 2 ? launch_pair_l<PAIR_F32_STEREO>(p, tpf, stream)
                         : launch_pair_l<PAIR_F32_EVEN>(p, tpf, stream);
}

// -------------------------------------------------------------- true peak
//
// Second pass over the pair maxima.  The true peak is a maximum, and a
// polyphase output cannot exceed ||c||_1 * max|x| over its taps' window; the
// channel's true peak is at least its sample peak (final by now).  So only
// pairs whose bound (over the pair and the pair before it, which holds the
// taps' history) exceeds the channel's sample peak can matter -- a few percent
// of programme material.  They are collected in a queue and evaluated one per
// thread, re-reading their frames from the PCM (L2 / HBM).  The result is
// identical to evaluating every frame.

constexpr int kTp2Threads = 128;

template <int FMT>
__device__ __forceinline__ float pcm_at(const unsigned char* p) {
  if (FMT == FMT_S16) return (float) (int) *reinterpret_cast<const short*>(p);
  return *reinterpret_cast<const float*>(p);
}

// One queued pair: both of its iterations share one window of NT history
// frames + 24 frames.  Stereo frames are read with 16-byte loads (the window
// starts on a 16-byte boundary of the track: lane origins and pairs are
// multiples of four frames); other layouts and windows that touch the track's
// ends read sample by sample.
template <int FMT, int TPF>
__device__ __forceinline__ void tp_pair_evaluate(const SweepParams& P, const uint4 cd) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  constexpr int NW = NT + kPairFrames;
  const Track& tr = P.tracks[cd.x];
  const long long frames = (long long) tr.frames;
  const long long t0 = (long long) ((unsigned long long) cd.z | ((unsigned long long) cd.w << 32));
  const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm);
  float win[NW];
  const bool inside = t0 >= NT && t0 + kPairFrames <= frames;
  if (inside && P.channels == 2) {
    const unsigned char* q = pcm + (t0 - NT) * (long long) P.fb;
    if (FMT == FMT_S16) {
      const uint32_t sel = cd.y ? 0xBB32u : 0x9910u;
#pragma unroll
      for (int k = 0; k < NW / 4; ++k) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(q) + k);
        win[4 * k + 0] = (float) sext_half(v.x, sel);
        win[4 * k + 1] = (float) sext_half(v.y, sel);
        win[4 * k + 2] = (float) sext_half(v.z, sel);
        win[4 * k + 3] = (float) sext_half(v.w, sel);
      }
    } else {
#pragma unroll
      for (int k = 0; k < NW / 2; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(q) + k);
        win[2 * k + 0] = cd.y ? v.y : v.x;
        win[2 * k + 1] = cd.y ? v.w : v.z;
      }
    }
  } else {
    const unsigned char* pc = pcm + cd.y * (FMT == FMT_S16 ? 2u : 4u);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const long long t = t0 - NT + k;
      win[k] = (t >= 0 && t < frames) ? pcm_at<FMT>(pc + t * (long long) P.fb) : 0.0f;
    }
  }
  // the reference produces no output beyond the last frame it was given
  const long long left = frames - t0;
  const int nvalid = left > kPairFrames ? kPairFrames : (int) left;
  float m = 0.0f;
  if (inside) {
#pragma unroll
    for (int i = 0; i < kPairFrames; ++i) m = fmaxf(m, tp_frame<TPF>(win, NT + i));
  } else {
#pragma unroll
    for (int i = 0; i < kPairFrames; ++i)
      if (i < nvalid) m = fmaxf(m, tp_frame<TPF>(win, NT + i));
  }
  uint32_t* cell = P.peaks + 2 * (tr.peak_base + cd.y) + 1;
  if (__float_as_uint(m) > __ldcg(cell)) atomicMax(cell, __float_as_uint(m));
}

// The pass runs as two kernels.  tp_scan_pair_kernel streams over the pair
// maxima (one warp per run of pairs of one sweep warp, scanning lane = sweep
// lane) and appends the pairs that can still matter to a global queue as
// (sweep warp, lane, channel, pair); tp_eval_pair_kernel evaluates the queue,
// one candidate per thread, grid-stride.  Candidates cluster in the loud
// passages: evaluating them where they are found leaves most SMs idle while a
// few warps work through dense runs; the queue spreads them evenly.
constexpr int kTpScanThreads = 256;
constexpr int kTpScanWarps = kTpScanThreads / 32;
constexpr int kTpSeg = 16;                  // pairs per scan item (codes held in registers)

__device__ __forceinline__ uint64_t tp_entry(uint32_t w, uint32_t lane, uint32_t ch1, uint32_t pair) {
  return (uint64_t) w | ((uint64_t) (pair | (lane << 16) | (ch1 << 21)) << 32);
}

template <int FMT>
__global__ void __launch_bounds__(kTpScanThreads)
tp_scan_pair_kernel(const __grid_constant__ SweepParams P, const uint32_t seg_pairs) {
  // seg_pairs <= kTpSeg: a lane keeps its item's codes in registers, the hits
  // are counted first and written second, with ONE queue reservation per CTA
  // (a reservation per warp and flush serialises on the counter's address).
  __shared__ uint32_t s_count[kTpScanWarps + 1];
  const uint32_t npairs = P.npairs;
  const uint32_t lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
  const uint32_t nseg = (npairs + seg_pairs - 1) / seg_pairs;
  const uint64_t item = (uint64_t) blockIdx.x * kTpScanWarps + wic;
  const bool live = item < (uint64_t) P.nwarps * nseg;
  const uint32_t w = live ? (uint32_t) (item / nseg) : 0u;
  const uint32_t p_begin = live ? (uint32_t) (item - (uint64_t) w * nseg) * seg_pairs : 0u;
  const uint32_t p_end = !live ? 0u : (p_begin + seg_pairs < npairs ? p_begin + seg_pairs : npairs);
  // the sweep lane's place in its track
  const WarpWork ww = P.work[w];
  const Track& tr = P.tracks[ww.track];
  const uint32_t slot = lane / P.lpc;
  const uint32_t ch = 2u * (lane - slot * P.lpc);
  const uint32_t chunk = ww.first_chunk + slot;
  const bool ok = live && slot < P.cpw && chunk < tr.nchunks;
  const LaneGeom g = lane_geometry((long long) tr.frames, P.L, P.W, P.aq, chunk);
  const int f_lo = P.W + g.o;
  const long long left = (long long) tr.frames - g.a;
  const int f_hi = f_lo + P.L;
  const int f_end = left < (long long) f_hi ? (left < 0 ? 0 : (int) left) : f_hi;
  const long long lead_in = (long long) tr.lead_in;

  const uint32_t* codes = P.mrec + ((size_t) w * npairs) * 32u + lane;
  const uint32_t* cell = P.peaks + 2 * ((ok ? tr.peak_base : 0) + (ok ? ch : 0));
  // the channel's true peak is at least its sample peak (final by now)
  const float floor0 = __uint_as_float(__ldcg(cell));
  const float floor1 = __uint_as_float(__ldcg(cell + 2));
  // Pairs this lane owns: those that overlap its own chunk frames [f_lo, f_end)
  // and do not lie entirely inside the track's lead-in.
  int p_lo = f_lo / kPairFrames, p_hi = (f_end + kPairFrames - 1) / kPairFrames;
  {
    const long long lead_local = lead_in - g.a;            // lane-local frame where the lead-in ends
    if (lead_local > 0) {
      const long long q = lead_local / kPairFrames;        // first pair with frames past the lead-in
      if (q > p_lo) p_lo = q > 0x7fffffff ? 0x7fffffff : (int) q;
    }
    if (!ok) p_hi = 0;
  }
  // all codes of the item, and the one before it (history of the first pair)
  uint32_t code[kTpSeg];
#pragma unroll
  for (int j = 0; j < kTpSeg; ++j)
    code[j] = p_begin + j < p_end ? __ldcs(codes + (size_t) (p_begin + j) * 32u) : 0u;
  const uint32_t before = (live && p_begin) ? __ldcs(codes + (size_t) (p_begin - 1) * 32u) : 0u;

  // per pair: which of the lane's two channels can still raise its peak
  auto hits = [&](int j, uint32_t prev, bool& hit0, bool& hit1) {
    const int p = (int) p_begin + j;
    const uint32_t cm = __vmaxu2(code[j], prev);     // per channel: this pair and its history
    const bool own = p >= p_lo && p < p_hi && p < (int) p_end;
    hit0 = own && P.tp_bound * pair_code_value<FMT>(cm & 0xffffu) > floor0;
    hit1 = own && P.tp_bound * pair_code_value<FMT>(cm >> 16) > floor1;
  };
  uint32_t mine = 0;                     // hits of this lane
  {
    uint32_t prev = before;
#pragma unroll
    for (int j = 0; j < kTpSeg; ++j) {
      bool h0, h1;
      hits(j, prev, h0, h1);
      prev = code[j];
      mine += (h0 ? 1u : 0u) + (h1 ? 1u : 0u);
    }
  }
  // exclusive prefix of the lanes' counts within the warp, warps within the CTA
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t) o) incl += v;
  }
  const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
  if (lane == 0) s_count[wic] = warp_total;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int k = 0; k < kTpScanWarps; ++k) { const uint32_t c = s_count[k]; s_count[k] = total; total += c; }
    s_count[kTpScanWarps] = total ? atomicAdd(P.tp_ticket, total) : 0u;
  }
  __syncthreads();
  if (!warp_total) return;
  uint64_t* out = P.tp_queue + s_count[kTpScanWarps] + s_count[wic] + (incl - mine);
  uint32_t prev = before;
#pragma unroll
  for (int j = 0; j < kTpSeg; ++j) {
    bool h0, h1;
    hits(j, prev, h0, h1);
    prev = code[j];
    if (h0) *out++ = tp_entry(w, lane, 0u, p_begin + j);
    if (h1) *out++ = tp_entry(w, lane, 1u, p_begin + j);
  }
}

template <int FMT, int TPF>
__global__ void __launch_bounds__(kTp2Threads, 4)
tp_eval_pair_kernel(const __grid_constant__ SweepParams P) {
  const uint32_t count = __ldcg(P.tp_ticket);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const uint64_t e = P.tp_queue[i];
    const uint32_t hi = (uint32_t) (e >> 32);
    const uint32_t pair = hi & 0xffffu, lane = (hi >> 16) & 31u, ch1 = hi >> 21;
    const WarpWork ww = P.work[(uint32_t) e];
    const Track& tr = P.tracks[ww.track];
    const uint32_t slot = lane / P.lpc;
    const uint32_t ch = 2u * (lane - slot * P.lpc) + ch1;
    const LaneGeom g = lane_geometry((long long) tr.frames, P.L, P.W, P.aq, ww.first_chunk + slot);
    const long long t0 = g.a + (long long) pair * kPairFrames;
    tp_pair_evaluate<FMT, TPF>(P, make_uint4(ww.track, ch, (uint32_t) (unsigned long long) t0,
                                             (uint32_t) ((unsigned long long) t0 >> 32)));
  }
}

template <int FMT, int TPF>
static cudaError_t launch_truepeak_pair_t(const SweepParams& p, uint32_t sms, cudaStream_t stream,
                                          cudaEvent_t hold) {
  if (p.npairs > 0xffffu) return cudaErrorInvalidValue;      // pair index is a 16-bit field
  static int per_sm = 0;
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_eval_pair_kernel<FMT, TPF>,
                                                      kTp2Threads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 4;
  }
  // scan: runs of kTpSeg pairs per warp (fewer for small batches, so that every SM gets several)
  const uint64_t nscan = (uint64_t) sms * 8 * kTpScanWarps;
  uint32_t seg = kTpSeg;
  while (seg > 4 && (uint64_t) p.nwarps * ((p.npairs + seg - 1) / seg) < nscan) seg >>= 1;
  const uint64_t nitems = (uint64_t) p.nwarps * ((p.npairs + seg - 1) / seg);
  const uint64_t ctas = (nitems + kTpScanWarps - 1) / kTpScanWarps;
  tp_scan_pair_kernel<FMT><<<(unsigned) ctas, kTpScanThreads, 0, stream>>>(p, seg);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && hold) e = cudaStreamWaitEvent(stream, hold, 0);
  if (e != cudaSuccess) return e;
  tp_eval_pair_kernel<FMT, TPF><<<sms * per_sm, kTp2Threads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_truepeak_pair(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                                 cudaStream_t stream, cudaEvent_t hold) {
  if (p.nwarps == 0 || tpf == 0) return cudaSuccess;
  if (format == FMT_S16)
    return tpf == 4 ? launch_truepeak_pair_t<FMT_S16, 4>(p, sms, stream, hold)
                    : launch_truepeak_pair_t<FMT_S16, 2>(p, sms, stream, hold);
  return tpf == 4 ? launch_truepeak_pair_t<FMT_F32, 4>(p, sms, stream, hold)
                  : launch_truepeak_pair_t<FMT_F32, 2>(p, sms, stream, hold);
}

}  // namespace lg
