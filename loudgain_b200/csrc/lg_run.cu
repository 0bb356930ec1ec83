// lg_run.cu -- the run sweep: fused K-weighting + chunk energies + sample peak +
// true-peak screening for STEREO tracks (what /root/reference/src/scan.c:448
// feeds through ebur128_add_frames_short for almost every file).
//
// A track is cut into runs of R chunks (R * L frames, lg_plan.h: choose_run).
// One lane filters one run in one go: it starts Wp frames early (warm-up), so
// only its high-pass state is still wrong when the run begins, and what that
// costs the output is a decaying mode that lg_post.cuh removes exactly from the
// chunk energies (the lane leaves sum y^2, the mode sums and state snapshots per
// chunk).  The error decays, so the mode sums stop xi_frames into the run: the
// rest of a long run costs 9 packed FP32 operations per frame instead of 11,
// and the warm-up is paid once per run instead of once per chunk.
//
// One persistent CTA per SM (run_warps_per_sm autonomous warps, no CTA barrier
// after the start): CTA b owns the work items b, b + grid, b + 2 grid, ... -- an
// item is 32 consecutive runs of one track, lane l = run first_run + l -- and its
// warps draw them from a ticket in shared memory, so every SM gets the same
// number of items (+-1) whatever order the hardware places CTAs in.  A track is a 2-D tensor (row y = run y, pitch R * L frames, a multiple
// of 16 bytes): per stage ONE tensor copy (cp.async.bulk.tensor.2d, completion on
// an mbarrier) lands the next piece of all 32 rows in shared memory, rows padded
// to an odd number of 16-byte units so that LDS.128 of different rows do not
// collide; the warm-up pieces are the tail of the rows one above (row -1 and
// everything past the track read as zeros: the tensor copy fills them in).  The
// last item of a track gets its complete rows from a second map with a smaller
// box and its partial row from a one-row map that ends with the track.
//
// Both channels of a frame run through FFMA2 / FADD2 (lg_packed.cuh).  16-bit
// frames are one 32-bit word; its halves are converted as (int) (w << 16) and
// (int) (w & 0xffff0000), i.e. in units of 1/65536 sample (exact; the scale is
// folded into the gain and the peaks).
//
// True peak is a maximum: a polyphase output cannot exceed ||c||_1 * max|x|
// over its window, and the channel's true peak is at least the largest sample
// seen so far.  The sweep tests every 24-frame pair against the channel's
// current peak cell; what survives goes into the CTA's own DENSE stretch of the
// candidate queue -- once per stage a warp adds up its lanes' survivors, takes
// that many slots from the CTA's ticket in shared memory and every lane stores
// its entries behind those of the lanes before it.  tp_eval_run_kernel then
// spreads all CTAs' candidates over the whole GPU, drops those the channel's
// FINAL sample peak rules out and evaluates the FIR on the rest: the same
// maximum as evaluating every frame (a window that is evaluated needlessly
// cannot raise it above the true value, one that is skipped cannot reach it).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "lg_common.h"
#include "lg_device.cuh"
#include "lg_kernels.h"
#include "lg_packed.cuh"
#include "lg_sweep.cuh"

namespace lg {

// Warps of the persistent CTA the kernel is compiled for (register budget:
// 64 K / (32 * warps)); fewer may be launched (SweepParams::run_warps_per_sm).
#ifndef LG_RUN_MAX_WARPS
#define LG_RUN_MAX_WARPS 16
#endif
constexpr uint32_t kRunMaxWarps = LG_RUN_MAX_WARPS;
constexpr size_t kRunMaxSmem = 227u * 1024u - 1024u;   // dynamic; the kernel also has a few static bytes

template <int FMT> struct RunGeom {
  static constexpr uint32_t kFB = FMT == FMT_S16 ? 4u : 8u;          // bytes per stereo frame
  static constexpr uint32_t kSF = run_stage_frames((uint32_t) FMT);  // frames per stage
  static constexpr uint32_t kRowBytes = kSF * kFB + 16u;              // odd number of 16-byte units
  static constexpr uint32_t kAuxOff = 32u * kRowBytes;                // partial row of a track's last item
  static constexpr uint32_t kSlotBytes = kAuxOff + 256u;
  static constexpr uint32_t kWarpBytes = kRunRing * kSlotBytes + 128u; // ring + mbarriers
  static_assert((kSF * kFB / 16u) % 2u == 0u, "row pitch must be an odd number of 16-byte units");
  static_assert(kSF % kPairFrames == 0u && kRunStageFrames % kSF == 0u, "stages are whole pairs; the warm-up whole stages");
  static_assert(kSlotBytes % 128u == 0u, "tensor copies land on 128-byte boundaries");
};

// Filter state and sums of one lane: .x = left, .y = right.
struct RunCtx {
  float2 xp, d1, w1, w2, v1, v2;
  float2 yr, yi;
  double e0x, e0y;
  float2 pd, pw;
};

// 12 frames of the lane's row -> float2 frames, in the sweep's unit.
template <int FMT>
__device__ __forceinline__ void run_load12(uint32_t rowp, float2* x) {
  if constexpr (FMT == FMT_S16) {
#pragma unroll
    for (int u = 0; u < kIter / 4; ++u) {
      const uint4 v = lds128(rowp + 16u * u);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        x[4 * u + i] = make_float2((float) (int) (w[i] << 16), (float) (int) (w[i] & 0xffff0000u));
    }
  } else {
#pragma unroll
    for (int u = 0; u < kIter / 2; ++u) {
      const uint4 v = lds128(rowp + 16u * u);
      x[2 * u] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
      x[2 * u + 1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
    }
  }
}

// max |x| per channel over 12 frames, folded into m.
__device__ __forceinline__ void run_max12(const float2* x, float2& m) {
#pragma unroll
  for (int i = 0; i < kIter; i += 2) {
    m.x = fmaxf(m.x, fmaxf(fabsf(x[i].x), fabsf(x[i + 1].x)));
    m.y = fmaxf(m.y, fmaxf(fabsf(x[i].y), fabsf(x[i + 1].y)));
  }
}

// One iteration inside a chunk, mode sums on (11 operations per frame).
__device__ __forceinline__ void run_iter_xi(RunCtx& c, const SweepParams& k, const float2* x) {
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
  mode_accumulate2(c.yr, c.yi, k, sr, si);
}

// ... past the mode-sum horizon (9 operations per frame).
__device__ __forceinline__ void run_iter_lean(RunCtx& c, const SweepParams& k, const float2* x) {
  float2 e = bc2(0.0f);
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const float2 y = k_step2(c, x[i], k);
    e = __ffma2_rn(y, y, e);
  }
  c.e0x += (double) e.x;
  c.e0y += (double) e.y;
}

__device__ __forceinline__ void run_iter_warm(RunCtx& c, const SweepParams& k, const float2* x) {
#pragma unroll
  for (int i = 0; i < kIter; ++i) (void) k_step2(c, x[i], k);
}

// Queue entry of a true-peak candidate (64 bits): high word = channel (1) | pair (15) |
// code of the bound (16; lg_sweep.cuh: peak_code of max |x| in the sweep's unit, rounded
// up), low word = item * 32 + lane.  (Entries that carry the window's address instead,
// so that the evaluation need not look the item and the track up, made the 16-bit sweep
// 10 % slower and the evaluation no faster: profiles/r02_tuning.txt.)
template <int FMT, bool TP>
__global__ void __launch_bounds__(kRunMaxWarps * 32, 1)
run_sweep_kernel(const __grid_constant__ SweepParams P) {
  using G = RunGeom<FMT>;
  constexpr uint32_t SF = G::kSF, FB = G::kFB, WPF = FB / 4u;
  constexpr uint32_t kPairsPerStageRun = SF / kPairFrames;
  static_assert(kPairsPerStageRun <= 2u, "a lane keeps the survivors of at most two pairs per stage");
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ uint32_t s_ticket, s_qticket, s_done;
  const uint32_t lane = pin(threadIdx.x & 31u);
  const uint32_t sm_base = (uint32_t) __cvta_generic_to_shared(smem_all) + (threadIdx.x >> 5) * G::kWarpBytes;
  const uint32_t bar0 = sm_base + kRunRing * G::kSlotBytes;
  if (threadIdx.x == 0) { s_ticket = 0u; s_qticket = 0u; s_done = 0u; }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kRunRing; ++i) mbar_init(bar0 + 8u * i, 1u);
    mbar_fence_init();
  }
  __syncthreads();                                 // the only CTA-wide barrier: warps are autonomous from here
  uint32_t phase = 0;                              // bit i: parity the next wait on slot i expects

  const int L = P.L, Wp = P.Wp;
  const uint32_t R = (uint32_t) P.R;
  const uint32_t nstages = P.run_nstages, niters = (uint32_t) P.niters;
  const uint32_t warm_stages = (uint32_t) Wp / SF;
#ifdef LG_RUN_NOLOAD       // ablation: compute on zeros (LG_RUN_NOLOAD=2: on pseudo-random samples)
  for (uint32_t i = lane; i < kRunRing * G::kSlotBytes / 4u; i += 32u) {
    uint32_t h = (i + 977u * threadIdx.x + 131071u * blockIdx.x) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // two 16-bit samples of about -12 dBFS peak
    reinterpret_cast<uint32_t*>(smem_all + (threadIdx.x >> 5) * G::kWarpBytes)[i] =
        LG_RUN_NOLOAD == 2 ? (((h & 0x3fffu) - 0x2000u) & 0xffffu) | ((((h >> 16) & 0x3fffu) - 0x2000u) << 16) : 0u;
  }
  __syncwarp();
#endif
  const uint32_t xi_iters = (uint32_t) P.xi_iters;
  const float thr_scale = 0.999f / (P.tp_bound * P.peak_scale);   // peak cell (raw units) -> |x| threshold
  // this CTA's dense stretch of the candidate queue (64-bit entries, see tp_eval_run_kernel)
  unsigned long long* const qcta =
      reinterpret_cast<unsigned long long*>(P.run_queue) + (size_t) blockIdx.x * P.run_cta_cap;

  // this CTA's items: blockIdx.x + k * gridDim.x, k drawn from the CTA's ticket
  uint32_t item = 0;
  if (lane == 0) item = atomicAdd(&s_ticket, 1u);
  item = blockIdx.x + __shfl_sync(0xffffffffu, item, 0) * gridDim.x;

  while (item < P.nitems) {
    const RunItem it = P.items[item];
    const Track& tr = P.tracks[it.track];
    const uint32_t run = it.first_run + lane;
    const bool active = run < tr.nruns;
    const uint32_t nchunks = tr.nchunks;
    const bool tail_item = it.tail_rows < 32u;
    const bool has_partial = tail_item && it.first_run + it.tail_rows < tr.nruns;
    const bool partial_lane = has_partial && lane == it.tail_rows;
    const unsigned char* maps = reinterpret_cast<const unsigned char*>(P.tmaps) + (size_t) it.track * kTmaMaxM * 128u;
    ChunkRec* recs = P.recs + tr.rec_base;
    uint32_t* cell = P.peaks + 2 * tr.peak_base;               // [ch][sample peak, true peak]

    // ---- staging.  The prefetch cursor walks the item's stages: the warm-up ones are the
    // tail of the rows one above (rows first_run - 1 ... first_run + 30), then the rows
    // themselves from their first frame on.
    int pf_x = (P.Lr - Wp) * (int) WPF, pf_y = (int) it.first_run - 1;
    uint32_t pf_s = 0, pf_slot = 0;
    const uint32_t tail_bytes = (it.tail_rows + (has_partial ? 1u : 0u)) * G::kRowBytes;
    auto issue = [&]() {
#ifndef LG_RUN_NOLOAD       // (ablation: no HBM traffic, compute on whatever is in shared memory)
      if (elect_one()) {
        const uint32_t bar = bar0 + 8u * pf_slot;
        const uint32_t dst = sm_base + pf_slot * G::kSlotBytes;
        if (!tail_item || pf_s < warm_stages) {
          mbar_arrive_expect_tx(bar, 32u * G::kRowBytes);
          tma_load_2d(dst, maps, pf_x, pf_y, bar);
        } else {
          // a track's last item: its complete rows (smaller box), its partial row (one-row map)
          mbar_arrive_expect_tx(bar, tail_bytes);
          if (it.tail_rows) tma_load_2d(dst, maps + 128, pf_x, pf_y, bar);
          if (has_partial) tma_load_2d(dst + G::kAuxOff, maps + 256, pf_x, 0, bar);
        }
      }
#endif
      pf_x += (int) (SF * WPF);
      if (++pf_s == warm_stages) { pf_x = 0; ++pf_y; }
      if (++pf_slot == (uint32_t) kRunRing) pf_slot = 0;
    };

    RunCtx c;
    c.xp = c.d1 = c.w1 = c.w2 = c.v1 = c.v2 = bc2(0.0f);
    c.yr = c.yi = bc2(0.0f);
    c.e0x = c.e0y = 0.0;
    c.pd = c.pw = bc2(0.0f);
    uint32_t j = 0;                                // chunk of the run being filled (warp-uniform)
    int nb = Wp + L;                               // lane-local frame where it ends
    float2 sp = bc2(0.0f), prev_pm = bc2(0.0f);    // lane's sample peak, previous pair's maxima
    // screening thresholds from the channel's peak cells (refreshed per stage)
    // The cells are fetched with 4-byte cp.async copies into the warp's shared memory and read
    // from there at the next refresh: a register load in flight across stages is waited for at
    // the top of the stage loop (the scoreboards do not survive the back-edge), a stall per refresh.
    const uint32_t cellbuf = bar0 + 64u;
    if (TP && lane < 2u) cp_async4(cellbuf + 4u * lane, cell + 2u * lane);
    cp_async_commit();
    float2 thr = bc2(0.0f);
    uint32_t seen_x = 0, seen_y = 0;               // what this warp last published
    // the survivors of the stage being worked on: up to two pairs x two channels per lane
    // (the bounds as float bits; turned into queue entries when they are stored)
    uint32_t cand[4];
    uint32_t cmask = 0;
    if (!active) thr = make_float2(3.0e38f, 3.0e38f);          // lanes past the track's end queue nothing
    // (until the first refresh, at the end of the first stage, every pair of an active lane is queued)

    auto close_chunk = [&]() {
      const uint32_t chunk = run * R + j;
      if (active && chunk < nchunks) {
        uint4* out = reinterpret_cast<uint4*>(recs + (size_t) chunk * 2u);
        const unsigned long long ex = (unsigned long long) __double_as_longlong(c.e0x);
        const unsigned long long ey = (unsigned long long) __double_as_longlong(c.e0y);
        out[0] = make_uint4((uint32_t) ex, (uint32_t) (ex >> 32), __float_as_uint(c.yr.x), __float_as_uint(c.yi.x));
        out[1] = make_uint4(__float_as_uint(c.pd.x), __float_as_uint(c.pw.x), __float_as_uint(c.d1.x),
                            __float_as_uint(c.w2.x));
        out[2] = make_uint4((uint32_t) ey, (uint32_t) (ey >> 32), __float_as_uint(c.yr.y), __float_as_uint(c.yi.y));
        out[3] = make_uint4(__float_as_uint(c.pd.y), __float_as_uint(c.pw.y), __float_as_uint(c.d1.y),
                            __float_as_uint(c.w2.y));
      }
      c.e0x = c.e0y = 0.0;
      c.yr = c.yi = bc2(0.0f);
      c.pd = c.d1; c.pw = c.w2;
      ++j;
      nb += L;
    };

#pragma unroll
    for (int i = 0; i < kRunRing - 1; ++i)
      if ((uint32_t) i < nstages) issue();

    uint32_t slot = 0;
    float2 half_pm = bc2(0.0f);
    for (uint32_t s = 0; s < nstages; ++s) {
#ifndef LG_RUN_NOLOAD
      mbar_wait(bar0 + 8u * slot, (phase >> slot) & 1u);
#endif
      phase ^= 1u << slot;
      __syncwarp();                    // the previous stage is consumed by every lane
      // (Asking for the next but one stage as soon as this stage's last frames are in
      // registers -- half a stage earlier -- made the kernel slower, 0.145 -> 0.159 ms: the
      // barrier and the copy's uniform-datapath code in the middle of the stage keep the
      // compiler from interleaving the two pairs' work.  profiles/r02_tuning.txt C.)
      if (pf_s < nstages) issue();
      // the partial row of a track's last item comes from its own buffer in the main stages
      const uint32_t rowp = sm_base + slot * G::kSlotBytes +
                            ((partial_lane && s >= warm_stages) ? G::kAuxOff : lane * G::kRowBytes);
      if (++slot == (uint32_t) kRunRing) slot = 0;
#ifndef LG_RUN_NOCOMP       // (ablation: staging only)
      // sample peak + true-peak screening of one pair: a pair (and the history
      // before it) that cannot beat the channel's peak is done
      auto screen = [&](const float2 pm, const uint32_t which /* pair of the stage */) {
        sp.x = fmaxf(sp.x, pm.x);
        sp.y = fmaxf(sp.y, pm.y);
        if (TP) {
          // (only the bounds are kept here; the queue entries are put together when -- and
          // if -- they are stored)
          const float cx = fmaxf(pm.x, prev_pm.x), cy = fmaxf(pm.y, prev_pm.y);
          prev_pm = pm;
          const bool hx = cx > thr.x, hy = cy > thr.y;
          if (which == 0u) {
            cand[0] = __float_as_uint(cx); cand[1] = __float_as_uint(cy);
            cmask |= (hx ? 1u : 0u) | (hy ? 2u : 0u);
          } else {
            cand[2] = __float_as_uint(cx); cand[3] = __float_as_uint(cy);
            cmask |= (hx ? 4u : 0u) | (hy ? 8u : 0u);
          }
        }
      };
      const uint32_t pair0 = s * kPairsPerStageRun;
      const int fe = (int) ((pair0 + kPairsPerStageRun) * kPairFrames);     // lane-local end of the stage
      if (s < warm_stages) {
        // ---- warm-up: filter state only (the warm-up is a whole number of stages)
#pragma unroll
        for (uint32_t pr = 0; pr < kPairsPerStageRun; ++pr) {
          const uint32_t buf = rowp + pr * kPairFrames * FB;
          float2 x0[kIter], x1[kIter], pm = bc2(0.0f);
          run_load12<FMT>(buf, x0);
          run_load12<FMT>(buf + kIter * FB, x1);
          run_max12(x0, pm);
          run_max12(x1, pm);
          if (pr == 0 && s == 0) c.xp = x0[0];               // lg_sweep.cuh: lane_start
          run_iter_warm(c, P, x0);
          run_iter_warm(c, P, x1);
          // the warm-up's pairs belong to the lane of the run before: peaks only
          sp.x = fmaxf(sp.x, pm.x);
          sp.y = fmaxf(sp.y, pm.y);
          prev_pm = pm;
        }
        if (s + 1 == warm_stages) { c.pd = c.d1; c.pw = c.w2; }
      } else if (nb >= fe && 2u * (pair0 + kPairsPerStageRun) <= niters &&
                 (2u * (pair0 + kPairsPerStageRun) <= xi_iters || 2u * pair0 >= xi_iters)) {
        // ---- the whole stage inside one chunk: straight-line code
        // (the pairs' maxima are screened after the stage's arithmetic: a vote and a
        // branch per pair inside it would cut the filter's instruction stream in two)
        float2 pms[kPairsPerStageRun];
        if (2u * pair0 < xi_iters) {
#pragma unroll
          for (uint32_t pr = 0; pr < kPairsPerStageRun; ++pr) {
            const uint32_t buf = rowp + pr * kPairFrames * FB;
            float2 x0[kIter], x1[kIter];
            pms[pr] = bc2(0.0f);
            run_load12<FMT>(buf, x0);
            run_load12<FMT>(buf + kIter * FB, x1);
            run_max12(x0, pms[pr]);
            run_max12(x1, pms[pr]);
            run_iter_xi(c, P, x0);
            run_iter_xi(c, P, x1);
          }
        } else {
#pragma unroll
          for (uint32_t pr = 0; pr < kPairsPerStageRun; ++pr) {
            const uint32_t buf = rowp + pr * kPairFrames * FB;
            float2 x0[kIter], x1[kIter];
            pms[pr] = bc2(0.0f);
            run_load12<FMT>(buf, x0);
            run_load12<FMT>(buf + kIter * FB, x1);
            run_max12(x0, pms[pr]);
            run_max12(x1, pms[pr]);
            run_iter_lean(c, P, x0);
            run_iter_lean(c, P, x1);
          }
        }
#pragma unroll
        for (uint32_t pr = 0; pr < kPairsPerStageRun; ++pr) screen(pms[pr], pr);
        if (nb == fe) close_chunk();
      } else {
        // ---- a chunk (or the mode sums, or the run) ends inside the stage: one
        // iteration at a time, split at the boundary (tests/emu: run_group_runs)
#pragma unroll 1
        for (uint32_t it2 = 0; it2 < 2u * kPairsPerStageRun; ++it2) {
          const uint32_t iter = 2u * pair0 + it2;
          if (iter >= niters) break;
          const int g0 = (int) (iter * kIter);
          float2 x[kIter], pm = bc2(0.0f);
          run_load12<FMT>(rowp + it2 * kIter * FB, x);
          run_max12(x, pm);
          const bool xi_on = iter < xi_iters;
          const int ib = nb >= g0 + kIter ? kIter : nb - g0;
          float2 e = bc2(0.0f), sr = bc2(0.0f), si = bc2(0.0f);
#pragma unroll
          for (int i = 0; i < kIter; ++i) {
            if (i == ib) {
              c.e0x += (double) e.x;
              c.e0y += (double) e.y;
              if (xi_on) mode_accumulate2(c.yr, c.yi, P, sr, si);
              close_chunk();
              e = sr = si = bc2(0.0f);
            }
            const float2 y = k_step2(c, x[i], P);
            e = __ffma2_rn(y, y, e);
            if (xi_on) {
              sr = __ffma2_rn(y, bc2(P.lam_re[i]), sr);
              si = __ffma2_rn(y, bc2(P.lam_im[i]), si);
            }
          }
          c.e0x += (double) e.x;
          c.e0y += (double) e.y;
          if (xi_on) mode_accumulate2(c.yr, c.yi, P, sr, si);
          if (nb == g0 + kIter) close_chunk();
          // screening works on pairs: fold the first iteration's maxima into the second's
          if (it2 & 1u) screen(make_float2(fmaxf(pm.x, half_pm.x), fmaxf(pm.y, half_pm.y)), it2 >> 1);
          else if (iter + 1u >= niters) screen(pm, it2 >> 1);
          else half_pm = pm;
        }
      }
      if (TP) {
        // ---- the stage's survivors -> the CTA's dense queue: one reservation per warp
        const uint32_t mine = __popc(cmask);
        const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
        if (total) {                                   // warp-uniform
          uint32_t incl = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t) o) incl += v;
          }
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&s_qticket, total);
          base = __shfl_sync(0xffffffffu, base, 0);
#ifndef LG_RUN_NOSTORE      // (ablation: screening without the queue)
          unsigned long long* out = qcta + base + (incl - mine);
          const uint32_t slot_id = item * 32u + lane;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!(cmask & (1u << k))) continue;
            // channel | pair | code of the bound
            const uint32_t w = ((uint32_t) (k & 1) << 31) | ((pair0 + (uint32_t) (k >> 1)) << 16) |
                               peak_code(__uint_as_float(cand[k]));
            *out++ = ((unsigned long long) w << 32) | slot_id;
          }
#endif
        }
        cmask = 0;
      }
#endif
#ifndef LG_RUN_NOFLOOR        // (ablation: thresholds from the item's start only)
      if (TP && ((s & 7u) == 7u || s < 2u)) {
        // after the first two stages, then every eighth: publish this warp's sample peak when it raises the cell, and
        // screen on against the larger of the two.  The cell values are the ones loaded at
        // the refresh before; the next load is issued here and used at the next refresh.
        const uint32_t ux = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(sp.x * P.peak_scale) : 0u);
        const uint32_t uy = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(sp.y * P.peak_scale) : 0u);
        cp_async_wait<0>();
        __syncwarp();
        const uint32_t cellx = lds32(cellbuf), celly = lds32(cellbuf + 4u);
        if (lane == 0) {
          if (ux > cellx && ux > seen_x) atomicMax(cell, ux);
          if (uy > celly && uy > seen_y) atomicMax(cell + 2, uy);
        }
        seen_x = max(seen_x, ux); seen_y = max(seen_y, uy);
        if (active)
          thr = make_float2(__uint_as_float(max(cellx, ux)) * thr_scale, __uint_as_float(max(celly, uy)) * thr_scale);
        __syncwarp();                          // everyone has read the buffer before it is refilled
        if (lane < 2u) cp_async4(cellbuf + 4u * lane, cell + 2u * lane);
        cp_async_commit();
      }
#endif
    }
    {
      const uint32_t ux = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(sp.x * P.peak_scale) : 0u);
      const uint32_t uy = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(sp.y * P.peak_scale) : 0u);
      if (lane == 0) {
        atomicMax(cell, ux);
        atomicMax(cell + 2, uy);
      }
    }
    cp_async_wait<0>();                // no cell fetch of this item's track may land in the next item's buffer
    __syncwarp();                      // every lane is done with the ring before the next item refills it
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(&s_ticket, 1u);
    item = blockIdx.x + __shfl_sync(0xffffffffu, next, 0) * gridDim.x;
  }
  if (TP && lane == 0) {
    // the warp that leaves last tells the evaluation how much this CTA queued
    __threadfence_block();
    if (atomicAdd(&s_done, 1u) == (blockDim.x >> 5) - 1u) P.run_counts[blockIdx.x] = atomicAdd(&s_qticket, 0u);
  }
}

// CTAs of the sweep (the planner's figure; the evaluation and the host read the counts of as many)
uint32_t run_sweep_grid(const SweepParams& p, uint32_t sms) {
  const uint32_t most = sms < p.nitems ? sms : p.nitems;
  return p.run_grid && p.run_grid < most ? p.run_grid : most;
}

template <int FMT, bool TP>
static cudaError_t launch_run_k(const SweepParams& p, uint32_t sms, cudaStream_t stream) {
  using G = RunGeom<FMT>;
  uint32_t warps = p.run_warps_per_sm ? p.run_warps_per_sm : kRunWarpsPerSM;
  if (warps > kRunMaxWarps) warps = kRunMaxWarps;
  while (warps > 1 && (size_t) warps * G::kWarpBytes > kRunMaxSmem) --warps;
  // one CTA per SM: more than half of an SM's shared memory, whatever the ring needs
  size_t smem = (size_t) warps * G::kWarpBytes;
  if (smem < 120u * 1024u) smem = 120u * 1024u;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(run_sweep_kernel<FMT, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) kRunMaxSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(run_sweep_kernel<FMT, TP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const uint32_t blocks = run_sweep_grid(p, sms);
  run_sweep_kernel<FMT, TP><<<blocks, warps * 32u, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_sweep_run(const SweepParams& p, uint32_t format, int tpf, uint32_t sms, cudaStream_t stream) {
  if (p.nitems == 0) return cudaSuccess;
#ifdef LG_RUN_NOTP         // ablation: no true-peak screening
  tpf = 0;
#endif
  if (format == FMT_S16)
    return tpf ? launch_run_k<FMT_S16, true>(p, sms, stream) : launch_run_k<FMT_S16, false>(p, sms, stream);
  return tpf ? launch_run_k<FMT_F32, true>(p, sms, stream) : launch_run_k<FMT_F32, false>(p, sms, stream);
}

// -------------------------------------------------------------- true peak
//
// The candidates the sweep left behind: sweep CTA b queued run_counts[b] entries
// (64 bits each: channel | pair | code, item * 32 + lane) at run_queue + b * run_cta_cap.
// tp_eval_run_kernel lays the CTAs'
// stretches end to end (a prefix over at most 148 counts, in shared memory) and
// deals the candidates out over the whole GPU, one per thread, grid-stride:
// candidates cluster in the loud passages, so evaluating them where they were
// found would leave most SMs idle behind a few busy warps.  A candidate whose
// bound does not exceed the channel's FINAL sample peak is dropped before its
// window is fetched.
constexpr int kTpRunThreads = 128;
constexpr uint32_t kTpMaxCtas = 1024;      // sweep CTAs (one per SM) the prefix has room for

// One queue entry, looked up: where its window lies and what it has to beat.
struct TpCand {
  uint32_t kind;                 // 0: nothing to do, 1: window inside the track (staged), 2: at the track's ends
  uint32_t ch;
  uint32_t* cell;                // the channel's true-peak cell
  uint32_t seen;                 // what the window has to beat (float bits)
  const unsigned char* q;        // kind 1: first byte of the window (NT frames of history + the pair)
};
constexpr int kTpBatch = 4;      // entries a thread looks up together
constexpr int kTpCellWords = 1024;   // peak-cell words (two per channel) a CTA keeps in shared memory

template <int FMT, int TPF>
__global__ void __launch_bounds__(kTpRunThreads, FMT == FMT_S16 ? 6 : 4)
tp_eval_run_kernel(const __grid_constant__ SweepParams P, const uint32_t nctas) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  constexpr int NW = NT + kPairFrames;
  constexpr uint32_t FB = FMT == FMT_S16 ? 4u : 8u;            // bytes per stereo frame
  constexpr uint32_t WB = (uint32_t) NW * FB;                  // bytes of a window: a multiple of 16
  static_assert(WB % 16u == 0u, "windows are staged in 16-byte copies");
  // A thread's window is staged in its own stretch of shared memory by cp.async: all its
  // 16-byte copies are in flight at once (as register loads the compiler spread them over
  // the FIR, a DRAM round trip each), and the NEXT candidate's window is requested as soon as
  // this one sits in registers, so that it arrives while the FIR runs.  The entries themselves
  // are looked up kTpBatch at a time, stage by stage (queue entry -> work item -> track ->
  // peak cells), so that the four dependent round trips are paid once per batch.
  extern __shared__ __align__(16) unsigned char tp_stage[];
  __shared__ uint32_t s_off[kTpMaxCtas + 1];
  // The peak cells of a batch with few channels, per CTA.  An album's dozen tracks share two
  // cache lines of cells: read from L2 for every candidate (hundreds of thousands of reads of
  // the same few sectors) they were the slowest link of the look-up.  The sample peaks are
  // final; the true-peak cells are floors, raised locally by shared-memory atomics (only a
  // value that beats the CTA's floor goes out to the global cell) and refreshed per batch.
  __shared__ uint32_t s_cells[kTpCellWords];
  const bool cells_cached = P.npeak_words <= (uint32_t) kTpCellWords;
  if (cells_cached)
    for (uint32_t w = threadIdx.x; w < P.npeak_words; w += blockDim.x) s_cells[w] = __ldcg(P.peaks + w);
  // exclusive prefix of the CTAs' counts (nctas <= 148 in practice): loaded by all threads,
  // summed by one warp
  for (uint32_t c = threadIdx.x; c < nctas; c += blockDim.x) s_off[c + 1] = __ldcg(P.run_counts + c);
  __syncthreads();
  if (threadIdx.x < 32u) {
    uint32_t run = 0;
    for (uint32_t c0 = 0; c0 < nctas; c0 += 32u) {
      const uint32_t c = c0 + threadIdx.x;
      const uint32_t v = c < nctas ? s_off[c + 1] : 0u;
      uint32_t incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (threadIdx.x >= (uint32_t) o) incl += u;
      }
      __syncwarp();
      if (c < nctas) s_off[c + 1] = run + incl;        // inclusive: s_off[c + 1] = entries before stretch c + 1
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_off[0] = 0u;
  }
  __syncthreads();
  const uint32_t count = s_off[nctas];
  const unsigned long long* queue = reinterpret_cast<const unsigned long long*>(P.run_queue);
  const uint32_t stage = (uint32_t) __cvta_generic_to_shared(tp_stage) + threadIdx.x * WB;
  // ... and its looked-up entries behind all the windows, 16 bytes each: window address, index of
  // the true-peak cell | channel << 29 | kind << 30, what the window has to beat
  const uint32_t cands = (uint32_t) __cvta_generic_to_shared(tp_stage) + blockDim.x * WB +
                         threadIdx.x * (uint32_t) (kTpBatch * 16);
  const uint32_t stride = gridDim.x * blockDim.x;

  // the window -> the thread's stage
  auto request = [&](const TpCand& c) {
    if (c.kind == 1u) {
#pragma unroll
      for (uint32_t k = 0; k < WB / 16u; ++k) cp_async16(stage + 16u * k, c.q + 16u * k);
    }
    cp_async_commit();
  };
  // a window at the track's ends or across the lead-in (rare): everything looked up again
  auto slow_window = [&](uint32_t i) -> float {
    uint32_t lo = 0, hi = nctas;
    while (hi - lo > 1u) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= i) lo = mid; else hi = mid;
    }
    const unsigned long long e = __ldcs(queue + (size_t) lo * P.run_cta_cap + (i - s_off[lo]));
    const uint32_t slot = (uint32_t) e, w = (uint32_t) (e >> 32);
    const uint32_t ch = w >> 31, pair = (w >> 16) & 0x7fffu;
    const RunItem it = P.items[slot >> 5];
    const Track& tr = P.tracks[it.track];
    const long long frames = (long long) tr.frames;
    const long long t0 = (long long) (it.first_run + (slot & 31u)) * P.Lr - P.Wp + (long long) pair * kPairFrames;
    const unsigned char* pcm = reinterpret_cast<const unsigned char*>(tr.pcm);
    float sw[NW];                                   // (its own array, in local memory: the
#pragma unroll 1                                    // loop is not unrolled)
    for (int k = 0; k < NW; ++k) {
      const long long t = t0 - NT + k;
      float v = 0.0f;
      if (t >= 0 && t < frames) {
        if (FMT == FMT_S16) v = (float) (int) reinterpret_cast<const short*>(pcm)[t * 2 + ch];
        else v = reinterpret_cast<const float*>(pcm)[t * 2 + ch];
      }
      sw[k] = v;
    }
    // the reference produces no output beyond the last frame it was given
    const long long left = frames - t0;
    const int nvalid = left > kPairFrames ? kPairFrames : (int) left;
    float m = 0.0f;
#pragma unroll 1
    for (int k = 0; k < nvalid; ++k) m = fmaxf(m, tp_frame<TPF>(sw, NT + k));
    return m;
  };

  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < count; base += (uint32_t) kTpBatch * stride) {
    TpCand c[kTpBatch];
    if (cells_cached && base >= (uint32_t) kTpBatch * stride)       // later batches: the floors as they are now
      for (uint32_t w = 2u * threadIdx.x + 1u; w < P.npeak_words; w += 2u * blockDim.x)
        atomicMax(&s_cells[w], __ldcg(P.peaks + w));
    // ---- look the batch's entries up, one stage for all of them at a time
    unsigned long long e[kTpBatch];
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      const uint32_t i = base + (uint32_t) j * stride;
      e[j] = 0ull;
      c[j].kind = 0u; c[j].ch = 0u; c[j].cell = nullptr; c[j].seen = 0u; c[j].q = nullptr;
      if (i < count && i >= base) {                 // (i >= base: no wrap-around)
        uint32_t lo = 0, hi = nctas;                // the stretch that holds entry i
        while (hi - lo > 1u) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_off[mid] <= i) lo = mid; else hi = mid;
        }
        e[j] = __ldcs(queue + (size_t) lo * P.run_cta_cap + (i - s_off[lo]));
        c[j].kind = 3u;                             // (an entry: resolved below)
      }
    }
    uint32_t trk[kTpBatch], run0[kTpBatch];
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      trk[j] = 0u; run0[j] = 0u;
      if (c[j].kind) {
        const RunItem* it = P.items + ((uint32_t) e[j] >> 5);
        trk[j] = it->track; run0[j] = it->first_run;
      }
    }
    uint32_t pkb[kTpBatch], lead[kTpBatch];
    unsigned long long frames[kTpBatch];
    const unsigned char* pcm[kTpBatch];
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      pkb[j] = 0u; lead[j] = 0u; frames[j] = 0ull; pcm[j] = nullptr;
      if (c[j].kind) {
        const Track& tr = P.tracks[trk[j]];
        pkb[j] = tr.peak_base; lead[j] = (uint32_t) tr.lead_in; frames[j] = tr.frames;
        pcm[j] = reinterpret_cast<const unsigned char*>(tr.pcm);
      }
    }
    uint32_t spk[kTpBatch];
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      spk[j] = 0u;
      if (c[j].kind) {
        const uint32_t w = (uint32_t) (e[j] >> 32);
        c[j].ch = w >> 31;
        const uint32_t ci = 2u * (pkb[j] + c[j].ch);
        uint32_t* cell = P.peaks + ci;
        spk[j] = cells_cached ? s_cells[ci] : __ldcg(cell);
        c[j].seen = cells_cached ? s_cells[ci + 1u] : __ldcg(cell + 1);
        c[j].cell = cell + 1;
      }
    }
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      if (!c[j].kind) continue;
      const uint32_t slot = (uint32_t) e[j], w = (uint32_t) (e[j] >> 32);
      const uint32_t pair = (w >> 16) & 0x7fffu;
      // What a window has to beat to matter: the channel's sample peak (final; the true peak
      // is reported as the larger of the two, lgb_batch_fetch) and the true-peak cell as it
      // was when the entry was looked up (cells only grow).  Without the first every window
      // of the first round beats an empty cell, and a hundred thousand atomics queue up on a
      // few addresses.
      c[j].seen = max(c[j].seen, spk[j]);
      // a candidate whose bound does not exceed the channel's FINAL sample peak is dropped
      // before its window is fetched
      if (!(P.tp_bound * (peak_code_value(w & 0xffffu) * P.peak_scale) > __uint_as_float(spk[j]))) {
        c[j].kind = 0u;
        continue;
      }
      const long long fr = (long long) frames[j];
      const long long t0 = (long long) (run0[j] + (slot & 31u)) * P.Lr - P.Wp + (long long) pair * kPairFrames;
      if (t0 >= NT && t0 + kPairFrames <= fr && t0 >= (long long) lead[j]) {
        c[j].kind = 1u;
        c[j].q = pcm[j] + (t0 - NT) * (long long) FB;
      } else if (!(t0 >= fr || t0 + kPairFrames <= (long long) lead[j])) {
        c[j].kind = 2u;
      } else {
        c[j].kind = 0u;
      }
    }
    // (kept in shared memory, so that the evaluation below is a loop and not four copies of the FIR)
#pragma unroll
    for (int j = 0; j < kTpBatch; ++j) {
      const unsigned long long qa = (unsigned long long) (uintptr_t) c[j].q;
      const uint32_t ci = c[j].kind ? (uint32_t) (c[j].cell - P.peaks) | (c[j].ch << 29) | (c[j].kind << 30) : 0u;
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(cands + 16u * j), "r"((uint32_t) qa),
                   "r"((uint32_t) (qa >> 32)), "r"(ci), "r"(c[j].seen) : "memory");
    }
    // ---- evaluate them; candidate j + 1's window travels while j's FIR runs
    request(c[0]);
#pragma unroll 1
    for (int j = 0; j < kTpBatch; ++j) {
      const uint4 cd = lds128(cands + 16u * j);
      const uint32_t kind = cd.z >> 30, ch = (cd.z >> 29) & 1u;
      cp_async_wait<0>();
      float win[NW];
      float m = 0.0f;
      if (kind == 1u) {
        if (FMT == FMT_S16) {
          const uint32_t sel = ch ? 0xBB32u : 0x9910u;
#pragma unroll
          for (int k = 0; k < NW / 4; ++k) {
            const uint4 v = lds128(stage + 16u * k);
            win[4 * k + 0] = (float) sext_half(v.x, sel);
            win[4 * k + 1] = (float) sext_half(v.y, sel);
            win[4 * k + 2] = (float) sext_half(v.z, sel);
            win[4 * k + 3] = (float) sext_half(v.w, sel);
          }
        } else {
#pragma unroll
          for (int k = 0; k < NW / 2; ++k) {
            const uint4 v = lds128(stage + 16u * k);
            win[2 * k + 0] = __uint_as_float(ch ? v.y : v.x);
            win[2 * k + 1] = __uint_as_float(ch ? v.w : v.z);
          }
        }
      }
      // (the stage's contents are in registers: the values above depend on every load)
      if (j + 1 < kTpBatch) {
        const uint4 nd = lds128(cands + 16u * (j + 1));
        TpCand nx;
        nx.kind = nd.z >> 30;
        nx.q = reinterpret_cast<const unsigned char*>((uintptr_t) (((unsigned long long) nd.y << 32) | nd.x));
        request(nx);
      }
      if (kind == 1u) {
#pragma unroll
        for (int k = 0; k < kPairFrames; ++k) m = fmaxf(m, tp_frame<TPF>(win, NT + k));
      } else if (kind == 2u) {
        m = slow_window(base + (uint32_t) j * stride);
      }
      if (kind && __float_as_uint(m) > cd.w) {
        const uint32_t ci = cd.z & 0x1fffffffu, mb = __float_as_uint(m);
        if (!cells_cached || mb > atomicMax(&s_cells[ci], mb)) atomicMax(P.peaks + ci, mb);
      }
    }
  }
  cp_async_wait<0>();
}

template <int FMT, int TPF>
static cudaError_t launch_tp_run_t(const SweepParams& p, uint32_t sms, cudaStream_t stream, cudaEvent_t hold,
                                   uint32_t cta_cap) {
  // Resident CTAs per SM of the evaluation, capped at 6 (48 K of the 64 K registers): its
  // CTAs stay for the whole kernel, and the small post-processing kernels that run next to
  // it must always find room on an SM.
  constexpr size_t smem = (size_t) kTpRunThreads * ((TpTraits<TPF>::kTaps + kPairFrames) * (FMT == FMT_S16 ? 4u : 8u) +
                                                    (size_t) kTpBatch * 16u);
  static int per_sm = 0;
  if (!per_sm) {
    int cap = 6;
    if (const char* e = getenv("LOUDGAIN_B200_TPEVAL_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;   // tuning
    // the carve-out of the sweep before it (no reconfiguration of the SMs between the two), which
    // holds the stages of `cap` CTAs; the default one need not
    const int carve = 100;
    cudaError_t e = cudaFuncSetAttribute(tp_eval_run_kernel<FMT, TPF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tp_eval_run_kernel<FMT, TPF>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_eval_run_kernel<FMT, TPF>, kTpRunThreads, smem) !=
            cudaSuccess || per_sm < 1)
      per_sm = 4;
    if (getenv("LOUDGAIN_B200_VERBOSE"))
      fprintf(stderr, "[lgb] tp_eval_run_kernel<%d,%d>: %d CTAs per SM by occupancy, cap %d, %zu B staged per CTA\n",
              FMT, TPF, per_sm, cap, smem);
    if (per_sm > cap) per_sm = cap;
  }
  // `hold`: the evaluation floods the memory system with gathers; started right behind the
  // sweep it slows the FP64 fix-up next to it down by more than it gains (tuning log)
  if (hold) {
    const cudaError_t e = cudaStreamWaitEvent(stream, hold, 0);
    if (e != cudaSuccess) return e;
  }
  const uint32_t nctas = run_sweep_grid(p, sms);                // the sweep's grid (launch_run_k)
  if (nctas > kTpMaxCtas) return cudaErrorInvalidValue;
  const uint32_t ctas = cta_cap && cta_cap < (uint32_t) per_sm ? cta_cap : (uint32_t) per_sm;
  tp_eval_run_kernel<FMT, TPF><<<sms * ctas, kTpRunThreads, smem, stream>>>(p, nctas);
  return cudaGetLastError();
}

cudaError_t launch_truepeak_run(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                                cudaStream_t stream, cudaEvent_t hold, uint32_t cta_cap) {
  if (p.nitems == 0 || tpf == 0) return cudaSuccess;
  if (format == FMT_S16)
    return tpf == 4 ? launch_tp_run_t<FMT_S16, 4>(p, sms, stream, hold, cta_cap)
                    : launch_tp_run_t<FMT_S16, 2>(p, sms, stream, hold, cta_cap);
  return tpf == 4 ? launch_tp_run_t<FMT_F32, 4>(p, sms, stream, hold, cta_cap)
                  : launch_tp_run_t<FMT_F32, 2>(p, sms, stream, hold, cta_cap);
}

}  // namespace lg
