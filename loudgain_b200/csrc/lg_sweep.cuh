// lg_sweep.cuh -- the per-lane body of the fused sweep: K-weighting + chunk
// energy + sample peak + polyphase true peak over one time chunk of one
// channel pair, started from zero filter state before the chunk.
//
// This is the B200 re-design of what the reference path does serially inside
// ebur128_add_frames_short (/root/reference/src/scan.c:448; behaviour per
// SURVEY.md A.2-A.5).  Time is cut into chunks of L = s100/k frames; every
// chunk is filtered independently, so a single stream spreads over all SMs.
// What a zero start state gets wrong is linear in the true high-pass state at
// the chunk start, so the sweep also accumulates the two cross terms
// (sum y*alpha, sum y*beta) and snapshots its own high-pass state at the
// chunk's first and last frame; lg_post.cuh composes the transition matrices
// across chunks and applies the exact energy correction in FP64.
//
// A lane works in iterations of kIter frames, handed to it as raw samples
// x[channel][frame]; how an iteration is run (warm-up / fast / masked) is
// decided uniformly per warp by iter_kind() in lg_common.h.
//
// All arithmetic here is explicit fmaf/add on floats so that the host
// compile used by tests/emu reproduces the device bit for bit.
#pragma once

#include <math.h>

#include "lg_common.h"
#include "lg_tp_coefs.h"

namespace lg {

template <int TPF> struct TpTraits;
template <> struct TpTraits<4> { static constexpr int kTaps = 12; };
template <> struct TpTraits<2> { static constexpr int kTaps = 24; };
template <> struct TpTraits<0> { static constexpr int kTaps = 0; };

// ----- filter state of one channel ----------------------------------------
struct KState {
  float d1, w1, w2;   // high-pass: last difference, last two integrator values
  float v1, v2;       // shelf
};

// The sweep's view of a coefficient set (registers / uniform registers).
struct KCoef {
  float c, ne2, np1, np2, q1, q2;
};

LG_HD KCoef load_kcoef(const CoefSet& cs) {
  KCoef k;
  k.c = cs.c; k.ne2 = -cs.e2; k.np1 = -cs.p1; k.np2 = -cs.p2; k.q1 = cs.q1; k.q2 = cs.q2;
  return k;
}

// One frame of K-weighting.  Returns the (unnormalised) K-weighted sample.
LG_HD float k_step(KState& s, float x, const KCoef& k) {
  const float t = fmaf(k.ne2, s.w2, x);
  const float d = fmaf(k.c, s.d1, t);
  const float w = s.w1 + d;
  const float yh = d - s.d1;
  const float u = fmaf(k.np2, s.v2, yh);
  const float v = fmaf(k.np1, s.v1, u);
  const float y = fmaf(k.q2, s.v2, fmaf(k.q1, s.v1, v));
  s.w2 = s.w1; s.w1 = w; s.d1 = d;
  s.v2 = s.v1; s.v1 = v;
  return y;
}

// max |phase outputs| for the newest frame at win[idx]; taps ascending = newest
// sample first, the order the reference accumulates in.
template <int TPF>
LG_HD float tp_frame(const float* win, int idx) {
  float m = 0.0f;
  if (TPF == 4) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float acc = 0.0f;
#pragma unroll
      for (int t = 0; t < 12; ++t) acc = fmaf(win[idx - t], kTp4f[p][t], acc);
      m = fmaxf(m, fabsf(acc));
    }
  } else if (TPF == 2) {
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < 24; ++t) acc = fmaf(win[idx - t], kTp2f[0][t], acc);
    m = fabsf(acc);
  }
  return m;
}

// Everything one lane carries through its chunk, for up to two channels.
template <int TPF>
struct LaneCtx {
  static constexpr int NT = TpTraits<TPF>::kTaps;
  static constexpr int WIN = NT + kIter;
  KState st[2];
  float win[2][WIN > 0 ? WIN : 1];   // [0, NT) history, [NT, WIN) this iteration
  float sp[2], tp[2];                // raw-unit sample / true peak
  float xa[2], xb[2];
  double e0[2];
  float pd[2], pw[2], qd[2], qw[2];  // state snapshots
  int f_lo, f_hi, f_tp;              // energy range, true-peak limit (lane-local)
};

template <int TPF>
LG_HD void lane_init(LaneCtx<TPF>& c, int W, int L, const LaneGeom& g) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    c.st[h].d1 = c.st[h].w1 = c.st[h].w2 = c.st[h].v1 = c.st[h].v2 = 0.0f;
#pragma unroll
    for (int i = 0; i < LaneCtx<TPF>::WIN; ++i) c.win[h][i] = 0.0f;
    c.sp[h] = c.tp[h] = c.xa[h] = c.xb[h] = 0.0f;
    c.e0[h] = 0.0;
    c.pd[h] = c.pw[h] = c.qd[h] = c.qw[h] = 0.0f;
  }
  c.f_lo = W + g.o;
  c.f_hi = c.f_lo + L;
  c.f_tp = c.f_lo + g.l_valid;
}

template <int TPF>
LG_HD void win_push(LaneCtx<TPF>& c, int h, const float* x) {
  constexpr int NT = LaneCtx<TPF>::NT;
  if (NT > 0) {
#pragma unroll
    for (int i = 0; i < kIter; ++i) c.win[h][NT + i] = x[i];
  }
}

template <int TPF>
LG_HD void win_slide(LaneCtx<TPF>& c, int h) {
  constexpr int NT = LaneCtx<TPF>::NT;
  if (NT > 0) {
#pragma unroll
    for (int i = 0; i < NT; ++i) c.win[h][i] = c.win[h][i + kIter];
  }
}

// Warm-up iteration: filter state and true-peak history only.
template <int TPF, int NCH>
LG_HD void iter_warm(LaneCtx<TPF>& c, const KCoef& k, const float x[2][kIter]) {
#pragma unroll
  for (int h = 0; h < NCH; ++h) {
#pragma unroll
    for (int i = 0; i < kIter; ++i) (void) k_step(c.st[h], x[h][i], k);
    win_push(c, h, x[h]);
    win_slide(c, h);
    // state before the first chunk frame of a lane with offset 0; lanes with
    // a larger offset overwrite it in their first masked iteration
    c.pd[h] = c.st[h].d1; c.pw[h] = c.st[h].w2;
  }
}

// Fast iteration: all kIter frames lie inside every lane's chunk.
// ab = alpha/beta of frames f0.. as float2 pairs.
template <int TPF, int NCH>
LG_HD void iter_fast(LaneCtx<TPF>& c, const KCoef& k, const float x[2][kIter],
                     const float* ab, int f0) {
  constexpr int NT = LaneCtx<TPF>::NT;
#pragma unroll
  for (int h = 0; h < NCH; ++h) {
    float e = 0.0f, sa = 0.0f, sb = 0.0f;
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const float y = k_step(c.st[h], x[h][i], k);
      e = fmaf(y, y, e);
      sa = fmaf(y, ab[2 * i], sa);
      sb = fmaf(y, ab[2 * i + 1], sb);
      c.sp[h] = fmaxf(c.sp[h], fabsf(x[h][i]));
    }
    c.e0[h] += (double) e;
    c.xa[h] += sa;
    c.xb[h] += sb;
    if (NT > 0) {
      win_push(c, h, x[h]);
#pragma unroll
      for (int i = 0; i < kIter; ++i) c.tp[h] = fmaxf(c.tp[h], tp_frame<TPF>(c.win[h], NT + i));
      win_slide(c, h);
    }
    if (f0 + kIter == c.f_hi) { c.qd[h] = c.st[h].d1; c.qw[h] = c.st[h].w2; }
  }
}

// Masked iteration: frames are tested one by one against the lane's ranges;
// also takes the state snapshots at the chunk's first and last frame.
template <int TPF, int NCH>
LG_HD void iter_masked(LaneCtx<TPF>& c, const KCoef& k, const float x[2][kIter],
                       const float* ab, int f0) {
  constexpr int NT = LaneCtx<TPF>::NT;
#pragma unroll
  for (int h = 0; h < NCH; ++h) {
    float e = 0.0f, sa = 0.0f, sb = 0.0f;
    win_push(c, h, x[h]);
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int f = f0 + i;
      if (f == c.f_lo) { c.pd[h] = c.st[h].d1; c.pw[h] = c.st[h].w2; }
      const float y = k_step(c.st[h], x[h][i], k);
      if (f >= c.f_lo && f < c.f_hi) {
        e = fmaf(y, y, e);
        sa = fmaf(y, ab[2 * i], sa);
        sb = fmaf(y, ab[2 * i + 1], sb);
      }
      if (f + 1 == c.f_hi) { c.qd[h] = c.st[h].d1; c.qw[h] = c.st[h].w2; }
      if (f >= c.f_lo && f < c.f_tp) {
        c.sp[h] = fmaxf(c.sp[h], fabsf(x[h][i]));
        if (NT > 0) c.tp[h] = fmaxf(c.tp[h], tp_frame<TPF>(c.win[h], NT + i));
      }
    }
    c.e0[h] += (double) e;
    c.xa[h] += sa;
    c.xb[h] += sb;
    win_slide(c, h);
  }
}

// ----- host-side sample access (tests/emu) ----------------------------------
// Raw samples of lane-local frames [f0, f0 + kIter) of a channel pair; frames
// outside the track read as zero, exactly what the device's zero-filling
// cp.async stages.
template <int FMT>
inline void host_load_iter(const void* pcm, long long frames, int channels, long long a, int f0,
                           int ch0, int nch, float x[2][kIter]) {
  for (int i = 0; i < kIter; ++i) {
    const long long t = a + f0 + i;
    x[0][i] = 0.0f; x[1][i] = 0.0f;
    if (t < 0 || t >= frames) continue;
    for (int h = 0; h < nch; ++h) {
      if (FMT == FMT_S16) x[h][i] = (float) ((const short*) pcm)[t * channels + ch0 + h];
      else x[h][i] = ((const float*) pcm)[t * channels + ch0 + h];
    }
  }
}

}  // namespace lg
