// lg_sweep.cuh -- the per-lane body of the fused sweep: K-weighting + chunk
// energy + sample peak + polyphase true peak over one time chunk of ONE
// channel, started from zero filter state before the chunk.
//
// This is the B200 re-design of what the reference path does serially inside
// ebur128_add_frames_short (/root/reference/src/scan.c:448; behaviour per
// SURVEY.md A.2-A.5).  Time is cut into chunks of L = s100/k frames; every
// (chunk, channel) is filtered independently by one lane, so a single stream
// spreads over all SMs.  What a zero start state gets wrong is linear in the
// true high-pass state at the chunk start, so the sweep also accumulates the
// two cross terms (sum y*alpha, sum y*beta) and snapshots its own high-pass
// state at the chunk's first and last frame; lg_post.cuh composes the
// transition matrices across chunks and applies the exact energy correction
// in FP64.
//
// True peak: the maximum over all polyphase outputs.  An output can never
// exceed ||c||_1 * max|x| over the taps' window, so a window whose bound is
// not above a value the channel's peak is already known to reach cannot move
// the maximum and need not be evaluated.  The device kernel uses that to
// defer only the "candidate" windows to a dense evaluation pass; this file
// provides the window evaluation itself (tp_window) and the exhaustive
// per-frame form used at chunk edges and by the host emulation.  Both give
// the same maximum, bit for bit.
//
// A lane works in iterations of kIter frames, handed to it as raw samples
// x[frame]; how an iteration is run (warm-up / fast / masked) is decided
// uniformly per warp by iter_kind() in lg_common.h.
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

// Upper bounds of the phases' tap L1 norms (1.8642 for 4x, 2.3068 for 2x),
// inflated to cover FP32 rounding of the evaluation.
template <int TPF> LG_HD float tp_gain_bound() { return TPF == 4 ? 1.8645f : 2.3072f; }

// ----- filter state of one channel ----------------------------------------
struct KState {
  float xp;           // previous input sample
  float d1, w1, w2;   // high-pass: last difference, last two integrator values
  float v1, v2;       // shelf
};

// The sweep reads its filter constants from a SweepParams (kernel parameter /
// constant bank on the device).
typedef SweepParams KCoef;

// Host-side: the filter part of a SweepParams from a coefficient set.
inline void fill_kcoef(const CoefSet& cs, SweepParams& k) {
  k.c = cs.c; k.ne2 = -cs.e2; k.np1 = -cs.p1; k.np2 = -cs.p2; k.r1 = cs.r1; k.r2 = cs.r2;
  for (int i = 0; i < kIter; ++i) { k.lam_re[i] = cs.lam_re[i]; k.lam_im[i] = cs.lam_im[i]; }
  k.rot_re = cs.rot_re; k.rot_im = cs.rot_im;
}

// One frame of K-weighting.  Returns the (unnormalised) K-weighted sample.
// High-pass: the numerator (1 - z^-1)^2 is split around the recursion -- the
// first difference is taken on the INPUT (exact for 16-bit samples, and it
// removes any constant offset before it can reach a state variable), the
// second one falls out of the integrator form for free (w[n] - w[n-1] = d[n]):
//   q = x - x[n-1] ; t = q - e2*w[n-2] ; d = c*d[n-1] + t ; w = w[n-1] + d ; yh = d
// so no state ever grows with the input's offset (a direct-form state sits at
// offset / e2, ~3e4 times the offset).
LG_HD float k_step(KState& s, float x, const KCoef& k) {
  const float q = x - s.xp;
  const float t = fmaf(k.ne2, s.w2, q);
  const float d = fmaf(k.c, s.d1, t);
  const float w = s.w1 + d;
  // shelf: v = d - p1 v1 - p2 v2 ; y = v + q1 v1 + q2 v2 = d + (q1 - p1) v1 + (q2 - p2) v2:
  // the output does not wait for the new state (two short chains instead of one of four)
  const float u = fmaf(k.np2, s.v2, d);
  const float v = fmaf(k.np1, s.v1, u);
  const float y = fmaf(k.r2, s.v2, fmaf(k.r1, s.v1, d));
  s.xp = x;
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

// max |phase outputs| over the kIter newest frames of a window
// win[0, NT) = history, win[NT, NT + kIter) = the iteration's frames.
template <int TPF>
LG_HD float tp_window(const float* win) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < kIter; ++i) m = fmaxf(m, tp_frame<TPF>(win, NT + i));
  return m;
}

// Everything one lane carries through its chunk.
struct LaneCtx {
  KState st;
  float sp;                      // raw-unit sample peak of everything the lane has read
  float yr, yi;                  // running sum y[f] lambda^(f - f0 of this iteration)
  double e0;
  float pd, pw, qd, qw;          // state snapshots
  int f_lo, f_hi;                // energy range (lane-local frames)
};

LG_HD void lane_init(LaneCtx& c, int W, int L, const LaneGeom& g) {
  c.st.xp = c.st.d1 = c.st.w1 = c.st.w2 = c.st.v1 = c.st.v2 = 0.0f;
  c.sp = c.yr = c.yi = 0.0f;
  c.e0 = 0.0;
  c.pd = c.pw = c.qd = c.qw = 0.0f;
  c.f_lo = W + g.o;
  c.f_hi = c.f_lo + L;
}

LG_HD float max_abs12(const float* x) {
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < kIter; i += 2) m = fmaxf(m, fmaxf(fabsf(x[i]), fabsf(x[i + 1])));
  return m;
}

// ----- iteration maxima: what the sweep leaves behind for the true-peak pass.
// max |x| of an iteration is stored as a 16-bit code = the upper half of its
// float pattern, rounded UP (a bfloat16 that is never below the value), two
// iterations per 32-bit word.  Codes order like the values they bound.
LG_HD uint32_t peak_code(float m) {
  union { float f; uint32_t u; } v;
  v.f = m;
  return (v.u + 0xffffu) >> 16;
}
LG_HD float peak_code_value(uint32_t code) {
  union { float f; uint32_t u; } v;
  v.u = code << 16;
  return v.f;
}

// Packed sweep: one code per PAIR of iterations and channel.  16-bit PCM
// maxima are small integers and are stored exactly; float maxima as above.
template <int FMT>
LG_HD uint32_t pair_code(float m) { return FMT == FMT_S16 ? (uint32_t) (int) m : peak_code(m); }
template <int FMT>
LG_HD float pair_code_value(uint32_t code) {
  return FMT == FMT_S16 ? (float) (int) code : peak_code_value(code);
}

// Y <- Y * lambda^-kIter + S: keeps Y = sum y[f] lambda^(f - f0) relative to
// the current iteration's first frame, so the constants lambda^i stay O(1).
LG_HD void mode_accumulate(LaneCtx& c, const KCoef& k, float sr, float si) {
  const float nr = fmaf(c.yr, k.rot_re, fmaf(-c.yi, k.rot_im, sr));
  const float ni = fmaf(c.yr, k.rot_im, fmaf(c.yi, k.rot_re, si));
  c.yr = nr;
  c.yi = ni;
}

// Every iteration kind returns max |x| over its kIter frames and folds it
// into the lane's sample peak.  Frames the lane reads beyond its own chunk
// belong to a neighbouring chunk of the same channel, frames beyond the
// track read as zero: neither can change the channel's maximum.

// Start state of a lane: everything at rest and the input difference of the
// very first frame taken as zero (xp = x[0]).  Any start state is exact after
// the fix-up (lg_post.cuh works with the lane's own snapshots); this one adds
// no start transient for an input that rides on a constant offset.  Frames
// before the track read as zero: the first lane of a track starts from the
// reference's own zero state.
LG_HD void lane_start(KState& s, float x0) {
  s.xp = x0;
  s.d1 = s.w1 = s.w2 = s.v1 = s.v2 = 0.0f;
}

// Warm-up iteration: filter state only.  `first`: the lane's very first
// iteration (lane-local frame 0), which sets the start state.
LG_HD float iter_warm(LaneCtx& c, const KCoef& k, const float* x, bool first) {
  if (first) lane_start(c.st, x[0]);
#pragma unroll
  for (int i = 0; i < kIter; ++i) (void) k_step(c.st, x[i], k);
  // state before the first chunk frame of a lane with offset 0; lanes with a
  // larger offset overwrite it in their first masked iteration
  c.pd = c.st.d1; c.pw = c.st.w2;
  const float m = max_abs12(x);
  c.sp = fmaxf(c.sp, m);
  return m;
}

// Fast iteration: all kIter frames lie inside the lane's chunk.
LG_HD float iter_fast(LaneCtx& c, const KCoef& k, const float* x, int f0) {
  float e = 0.0f, sr = 0.0f, si = 0.0f;
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const float y = k_step(c.st, x[i], k);
    e = fmaf(y, y, e);
    sr = fmaf(y, k.lam_re[i], sr);
    si = fmaf(y, k.lam_im[i], si);
  }
  c.e0 += (double) e;
  mode_accumulate(c, k, sr, si);
  if (f0 + kIter == c.f_hi) { c.qd = c.st.d1; c.qw = c.st.w2; }
  const float m = max_abs12(x);
  c.sp = fmaxf(c.sp, m);
  return m;
}

// Masked iteration: frames are tested one by one against the lane's energy
// range; takes the state snapshots.
LG_HD float iter_masked(LaneCtx& c, const KCoef& k, const float* x, int f0) {
  float e = 0.0f, sr = 0.0f, si = 0.0f;
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const int f = f0 + i;
    if (f == c.f_lo) { c.pd = c.st.d1; c.pw = c.st.w2; }
    const float y = k_step(c.st, x[i], k);
    if (f >= c.f_lo && f < c.f_hi) {
      e = fmaf(y, y, e);
      sr = fmaf(y, k.lam_re[i], sr);
      si = fmaf(y, k.lam_im[i], si);
    }
    if (f + 1 == c.f_hi) { c.qd = c.st.d1; c.qw = c.st.w2; }
  }
  c.e0 += (double) e;
  mode_accumulate(c, k, sr, si);
  const float m = max_abs12(x);
  c.sp = fmaxf(c.sp, m);
  return m;
}

// ----- true peak of one iteration -----------------------------------------
// win[0, NT) = the NT frames before the iteration, win[NT, NT + kIter) = its
// frames; only the first `nvalid` frames exist in the track (the reference
// produces no output beyond the last frame it was given).
template <int TPF>
LG_HD float tp_window_valid(const float* win, int nvalid) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < kIter; ++i)
    if (i < nvalid) m = fmaxf(m, tp_frame<TPF>(win, NT + i));
  return m;
}

// Does the true-peak pass have to look at lane-local iteration `it` of a
// lane at all?  Only if it overlaps the lane's own chunk frames (everything
// else is covered by the neighbouring chunk's lane) and starts inside the
// track.
LG_BOTH bool tp_iter_owned(int it, int f_lo, int f_hi, long long a, long long frames) {
  const int f0 = it * kIter;
  return f0 + kIter > f_lo && f0 < f_hi && a + f0 < frames;
}

// ----- host-side sample access (tests/emu) ----------------------------------
// Raw samples of lane-local frames [f0, f0 + kIter) of one channel; frames
// outside the track read as zero, exactly what the device's zero-filling
// cp.async stages.
template <int FMT>
inline void host_load_iter(const void* pcm, long long frames, int channels, long long a, int f0,
                           int ch, float* x) {
  for (int i = 0; i < kIter; ++i) {
    const long long t = a + f0 + i;
    x[i] = 0.0f;
    if (t < 0 || t >= frames) continue;
    if (FMT == FMT_S16) x[i] = (float) ((const short*) pcm)[t * channels + ch];
    else x[i] = ((const float*) pcm)[t * channels + ch];
  }
}

}  // namespace lg
