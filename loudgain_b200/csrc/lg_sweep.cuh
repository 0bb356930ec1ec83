// lg_sweep.cuh -- the per-(chunk, channel-pair) body of the fused sweep:
// K-weighting + slot energy + sample peak + polyphase true peak over one time
// chunk of one track, started from zero filter state `W` frames early.
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
// All arithmetic here is explicit fmaf/add on floats so that the host
// compile used by tests/emu reproduces the device bit for bit.
#pragma once

#include <math.h>

#include "lg_common.h"
#include "lg_tp_coefs.h"

namespace lg {

// ----- sample access -------------------------------------------------------
// A Source hands out the raw (unscaled) sample pair of lane-local frame f;
// frames outside the track read as zero (the reference starts from a zeroed
// delay line and filter state).
template <int FMT>
struct GlobalSource {
  const void* pcm;     // frame 0 of the track
  long long frames;    // track length
  long long origin;    // track frame of lane-local frame 0 (may be negative)
  int channels;
  int ch0;             // first channel of the pair
  int nch;             // 1 or 2 live channels in the pair

  LG_HD void get(int f, float& a, float& b) const {
    const long long t = origin + f;
    a = 0.0f; b = 0.0f;
    if (t < 0 || t >= frames) return;
    if (FMT == FMT_S16) {
      const short* p = (const short*) pcm + t * channels + ch0;
#if defined(__CUDA_ARCH__)
      a = (float) __ldg(p);
      if (nch > 1) b = (float) __ldg(p + 1);
#else
      a = (float) p[0];
      if (nch > 1) b = (float) p[1];
#endif
    } else {
      const float* p = (const float*) pcm + t * channels + ch0;
#if defined(__CUDA_ARCH__)
      a = __ldg(p);
      if (nch > 1) b = __ldg(p + 1);
#else
      a = p[0];
      if (nch > 1) b = p[1];
#endif
    }
  }
};

LG_HD void basis_at(const float* basis, int f, float& al, float& be) {
#if defined(__CUDA_ARCH__)
  const float2 t = __ldg((const float2*) basis + f);
  al = t.x; be = t.y;
#else
  al = basis[2 * f]; be = basis[2 * f + 1];
#endif
}

// ----- filter state of one channel ----------------------------------------
struct KState {
  float d1, w1, w2;   // high-pass: last difference, last two integrator values
  float v1, v2;       // shelf
};

// One frame of K-weighting.  Returns the (unnormalised) K-weighted sample.
LG_HD float k_step(KState& s, float x, const CoefSet& cs) {
  const float t = fmaf(-cs.e2, s.w2, x);
  const float d = fmaf(cs.c, s.d1, t);
  const float w = s.w1 + d;
  const float yh = d - s.d1;
  const float u = fmaf(-cs.p2, s.v2, yh);
  const float v = fmaf(-cs.p1, s.v1, u);
  const float y = fmaf(cs.q2, s.v2, fmaf(cs.q1, s.v1, v));
  s.w2 = s.w1; s.w1 = w; s.d1 = d;
  s.v2 = s.v1; s.v1 = v;
  return y;
}

template <int TPF> struct TpTraits;
template <> struct TpTraits<4> { static constexpr int kTaps = 12, kPhases = 3; };
template <> struct TpTraits<2> { static constexpr int kTaps = 24, kPhases = 1; };
template <> struct TpTraits<0> { static constexpr int kTaps = 0, kPhases = 0; };

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

// Result of one chunk for one channel, before it is written out.
struct ChanOut {
  double e0;
  float xa, xb;
  float pd, pw, qd, qw;
  float sp, tp;
};

// Processes one chunk for a pair of channels.
//   L_energy : frames over which energy is accumulated (cs.L)
//   L_valid  : frames of the chunk that exist in the track (<= L_energy);
//              true-peak outputs beyond it are discarded.
template <int TPF, class Source>
LG_HD void sweep_chunk(const CoefSet& cs, const float* basis, const Source& src,
                       int L_energy, int L_valid, ChanOut out[2]) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  constexpr int WIN = NT + kIter;
  KState st[2];
  float win[2][WIN > 0 ? WIN : 1];
  float sp[2] = {0.0f, 0.0f}, tp[2] = {0.0f, 0.0f};
  float xa[2] = {0.0f, 0.0f}, xb[2] = {0.0f, 0.0f};
  double e0[2] = {0.0, 0.0};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    st[h].d1 = st[h].w1 = st[h].w2 = st[h].v1 = st[h].v2 = 0.0f;
#pragma unroll
    for (int i = 0; i < WIN; ++i) win[h][i] = 0.0f;
  }

  const int W = cs.W;
  // ---- warm-up: state only, keeps the true-peak window primed
  for (int f0 = 0; f0 < W; f0 += kIter) {
    float x[2][kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) src.get(f0 + i, x[0][i], x[1][i]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < kIter; ++i) (void) k_step(st[h], x[h][i], cs);
      if (NT > 0) {
        // same slide as the main loop: newest frames enter at the top, then
        // the window moves down so that win[0..NT) is the history
#pragma unroll
        for (int i = 0; i < kIter; ++i) win[h][NT + i] = x[h][i];
#pragma unroll
        for (int i = 0; i < NT; ++i) win[h][i] = win[h][i + kIter];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    out[h].pd = st[h].d1; out[h].pw = st[h].w2;
    out[h].qd = st[h].d1; out[h].qw = st[h].w2;
  }

  const int f_hi = W + L_energy;
  const int f_tp = W + L_valid;
  const int f_end = L_valid < L_energy ? f_tp : f_hi;   // partial chunks stop early
  const int n_full = (L_valid < L_energy ? L_valid : L_energy) / kIter;

  // ---- main iterations: every frame is inside the chunk
  int f0 = W;
  for (int it = 0; it < n_full; ++it, f0 += kIter) {
    float x[2][kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) src.get(f0 + i, x[0][i], x[1][i]);
    float al[kIter], be[kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) basis_at(basis, f0 + i, al[i], be[i]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float e = 0.0f, sa = 0.0f, sb = 0.0f;
#pragma unroll
      for (int i = 0; i < kIter; ++i) {
        const float y = k_step(st[h], x[h][i], cs);
        e = fmaf(y, y, e);
        sa = fmaf(y, al[i], sa);
        sb = fmaf(y, be[i], sb);
        sp[h] = fmaxf(sp[h], fabsf(x[h][i]));
      }
      e0[h] += (double) e;
      xa[h] += sa;
      xb[h] += sb;
      if (NT > 0) {
#pragma unroll
        for (int i = 0; i < kIter; ++i) win[h][NT + i] = x[h][i];
#pragma unroll
        for (int i = 0; i < kIter; ++i) tp[h] = fmaxf(tp[h], tp_frame<TPF>(win[h], NT + i));
#pragma unroll
        for (int i = 0; i < NT; ++i) win[h][i] = win[h][i + kIter];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) { out[h].qd = st[h].d1; out[h].qw = st[h].w2; }

  // ---- tail iterations: frames are masked one by one
  for (; f0 < f_end; f0 += kIter) {
    float x[2][kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) src.get(f0 + i, x[0][i], x[1][i]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float e = 0.0f, sa = 0.0f, sb = 0.0f;
      if (NT > 0) {
#pragma unroll
        for (int i = 0; i < kIter; ++i) win[h][NT + i] = x[h][i];
      }
#pragma unroll
      for (int i = 0; i < kIter; ++i) {
        const int f = f0 + i;
        const float y = k_step(st[h], x[h][i], cs);
        if (f < f_hi) {
          float al, be;
          basis_at(basis, f, al, be);
          e = fmaf(y, y, e);
          sa = fmaf(y, al, sa);
          sb = fmaf(y, be, sb);
        }
        if (f + 1 == f_hi) { out[h].qd = st[h].d1; out[h].qw = st[h].w2; }
        if (f < f_tp) {
          sp[h] = fmaxf(sp[h], fabsf(x[h][i]));
          if (NT > 0) tp[h] = fmaxf(tp[h], tp_frame<TPF>(win[h], NT + i));
        }
      }
      e0[h] += (double) e;
      xa[h] += sa;
      xb[h] += sb;
      if (NT > 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i) win[h][i] = win[h][i + kIter];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    out[h].e0 = e0[h]; out[h].xa = xa[h]; out[h].xb = xb[h];
    out[h].sp = sp[h]; out[h].tp = tp[h];
  }
}

}  // namespace lg
