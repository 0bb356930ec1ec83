// lg_design.h -- host-side (double precision) derivation of everything the
// sweep and the fix-up need for one sample rate: K-weighting coefficients in
// the sweep's FP32-friendly form, the high-pass state-transition powers, the
// alpha/beta correction basis and its Gram sums.
//
// The filter being realised is the K-weighting of SURVEY.md A.3 (what the
// reference path applies inside ebur128_add_frames_short, scan.c:448): a
// +4 dB shelf at 1681.97 Hz cascaded with a 38.1 Hz high-pass.  It is
// re-factored here, not transcribed: the high-pass runs as a leaky double
// integrator so that its near-unit-circle poles survive FP32, and its
// dependence on the state at the chunk start is removed afterwards by a
// linear correction (see lg_post.cuh).
#pragma once

#include <cmath>
#include <vector>

#include "lg_common.h"

namespace lg {

struct KDesign {
  double sb[3], sa[3];   // shelf numerator / denominator (sa[0] = 1)
  double ha[3];          // high-pass denominator (numerator is 1,-2,1)
  double e1, e2, c;      // e1 = 2 + ha1, e2 = 1 + ha1 + ha2, c = 1 - e1
  double shelf_pole_radius, hp_pole_radius;
};

inline KDesign k_design(unsigned long rate) {
  const double kPi = 3.14159265358979323846;
  KDesign k;
  {
    const double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    const double K = std::tan(kPi * f0 / (double) rate);
    const double Vh = std::pow(10.0, G / 20.0), Vb = std::pow(Vh, 0.4996667741545416);
    const double den = 1.0 + K / Q + K * K;
    k.sb[0] = (Vh + Vb * K / Q + K * K) / den;
    k.sb[1] = 2.0 * (K * K - Vh) / den;
    k.sb[2] = (Vh - Vb * K / Q + K * K) / den;
    k.sa[0] = 1.0;
    k.sa[1] = 2.0 * (K * K - 1.0) / den;
    k.sa[2] = (1.0 - K / Q + K * K) / den;
  }
  {
    const double f0 = 38.13547087602444, Q = 0.5003270373238773;
    const double K = std::tan(kPi * f0 / (double) rate);
    const double den = 1.0 + K / Q + K * K;
    k.ha[0] = 1.0;
    k.ha[1] = 2.0 * (K * K - 1.0) / den;
    k.ha[2] = (1.0 - K / Q + K * K) / den;
    // 2 + ha1 and 1 + ha1 + ha2 without cancellation:
    k.e1 = (2.0 * K / Q + 4.0 * K * K) / den;
    k.e2 = 4.0 * K * K / den;
    k.c = 1.0 - k.e1;
  }
  auto radius = [](const double a[3]) {
    const double disc = a[1] * a[1] - 4.0 * a[2];
    if (disc < 0) return std::sqrt(a[2]);
    const double r1 = std::fabs((-a[1] + std::sqrt(disc)) / 2.0);
    const double r2 = std::fabs((-a[1] - std::sqrt(disc)) / 2.0);
    return r1 > r2 ? r1 : r2;
  };
  k.shelf_pole_radius = radius(k.sa);
  k.hp_pole_radius = radius(k.ha);
  return k;
}

inline void mat2_mul(const double a[4], const double b[4], double out[4]) {
  double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3],
                 a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3]};
  for (int i = 0; i < 4; ++i) out[i] = r[i];
}

inline void mat2_pow(const double m[4], unsigned long n, double out[4]) {
  double base[4] = {m[0], m[1], m[2], m[3]};
  double acc[4] = {1, 0, 0, 1};
  while (n) {
    if (n & 1) mat2_mul(acc, base, acc);
    mat2_mul(base, base, base);
    n >>= 1;
  }
  for (int i = 0; i < 4; ++i) out[i] = acc[i];
}

// Warm-up length: long enough for the shelf's own (unknown) start state to
// decay below 1e-5 and to prime the true-peak window; multiple of kIter.
inline int warmup_frames(const KDesign& k) {
  int w = (int) std::ceil(std::log(1e-5) / std::log(k.shelf_pole_radius));
  if (w < kMaxTaps) w = kMaxTaps;
  return ((w + kIter - 1) / kIter) * kIter;
}

inline int true_peak_factor(unsigned long rate) {
  return rate < 96000 ? 4 : (rate < 192000 ? 2 : 0);
}

inline cplx c_mul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline cplx c_div(cplx a, cplx b) {
  const double d = b.re * b.re + b.im * b.im;
  return cplx{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
inline cplx c_add(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
inline cplx c_sub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
inline cplx c_pow(cplx a, long n) {      // integer power, n may be negative
  if (n < 0) { a = c_div(cplx{1, 0}, a); n = -n; }
  cplx r{1, 0};
  while (n) { if (n & 1) r = c_mul(r, a); a = c_mul(a, a); n >>= 1; }
  return r;
}

// Fills `cs` for (rate, chunks-per-slot k, input full scale).
inline void make_coefset(unsigned long rate, int k, double full_scale, CoefSet& cs) {
  const KDesign d = k_design(rate);
  cs = CoefSet();
  cs.s100 = (int32_t) ((rate + 5) / 10);
  cs.k = k;
  cs.L = cs.s100 / k;
  cs.W = warmup_frames(d);
  cs.tpf = true_peak_factor(rate);
  cs.c = (float) d.c;
  cs.e2 = (float) d.e2;
  cs.p1 = (float) d.sa[1];
  cs.p2 = (float) d.sa[2];
  const double q1 = d.sb[1] / d.sb[0], q2 = d.sb[2] / d.sb[0];
  cs.q1 = (float) q1;
  cs.q2 = (float) q2;
  cs.r1 = (float) (q1 - d.sa[1]);
  cs.r2 = (float) (q2 - d.sa[2]);
  cs.gain = (d.sb[0] / full_scale) * (d.sb[0] / full_scale);
  const double M[4] = {d.c, -d.e2, 1.0, 1.0};   // (d1, w2) -> one frame later
  mat2_pow(M, (unsigned long) cs.L, cs.ML);
  {
    const double det = M[0] * M[3] - M[1] * M[2];
    const double inv[4] = {M[3] / det, -M[1] / det, -M[2] / det, M[0] / det};
    for (int o = 0; o < kMaxAlign; ++o) mat2_pow(inv, (unsigned long) (cs.W + o), cs.MinvWo[o]);
  }
  {
    int h = (int) std::ceil(std::log(1e-18) / ((double) cs.L * std::log(d.hp_pole_radius))) + 1;
    cs.horner = h < 2 ? 2 : h;
  }
  // lambda: eigenvalue of M with Im > 0 (the high-pass poles are complex for
  // every supported rate: Q = 0.5003 > 0.5).
  const double tr = 1.0 + d.c, dt = d.c + d.e2, disc = tr * tr - 4.0 * dt;
  const cplx lam = disc < 0 ? cplx{tr / 2.0, std::sqrt(-disc) / 2.0}
                            : cplx{(tr + std::sqrt(disc)) / 2.0, 0.0};
  for (int i = 0; i < kIter; ++i) {
    const cplx p = c_pow(lam, i);
    cs.lam_re[i] = (float) p.re;
    cs.lam_im[i] = (float) p.im;
  }
  {
    const cplx r = c_pow(lam, -kIter);
    cs.rot_re = (float) r.re;
    cs.rot_im = (float) r.im;
  }
  // Response of the K-weighted output to a start state tau = (d1, w2) at
  // lane-local frame 0, once the shelf has settled: Re(A lambda^f) with
  //   tau = a v + conj(a v),  v = (lambda - 1, 1),
  //   A = 2 a (lambda - 1) lambda Hs(lambda),  Hs = shelf transfer function
  // (the state's d component at frame f is d[f-1] = 2 Re(a (lambda-1) lambda^f),
  // and the high-pass output is yh[f] = d[f]).
  {
    const cplx one{1, 0};
    const cplx lm1 = c_sub(lam, one);
    const cplx il = c_div(one, lam), il2 = c_mul(il, il);
    const cplx num = c_add(one, c_add(c_mul(cplx{q1, 0}, il), c_mul(cplx{q2, 0}, il2)));
    const cplx den = c_add(one, c_add(c_mul(cplx{d.sa[1], 0}, il), c_mul(cplx{d.sa[2], 0}, il2)));
    const cplx g = c_mul(c_mul(cplx{2, 0}, c_mul(lm1, lam)), c_div(num, den));
    // a = p + iq with p = tau_w / 2, q = (tau_w x - tau_d) / (2 y), lambda - 1 = x + iy
    const double x = lm1.re, y = lm1.im;
    const cplx a_d{0.0, -1.0 / (2.0 * y)};          // d a / d tau_d
    const cplx a_w{0.5, x / (2.0 * y)};             // d a / d tau_w
    cs.Ad = c_mul(g, a_d);
    cs.Aw = c_mul(g, a_w);
  }
  for (int j = 0; j < 4; ++j) {
    const int aq = 1 << j;
    cs.xi_scale[j] = c_pow(lam, (long) (sweep_iters(cs.W, cs.L, aq) - 1) * kIter);
  }
  {
    const double r2 = lam.re * lam.re + lam.im * lam.im;
    const cplx l2 = c_mul(lam, lam);
    for (int o = 0; o < kMaxAlign; ++o) {
      const long f0 = cs.W + o;
      cs.S1o[o] = std::pow(r2, (double) f0) * (1.0 - std::pow(r2, (double) cs.L)) / (1.0 - r2);
      const cplx top = c_mul(c_pow(l2, f0), c_sub(cplx{1, 0}, c_pow(l2, cs.L)));
      cs.S2o[o] = c_div(top, c_sub(cplx{1, 0}, l2));
    }
  }
}

// Run-sweep part of a coefficient set (lg_run.cu, lg_post.cuh): a lane filters
// R chunks in one go starting Wp frames early; chunk j of a run starts at
// run-local frame j * L.  Its mode sum is kept relative to the first frame of
// the last iteration that fed it (iterations are kIter frames from run-local
// frame -Wp, Wp a multiple of kIter) and only over run-local frames
// < xi_frames; xi[j] = lambda^(that iteration's first frame - j * L) refers it to
// the chunk's own first frame (0 if the chunk starts past xi_frames: no sum).
inline void make_run_coefs(const KDesign& d, int R, int Wp, int xi_frames, CoefSet& cs,
                           std::vector<cplx>& xi_table) {
  cs.run_chunks = R;
  cs.run_warm = Wp;
  cs.xi_frames = xi_frames;
  cs.xi_off = (uint32_t) xi_table.size();
  const double tr = 1.0 + d.c, dt = d.c + d.e2, disc = tr * tr - 4.0 * dt;
  const cplx lam = disc < 0 ? cplx{tr / 2.0, std::sqrt(-disc) / 2.0}
                            : cplx{(tr + std::sqrt(disc)) / 2.0, 0.0};
  for (int j = 0; j < R; ++j) {
    const long f = (long) j * cs.L;
    if (f >= xi_frames) { xi_table.push_back(cplx{0.0, 0.0}); continue; }
    const long end = f + cs.L < xi_frames ? f + cs.L : xi_frames;     // exclusive
    const long it_last = (end - 1) / kIter;
    xi_table.push_back(c_pow(lam, it_last * kIter - f));
  }
  const double r2 = lam.re * lam.re + lam.im * lam.im;
  const cplx l2 = c_mul(lam, lam);
  cs.S1run = (1.0 - std::pow(r2, (double) cs.L)) / (1.0 - r2);
  cs.S2run = c_div(c_sub(cplx{1, 0}, c_pow(l2, cs.L)), c_sub(cplx{1, 0}, l2));
}

inline uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

// Frames between consecutive 16-byte-aligned frame boundaries.
inline uint32_t align_quantum(uint32_t frame_bytes) { return 16u / gcd_u32(16u, frame_bytes); }

// Largest divisor k of s100 whose chunk length s100/k is still >= min_len.
inline int pick_chunks_per_slot(int s100, int min_len) {
  int best = 1;
  for (int k = 1; k <= s100; ++k) {
    if (s100 % k) continue;
    if (s100 / k < min_len) break;
    best = k;
  }
  return best;
}

// BS.1770 channel weights of the default map: 4 ch ->
// L R Ls Rs; 5 ch -> L R C Ls Rs; otherwise L R C (unused) Ls Rs, rest unused.
// Emitted as weight classes (lg_common.h Track::wclass): 1 -> 1.0, 2 -> 1.41.
inline void default_weight_classes(unsigned channels, uint8_t* w) {
  for (unsigned i = 0; i < channels; ++i) w[i] = 0;
  if (channels == 4) {
    w[0] = w[1] = 1; w[2] = w[3] = 2;
  } else if (channels == 5) {
    w[0] = w[1] = w[2] = 1; w[3] = w[4] = 2;
  } else {
    static const uint8_t six[6] = {1, 1, 1, 0, 2, 2};
    for (unsigned i = 0; i < channels && i < 6; ++i) w[i] = six[i];
  }
}

// Weight class of an explicit channel role (enum channel of ebur128.h), for
// states whose map was changed with ebur128_set_channel.
inline uint8_t weight_class_of_role(int role) {
  switch (role) {
    case 0: return 0;                       // UNUSED
    case 4: case 5:                         // Mp110 / Mm110 (surrounds)
    case 9: case 10: case 11: case 12:      // Mp060 Mm060 Mp090 Mm090
      return 2;
    case 6: return 3;                       // DUAL_MONO
    default: return 1;
  }
}

}  // namespace lg
