// lg_post.cuh -- FP64 post-processing of the sweep's chunk records:
//   * carry of the high-pass state across chunk boundaries by composing the
//     chunk transition matrix M^L (a short look-back suffices: |lambda|^L < 1),
//   * exact energy correction of each zero-started chunk,
//   * 100 ms slot energies, 400 ms gating blocks, 3 s short-term blocks
//     (block schedule of SURVEY.md A.2: block b covers slots b..b+3, short-term
//     block j covers slots 10j..10j+29).
// Per-item functions are host/device so tests/emu can run them on the CPU.
#pragma once

#include "lg_common.h"

namespace lg {

LG_HD void mat2_vec(const double m[4], double x, double y, double& ox, double& oy) {
  ox = m[0] * x + m[1] * y;
  oy = m[2] * x + m[3] * y;
}

// True K-weighted energy (still unscaled) of chunk j of one channel.
// recs points at the channel's record of chunk 0; consecutive chunks are
// `stride` records apart.  `o` is the lane's alignment offset and `aq_log2`
// the track's alignment class (lg_common.h, lane_geometry).
//
// Run sweep (cs.run_chunks > 0, lg_run.cu): the snapshots P, Q are taken at every
// chunk boundary of a run, so the carry is the same recurrence; the lane's
// error at the chunk's first frame is T - P itself, its response Re(A lambda^f)
// counts f from that frame, and the mode sum is referred to it by
// xi_table[cs.xi_off + (j mod R)] (lg_design.h: make_run_coefs).  Chunks that
// start past the run's mode-sum horizon have no cross term (the error has
// decayed by 1e-5 there); the quadratic term is always exact.
// T <- Q + M^L (T - P): the true high-pass state one chunk on, given the lane's own
// snapshots of that chunk.
LG_HD void carry_step(const CoefSet& cs, const ChunkRec& r, double& td, double& tw) {
  double mx, my;
  mat2_vec(cs.ML, td - (double) r.pd, tw - (double) r.pw, mx, my);
  td = (double) r.qd + mx;
  tw = (double) r.qw + my;
}

// Energy of chunk j (record r) given the true high-pass state (td, tw) at its first frame.
LG_HD double chunk_energy_at(const CoefSet& cs, const ChunkRec& r, double td, double tw, long long j, int o,
                             int aq_log2, const cplx* xi_table) {
  if (cs.run_chunks > 0) {
    const double ad = td - (double) r.pd, aw = tw - (double) r.pw;
    const double Are = ad * cs.Ad.re + aw * cs.Aw.re, Aim = ad * cs.Ad.im + aw * cs.Aw.im;
    const cplx sc = xi_table[cs.xi_off + (uint32_t) (j % cs.run_chunks)];
    const double xr = (double) r.yr * sc.re - (double) r.yi * sc.im;
    const double xi = (double) r.yr * sc.im + (double) r.yi * sc.re;
    const double cross = Are * xr - Aim * xi;
    const double a2r = Are * Are - Aim * Aim, a2i = 2.0 * Are * Aim;
    const double quad = 0.5 * (Are * Are + Aim * Aim) * cs.S1run +
                        0.5 * (a2r * cs.S2run.re - a2i * cs.S2run.im);
    return r.e0 + 2.0 * cross + quad;
  }
  // State the zero-started run was missing at its own first (warm-up) frame,
  // W + o frames before the chunk ...
  double ad, aw;
  mat2_vec(cs.MinvWo[o], td - (double) r.pd, tw - (double) r.pw, ad, aw);
  // ... whose response in the output is Re(A lambda^f) over the chunk.
  const double Are = ad * cs.Ad.re + aw * cs.Aw.re, Aim = ad * cs.Ad.im + aw * cs.Aw.im;
  const cplx sc = cs.xi_scale[aq_log2];
  const double xr = (double) r.yr * sc.re - (double) r.yi * sc.im;
  const double xi = (double) r.yr * sc.im + (double) r.yi * sc.re;
  const double cross = Are * xr - Aim * xi;                     // Re(A Xi)
  const double a2r = Are * Are - Aim * Aim, a2i = 2.0 * Are * Aim;
  const double quad = 0.5 * (Are * Are + Aim * Aim) * cs.S1o[o] +
                      0.5 * (a2r * cs.S2o[o].re - a2i * cs.S2o[o].im);
  return r.e0 + 2.0 * cross + quad;
}

LG_HD double chunk_true_energy(const CoefSet& cs, const ChunkRec* recs, long long stride,
                               long long j, int o, int aq_log2, const cplx* xi_table) {
  // T = true high-pass state at the first frame of chunk j.
  double td = 0.0, tw = 0.0;
  long long i = j - cs.horner;
  if (i < 0) i = 0;
  for (; i < j; ++i) carry_step(cs, recs[i * stride], td, tw);
  return chunk_energy_at(cs, recs[j * stride], td, tw, j, o, aq_log2, xi_table);
}

LG_HD double weight_of(uint8_t wclass) {
  return wclass == 1 ? 1.0 : (wclass == 2 ? 1.41 : (wclass == 3 ? 2.0 : 0.0));
}

// The snapshots of a chunk record: what the carry needs of it.
struct CarryRec { float pd, pw, qd, qw; };
LG_HD CarryRec carry_rec(const ChunkRec& r) { CarryRec c; c.pd = r.pd; c.pw = r.pw; c.qd = r.qd; c.qw = r.qw; return c; }
LG_HD void carry_step(const CoefSet& cs, const CarryRec& r, double& td, double& tw) {
  double mx, my;
  mat2_vec(cs.ML, td - (double) r.pd, tw - (double) r.pw, mx, my);
  td = (double) r.qd + mx;
  tw = (double) r.qw + my;
}

// Weighted, scaled energy (sum over frames, not yet the mean) of one 100 ms slot,
// straight from the sweep's records: per channel one pass of the state
// carry from `horner` chunks before the slot through its k chunks (what the
// fix-up and slot kernels did in two steps; the carry is shared by the slot's
// chunks instead of being restarted for each).
//
// The carry is a chain -- every step needs the state before it -- but the records it reads are
// not: one record load per step in the chain's shadow made the kernel a sequence of L2 round
// trips (15 us on the whole GPU, several times that on the few SMs a pipelined run leaves it).
// So two channels are carried side by side (independent chains), the look-back records are
// fetched four steps at a time before the steps are taken, and a slot chunk's record is
// fetched while the chunk before it is evaluated.  Per channel the arithmetic and its order
// are what they were.
LG_HD double slot_energy_fused(const Track& tr, const CoefSet& cs, const ChunkRec* recs, uint32_t slot,
                               const cplx* xi_table, int aq_log2) {
  double total = 0.0;
  const long long j0 = (long long) slot * cs.k;
  const long long stride = tr.channels;
  long long i0 = j0 - cs.horner;
  if (i0 < 0) i0 = 0;
  for (uint32_t c0 = 0; c0 < tr.channels; c0 += 2) {
    const bool two = c0 + 1 < tr.channels;
    const double wa = weight_of(tr.wclass[c0]), wb = two ? weight_of(tr.wclass[c0 + 1]) : 0.0;
    if (wa == 0.0 && wb == 0.0) continue;
    const ChunkRec* ra = recs + tr.rec_base + c0;
    const ChunkRec* rb = ra + (two ? 1 : 0);
    double tda = 0.0, twa = 0.0, tdb = 0.0, twb = 0.0;
    long long i = i0;
    for (; i + 4 <= j0; i += 4) {
      CarryRec a[4], b[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int u = 0; u < 4; ++u) {
        a[u] = carry_rec(ra[(i + u) * stride]);
        b[u] = carry_rec(rb[(i + u) * stride]);
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int u = 0; u < 4; ++u) {
        carry_step(cs, a[u], tda, twa);
        carry_step(cs, b[u], tdb, twb);
      }
    }
    for (; i < j0; ++i) {
      const CarryRec a = carry_rec(ra[i * stride]), b = carry_rec(rb[i * stride]);
      carry_step(cs, a, tda, twa);
      carry_step(cs, b, tdb, twb);
    }
    double sa = 0.0, sb = 0.0;
    ChunkRec na = ra[j0 * stride], nb = rb[j0 * stride];
    for (int q = 0; q < cs.k; ++q) {
      const long long j = j0 + q;
      const ChunkRec r_a = na, r_b = nb;
      if (q + 1 < cs.k) { na = ra[(j + 1) * stride]; nb = rb[(j + 1) * stride]; }
      int o = 0;
      if (cs.run_chunks == 0)
        o = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, j).o;
      sa += chunk_energy_at(cs, r_a, tda, twa, j, o, aq_log2, xi_table);
      carry_step(cs, r_a, tda, twa);
      if (two) {
        sb += chunk_energy_at(cs, r_b, tdb, twb, j, o, aq_log2, xi_table);
        carry_step(cs, r_b, tdb, twb);
      }
    }
    if (wa != 0.0) total += wa * sa;
    if (two && wb != 0.0) total += wb * sb;
  }
  return total * cs.gain;
}

LG_HD double gating_block(const double* eslot, int s100, uint64_t b) {
  const double s = (eslot[b] + eslot[b + 1]) + (eslot[b + 2] + eslot[b + 3]);
  return s / (4.0 * (double) s100);
}
LG_HD double gating_block(const double* eslot, const CoefSet& cs, uint32_t b) {
  return gating_block(eslot, cs.s100, b);
}

LG_HD double shortterm_block(const double* eslot, int s100, uint64_t j) {
  double s = 0.0;
  for (int i = 0; i < 30; ++i) s += eslot[10u * j + i];
  return s / (30.0 * (double) s100);
}
LG_HD double shortterm_block(const double* eslot, const CoefSet& cs, uint32_t j) {
  return shortterm_block(eslot, cs.s100, j);
}

// 10^((-70 + 0.691)/10): absolute gate as block energy.
LG_HD double abs_gate_energy() { return 1.1724653045822964e-07; }

LG_HD double energy_to_lufs(double e) { return 10.0 * log10(e) - 0.691; }

}  // namespace lg
