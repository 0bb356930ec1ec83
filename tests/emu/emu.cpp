// emu.cpp -- TEST-ONLY host compile of the per-thread device math.
//
// The sweep / fix-up / block functions in loudgain_b200/csrc/*.cuh are written
// with explicit fmaf so that a plain C++ build reproduces the GPU arithmetic
// bit for bit.  This file runs them sequentially over a plan so that CPU-only
// tests can check the kernel logic against the oracle before any GPU time is
// spent.  It is NOT part of the product: libebur128.so does not contain it and
// has no CPU path; nothing under loudgain_b200/ loads this library.
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ebur128_b200.h"
#include "../../loudgain_b200/csrc/lg_plan.h"
#include "../../loudgain_b200/csrc/lg_post.cuh"
#include "../../loudgain_b200/csrc/lg_sweep.cuh"

using namespace lg;

// One lane's iteration maxima, kept for the screened true-peak pass.
struct LaneCodes {
  uint32_t track, ch;
  LaneGeom geo;
  std::vector<uint32_t> code;   // per iteration (lg_sweep.cuh: peak_code)
  std::vector<float> iter_max;  // per iteration: max |x|
};

// peaks: [2 * total_peaks] = (sample peak, exhaustive true peak) per channel;
// tp_screened: [total_peaks] true peak of the two-pass scheme the device runs
// (sweep records iteration maxima, the true-peak pass evaluates only the
// iterations whose bound exceeds the channel's final sample peak).
// Plan options as lgb_batch_create sets them (lg_batch.cu), tuning variables included.
static PlanOptions emu_options(uint64_t target_tasks) {
  PlanOptions opt;
  opt.target_tasks = target_tasks;
  if (const char* e = getenv("LOUDGAIN_B200_RUN")) opt.use_run = atoi(e) != 0;
  if (const char* e = getenv("LOUDGAIN_B200_RUN_CHUNKS")) opt.force_run_chunks = atoi(e);
  if (const char* e = getenv("LOUDGAIN_B200_CHUNKS_PER_SLOT")) opt.force_k = atoi(e);
  if (const char* e = getenv("LOUDGAIN_B200_SMS")) opt.sms = (uint32_t) atoi(e);
  if (const char* e = getenv("LOUDGAIN_B200_TAIL_FRAC")) opt.tail_frac = atof(e);
  if (const char* e = getenv("LOUDGAIN_B200_TAIL_DIV")) opt.tail_div = atoi(e);
  return opt;
}

template <int FMT, int TPF>
static void run_group(const Plan& p, const SweepGroup& g, std::vector<ChunkRec>& recs,
                      std::vector<float>& peaks, std::vector<float>& tp_screened) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  const SweepParams& k = g.params;
  std::vector<LaneCodes> lanes;
  for (uint32_t w = 0; w < g.nwarps; ++w) {
    const WarpWork ww = p.work[g.first_warp + w];
    const Track& tr = p.tracks[ww.track];
    const CoefSet& cs = p.coefs[tr.coef];
    const uint32_t C = tr.channels, lpc = k.lpc, cpw = k.cpw;
    // a packed lane (even channel counts, lg_pair.cu) holds channels 2j and 2j+1
    for (uint32_t vlane = 0; vlane < (k.packed ? 64u : 32u); ++vlane) {
      const uint32_t lane = k.packed ? vlane / 2u : vlane;
      const uint32_t slot = lane / lpc, chl = lane - slot * lpc;
      const uint32_t ch = k.packed ? 2u * chl + (vlane & 1u) : ww.ch_base + chl;
      const uint32_t chunk = ww.first_chunk + slot;
      if (!(slot < cpw && ch < C && chunk < tr.nchunks)) continue;
      const LaneGeom geo = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, chunk);
      LaneCtx c;
      lane_init(c, cs.W, cs.L, geo);
      LaneCodes lc{ww.track, ch, geo, {}, {}};
      float win[(NT > 0 ? NT : 1) + kIter] = {0};
      float tp = 0.0f;
      for (uint32_t it = 0; it < tr.niters; ++it) {
        const int f0 = (int) it * kIter;
        float* x = win + NT;
        host_load_iter<FMT>(tr.pcm, (long long) tr.frames, (int) C, geo.a, f0, (int) ch, x);
        const int kind = iter_kind(f0, cs.W, (int) tr.aq, cs.L, ww.lmin_valid);
        const float m = kind == ITER_WARM ? iter_warm(c, k, x, it == 0)
                      : kind == ITER_FAST ? iter_fast(c, k, x, f0) : iter_masked(c, k, x, f0);
        lc.code.push_back(peak_code(m));
        lc.iter_max.push_back(m);
        if (NT > 0) {
          if (tp_iter_owned((int) it, c.f_lo, c.f_hi, geo.a, (long long) tr.frames) &&
              geo.a + f0 + kIter > (long long) tr.lead_in) {
            const long long left = (long long) tr.frames - (geo.a + f0);
            tp = std::max(tp, tp_window_valid<TPF>(win, left > kIter ? kIter : (int) left));
          }
          for (int i = 0; i < NT; ++i) win[i] = win[i + kIter];   // NT >= kIter
        }
      }
      ChunkRec& r = recs[tr.rec_base + (uint64_t) chunk * C + ch];
      r.e0 = c.e0; r.yr = c.yr; r.yi = c.yi;
      r.pd = c.pd; r.pw = c.pw; r.qd = c.qd; r.qw = c.qw;
      float& sp = peaks[2 * (tr.peak_base + ch)];
      float& tpx = peaks[2 * (tr.peak_base + ch) + 1];
      sp = std::max(sp, c.sp);
      tpx = std::max(tpx, tp);
      if (NT > 0) lanes.push_back(std::move(lc));
    }
  }
  // ---- the device's second pass, on the recorded codes
  for (const LaneCodes& lc : lanes) {
    const Track& tr = p.tracks[lc.track];
    const CoefSet& cs = p.coefs[tr.coef];
    const int f_lo = cs.W + lc.geo.o, f_hi = f_lo + cs.L;
    const float floor_ = peaks[2 * (tr.peak_base + lc.ch)];
    float& out = tp_screened[tr.peak_base + lc.ch];
    if (k.packed) {
      // truepeak_pair_kernel: one code per pair of iterations, screened over
      // the pair and the pair before it; both iterations are evaluated
      const long long left0 = (long long) tr.frames - lc.geo.a;
      const int f_end = left0 < (long long) f_hi ? (left0 < 0 ? 0 : (int) left0) : f_hi;
      auto pcode = [&](uint32_t pr) {
        uint32_t c0 = pair_code<FMT>(lc.iter_max[2 * pr]);
        if (2 * pr + 1 < tr.niters) c0 = std::max(c0, pair_code<FMT>(lc.iter_max[2 * pr + 1]));
        return c0;
      };
      for (uint32_t pr = 0; pr < k.npairs; ++pr) {
        uint32_t cm = pcode(pr);
        if (pr) cm = std::max(cm, pcode(pr - 1));
        const int f0 = (int) pr * kPairFrames;
        if (!(f0 + kPairFrames > f_lo && f0 < f_end)) continue;
        if (!(lc.geo.a + f0 + kPairFrames > (long long) tr.lead_in)) continue;
        if (!(k.tp_bound * pair_code_value<FMT>(cm) > floor_)) continue;
        for (int h = 0; h < 2; ++h) {
          const long long t0 = lc.geo.a + f0 + h * kIter;
          if (t0 >= (long long) tr.frames) break;
          float win[(NT > 0 ? NT : 1) + kIter];
          for (int q = 0; q < NT + kIter; ++q) {
            const long long t = t0 - NT + q;
            win[q] = 0.0f;
            if (t < 0 || t >= (long long) tr.frames) continue;
            if (FMT == FMT_S16) win[q] = (float) ((const short*) tr.pcm)[t * tr.channels + lc.ch];
            else win[q] = ((const float*) tr.pcm)[t * tr.channels + lc.ch];
          }
          const long long left = (long long) tr.frames - t0;
          out = std::max(out, tp_window_valid<TPF>(win, left > kIter ? kIter : (int) left));
        }
      }
      continue;
    }
    for (uint32_t it = 0; it < tr.niters; ++it) {
      uint32_t cm = lc.code[it];
      for (uint32_t b = 1; b <= (uint32_t) NT / kIter && b <= it; ++b) cm = std::max(cm, lc.code[it - b]);
      if (!(k.tp_bound * peak_code_value(cm) > floor_)) continue;
      if (!tp_iter_owned((int) it, f_lo, f_hi, lc.geo.a, (long long) tr.frames)) continue;
      if (!(lc.geo.a + (long long) it * kIter + kIter > (long long) tr.lead_in)) continue;
      float win[(NT > 0 ? NT : 1) + kIter];
      const long long t0 = lc.geo.a + (long long) it * kIter;
      for (int q = 0; q < NT + kIter; ++q) {
        const long long t = t0 - NT + q;
        win[q] = 0.0f;
        if (t < 0 || t >= (long long) tr.frames) continue;
        if (FMT == FMT_S16) win[q] = (float) ((const short*) tr.pcm)[t * tr.channels + lc.ch];
        else win[q] = ((const float*) tr.pcm)[t * tr.channels + lc.ch];
      }
      const long long left = (long long) tr.frames - t0;
      out = std::max(out, tp_window_valid<TPF>(win, left > kIter ? kIter : (int) left));
    }
  }
}

// ---- run sweep (lg_run.cu): one lane = one run of R chunks, both channels of a
// stereo track; the scalar restatement of the kernel's per-channel arithmetic.
template <int FMT, int TPF>
static void run_group_runs(const Plan& p, const SweepGroup& g, std::vector<ChunkRec>& recs,
                           std::vector<float>& peaks, std::vector<float>& tp_screened) {
  constexpr int NT = TpTraits<TPF>::kTaps;
  const SweepParams& k = g.params;
  for (uint32_t ii = 0; ii < g.nitems; ++ii) {
    const RunItem item = p.items[g.first_item + ii];
    const Track& tr = p.tracks[item.track];
    const CoefSet& cs = p.coefs[tr.coef];
    const int L = cs.L, R = cs.run_chunks, Wp = cs.run_warm;
    const long long Lr = (long long) R * L;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      const uint32_t run = item.first_run + lane;
      if (run >= tr.nruns) continue;
      const long long a = (long long) run * Lr - Wp;       // track frame of lane-local frame 0
      for (uint32_t ch = 0; ch < 2; ++ch) {
        KState st;
        st.xp = st.d1 = st.w1 = st.w2 = st.v1 = st.v2 = 0.0f;
        double e = 0.0;
        LaneCtx c;                 // only yr / yi (mode_accumulate)
        c.yr = c.yi = 0.0f;
        float pd = 0.0f, pw = 0.0f, sp = 0.0f, tp = 0.0f;
        int j = 0;
        long long nb = Wp + L;
        std::vector<float> iter_max;
        float win[(NT > 0 ? NT : 1) + kIter] = {0};
        auto close_chunk = [&]() {
          const uint64_t chunk = (uint64_t) run * R + j;
          if (chunk < tr.nchunks) {
            ChunkRec& r = recs[tr.rec_base + chunk * 2 + ch];
            r.e0 = e; r.yr = c.yr; r.yi = c.yi; r.pd = pd; r.pw = pw; r.qd = st.d1; r.qw = st.w2;
          }
          e = 0.0; c.yr = c.yi = 0.0f;
          pd = st.d1; pw = st.w2;
          ++j; nb += L;
        };
        for (uint32_t it = 0; it < tr.niters; ++it) {
          const int g0 = (int) it * kIter;
          float* x = win + NT;
          host_load_iter<FMT>(tr.pcm, (long long) tr.frames, 2, a, g0, (int) ch, x);
          const float m = max_abs12(x);
          sp = std::max(sp, m);
          iter_max.push_back(m);
          // the kernel filters 16-bit samples in units of 1 / 65536 (exact scaling)
          float xs[kIter];
          for (int i = 0; i < kIter; ++i) xs[i] = FMT == FMT_S16 ? x[i] * kRunS16Scale : x[i];
          x = xs;
          if (it == 0) lane_start(st, x[0]);
          if (g0 + kIter <= Wp) {
            for (int i = 0; i < kIter; ++i) (void) k_step(st, x[i], k);
            if (g0 + kIter == Wp) { pd = st.d1; pw = st.w2; }
          } else {
            const bool xi_on = (int) it < k.xi_iters;
            const int ib = nb >= g0 + kIter ? kIter : (int) (nb - g0);    // split position (kIter: none)
            float ef = 0.0f, sr = 0.0f, si = 0.0f;
            for (int i = 0; i < kIter; ++i) {
              if (i == ib) {
                e += (double) ef;
                if (xi_on) mode_accumulate(c, k, sr, si);
                close_chunk();
                ef = sr = si = 0.0f;
              }
              const float y = k_step(st, x[i], k);
              ef = fmaf(y, y, ef);
              if (xi_on) { sr = fmaf(y, k.lam_re[i], sr); si = fmaf(y, k.lam_im[i], si); }
            }
            e += (double) ef;
            if (xi_on) mode_accumulate(c, k, sr, si);
            if (nb == g0 + kIter) close_chunk();
          }
          if (NT > 0) {
            // exhaustive true peak over the pairs the lane owns (pair granularity, as the device)
            const int p0 = (g0 / kPairFrames) * kPairFrames;
            if (p0 >= Wp && a + p0 < (long long) tr.frames && a + p0 + kPairFrames > (long long) tr.lead_in) {
              const long long left = (long long) tr.frames - (a + g0);
              if (left > 0) tp = std::max(tp, tp_window_valid<TPF>(win, left > kIter ? kIter : (int) left));
            }
            for (int i = 0; i < NT; ++i) win[i] = win[i + kIter];
          }
        }
        float& spo = peaks[2 * (tr.peak_base + ch)];
        float& tpo = peaks[2 * (tr.peak_base + ch) + 1];
        spo = std::max(spo, sp);
        tpo = std::max(tpo, tp);
        if (NT > 0) {
          // deferred: needs the channel's final sample peak -> second loop below
        }
      }
    }
  }
  if (NT == 0) return;
  // ---- the device's screened pass: a pair is evaluated iff the bound over it
  // and the pair before it exceeds the channel's final sample peak
  for (uint32_t ii = 0; ii < g.nitems; ++ii) {
    const RunItem item = p.items[g.first_item + ii];
    const Track& tr = p.tracks[item.track];
    const CoefSet& cs = p.coefs[tr.coef];
    const int Wp = cs.run_warm;
    const long long Lr = (long long) cs.run_chunks * cs.L;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      const uint32_t run = item.first_run + lane;
      if (run >= tr.nruns) continue;
      const long long a = (long long) run * Lr - Wp;
      for (uint32_t ch = 0; ch < 2; ++ch) {
        const float floor_ = peaks[2 * (tr.peak_base + ch)];
        float& out = tp_screened[tr.peak_base + ch];
        auto pair_max = [&](uint32_t pr) {
          float m = 0.0f;
          for (int h = 0; h < 2; ++h) {
            const uint32_t it = 2 * pr + h;
            if (it >= tr.niters) break;
            float x[kIter];
            host_load_iter<FMT>(tr.pcm, (long long) tr.frames, 2, a, (int) it * kIter, (int) ch, x);
            m = std::max(m, max_abs12(x));
          }
          return m;
        };
        for (uint32_t pr = (uint32_t) Wp / kPairFrames; pr < k.npairs; ++pr) {
          uint32_t cm = pair_code<FMT>(pair_max(pr));
          if (pr) cm = std::max(cm, pair_code<FMT>(pair_max(pr - 1)));
          const long long t0 = a + (long long) pr * kPairFrames;
          if (t0 >= (long long) tr.frames || t0 + kPairFrames <= (long long) tr.lead_in) continue;
          if (!(k.tp_bound * pair_code_value<FMT>(cm) > floor_)) continue;
          float win[(NT > 0 ? NT : 1) + kPairFrames];
          for (int q = 0; q < NT + kPairFrames; ++q) {
            const long long t = t0 - NT + q;
            win[q] = 0.0f;
            if (t < 0 || t >= (long long) tr.frames) continue;
            if (FMT == FMT_S16) win[q] = (float) ((const short*) tr.pcm)[t * 2 + ch];
            else win[q] = ((const float*) tr.pcm)[t * 2 + ch];
          }
          const long long left = (long long) tr.frames - t0;
          const int nvalid = left > kPairFrames ? kPairFrames : (int) left;
          for (int i = 0; i < nvalid; ++i) out = std::max(out, tp_frame<TPF>(win, NT + i));
        }
      }
    }
  }
}

static int log2u(uint32_t v) { int r = 0; while (v > 1) { v >>= 1; ++r; } return r; }

static void run_query(const Plan& p, const Query& q, const std::vector<double>& zblock,
                      const std::vector<double>& zst, double abs_gate, lgb_result& r) {
  memset(&r, 0, sizeof(r));
  r.loudness = -HUGE_VAL;
  double s = 0; uint64_t n = 0;
  for (uint32_t m = 0; m < q.count; ++m) {
    const Track& tr = p.tracks[p.members[q.first + m]];
    for (uint32_t i = 0; i < tr.nblocks; ++i) {
      const double e = zblock[tr.block_base + i];
      if (e >= abs_gate) { s += e; ++n; }
    }
  }
  r.sum_abs = s; r.n_abs = n;
  if (n) {
    const double thr = s / (double) n * 0.1;
    r.rel_threshold = thr;
    s = 0; n = 0;
    for (uint32_t m = 0; m < q.count; ++m) {
      const Track& tr = p.tracks[p.members[q.first + m]];
      for (uint32_t i = 0; i < tr.nblocks; ++i) {
        const double e = zblock[tr.block_base + i];
        if (e >= abs_gate && e >= thr) { s += e; ++n; }
      }
    }
    r.sum_rel = s; r.n_rel = n;
    if (n) r.loudness = energy_to_lufs(s / (double) n);
  }
  std::vector<double> v;
  for (uint32_t m = 0; m < q.count; ++m) {
    const Track& tr = p.tracks[p.members[q.first + m]];
    for (uint32_t i = 0; i < tr.nst; ++i)
      if (zst[tr.st_base + i] >= abs_gate) v.push_back(zst[tr.st_base + i]);
  }
  r.n_shortterm = v.size();
  if (!v.empty()) {
    std::sort(v.begin(), v.end());
    double mean = 0;
    for (double e : v) mean += e;
    mean /= (double) v.size();
    const double fl = 0.01 * mean;
    size_t first = 0;
    while (first < v.size() && v[first] < fl) ++first;
    const size_t m = v.size() - first;
    if (m) {
      const double hi = v[first + (size_t) ((double) (m - 1) * 0.95 + 0.5)];
      const double lo = v[first + (size_t) ((double) (m - 1) * 0.1 + 0.5)];
      r.range = energy_to_lufs(hi) - energy_to_lufs(lo);
    }
  }
}

// Host pointers in `tracks[i].pcm`.  blocks_out / st_out (optional) receive the
// concatenated gating / short-term block energies of all tracks.
extern "C" int emu_measure(const lgb_track* tracks, size_t ntracks, uint32_t nalbums,
                           uint64_t target_tasks, lgb_result* track_results,
                           lgb_result* album_results, double* sample_peaks, double* true_peaks,
                           double* blocks_out, double* st_out, int32_t* chunk_len_out,
                           double* true_peaks_screened, double* slots_out) {
  std::vector<TrackIn> in(ntracks);
  for (size_t i = 0; i < ntracks; ++i)
    in[i] = TrackIn{tracks[i].pcm, tracks[i].frames, tracks[i].channels, tracks[i].samplerate,
                    tracks[i].format, tracks[i].album, tracks[i].weight_class, tracks[i].lead_in};
  Plan p;
  const PlanOptions opt = emu_options(target_tasks);
  build_plan(in.data(), ntracks, nalbums, opt, p);
  std::vector<ChunkRec> recs(p.total_recs);
  std::vector<float> peaks(2 * p.total_peaks, 0.0f), tps(p.total_peaks, 0.0f);
  for (const SweepGroup& g : p.groups) {
    if (g.run) {
      if (g.format == FMT_S16) {
        if (g.tpf == 4) run_group_runs<FMT_S16, 4>(p, g, recs, peaks, tps);
        else if (g.tpf == 2) run_group_runs<FMT_S16, 2>(p, g, recs, peaks, tps);
        else run_group_runs<FMT_S16, 0>(p, g, recs, peaks, tps);
      } else {
        if (g.tpf == 4) run_group_runs<FMT_F32, 4>(p, g, recs, peaks, tps);
        else if (g.tpf == 2) run_group_runs<FMT_F32, 2>(p, g, recs, peaks, tps);
        else run_group_runs<FMT_F32, 0>(p, g, recs, peaks, tps);
      }
      continue;
    }
    if (g.format == FMT_S16) {
      if (g.tpf == 4) run_group<FMT_S16, 4>(p, g, recs, peaks, tps);
      else if (g.tpf == 2) run_group<FMT_S16, 2>(p, g, recs, peaks, tps);
      else run_group<FMT_S16, 0>(p, g, recs, peaks, tps);
    } else {
      if (g.tpf == 4) run_group<FMT_F32, 4>(p, g, recs, peaks, tps);
      else if (g.tpf == 2) run_group<FMT_F32, 2>(p, g, recs, peaks, tps);
      else run_group<FMT_F32, 0>(p, g, recs, peaks, tps);
    }
  }
  std::vector<double> eslot(p.total_slots), zblock(p.total_blocks),
      zst(p.total_st);
  for (size_t ti = 0; ti < ntracks; ++ti) {
    const Track& tr = p.tracks[ti];
    const CoefSet& cs = p.coefs[tr.coef];
    if (chunk_len_out) chunk_len_out[ti] = cs.L;
    for (uint32_t s = 0; s < tr.nslots; ++s)
      eslot[tr.slot_base + s] = slot_energy_fused(tr, cs, recs.data(), s, p.xi_table.data(), log2u(tr.aq));
    for (uint32_t b = 0; b < tr.nblocks; ++b)
      zblock[tr.block_base + b] = gating_block(eslot.data() + tr.slot_base, cs, b);
    for (uint32_t j = 0; j < tr.nst; ++j)
      zst[tr.st_base + j] = shortterm_block(eslot.data() + tr.slot_base, cs, j);
  }
  const double abs_gate = pow(10.0, (-70.0 + 0.691) / 10.0);
  for (size_t i = 0; i < ntracks; ++i)
    if (track_results) run_query(p, p.queries[i], zblock, zst, abs_gate, track_results[i]);
  for (uint32_t a = 0; a < nalbums; ++a)
    if (album_results) run_query(p, p.queries[ntracks + a], zblock, zst, abs_gate, album_results[a]);
  for (size_t i = 0; i < ntracks; ++i) {
    const Track& tr = p.tracks[i];
    const double scale = tr.format == FMT_S16 ? 32768.0 : 1.0;
    for (uint32_t c = 0; c < tr.channels; ++c) {
      const double s = (double) peaks[2 * (tr.peak_base + c)] / scale;
      const double t = (double) peaks[2 * (tr.peak_base + c) + 1] / scale;
      if (sample_peaks) sample_peaks[tr.peak_base + c] = s;
      if (true_peaks) true_peaks[tr.peak_base + c] = t > s ? t : s;
      if (true_peaks_screened) {
        const double u = (double) tps[tr.peak_base + c] / scale;
        true_peaks_screened[tr.peak_base + c] = u > s ? u : s;
      }
    }
  }
  if (slots_out) std::copy(eslot.begin(), eslot.end(), slots_out);   // [sum of frames / s100]
  if (blocks_out) std::copy(zblock.begin(), zblock.end(), blocks_out);
  if (st_out) std::copy(zst.begin(), zst.end(), st_out);
  return 0;
}

extern "C" void emu_plan_sizes(const lgb_track* tracks, size_t ntracks, uint64_t target_tasks,
                               uint64_t* total_blocks, uint64_t* total_st) {
  std::vector<TrackIn> in(ntracks);
  for (size_t i = 0; i < ntracks; ++i)
    in[i] = TrackIn{tracks[i].pcm, tracks[i].frames, tracks[i].channels, tracks[i].samplerate,
                    tracks[i].format, LGB_NO_ALBUM, nullptr};
  Plan p;
  const PlanOptions opt = emu_options(target_tasks);
  build_plan(in.data(), ntracks, 0, opt, p);
  *total_blocks = p.total_blocks;
  *total_st = p.total_st;
}

// Checks the run sweep's view of a plan (lg_run.cu stages rows of the track's
// 2-D tensor: row y = run y, 16-byte pitched): every item's rows start on
// 16-byte boundaries, the stages cover warm-up + run exactly, the warm-up is a
// whole number of stages (so that its boxes come from row y - 1 and end at the
// row's end), complete / partial rows are classified right, and every chunk of
// every track is covered by exactly one lane.  Returns the number of items, or
// -1 - (index of the first bad item).
// lg_common.h: run_grid_ctas, for the planner's test
extern "C" uint32_t emu_run_grid(uint32_t nitems, uint32_t sms, uint32_t warps, int spare) {
  return run_grid_ctas(nitems, sms, warps, spare != 0);
}

extern "C" long long emu_check_run_view(const lgb_track* tracks, size_t ntracks, uint64_t target_tasks,
                                        long long* full_items) {
  std::vector<TrackIn> in(ntracks);
  for (size_t i = 0; i < ntracks; ++i)
    in[i] = TrackIn{tracks[i].pcm, tracks[i].frames, tracks[i].channels, tracks[i].samplerate,
                    tracks[i].format, LGB_NO_ALBUM, nullptr, tracks[i].lead_in};
  Plan p;
  const PlanOptions opt = emu_options(target_tasks);
  build_plan(in.data(), ntracks, 0, opt, p);
  long long nitems = 0;
  *full_items = 0;
  std::vector<std::vector<int>> covered(ntracks);
  for (size_t i = 0; i < ntracks; ++i) covered[i].assign(p.tracks[i].nchunks, 0);
  for (const SweepGroup& g : p.groups) {
    if (!g.run) continue;
    const SweepParams& sp = g.params;
    for (uint32_t ii = 0; ii < g.nitems; ++ii) {
      const RunItem it = p.items[g.first_item + ii];
      const Track& tr = p.tracks[it.track];
      const CoefSet& cs = p.coefs[tr.coef];
      const long long Lr = (long long) cs.run_chunks * cs.L;
      bool ok = sp.Lr == Lr && sp.Wp == cs.run_warm && sp.R == cs.run_chunks &&
                (Lr * tr.fb) % 16 == 0 && ((long long) sp.Wp * tr.fb) % 16 == 0 &&
                sp.Wp % (int) sp.run_stage_frames == 0 && sp.run_stage_frames % kPairFrames == 0 &&
                Lr % kIter == 0 && sp.Wp % kPairFrames == 0 && Lr >= sp.Wp &&
                (long long) sp.run_nstages * sp.run_stage_frames >= sp.Wp + Lr &&
                (long long) (sp.run_nstages - 1) * sp.run_stage_frames < sp.Wp + Lr &&
                (long long) sp.niters * kIter == sp.Wp + Lr && sp.xi_iters * kIter == sp.Wp + cs.xi_frames &&
                it.first_run % 32 == 0 && it.first_run < tr.nruns &&
                tr.nfull == tr.frames / (uint64_t) Lr && tr.nruns == (tr.frames + Lr - 1) / (uint64_t) Lr;
      // rows [0, tail_rows) complete, row tail_rows partial or absent
      for (uint32_t lane = 0; lane < 32 && ok; ++lane) {
        const uint64_t run = it.first_run + lane;
        const bool complete = (run + 1) * (uint64_t) Lr <= tr.frames;
        ok = complete == (lane < it.tail_rows);
        if (run < tr.nruns)
          for (int j = 0; j < cs.run_chunks; ++j) {
            const uint64_t c = run * cs.run_chunks + j;
            if (c < tr.nchunks) ++covered[it.track][c];
          }
      }
      // the launch grid and the CTAs' stretches of the candidate queue (lg_common.h: run_grid_ctas;
      // CTA b owns the items b, b + grid, ...: the fullest CTA's lanes must fit its stretch)
      const uint32_t grid = sp.run_grid;
      ok = ok && grid >= 1 && grid <= g.nitems && grid <= opt.sms &&
           grid == run_grid_ctas(g.nitems, opt.sms, sp.run_warps_per_sm, opt.spare_sms) &&
           (uint64_t) sp.run_cta_cap >= (uint64_t) ((g.nitems + grid - 1) / grid) * 32u * sp.run_lane_stride &&
           sp.run_lane_stride >= 2u * sp.npairs &&
           (cs.tpf == 0 || g.queue_cap == (uint64_t) sp.run_cta_cap * grid);
      if (!ok) return -1 - (long long) (g.first_item + ii);
      ++nitems;
      if (it.tail_rows == 32) ++*full_items;
    }
  }
  for (size_t i = 0; i < ntracks; ++i)
    if (p.tracks[i].nruns)
      for (int c : covered[i]) if (c != 1) return -1000000 - (long long) i;
  return nitems;
}
