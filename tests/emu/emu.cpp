// emu.cpp -- TEST-ONLY host compile of the per-thread device math.
//
// The sweep / fix-up / block functions in loudgain_b200/csrc/*.cuh are written
// with explicit fmaf so that a plain C++ build reproduces the GPU arithmetic
// bit for bit.  This file runs them sequentially over a plan so that CPU-only
// tests can check the kernel logic against the oracle before any GPU time is
// spent.  It is NOT part of the product: libebur128.so does not contain it and
// has no CPU path; nothing under loudgain_b200/ loads this library.
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

template <int FMT, int TPF>
static void run_group(const Plan& p, const SweepGroup& g, std::vector<ChunkRec>& recs,
                      std::vector<float>& peaks) {
  const SweepParams& k = g.params;
  for (uint32_t w = 0; w < g.nwarps; ++w) {
    const WarpWork ww = p.work[g.first_warp + w];
    const Track& tr = p.tracks[ww.track];
    const CoefSet& cs = p.coefs[tr.coef];
    const uint32_t C = tr.channels, lpc = k.lpc, cpw = k.cpw;
    const LaneGeom glast = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq,
                                         ww.first_chunk + cpw - 1);
    const long long tp_safe = (long long) tr.frames - glast.a;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      const uint32_t slot = lane / lpc, chl = lane - slot * lpc, ch = ww.ch_base + chl;
      const uint32_t chunk = ww.first_chunk + slot;
      if (!(slot < cpw && ch < C && chunk < tr.nchunks)) continue;
      const LaneGeom geo = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, chunk);
      LaneCtx<TPF> c;
      lane_init(c, cs.W, cs.L, geo);
      for (uint32_t it = 0; it < tr.niters; ++it) {
        const int f0 = (int) it * kIter;
        float x[kIter];
        host_load_iter<FMT>(tr.pcm, (long long) tr.frames, (int) C, geo.a, f0, (int) ch, x);
        const int kind = iter_kind(f0, cs.W, (int) tr.aq, cs.L, ww.lmin_valid);
        if (kind == ITER_WARM) {
          iter_warm<TPF>(c, k, x);
          continue;
        }
        if (kind == ITER_FAST) {
          (void) iter_fast_energy<TPF>(c, k, x, f0);
          iter_peaks_all<TPF>(c, x);        // the device defers this to its candidate queue
        } else {
          iter_masked_energy<TPF>(c, k, x, f0);
          if ((long long) f0 + kIter <= tp_safe) iter_peaks_all<TPF>(c, x);
          else iter_peaks_masked<TPF>(c, x, f0);
        }
        hist_advance(c, x);
      }
      ChunkRec& r = recs[tr.rec_base + (uint64_t) chunk * C + ch];
      r.e0 = c.e0; r.yr = c.yr; r.yi = c.yi;
      r.pd = c.pd; r.pw = c.pw; r.qd = c.qd; r.qw = c.qw;
      float& sp = peaks[2 * (tr.peak_base + ch)];
      float& tp = peaks[2 * (tr.peak_base + ch) + 1];
      sp = std::max(sp, c.sp);
      tp = std::max(tp, c.tp);
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
                           double* blocks_out, double* st_out, int32_t* chunk_len_out) {
  std::vector<TrackIn> in(ntracks);
  for (size_t i = 0; i < ntracks; ++i)
    in[i] = TrackIn{tracks[i].pcm, tracks[i].frames, tracks[i].channels, tracks[i].samplerate,
                    tracks[i].format, tracks[i].album, tracks[i].weight_class};
  Plan p;
  build_plan(in.data(), ntracks, nalbums, target_tasks, p);
  std::vector<ChunkRec> recs(p.total_recs);
  std::vector<float> peaks(2 * p.total_peaks, 0.0f);
  for (const SweepGroup& g : p.groups) {
    if (g.format == FMT_S16) {
      if (g.tpf == 4) run_group<FMT_S16, 4>(p, g, recs, peaks);
      else if (g.tpf == 2) run_group<FMT_S16, 2>(p, g, recs, peaks);
      else run_group<FMT_S16, 0>(p, g, recs, peaks);
    } else {
      if (g.tpf == 4) run_group<FMT_F32, 4>(p, g, recs, peaks);
      else if (g.tpf == 2) run_group<FMT_F32, 2>(p, g, recs, peaks);
      else run_group<FMT_F32, 0>(p, g, recs, peaks);
    }
  }
  std::vector<double> echunk(p.total_recs, 0.0), eslot(p.total_slots), zblock(p.total_blocks),
      zst(p.total_st);
  for (size_t ti = 0; ti < ntracks; ++ti) {
    const Track& tr = p.tracks[ti];
    const CoefSet& cs = p.coefs[tr.coef];
    if (chunk_len_out) chunk_len_out[ti] = cs.L;
    for (uint64_t chunk = 0; chunk < (uint64_t) tr.nslots * cs.k; ++chunk) {
      const LaneGeom geo = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, (long long) chunk);
      for (uint32_t ch = 0; ch < tr.channels; ++ch)
        echunk[tr.rec_base + chunk * tr.channels + ch] =
            chunk_true_energy(cs, recs.data() + tr.rec_base + ch, tr.channels, (long long) chunk,
                              geo.o, log2u(tr.aq));
    }
    for (uint32_t s = 0; s < tr.nslots; ++s)
      eslot[tr.slot_base + s] = slot_energy(tr, cs, echunk.data(), s);
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
    }
  }
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
  build_plan(in.data(), ntracks, 0, target_tasks, p);
  *total_blocks = p.total_blocks;
  *total_st = p.total_st;
}
