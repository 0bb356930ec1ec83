// lg_plan.h -- host-side planner: turns a list of tracks (interleaved PCM
// already in HBM) into the tables the kernels consume: coefficient sets, track
// descriptors with their record/slot/block offsets, the warp work list of the
// sweep and the query (track / album) member lists.
//
// Host-only, no CUDA calls: lg_batch.cu uploads the result, tests/emu runs it
// on the CPU.
#pragma once

#include <math.h>

#include <map>
#include <tuple>
#include <vector>

#include "lg_common.h"
#include "lg_design.h"
#include "lg_sweep.cuh"

namespace lg {

struct TrackIn {
  const void* pcm;
  uint64_t frames;
  uint32_t channels;
  uint32_t samplerate;
  uint32_t format;          // Format
  uint32_t album;           // album index, or kNoAlbum
  const uint8_t* wclass;    // optional explicit weight classes [channels]
  uint64_t lead_in = 0;     // leading context frames (segment of a longer stream)
  uint32_t flags = 0;       // Track::flags
};

constexpr uint32_t kNoAlbum = 0xffffffffu;

// Sweep launch group: all warps that run one kernel launch = one (format,
// coefficient set, channel count); everything uniform over the launch lives in
// `params` (device pointers are filled in by lg_batch.cu).
struct SweepGroup {
  uint32_t format;
  int32_t tpf;
  uint32_t coef;
  uint32_t channels;
  uint32_t first_warp;   // into Plan::work
  uint32_t nwarps;
  uint64_t mrec_base;    // first word of the group's iteration maxima (tpf != 0)
  bool run = false;      // run sweep (lg_run.cu): work items are Plan::items[first_item, +nitems)
  uint32_t first_item = 0, nitems = 0;
  uint64_t queue_base = 0, queue_cap = 0;   // true-peak candidate queue of a run group (entries)
  SweepParams params;
};

struct Plan {
  std::vector<CoefSet> coefs;
  std::vector<Track> tracks;
  std::vector<WarpWork> work;
  std::vector<SweepGroup> groups;
  std::vector<Query> queries;        // [ntracks] track queries, then [nalbums]
  std::vector<uint32_t> members;     // track indices referenced by queries
  std::vector<RunItem> items;        // work items of the run-sweep groups
  std::vector<cplx> xi_table;        // per coefficient set with run_chunks > 0: lambda^(n_j), j < run_chunks
  uint64_t total_recs = 0, total_slots = 0, total_blocks = 0, total_st = 0,
           total_peaks = 0, total_mrec = 0;
  uint64_t total_queue = 0;          // candidate queue entries over all run groups
  uint64_t total_samples = 0;        // frames * channels over all tracks
  uint32_t nalbums = 0;
};

// Even channel counts run the packed sweep (lg_pair.cu): a lane holds one
// chunk and one channel pair.
inline bool track_is_packed(uint32_t channels, bool allow_packed) {
  return allow_packed && channels >= 2 && (channels & 1u) == 0 && channels <= 64;
}

// Resident sweep warps per SM (launch bounds of the two kernels), for the
// chunk-length choice below.
constexpr uint32_t kScalarWarpsPerSM = 32, kPackedWarpsPerSM = 16;

struct PlanOptions {
  uint64_t target_tasks = 0;   // about 2048 x SM count; 0 = shortest chunks
  int force_k = 0;             // > 0 pins the chunks per 100 ms slot (tuning / tests)
  bool allow_packed = true;
  bool use_tma = true;         // 2-D TMA staging for packed stereo groups
  bool use_run = true;         // stereo tracks go through the run sweep (lg_run.cu)
  uint32_t sms = 148;          // SMs of the device (run sweep: one wave of work items when the batch is small)
  uint32_t run_warps_per_sm = 0;   // resident warps per SM of the run sweep (0 = kRunWarpsPerSM)
  int force_run_chunks = 0;    // > 0 pins the chunks per run (tuning / tests)
  bool spare_sms = true;       // a one-round run sweep leaves the SMs it does not need free (run_grid_ctas)
  // Tail filler (profiles/r01_pair_tuning.txt I/J): the tracks that hold the last
  // `tail_frac` of the batch's lane-frames get chunks `tail_div` times shorter.
  // Their launch group follows the main one on a second stream, and its short
  // work items keep the SMs full while the main group's long ones run out.
  double tail_frac = 0.0;      // 0 = off
  int tail_div = 3;
};

// Chunks per 100 ms slot of one track: the longest chunk (a divisor of the
// slot) not above the wanted length, but never below four warm-ups.
inline int chunks_per_slot_for(int s100, int W, uint64_t want_len, int force_k) {
  int best = 1;
  for (int k = 1; k <= s100; ++k) {
    if (s100 % k) continue;
    const int L = s100 / k;
    if (L < 4 * W) break;
    best = k;
    if (force_k > 0 ? k >= force_k : (uint64_t) L <= want_len) break;
  }
  return best;
}

// ---- run sweep planning ------------------------------------------------------
// A stereo track is cut into RUNS of R chunks (R * L frames); one lane filters
// one run in one go, starting Wp frames early.  Longer runs mean less warm-up
// and fewer frames that need the mode sums (they stop xi_frames into the run),
// but fewer lanes: (k, R) is chosen so that the group's work items fill the
// resident warps in as few equal waves as possible.

inline bool track_is_run(uint32_t channels, bool use_run) { return use_run && channels == 2; }

// Frames after which a start-state error has decayed by 1e-5 in amplitude
// (its energy by 1e-10), rounded up to whole iterations.
inline int xi_horizon(const KDesign& d) {
  const int h = (int) std::ceil(std::log(1e-5) / std::log(d.hp_pole_radius));
  return ((h + kIter - 1) / kIter) * kIter;
}

inline int run_warmup(const KDesign& d) {
  const int w = warmup_frames(d);
  return ((w + kRunStageFrames - 1) / kRunStageFrames) * kRunStageFrames;
}

struct RunChoice { int k = 1, R = 0; };

inline uint64_t run_items_of(uint64_t frames, uint64_t Lr) {
  const uint64_t runs = (frames + Lr - 1) / Lr;
  return (runs + 31) / 32;
}

inline RunChoice choose_run(uint32_t rate, const std::vector<uint64_t>& frames, uint32_t aq,
                            uint64_t slots, int force_k, int force_R) {
  const KDesign d = k_design(rate);
  const int s100 = (int) ((rate + 5) / 10);
  const int Wp = run_warmup(d), H = xi_horizon(d);
  RunChoice best;
  double best_cost = 0.0;
  for (int k = 1; k <= s100; ++k) {
    if (s100 % k) continue;
    const int L = s100 / k;
    if (force_k > 0 ? k != force_k : (L > 2400 && k < s100)) continue;   // chunk records every <= 2400 frames
    if (force_k <= 0 && L < 256 && k > 1) break;
    int rstep = 12 / (int) gcd_u32((uint32_t) L, 12u);
    while (((long long) rstep * L) % aq) rstep += 12 / (int) gcd_u32((uint32_t) L, 12u);
    for (int R = rstep; R <= 4096; R += rstep) {
      if (force_R > 0 && R < force_R) continue;
      const long long Lr = (long long) R * L;
      if (Lr < Wp) continue;
      if (Lr + Wp > 24 * 32000ll) break;               // pair index is a 15-bit field of the queue entries
      uint64_t nitems = 0;
      for (uint64_t f : frames) nitems += run_items_of(f, (uint64_t) Lr);
      const double ops = 9.0 + 2.0 * (H < Lr ? (double) H / (double) Lr : 1.0);
      const double item = (double) (Wp + Lr) * ops + 4000.0;
      const double waves = nitems <= 6 * slots ? (double) ((nitems + slots - 1) / slots)
                                               : (double) nitems / (double) slots + 1.0;
      const double cost = waves * item;
      if (best.R == 0 || cost < best_cost * 0.999 ) { best.k = k; best.R = R; best_cost = cost; }
      if (force_R > 0) break;
      if (nitems <= slots / 2) break;                  // longer runs would leave warps idle
    }
  }
  if (best.R == 0) { best.k = 1; best.R = 12; }
  return best;
}

inline void build_plan(const TrackIn* in, size_t n, uint32_t nalbums, const PlanOptions& opt, Plan& p) {
  p = Plan();
  p.nalbums = nalbums;
  // -- chunk length.  Measured on the B200 (profiles/r01_pair_tuning.txt): the
  // sweep time is flat once a batch fills about two waves of resident warps and
  // rises below that, while longer chunks only save warm-up frames.  So: the
  // longest chunk that still gives every launch group ~1.8 waves.
  uint64_t want_len = 0;
  double total_work = 0.0;             // lane-frames of the batch
  for (size_t i = 0; i < n; ++i)
    total_work += (double) in[i].frames *
                  (track_is_packed(in[i].channels, opt.allow_packed) ? in[i].channels / 2.0 : (double) in[i].channels);
  if (opt.force_k <= 0 && opt.target_tasks) {
    const uint32_t sms = (uint32_t) (opt.target_tasks / 2048u);
    if (sms) {
      double waves_frames = 0.0;       // sum over tracks of frames / (lane capacity of its kernel)
      for (size_t i = 0; i < n; ++i) {
        const bool packed = track_is_packed(in[i].channels, opt.allow_packed);
        const double lanes_per_frame = packed ? in[i].channels / 2.0 : (double) in[i].channels;
        const double cap = (double) sms * (packed ? kPackedWarpsPerSM : kScalarWarpsPerSM) * 32.0;
        waves_frames += (double) in[i].frames * lanes_per_frame / cap;
      }
      want_len = (uint64_t) (waves_frames / 1.8);
    } else {
      uint64_t samples = 0;
      for (size_t i = 0; i < n; ++i) samples += in[i].frames * (uint64_t) in[i].channels;
      want_len = samples / opt.target_tasks;
    }
  } else if (opt.force_k <= 0) {
    want_len = 0;                       // shortest chunks
  }

  // -- run sweep: one (k, R) per (rate, format) of the stereo tracks
  std::map<std::pair<uint32_t, uint32_t>, RunChoice> run_choice;
  {
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint64_t>> by_rate;
    for (size_t i = 0; i < n; ++i)
      if (track_is_run(in[i].channels, opt.use_run))
        by_rate[std::make_pair(in[i].samplerate, in[i].format)].push_back(in[i].frames);
    const uint64_t slots = (uint64_t) opt.sms * (opt.run_warps_per_sm ? opt.run_warps_per_sm : kRunWarpsPerSM);
    for (const auto& kv : by_rate)
      run_choice[kv.first] = choose_run(kv.first.first, kv.second,
                                        align_quantum(2u * (kv.first.second == FMT_S16 ? 2u : 4u)), slots,
                                        opt.force_k, opt.force_run_chunks);
  }
  std::map<std::tuple<uint32_t, int, uint32_t, int>, uint32_t> coef_index;
  p.tracks.resize(n);
  double work_before = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const TrackIn& t = in[i];
    Track& tr = p.tracks[i];
    // a tail track: starts inside the last tail_frac of the batch (never the first track)
    const bool tail = opt.tail_frac > 0.0 && opt.force_k <= 0 && want_len && i > 0 &&
                      work_before >= (1.0 - opt.tail_frac) * total_work;
    work_before += (double) t.frames *
                   (track_is_packed(t.channels, opt.allow_packed) ? t.channels / 2.0 : (double) t.channels);
    tr = Track();
    tr.pcm = t.pcm;
    tr.frames = t.frames;
    tr.channels = t.channels;
    tr.format = t.format;
    tr.album = t.album;
    tr.lead_in = t.lead_in;
    tr.flags = t.flags;
    if (t.wclass) for (uint32_t c = 0; c < t.channels; ++c) tr.wclass[c] = t.wclass[c];
    else default_weight_classes(t.channels, tr.wclass);
    const int s100 = (int) ((t.samplerate + 5) / 10);
    const KDesign kd = k_design(t.samplerate);
    const bool is_run = track_is_run(t.channels, opt.use_run);
    int k, R = 0;
    if (is_run) {
      const RunChoice rc = run_choice[std::make_pair(t.samplerate, t.format)];
      k = rc.k; R = rc.R;
    } else {
      k = chunks_per_slot_for(s100, warmup_frames(kd),
                              tail ? want_len / (uint64_t) (opt.tail_div > 1 ? opt.tail_div : 1) : want_len,
                              opt.force_k);
    }
    const auto key = std::make_tuple(t.samplerate, k, t.format, R);
    auto it = coef_index.find(key);
    if (it == coef_index.end()) {
      CoefSet cs;
      const double unit = (t.format == FMT_S16 ? 32768.0 : 1.0) *
                          (is_run && t.format == FMT_S16 ? (double) kRunS16Scale : 1.0);
      make_coefset(t.samplerate, k, unit, cs);
      if (is_run) make_run_coefs(kd, R, run_warmup(kd), xi_horizon(kd), cs, p.xi_table);
      it = coef_index.emplace(key, (uint32_t) p.coefs.size()).first;
      p.coefs.push_back(cs);
    }
    tr.coef = it->second;
    const CoefSet& cs = p.coefs[tr.coef];
    tr.fb = t.channels * (t.format == FMT_S16 ? 2u : 4u);
    tr.aq = align_quantum(tr.fb);
    tr.niters = (uint32_t) sweep_iters(cs.W, cs.L, (int) tr.aq);
    if (is_run) {
      const uint64_t Lr = (uint64_t) cs.run_chunks * (uint64_t) cs.L;
      tr.niters = (uint32_t) ((cs.run_warm + (int64_t) Lr) / kIter);
      tr.nruns = (uint32_t) ((t.frames + Lr - 1) / Lr);
      tr.nfull = (uint32_t) (t.frames / Lr);
    }
    tr.nslots = (uint32_t) (t.frames / (uint64_t) s100);
    tr.nchunks = (uint32_t) ((t.frames + cs.L - 1) / (uint64_t) cs.L);
    tr.nblocks = tr.nslots >= 4 ? tr.nslots - 3 : 0;
    tr.nst = tr.nslots >= 30 ? (tr.nslots - 30) / 10 + 1 : 0;
    tr.rec_base = p.total_recs;
    tr.slot_base = p.total_slots;
    tr.block_base = p.total_blocks;
    tr.st_base = p.total_st;
    tr.peak_base = p.total_peaks;
    p.total_recs += (uint64_t) tr.nchunks * tr.channels;
    p.total_slots += tr.nslots;
    p.total_blocks += tr.nblocks;
    p.total_st += tr.nst;
    p.total_peaks += tr.channels;
    p.total_samples += t.frames * t.channels;
  }
  // -- sweep work list, one launch group per (format, coefficient set, channels)
  std::map<std::tuple<uint32_t, uint32_t, uint32_t>, std::vector<uint32_t>> by_group;
  for (size_t i = 0; i < n; ++i)
    by_group[std::make_tuple(p.tracks[i].format, p.tracks[i].coef, p.tracks[i].channels)]
        .push_back((uint32_t) i);
  for (const auto& kv : by_group) {
    const Track& t0 = p.tracks[kv.second[0]];
    const CoefSet& cs = p.coefs[t0.coef];
    SweepGroup g;
    g.format = t0.format; g.tpf = cs.tpf; g.coef = t0.coef; g.channels = t0.channels;
    g.first_warp = (uint32_t) p.work.size();
    SweepParams& sp = g.params;
    sp = SweepParams();
    fill_kcoef(cs, sp);
    sp.tp_bound = cs.tpf == 4 ? 1.8645f : 2.3072f;   // > ||taps||_1 (1.8642 / 2.3068)
    sp.tp_taps = cs.tpf == 4 ? 12u : (cs.tpf == 2 ? 24u : 0u);
    sp.W = cs.W; sp.L = cs.L; sp.niters = (int32_t) t0.niters; sp.aq = (int32_t) t0.aq;
    sp.npairs = (t0.niters + 1u) / 2u;
    sp.channels = t0.channels; sp.fb = t0.fb;
    if (cs.run_chunks) {
      // ---- run sweep group: work items of 32 consecutive runs
      g.run = true;
      sp.R = cs.run_chunks; sp.Lr = cs.run_chunks * cs.L; sp.Wp = cs.run_warm;
      sp.xi_iters = (cs.run_warm + cs.xi_frames) / kIter;
      sp.niters = (cs.run_warm + sp.Lr) / kIter;
      sp.npairs = ((uint32_t) sp.niters + 1u) / 2u;
      sp.run_stage_frames = run_stage_frames(t0.format);
      sp.run_nstages = ((uint32_t) (sp.Wp + sp.Lr) + sp.run_stage_frames - 1u) / sp.run_stage_frames;
      sp.run_warps_per_sm = opt.run_warps_per_sm ? opt.run_warps_per_sm : kRunWarpsPerSM;
      sp.peak_scale = t0.format == FMT_S16 ? 1.0f / (float) kRunS16Scale : 1.0f;
      sp.packed = 1u; sp.lpc = 1u; sp.cpw = 32u;
      g.first_item = (uint32_t) p.items.size();
      for (uint32_t i : kv.second) {
        const Track& tr = p.tracks[i];
        for (uint32_t r0 = 0; r0 < tr.nruns; r0 += 32u) {
          const uint32_t full = tr.nfull > r0 ? tr.nfull - r0 : 0u;
          p.items.push_back(RunItem{i, r0, full < 32u ? full : 32u, 0u});
        }
      }
      g.nitems = (uint32_t) p.items.size() - g.first_item;
      sp.nitems = g.nitems;
      g.nwarps = g.nitems;
      sp.nwarps = g.nitems;
      g.queue_base = p.total_queue;
      sp.run_lane_stride = (sp.npairs * 2u + 3u) & ~3u;        // 16-byte multiples
      {
        // one dense stretch per sweep CTA (lg_run.cu launches run_grid_ctas() of them, CTA b
        // owning the items b, b + grid, ...): room for every pair of every lane of its items
        const uint32_t grid = run_grid_ctas(g.nitems, opt.sms, sp.run_warps_per_sm, opt.spare_sms);
        sp.run_grid = grid;
        const uint64_t per_cta = grid ? (uint64_t) ((g.nitems + grid - 1) / grid) * 32u * sp.run_lane_stride : 0u;
        sp.run_cta_cap = (uint32_t) per_cta;
        g.queue_cap = cs.tpf ? per_cta * grid : 0u;
      }
      p.total_queue += g.queue_cap;
      if (g.nitems) p.groups.push_back(g);
      continue;
    }
    const bool packed = track_is_packed(t0.channels, opt.allow_packed);
    sp.packed = packed ? 1u : 0u;
    sp.lpc = packed ? t0.channels / 2u : (t0.channels < 32u ? t0.channels : 32u);
    sp.cpw = 32u / sp.lpc;
    const uint32_t pps = packed ? (uint32_t) pair_pps(t0.format) : (uint32_t) kPairsPerStage;
    const uint32_t ring = packed ? (uint32_t) kPairRing : (uint32_t) kRing;
    sp.stage_row_bytes = pps * kPairFrames * t0.fb;
    sp.units = sp.stage_row_bytes >> 4;
    sp.row_stride = (sp.units | 1u) << 4;
    sp.stage_bytes = sp.cpw * sp.row_stride;
    sp.ring_bytes = ring * sp.stage_bytes;
    sp.kcopies = (sp.units + sp.lpc - 1) / sp.lpc;
    sp.warp_smem = sp.ring_bytes;
    // 2-D TMA staging: stereo, one lane per row, rows of a class 16-byte pitched
    const bool stereo = packed && t0.channels == 2;
    const int tma_m = (opt.use_tma && stereo) ? tma_interleave(cs.L, (int) t0.aq) : 0;
    sp.tma_m = 0;
    if (tma_m && (sp.units | 1u) == sp.units + 1u &&
        (long long) tma_m * cs.L >= (long long) ((sp.npairs + pps - 1) / pps) * pps * kPairFrames + 8) {
      sp.tma_m = (uint32_t) tma_m;
      sp.tma_class_bytes = (32u / (uint32_t) tma_m) * sp.row_stride;
      sp.tma_shift = 0;
      for (int r = 0; r < tma_m; ++r)
        if (tma_class(cs.L, cs.W, (int) t0.aq, tma_m, r).shift) sp.tma_shift |= 1u << r;
      // ring, then one mbarrier per stage; TMA destinations are 128-byte aligned
      sp.warp_smem = (sp.ring_bytes + ring * 8u + 127u) & ~127u;
    }
    const long long stage_frames = (long long) ((sp.npairs + pps - 1) / pps) * pps * kPairFrames;
    for (uint32_t i : kv.second) {
      const Track& tr = p.tracks[i];
      for (uint32_t c = 0; c < tr.nchunks; c += sp.cpw) {
        const uint32_t last = c + sp.cpw - 1;
        const LaneGeom g0 = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, c);
        const LaneGeom g1 = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq, last);
        int lmin = cs.L;
        if (last >= tr.nchunks - 1) {
          const LaneGeom gl = lane_geometry((long long) tr.frames, cs.L, cs.W, (int) tr.aq,
                                            tr.nchunks - 1);
          lmin = gl.l_valid;
        }
        const bool interior = g0.a >= 0 && g1.a + stage_frames <= (long long) tr.frames;
        uint16_t mode = interior ? 1 : 0;
        if (interior && sp.tma_m) {
          // every class's rows of this warp must exist in the class's tensor
          bool ok = true;
          const long long m = sp.tma_m, mL = m * cs.L;
          for (int r = 0; r < (int) m && ok; ++r) {
            const TmaClass tc = tma_class(cs.L, cs.W, (int) tr.aq, (int) m, r);
            const long long nrows = (long long) tr.frames > tc.base_frame
                                        ? ((long long) tr.frames - tc.base_frame) / mL : 0;
            const long long first = ((long long) c + r) / m - tc.shift;
            ok = first >= 0 && first + 32 / m <= nrows;
          }
          if (ok) mode = 2;
        }
        for (uint32_t cb = 0; cb < (packed ? 1u : tr.channels); cb += 32)
          p.work.push_back(WarpWork{i, c, lmin, mode, (uint16_t) cb});
      }
    }
    g.nwarps = (uint32_t) p.work.size() - g.first_warp;
    sp.nwarps = g.nwarps;
    g.mrec_base = p.total_mrec;
    if (cs.tpf) p.total_mrec += (uint64_t) g.nwarps * sp.npairs * 32u;
    if (g.nwarps) p.groups.push_back(g);
  }
  // -- queries: one per track, then one per album
  p.queries.resize(n + nalbums);
  for (size_t i = 0; i < n; ++i) {
    p.queries[i] = Query{(uint32_t) p.members.size(), 1};
    p.members.push_back((uint32_t) i);
  }
  for (uint32_t a = 0; a < nalbums; ++a) {
    Query q{(uint32_t) p.members.size(), 0};
    for (size_t i = 0; i < n; ++i)
      if (in[i].album == a) { p.members.push_back((uint32_t) i); ++q.count; }
    p.queries[n + a] = q;
  }
}

}  // namespace lg
