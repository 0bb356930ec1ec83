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
  SweepParams params;
};

struct Plan {
  std::vector<CoefSet> coefs;
  std::vector<Track> tracks;
  std::vector<WarpWork> work;
  std::vector<SweepGroup> groups;
  std::vector<Query> queries;        // [ntracks] track queries, then [nalbums]
  std::vector<uint32_t> members;     // track indices referenced by queries
  uint64_t total_recs = 0, total_slots = 0, total_blocks = 0, total_st = 0,
           total_peaks = 0, total_mrec = 0;
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

  std::map<std::tuple<uint32_t, int, uint32_t>, uint32_t> coef_index;
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
    if (t.wclass) for (uint32_t c = 0; c < t.channels; ++c) tr.wclass[c] = t.wclass[c];
    else default_weight_classes(t.channels, tr.wclass);
    const int s100 = (int) ((t.samplerate + 5) / 10);
    const KDesign kd = k_design(t.samplerate);
    const int k = chunks_per_slot_for(s100, warmup_frames(kd),
                                      tail ? want_len / (uint64_t) (opt.tail_div > 1 ? opt.tail_div : 1) : want_len,
                                      opt.force_k);
    const auto key = std::make_tuple(t.samplerate, k, t.format);
    auto it = coef_index.find(key);
    if (it == coef_index.end()) {
      CoefSet cs;
      make_coefset(t.samplerate, k, t.format == FMT_S16 ? 32768.0 : 1.0, cs);
      it = coef_index.emplace(key, (uint32_t) p.coefs.size()).first;
      p.coefs.push_back(cs);
    }
    tr.coef = it->second;
    const CoefSet& cs = p.coefs[tr.coef];
    tr.fb = t.channels * (t.format == FMT_S16 ? 2u : 4u);
    tr.aq = align_quantum(tr.fb);
    tr.niters = (uint32_t) sweep_iters(cs.W, cs.L, (int) tr.aq);
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
    sp.W = cs.W; sp.L = cs.L; sp.niters = (int32_t) t0.niters; sp.aq = (int32_t) t0.aq;
    sp.npairs = (t0.niters + 1u) / 2u;
    sp.channels = t0.channels; sp.fb = t0.fb;
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
