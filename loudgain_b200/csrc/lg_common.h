// lg_common.h -- plain-old-data shared by the host planner and the sm_100a
// kernels of the loudness sweep (the ebur128_add_frames_* path that
// /root/reference/src/scan.c:448 drives, plus the query-time reductions of
// scan.c:294-307,383-391).
#pragma once

#include <stddef.h>
#include <stdint.h>

// Per-thread math is written once: device code under nvcc, plain inline host
// code when a C++ compiler builds the test-only emulation (tests/emu).
#if defined(__CUDACC__)
#define LG_HD __device__ __forceinline__
#define LG_BOTH __host__ __device__ __forceinline__   // also used by the host planner
#else
#define LG_HD inline
#define LG_BOTH inline
#endif

namespace lg {

// Sample formats the sweep reads straight from HBM.
enum Format : uint32_t { FMT_S16 = 0, FMT_F32 = 1 };

// Frames the sweep advances per unrolled iteration (= tap count of a 4x
// true-peak phase, so that window slides by exactly one phase length).
constexpr int kIter = 12;
// The sweep computes in pairs of iterations (24 frames: a whole number of
// 16-byte units for any even-sized frame) and stages kPairsPerStage pairs per
// row and stage from HBM.
constexpr int kPairFrames = 2 * kIter;
#ifndef LG_PAIRS_PER_STAGE
#define LG_PAIRS_PER_STAGE 1
#endif
#ifndef LG_RING
#define LG_RING 3
#endif
constexpr int kPairsPerStage = LG_PAIRS_PER_STAGE;
constexpr int kItersPerStage = 2 * kPairsPerStage;
constexpr int kStageFrames = kIter * kItersPerStage;
constexpr int kRing = LG_RING;      // staging ring depth (stages in flight + 1)
// The packed sweep (lg_pair.cu: even channel counts, one lane = one chunk and
// one channel PAIR) stages kPairPPS iteration pairs per row and stage through
// a ring of kPairRing stages.
#ifndef LG_PAIR_PPS
#define LG_PAIR_PPS 2
#endif
#ifndef LG_PAIR_RING
#define LG_PAIR_RING 2
#endif
constexpr int kPairPPS = LG_PAIR_PPS;
constexpr int kPairRing = LG_PAIR_RING;
// Float frames are twice as large: half the pairs per stage keeps the stage (and
// with it the resident warps per SM) the same as for 16-bit input.
LG_BOTH constexpr int pair_pps(uint32_t format) {
  return format == FMT_S16 ? kPairPPS : (kPairPPS > 1 ? kPairPPS / 2 : 1);
}
// Run sweep (lg_run.cu): frames per staged piece of a row, ring depth, resident
// one-warp CTAs per SM, and the unit 16-bit samples are converted in: a frame
// word's halves become (int) (w << 16) and (int) (w & 0xffff0000), i.e. 65536
// times the sample (exact in FP32; one shift / mask per sample instead of a
// sign-extending byte permute).
#ifndef LG_RUN_STAGE_FRAMES
#define LG_RUN_STAGE_FRAMES 48
#endif
#ifndef LG_RUN_RING
#define LG_RUN_RING 2
#endif
#ifndef LG_RUN_WARPS_PER_SM
#define LG_RUN_WARPS_PER_SM 16
#endif
constexpr uint32_t kRunStageFrames = LG_RUN_STAGE_FRAMES;   // 16-bit input; a multiple of 2 * kPairFrames
// float frames are twice as large: half the frames per stage, the same bytes
LG_BOTH constexpr uint32_t run_stage_frames(uint32_t format) {
  return (format == FMT_S16 || kRunStageFrames < 2u * kPairFrames) ? kRunStageFrames : kRunStageFrames / 2u;
}
constexpr int kRunRing = LG_RUN_RING;
constexpr uint32_t kRunWarpsPerSM = LG_RUN_WARPS_PER_SM;
constexpr float kRunS16Scale = 65536.0f;
// Longest per-phase tap count (49-tap prototype / factor 2, zero taps dropped).
constexpr int kMaxTaps = 24;
// A lane starts on a 16-byte boundary of the track, i.e. on a multiple of
// 16 / gcd(16, frame bytes) frames: at most 8 (mono / odd-channel S16).
constexpr int kMaxAlign = 8;

constexpr int kMaxChannels = 64;

struct cplx { double re, im; };

// One K-weighting coefficient set: a (sample rate, chunk length, warm-up,
// input scale) combination.  Floats feed the sweep, doubles the fix-up.
struct CoefSet {
  // --- sweep, FP32.  High-pass in "leaky double integrator" form
  //   q = x - x1 ; t = q - e2*w2 ; d = c*d1 + t ; w = w1 + d ; yh = d
  // (algebraically w[n] = q[n] - a1 w[n-1] - a2 w[n-2], yh = w - w1, q = x - x1)
  // followed by the shelf  v = yh - p1 v1 - p2 v2 ; y = v + q1 v1 + q2 v2
  // (computed as y = yh + r1 v1 + r2 v2, r = q - p).
  float c, e2, p1, p2, q1, q2;
  float r1, r2;     // q1 - p1, q2 - p2: the shelf's output taps on its OLD state
  // lambda = the high-pass pole (Im > 0).  The response of the K-weighted
  // output to a high-pass start state is, after the warm-up, Re(A lambda^f);
  // the sweep accumulates Xi = sum y[f] lambda^f with these constants.
  float lam_re[kIter], lam_im[kIter];   // lambda^i, i = 0..kIter-1
  float rot_re, rot_im;                 // lambda^-kIter
  int32_t s100;     // frames per 100 ms = (rate+5)/10
  int32_t k;        // chunks per 100 ms slot (k divides s100)
  int32_t L;        // frames per chunk = s100 / k
  int32_t W;        // warm-up frames run before each chunk (multiple of kIter)
  int32_t tpf;      // true-peak oversampling factor: 4, 2 or 0 (none)
  int32_t horner;   // number of previous chunks the state carry looks back
  // --- fix-up, FP64
  double ML[4];                   // M^L, M = one-frame transition of (d1, w2)
  double MinvWo[kMaxAlign][4];    // M^-(W+o), o = lane alignment offset
  cplx Ad, Aw;                    // A = tau_d * Ad + tau_w * Aw for start state tau
  cplx xi_scale[4];               // lambda^((niters-1)*kIter) for aq = 1, 2, 4, 8
  double S1o[kMaxAlign];          // sum over the chunk of |lambda|^(2f), per offset o
  cplx S2o[kMaxAlign];            // sum over the chunk of lambda^(2f), per offset o
  double gain;      // (shelf b0 / input full scale)^2: raw energy -> K-weighted
  // --- run sweep (lg_run.cu; run_chunks > 0): a lane filters run_chunks
  // consecutive chunks in one go, the mode sums restart at every chunk's first
  // frame and stop xi_frames after the run's first frame
  int32_t run_chunks;     // R: chunks per run (0: one chunk per lane, kernels of lg_kernels.cu / lg_pair.cu)
  int32_t run_warm;       // Wp: warm-up frames before a run (multiple of the stage length)
  int32_t xi_frames;      // mode sums are accumulated for run-local frames < xi_frames (multiple of kIter)
  uint32_t xi_off;        // first of run_chunks entries in the plan's xi table (lg_plan.h)
  double S1run;           // sum over a chunk of |lambda|^(2f), f from the chunk's first frame
  cplx S2run;             // sum over a chunk of lambda^(2f)
};

// Sweep iterations per chunk for a lane alignment quantum aq.
LG_BOTH int sweep_iters(int W, int L, int aq) { return (W + aq - 1 + L + kIter - 1) / kIter; }

// One audio track resident in HBM as interleaved PCM [frames][channels].
struct Track {
  const void* pcm;     // device pointer to frame 0 (16-byte aligned)
  uint64_t frames;
  uint32_t channels;
  uint32_t format;     // Format
  uint32_t coef;       // index into the CoefSet table
  uint32_t fb;         // bytes per frame
  uint32_t aq;         // lane alignment quantum in frames (16 / gcd(16, fb))
  uint32_t niters;     // sweep iterations per chunk (uniform over the track)
  uint32_t nslots;     // complete 100 ms slots
  uint32_t nchunks;    // chunks incl. the tail after the last complete slot
  uint32_t nblocks;    // 400 ms gating blocks
  uint32_t nst;        // 3 s short-term blocks
  uint32_t album;      // album (query group) index
  uint32_t flags;      // bit 0: EBUR128_MODE_HISTOGRAM semantics (block energies become 0.1 LU bin centres)
  uint64_t rec_base;   // first chunk record (index = rec_base + chunk*channels + ch)
  uint64_t slot_base;  // first slot energy
  uint64_t block_base; // first gating block
  uint64_t st_base;    // first short-term block
  uint64_t peak_base;  // first per-channel peak cell
  uint64_t lead_in;    // leading context frames: no true-peak output is taken inside them
  uint32_t nruns;      // run sweep: runs of run_chunks chunks (the last one may be partial)
  uint32_t nfull;      // run sweep: runs that lie completely inside the track
  // BS.1770 channel weight class: 0 unused, 1 -> 1.0, 2 -> 1.41, 3 -> 2.0
  uint8_t wclass[kMaxChannels];
};

// What the sweep leaves behind for one (chunk, channel).
struct ChunkRec {
  double e0;        // sum y^2 over the chunk, filter started from zero state
  float yr, yi;     // sum y[f] lambda^f, scaled by lambda^-(start of last iteration)
  float pd, pw;     // high-pass state (d1, w2) at the chunk's first frame
  float qd, qw;     // ... and after its last frame
};
static_assert(sizeof(ChunkRec) == 32, "ChunkRec is written as two 16-byte stores");

// Where one track's block energies live (device pointers).
struct BlockList {
  const double* z;    // 400 ms gating block energies
  const double* st;   // 3 s short-term block energies
  uint32_t nz, nst;
};

// A query: integrated loudness + range over the union of member block lists.
struct Query {
  uint32_t first;   // index of the first member in the member list
  uint32_t count;   // number of members
};

struct QueryResult {
  double loudness;  // LUFS, -inf if nothing passed the gates
  double range;     // LU
  double rel_thr;   // relative gate energy (diagnostic)
  double sum1;      // sum of abs-gated block energies
  double sum2;      // sum of rel-gated block energies
  uint64_t n1;      // abs-gated block count
  uint64_t n2;      // rel-gated block count
  uint64_t nst;     // abs-gated short-term block count
};

// ---- album queries over tracks that live on several GPUs (lg_kernels.cu:
// xchg_*_kernel, lg_batch.cu: lgb_exchange_*) --------------------------------
// Every rank keeps one exchange region in its own HBM, mapped into every peer
// (CUDA IPC, NVLink peer access); the layout is the same on every rank:
//   flags  [2 phases][kMaxWorld] uint64   step number rank r has completed phase p of
//   hdr    [2 parities][world][nalbums]   XchgHdr: rank r's share of album a
//   st     [2 parities][world][st_cap]    rank r's short-term energies, album after album
// A rank WRITES its own row into every peer's region (and its own) and READS
// only its local region.  Steps alternate between the two parities, so step
// k + 1 never overwrites what a slower peer still reads for step k.
constexpr uint32_t kMaxWorld = 16;

struct XchgHdr {
  double s1; uint64_t n1;     // gating blocks above the absolute gate: energy sum, count
  double sst; uint64_t nst;   // short-term blocks above the absolute gate
  double s2; uint64_t n2;     // gating blocks above the relative gate (second phase)
  uint32_t st_off, st_cnt;    // where the rank's short-term energies of the album are in its stretch
  uint64_t pad_;
};
static_assert(sizeof(XchgHdr) == 64, "one header is four 16-byte stores");

struct XchgParams {
  uint32_t world, rank, nalbums, first_query;   // album a is query first_query + a of the batch
  uint64_t st_cap;                              // doubles per rank and parity
  unsigned char* peer[kMaxWorld];               // region of rank p as mapped here (peer[rank] = local)
  const uint32_t* st_off;                       // [nalbums + 1] local prefix of the albums' short-term counts
  unsigned long long* ctl;                      // local: [0] step, [1..3] CTAs done per phase, [4] time-outs
};

LG_BOTH size_t xchg_flags_bytes() { return 2u * kMaxWorld * sizeof(uint64_t); }
LG_BOTH size_t xchg_hdr_off(uint32_t world, uint32_t nalbums, uint32_t parity, uint32_t r, uint32_t a) {
  return xchg_flags_bytes() + (((size_t) parity * world + r) * nalbums + a) * sizeof(XchgHdr);
}
LG_BOTH size_t xchg_st_off(uint32_t world, uint32_t nalbums, uint64_t st_cap, uint32_t parity, uint32_t r) {
  const size_t hdr_end = xchg_flags_bytes() + (size_t) 2 * world * nalbums * sizeof(XchgHdr);
  return ((hdr_end + 255u) & ~(size_t) 255u) + ((size_t) parity * world + r) * st_cap * sizeof(double);
}
LG_BOTH size_t xchg_region_bytes(uint32_t world, uint32_t nalbums, uint64_t st_cap) {
  return xchg_st_off(world, nalbums, st_cap, 2, 0);
}

// Work descriptor: one warp of the sweep = up to 32/min(C,32) consecutive
// chunks of one track, for channels [ch_base, ch_base + 32).
struct WarpWork {
  uint32_t track;
  uint32_t first_chunk;
  int32_t lmin_valid;   // shortest valid chunk length among the warp's chunks
  uint16_t interior;    // 1: every byte the warp stages lies inside the track
  uint16_t ch_base;     // first channel this warp handles (0 unless C > 32)
};

// Work item of the run sweep (lg_run.cu): one warp = 32 consecutive runs of one
// stereo track, lane l = run first_run + l.
struct RunItem {
  uint32_t track;
  uint32_t first_run;
  uint32_t tail_rows;   // 32: every row is a complete run of the track's 2-D tensor view;
                        // < 32: the track's last item -- rows [0, tail_rows) complete, row
                        // tail_rows partial (or absent), the rest beyond the track
  uint32_t pad_;
};

LG_BOTH uint32_t chunks_per_warp(uint32_t channels) { return 32u / (channels < 32u ? channels : 32u); }

// Everything that is uniform over one sweep launch (one format, coefficient
// set and channel count).  Passed by value as the kernel parameter, so the
// filter constants reach the FMAs straight from the constant bank.
struct SweepParams {
  float c, ne2, np1, np2, r1, r2;
  float lam_re[kIter], lam_im[kIter];
  float rot_re, rot_im;
  float tp_bound;          // ||taps||_1 bound used for true-peak screening
  uint32_t tp_taps;        // taps per interpolator phase (12 at 4x, 24 at 2x, 0: no interpolator)
  int32_t W, L, niters, aq;
  uint32_t npairs;         // (niters + 1) / 2: iteration pairs per lane
  uint32_t channels, fb;
  uint32_t packed;         // 1: packed sweep, a lane holds channels (2j, 2j+1) of its chunk
  uint32_t lpc;            // lanes per chunk = min(channels, 32); channels / 2 when packed
  uint32_t cpw;            // chunks per warp = 32 / lpc
  // staging: bytes one row advances per stage, 16-byte units of that piece,
  // padded row stride in shared memory (odd number of units), bytes per stage
  // buffer and per ring, copies per lane and stage (ceil(units / lpc))
  uint32_t stage_row_bytes, units, row_stride, stage_bytes, ring_bytes, kcopies;
  uint32_t warp_smem, nwarps;
  const Track* tracks;
  const WarpWork* work;
  ChunkRec* recs;
  uint32_t* peaks;
  // Iteration maxima for the true-peak pass: one word per (warp, pair, lane),
  // index (warp * npairs + pair) * 32 + lane, two 16-bit codes each: the
  // lane's two iterations of the pair (lg_sweep.cuh: peak_code), or, when
  // packed, the pair's maximum for each of the lane's two channels
  // (lg_sweep.cuh: pair_code).  Unused when the rate has no interpolator.
  uint32_t* mrec;
  uint32_t* tp_ticket;     // counter of the group's true-peak pass (zeroed per run): work
                           // items drawn (scalar pass) / candidates queued (packed pass)
  // 2-D TMA staging (packed stereo sweep, WarpWork::interior == 2): m tensor
  // maps per track (one per chunk class j mod m), kTmaMaxM slots of 128 bytes each
  const void* tmaps;
  uint32_t tma_m;
  uint32_t tma_shift;      // bit r: class r's chunks live one row earlier (tma_class)
  uint32_t tma_class_bytes;  // (32 / tma_m) * row_stride: one class's rows in a stage buffer
  uint64_t* tp_queue;        // packed pass: candidate queue, 2 entries per mrec word at most
  uint32_t ctas_per_sm;      // packed sweep: cap on resident CTAs per SM (0 = the kernel's own limit)
  // ---- run sweep (lg_run.cu)
  const RunItem* items;
  uint32_t nitems;
  int32_t R, Lr, Wp;         // chunks per run, frames per run (R * L), warm-up frames
  int32_t xi_iters;          // iterations (from run-local frame -Wp) that accumulate the mode sums
  uint32_t run_stage_frames; // frames per staged piece of a row
  uint32_t run_nstages;      // stages per item = ceil((Wp + Lr) / stage frames)
  uint32_t run_warps_per_sm; // resident one-warp CTAs per SM
  uint32_t* run_queue;       // true-peak candidates: 64-bit entries, one dense stretch per sweep CTA
  void* tp_dense;            // (unused)
  uint32_t* run_counts;      // [sweep CTAs] entries a CTA queued
  uint32_t run_lane_stride;  // most entries one lane can queue = npairs * 2 channels
  uint32_t run_cta_cap;      // entries reserved per sweep CTA (its items' lanes x run_lane_stride)
  float peak_scale;          // raw sample unit / the sweep's internal unit (16-bit input is scaled by 65536)
  uint32_t run_grid;         // persistent CTAs the sweep is launched with (run_grid_ctas)
  uint32_t npeak_words;      // words of the batch's peak-cell area (two per channel), set per launch
};

// CTAs of the run sweep (one per SM, `warps` autonomous warps each, warp w on sub-partition
// w % 4).  Several rounds of items: every SM.  One round: the launch lasts as long as the
// fullest sub-partition, so the SMs that are used are filled up to that level and the others
// are left free -- the small kernels behind the PREVIOUS run of a repeated batch (fix-up,
// blocks, queries, the album exchange) run there while this sweep is under way
// (lg_batch.cu: pipelined runs).  12-track album: 2200 items -> 138 CTAs x 16 warps.
LG_BOTH uint32_t run_grid_ctas(uint32_t nitems, uint32_t sms, uint32_t warps, bool spare) {
  const uint32_t grid = nitems < sms ? nitems : sms;
  if (!spare || nitems <= sms || nitems > sms * warps || warps < 4u) return grid;
  const uint32_t per_sm = (nitems + sms - 1u) / sms;
  uint32_t fill = ((per_sm + 3u) / 4u) * 4u;
  if (fill > warps) fill = warps;
  const uint32_t g = (nitems + fill - 1u) / fill;
  return g < grid ? g : grid;
}

// ---- 2-D TMA view of a track -------------------------------------------------
// Chunks j = i*m + r of one class r start m*L frames apart, a multiple of 16
// bytes when m is chosen so, and every lane of the class has the same
// alignment offset.  So class r is a 2-D tensor: row i = the m*L frames from
// lane-local frame 0 of chunk i*m + r.  Class 0 starts W frames before the
// track; it is shifted by one row (row i - 1 holds chunk i*m).
constexpr int kTmaMaxM = 4;     // stereo rows: 16-byte pitch at m = 2 or 4 (aq <= 4 frames)
constexpr int kTmaBoxPad = 4;       // 32-bit words appended to a box row (odd 16-byte pitch in shared memory)

struct TmaClass {
  long long base_frame;   // track frame of word 0 of row 0
  int shift;              // chunk i*m + r lives in row i - shift
};

LG_BOTH int tma_interleave(int L, int aq) {
  int m = 2;
  while (m < kTmaMaxM && ((long long) m * L) % aq) m <<= 1;
  return ((long long) m * L) % aq ? 0 : m;
}

LG_BOTH TmaClass tma_class(int L, int W, int aq, int m, int r) {
  const long long s = (long long) r * L - W;
  const long long a = s & ~(long long) (aq - 1);       // as lane_geometry: floor, also for s < 0
  TmaClass c;
  c.shift = a < 0 ? 1 : 0;
  c.base_frame = a + (c.shift ? (long long) m * L : 0);
  return c;
}

// ---- lane geometry ---------------------------------------------------------
// Chunk j covers track frames [j*L, j*L + L).  Its lane starts filtering from
// zero state at frame a = align_down(j*L - W, aq), so that every staged row
// begins on a 16-byte boundary; o = (j*L - W) - a is the lane's offset.
// Lane-local frame f is track frame a + f; energy is accumulated over
// [W + o, W + o + L).
struct LaneGeom {
  long long a;     // track frame of lane-local frame 0 (may be negative)
  int o;           // 0 <= o < aq
  int l_valid;     // frames of the chunk that exist in the track (0..L)
};

LG_BOTH LaneGeom lane_geometry(long long frames, int L, int W, int aq, long long chunk) {
  LaneGeom g;
  const long long b = chunk * (long long) L;
  const long long s = b - W;
  g.a = s & ~(long long) (aq - 1);     // aq is a power of two: floor, also for s < 0
  g.o = (int) (s - g.a);
  const long long left = frames - b;
  g.l_valid = left <= 0 ? 0 : (left < L ? (int) left : L);
  return g;
}

// How an iteration [f0, f0 + kIter) is run, uniformly for all lanes of a warp.
enum IterKind : int { ITER_WARM = 0, ITER_FAST = 1, ITER_MASKED = 2 };

LG_BOTH int iter_kind(int f0, int W, int aq, int L, int lmin_valid) {
  if (f0 + kIter <= W) return ITER_WARM;       // before every lane's chunk
  const int lfast = lmin_valid < L ? lmin_valid : L;
  if (f0 >= W + aq - 1 && f0 + kIter <= W + lfast) return ITER_FAST;
  return ITER_MASKED;
}

}  // namespace lg
