// lg_batch.cu -- host layer of the batch API (include/ebur128_b200.h): plan
// upload, workspace, launch sequencing on one CUDA stream, result fetch.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ebur128_b200.h"
#include "lg_batch.h"
#include "lg_kernels.h"
#include "lg_plan.h"

namespace lg {

static thread_local std::string g_error;

void set_error(const char* what, cudaError_t e) {
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
}
void set_error(const char* what) { g_error = what; }

// Device memory comes from the stream-ordered pool: batches are created and
// destroyed per album on the drop-in path, and pool allocations are cheap.
template <class T>
static bool upload(const std::vector<T>& v, T** dev, cudaStream_t s) {
  *dev = nullptr;
  const size_t bytes = (v.empty() ? 1 : v.size()) * sizeof(T);
  cudaError_t e = cudaMallocAsync((void**) dev, bytes, s);
  if (e != cudaSuccess) { set_error("cudaMallocAsync", e); return false; }
  if (!v.empty()) {
    e = cudaMemcpyAsync(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { set_error("cudaMemcpyAsync(H2D tables)", e); return false; }
  }
  return true;
}

template <class T>
static bool dalloc(T** dev, uint64_t n, cudaStream_t s) {
  *dev = nullptr;
  cudaError_t e = cudaMallocAsync((void**) dev, (n ? n : 1) * sizeof(T), s);
  if (e != cudaSuccess) { set_error("cudaMallocAsync", e); return false; }
  return true;
}

int query_lists_sync(const BlockList* lists, size_t n, cudaStream_t stream, QueryResult* out) {
  std::vector<BlockList> hl(lists, lists + n);
  std::vector<uint32_t> hm(n);
  for (size_t i = 0; i < n; ++i) hm[i] = (uint32_t) i;
  std::vector<Query> hq(1, Query{0, (uint32_t) n});
  BlockList* dl = nullptr; uint32_t* dm = nullptr; Query* dq = nullptr; QueryResult* dr = nullptr;
  bool ok = upload(hl, &dl, stream) && upload(hm, &dm, stream) && upload(hq, &dq, stream) &&
            dalloc(&dr, 1, stream);
  if (ok) {
    uint64_t blocks = 0;
    for (size_t i = 0; i < n; ++i) blocks += lists[i].nz;
    cudaError_t e = launch_queries(dl, dq, dm, 1, pow(10.0, (-70.0 + 0.691) / 10.0), dr, stream,
                                   query_cluster_size(blocks));
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dr, sizeof(QueryResult), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("query_lists_sync", e); ok = false; }
  }
  cudaFreeAsync(dl, stream); cudaFreeAsync(dm, stream); cudaFreeAsync(dq, stream);
  cudaFreeAsync(dr, stream);
  return ok ? 0 : 1;
}

int query_each_sync(const BlockList* lists, size_t n, cudaStream_t stream, QueryResult* out) {
  if (!n) return 0;
  std::vector<BlockList> hl(lists, lists + n);
  std::vector<uint32_t> hm(n);
  std::vector<Query> hq(n);
  uint64_t largest = 0;
  for (size_t i = 0; i < n; ++i) {
    hm[i] = (uint32_t) i;
    hq[i] = Query{(uint32_t) i, 1u};
    if (lists[i].nz > largest) largest = lists[i].nz;
  }
  BlockList* dl = nullptr; uint32_t* dm = nullptr; Query* dq = nullptr; QueryResult* dr = nullptr;
  bool ok = upload(hl, &dl, stream) && upload(hm, &dm, stream) && upload(hq, &dq, stream) &&
            dalloc(&dr, n, stream);
  if (ok) {
    cudaError_t e = launch_queries(dl, dq, dm, (uint32_t) n, pow(10.0, (-70.0 + 0.691) / 10.0), dr, stream,
                                   query_cluster_size(largest));
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dr, n * sizeof(QueryResult), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("query_each_sync", e); ok = false; }
  }
  cudaFreeAsync(dl, stream); cudaFreeAsync(dm, stream); cudaFreeAsync(dq, stream);
  cudaFreeAsync(dr, stream);
  return ok ? 0 : 1;
}

}  // namespace lg

using namespace lg;

static void to_result(const QueryResult& q, lgb_result& r) {
  r.loudness = q.loudness; r.range = q.range; r.rel_threshold = q.rel_thr;
  r.sum_abs = q.sum1; r.sum_rel = q.sum2; r.n_abs = q.n1; r.n_rel = q.n2; r.n_shortterm = q.nst;
}

// Exchange region of one rank for album queries over tracks on several GPUs
// (lg_common.h: XchgParams; kernels in lg_kernels.cu).
struct lgb_exchange {
  uint32_t world = 1, rank = 0, nalbums = 0;
  uint64_t st_cap = 0;
  unsigned char* region = nullptr;               // cudaMalloc (IPC-exportable), this rank's
  unsigned char* peer[kMaxWorld] = {nullptr};    // every rank's region as mapped here
  bool opened = false;
  unsigned long long* d_ctl = nullptr;           // [8]: step, CTA counters, time-outs
  unsigned long long* h_ctl = nullptr;           // pinned mirror of the time-out counter
};

// Result mirrors of a batch = runs that may be in flight at once (lgb_batch_set_max_in_flight).
constexpr int kMirrors = 3;

struct lgb_batch {
  Plan plan;
  cudaStream_t stream = nullptr;
  // device tables
  Track* d_tracks = nullptr;
  CoefSet* d_coefs = nullptr;
  WarpWork* d_work = nullptr;
  Query* d_queries = nullptr;
  uint32_t* d_members = nullptr;
  BlockList* d_lists = nullptr;
  ChunkRec* d_recs = nullptr;
  ChunkRec* d_recs_alt[kMirrors] = {nullptr};   // pipelined runs: the chunk records of mirrors 1, 2 (allocated with pstream)
  bool recs_per_mirror = false;
  uint32_t* d_mrec = nullptr;
  unsigned char* d_tmaps = nullptr;    // tensor maps of the TMA-staged groups, kTmaMaxM x 128 B per track
  uint64_t* d_tpq = nullptr;           // candidate queue of the packed true-peak pass
  RunItem* d_items = nullptr;          // work items of the run-sweep groups
  cplx* d_xi = nullptr;                // plan.xi_table
  uint32_t* d_runq = nullptr;          // candidate queues of the run-sweep groups (32-bit entries)
  uint32_t* d_runcnt = nullptr;        // [items * 32] candidates queued per sweep lane
  double* d_eslot = nullptr;
  double* d_zblock = nullptr;
  double* d_zst = nullptr;
  // With an album exchange attached, mirrors 1 and 2 get slot / block / short-term buffers (and
  // block-list tables) of their own: the exchange of run k waits for the slowest rank, and the
  // fix-up of run k + 1 must not wait for it in turn (lgb_batch_attach_exchange).
  bool blocks_per_mirror = false;
  double* d_eslot_m[kMirrors] = {nullptr};
  double* d_zblock_m[kMirrors] = {nullptr};
  double* d_zst_m[kMirrors] = {nullptr};
  BlockList* d_lists_m[kMirrors] = {nullptr};
  // [results][peak cells][per-group counters]: read back per step.  Two of them, used by
  // alternate runs: the sweep of run k + 1 raises its peak cells while the queries of run k
  // are still filling in their results (pipelined runs, below).
  uint32_t* d_out2[kMirrors] = {nullptr};
  // Pinned host mirrors of the (tiny) results, filled by the step itself.  Two of
  // them, used alternately: a second run may be enqueued before the first one's
  // results are fetched (the host turn-around between steps then overlaps the GPU).
  unsigned char* h_out[kMirrors] = {nullptr};
  size_t out_bytes = 0, peaks_off = 0;
  cudaEvent_t ev_done[kMirrors] = {nullptr};
  uint32_t max_in_flight = 2;
  // A batch that is run repeatedly replays its step as a CUDA graph (one per
  // mirror): the first run launches directly, the later ones capture / replay.
  cudaGraphExec_t graph[kMirrors] = {nullptr};
  bool graph_off = false;
  // Pipelined runs (a batch that is run again and again): only the sweep and the true-peak
  // evaluation stay on the caller's stream; the fix-up, the blocks, the queries (with their
  // album exchange) and the read-back of run k go to `pstream` and finish while the sweep of
  // run k + 1 is under way -- that sweep waits for nothing of run k (odd and even runs have
  // their own chunk records and result areas).  The small kernels find room on the
  // SMs the sweep leaves free (lg_common.h: run_grid_ctas).  Everything behind the block
  // kernel replays as a CUDA graph per mirror.
  bool pipeline = true, post_in_flight = false, pgraph_off = false;
  cudaStream_t pstream = nullptr;      // lowest priority
  // ... and the sweeps / the evaluation of pipelined runs go to a HIGH-priority stream of the
  // library (ordered behind the caller's stream at every run, the caller's stream behind it
  // again): a persistent sweep lasts as long as its last CTA, so when the evaluation of run k
  // ends its CTAs must get their SMs before the post-processing kernels of run k settle there.
  cudaStream_t mstream = nullptr, pq = nullptr, pq2 = nullptr;
  cudaStream_t fstream = nullptr;      // high priority: fix-up and blocks, next to the evaluation
  cudaEvent_t ev_in = nullptr;
  cudaStream_t ms = nullptr;           // where the current step's sweeps go (stream or mstream)
  // LOUDGAIN_B200_MTRACE (tuning): timing events around the sweep + evaluation of pipelined runs;
  // the fetch prints their duration and the idle time of mstream since the run before
  bool mtrace = false;
  cudaEvent_t tm0[kMirrors] = {nullptr}, tm1[kMirrors] = {nullptr}, tm2[kMirrors] = {nullptr};
  cudaEvent_t ev_swept = nullptr, ev_pidle = nullptr, ev_mdone[kMirrors] = {nullptr};
  cudaGraphExec_t pgraph[kMirrors] = {nullptr};
  size_t nmarks = 0;
  bool cells_clean[kMirrors] = {false};    // the mirror's peak cells were zeroed behind the last pipelined run that used them
  uint32_t runs = 0, fetched = 0;
  // The true-peak pass only feeds the peak cells and the fix-up / block / query
  // kernels never read them, so after the sweep the step forks: the small
  // post-processing kernels go to a high-priority side stream and slip in next
  // to the true-peak pass on the main stream; joined before the result copy.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_post = nullptr, ev_fix = nullptr;
  // Sweep launch groups after the first go round-robin to these streams, so
  // that the CTAs of the next group fill the SMs while the previous group's
  // last work items run out (the groups touch disjoint tracks).
  static constexpr int kGroupStreams = 3;
  cudaStream_t gstream[kGroupStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_gfork = nullptr, ev_gjoin[kGroupStreams] = {nullptr, nullptr, nullptr};
  int ngstreams = 0;
  uint32_t pair_ctas = 0;            // tuning: cap on resident CTAs per SM of the packed sweep (0 = none)
  cudaEvent_t ev_blocks = nullptr;   // block lists of the current run are complete (lgb_batch_wait_blocks)
  const double* hist_tab = nullptr;  // set when a track asks for histogram-mode block energies
  double abs_gate = 0.0;
  uint32_t launches = 0, sweep_launches = 0, sms = 148;
  uint32_t query_cluster = 1;        // CTAs per query (lg_kernels.cu: query_kernel)
  // album queries answered together with other ranks (lgb_batch_attach_exchange)
  lgb_exchange* xchg = nullptr;
  uint32_t* d_xstoff = nullptr;
  uint32_t xst_smem = 0;             // short-term energies the exchange's range CTA stages in shared memory
  uint32_t xcluster = 1;             // CTAs that share an album's local gating blocks in the exchange
  cudaStream_t qstream = nullptr;    // second and third query stream: the halves of the queries (and, with an
  cudaStream_t q2stream = nullptr;   // exchange, the albums' ranges) run next to each other
  cudaEvent_t ev_q0 = nullptr, ev_q1 = nullptr, ev_q2 = nullptr, ev_pub = nullptr;
  // LOUDGAIN_B200_STEP_TRACE (tuning): direct launches with a timing event behind every
  // stage of the step on the stream it runs on; the fetch prints them
  bool trace = false, trace_pipelined = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks2[kMirrors];    // per mirror
  int mark_set = 0;
  // optional sweep timing
  bool timing = false, timed_run_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  double sweep_ms_total = 0.0, tp_ms_total = 0.0;
  uint64_t sweep_runs = 0;

  QueryResult* d_results(int parity) const { return reinterpret_cast<QueryResult*>(d_out2[parity]); }
  uint32_t* d_peaks(int parity) const { return d_out2[parity] + 16 * plan.queries.size(); }
  DeviceTables tables(int parity) const {
    DeviceTables t;
    t.tracks = d_tracks; t.coefs = d_coefs; t.work = d_work;
    t.queries = d_queries; t.members = d_members; t.lists = d_lists; t.recs = (parity && recs_per_mirror) ? d_recs_alt[parity] : d_recs; t.peaks = d_peaks(parity);
    t.eslot = d_eslot; t.zblock = d_zblock; t.zst = d_zst;
    if (blocks_per_mirror && parity) {
      t.eslot = d_eslot_m[parity]; t.zblock = d_zblock_m[parity]; t.zst = d_zst_m[parity];
      t.lists = d_lists_m[parity];
    }
    t.results = d_results(parity); t.xi_table = d_xi; t.hist_tab = hist_tab;
    return t;
  }
};

// Tensor maps for the groups the planner gave 2-D TMA staging (lg_common.h:
// tma_class): per track and chunk class r one map over rows of m*L frames.
static bool make_tensor_maps(lgb_batch* b) {
  const Plan& p = b->plan;
  bool any = false;
  for (const SweepGroup& g : p.groups) any = any || g.params.tma_m != 0 || g.run;
  if (!any) return true;
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available");
      return false;
    }
    encode = (PFN_cuTensorMapEncodeTiled) fn;
  }
  std::vector<CUtensorMap> maps(p.tracks.size() * kTmaMaxM);
  memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
  auto fail = [](CUresult rc) {
    char msg[96];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d)", (int) rc);
    set_error(msg);
    return false;
  };
  // Run sweep (lg_run.cu): per track the 2-D view "row y = run y" with a 32-row box,
  // the same view with a box of the last item's complete rows, and the partial
  // last run as a one-row tensor that ends with the track.
  for (const SweepGroup& g : p.groups) {
    if (!g.run) continue;
    const SweepParams& sp = g.params;
    const cuuint32_t wpf = sp.fb / 4u, boxw = sp.run_stage_frames * wpf + 4u;
    const cuuint32_t estr[2] = {1, 1};
    for (uint32_t ii = 0; ii < g.nitems; ++ii) {
      const RunItem& it = p.items[g.first_item + ii];
      if (it.first_run != 0) continue;                     // once per track
      const Track& tr = p.tracks[it.track];
      CUtensorMap* m = &maps[(size_t) it.track * kTmaMaxM];
      const cuuint64_t pitch[1] = {(cuuint64_t) sp.Lr * sp.fb};
      {
        const cuuint64_t dims[2] = {tr.nfull ? (cuuint64_t) sp.Lr * wpf : (cuuint64_t) tr.frames * wpf,
                                    tr.nfull ? (cuuint64_t) tr.nfull : 1u};
        const cuuint32_t box[2] = {boxw, 32u};
        const CUresult rc = encode(&m[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*) tr.pcm, dims, pitch, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail(rc);
      }
      const uint32_t last_first = ((tr.nruns - 1u) / 32u) * 32u;
      const uint32_t tail_rows = tr.nfull > last_first ? tr.nfull - last_first : 0u;
      if (tail_rows > 0 && tail_rows < 32u) {
        const cuuint64_t dims[2] = {(cuuint64_t) sp.Lr * wpf, (cuuint64_t) tr.nfull};
        const cuuint32_t box[2] = {boxw, tail_rows};
        const CUresult rc = encode(&m[1], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*) tr.pcm, dims, pitch, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail(rc);
      }
      if (tr.nruns > tr.nfull) {
        const uint64_t done = (uint64_t) tr.nfull * (uint64_t) sp.Lr;
        const cuuint64_t dims[2] = {(cuuint64_t) (tr.frames - done) * wpf, 1u};
        const cuuint32_t box[2] = {boxw, 1u};
        void* base = (void*) ((const unsigned char*) tr.pcm + done * sp.fb);
        const CUresult rc = encode(&m[2], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, pitch, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail(rc);
      }
    }
  }
  for (const SweepGroup& g : p.groups) {
    const SweepParams& sp = g.params;
    if (!sp.tma_m) continue;
    const long long m = sp.tma_m, mL = m * sp.L;
    const uint32_t wpf = sp.fb / 4u;                        // 32-bit words per frame
    for (uint32_t w = 0; w < g.nwarps; ++w) {
      const WarpWork& ww = p.work[g.first_warp + w];
      if (ww.interior != 2 || maps[(size_t) ww.track * kTmaMaxM].opaque[0]) continue;   // done
      const Track& tr = p.tracks[ww.track];
      for (int r = 0; r < (int) m; ++r) {
        const TmaClass tc = tma_class(sp.L, sp.W, (int) tr.aq, (int) m, r);
        const long long nrows = ((long long) tr.frames - tc.base_frame) / mL;
        if (nrows < 1) continue;
        const cuuint64_t dims[2] = {(cuuint64_t) (mL * wpf), (cuuint64_t) nrows};
        const cuuint64_t strides[1] = {(cuuint64_t) (mL * sp.fb)};
        const cuuint32_t box[2] = {(cuuint32_t) (sp.stage_row_bytes / 4u + kTmaBoxPad), (cuuint32_t) (32 / m)};
        const cuuint32_t estr[2] = {1, 1};
        void* base = (void*) ((const unsigned char*) tr.pcm + tc.base_frame * (long long) sp.fb);
        const CUresult rc = encode(&maps[(size_t) ww.track * kTmaMaxM + r], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
                                   base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
          char msg[96];
          snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d)", (int) rc);
          set_error(msg);
          return false;
        }
      }
    }
  }
  if (cudaMallocAsync((void**) &b->d_tmaps, maps.size() * sizeof(CUtensorMap), b->stream) != cudaSuccess ||
      cudaMemcpyAsync(b->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice,
                      b->stream) != cudaSuccess ||
      cudaStreamSynchronize(b->stream) != cudaSuccess) {
    set_error("tensor map upload failed");
    return false;
  }
  return true;
}

extern "C" LG_EXPORT const char* lgb_last_error(void) { return g_error.c_str(); }

extern "C" LG_EXPORT lgb_batch* lgb_batch_create(const lgb_track* tracks, size_t ntracks, uint32_t nalbums,
                                       void* cuda_stream) {
  g_error.clear();
  std::vector<TrackIn> in(ntracks);
  bool any_hist = false;
  for (size_t i = 0; i < ntracks; ++i) {
    const lgb_track& t = tracks[i];
    if (t.channels == 0 || t.channels > (uint32_t) kMaxChannels || t.samplerate < 16 ||
        t.format > LGB_FORMAT_F32 || (t.frames && !t.pcm) ||
        (t.album != LGB_NO_ALBUM && t.album >= nalbums)) {
      set_error("lgb_batch_create: invalid track descriptor");
      return nullptr;
    }
    if (((uintptr_t) t.pcm & 15u) != 0) {
      set_error("lgb_batch_create: pcm must be 16-byte aligned");
      return nullptr;
    }
    const uint32_t s100 = (t.samplerate + 5) / 10;
    if (t.lead_in % s100 || t.lead_in > t.frames) {
      set_error("lgb_batch_create: lead_in must be a whole number of 100 ms slots within the track");
      return nullptr;
    }
    in[i] = TrackIn{t.pcm, t.frames, t.channels, t.samplerate, t.format, t.album, t.weight_class,
                    t.lead_in, t.flags & LGB_TRACK_HISTOGRAM};
    any_hist = any_hist || (t.flags & LGB_TRACK_HISTOGRAM);
  }
  lgb_batch* b = new lgb_batch();
  if (any_hist && !(b->hist_tab = hist_table())) {
    set_error("lgb_batch_create: the histogram table could not be created");
    delete b;
    return nullptr;
  }
  b->stream = (cudaStream_t) cuda_stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Two full waves of resident threads when the audio is long enough.
  PlanOptions opt;
  opt.target_tasks = (uint64_t) sms * 2048u;
  if (const char* e = getenv("LOUDGAIN_B200_CHUNKS_PER_SLOT")) opt.force_k = atoi(e);   // tuning
  if (const char* e = getenv("LOUDGAIN_B200_SCALAR_SWEEP")) opt.allow_packed = atoi(e) == 0;
  if (const char* e = getenv("LOUDGAIN_B200_TMA")) opt.use_tma = atoi(e) != 0;   // 0: cp.async staging only
  opt.sms = (uint32_t) sms;
  if (const char* e = getenv("LOUDGAIN_B200_RUN")) opt.use_run = atoi(e) != 0;   // 0: stereo through lg_pair.cu
  if (const char* e = getenv("LOUDGAIN_B200_RUN_CHUNKS")) opt.force_run_chunks = atoi(e);   // tuning
  if (const char* e = getenv("LOUDGAIN_B200_RUN_WARPS")) opt.run_warps_per_sm = (uint32_t) atoi(e);   // tuning
  if (const char* e = getenv("LOUDGAIN_B200_SPARE_SMS")) opt.spare_sms = atoi(e) != 0;   // 0: the sweep takes every SM
  if (const char* e = getenv("LOUDGAIN_B200_PIPELINE")) b->pipeline = atoi(e) != 0;      // 0: one run behind the other
  if (const char* e = getenv("LOUDGAIN_B200_PAIR_CTAS")) b->pair_ctas = (uint32_t) atoi(e);   // tuning
  b->mtrace = getenv("LOUDGAIN_B200_MTRACE") != nullptr;
  if (b->mtrace)
    for (int i = 0; i < kMirrors; ++i) { cudaEventCreate(&b->tm0[i]); cudaEventCreate(&b->tm1[i]); cudaEventCreate(&b->tm2[i]); }
  b->trace = getenv("LOUDGAIN_B200_STEP_TRACE") != nullptr;     // =2: the pipelined form of the step
  b->trace_pipelined = b->trace && atoi(getenv("LOUDGAIN_B200_STEP_TRACE")) == 2;
  if (b->trace_pipelined) b->pgraph_off = true;                 // (timing events cannot sit inside a graph)
  if (const char* e = getenv("LOUDGAIN_B200_TAIL_FRAC")) opt.tail_frac = atof(e);      // tuning
  if (const char* e = getenv("LOUDGAIN_B200_TAIL_DIV")) opt.tail_div = atoi(e);
  build_plan(in.data(), ntracks, nalbums, opt, b->plan);
  const Plan& p = b->plan;
  if (getenv("LOUDGAIN_B200_VERBOSE"))
    for (const SweepGroup& g : p.groups)
      fprintf(stderr, "[lgb] group fmt=%u ch=%u run=%d L=%d R=%d Wp=%d niters=%d xi_iters=%d items/warps=%u nstages=%u\n",
              g.format, g.channels, (int) g.run, g.params.L, g.params.R, g.params.Wp, g.params.niters,
              g.params.xi_iters, g.nwarps, g.params.run_nstages);
  b->abs_gate = pow(10.0, (-70.0 + 0.691) / 10.0);
  bool ok = upload(p.tracks, &b->d_tracks, b->stream) && upload(p.coefs, &b->d_coefs, b->stream) &&
            upload(p.work, &b->d_work, b->stream) &&
            upload(p.queries, &b->d_queries, b->stream) &&
            upload(p.members, &b->d_members, b->stream) &&
            dalloc(&b->d_recs, p.total_recs, b->stream) &&
            upload(p.items, &b->d_items, b->stream) && upload(p.xi_table, &b->d_xi, b->stream) &&
            dalloc(&b->d_runq, 2 * p.total_queue, b->stream) &&          // 64-bit entries
            dalloc(&b->d_runcnt, (uint64_t) p.items.size() * 32u, b->stream) &&
            dalloc(&b->d_out2[0], 16 * (uint64_t) p.queries.size() + 2 * p.total_peaks + p.groups.size() + 1,
                   b->stream) &&
            dalloc(&b->d_out2[1], 16 * (uint64_t) p.queries.size() + 2 * p.total_peaks + p.groups.size() + 1,
                   b->stream) &&
            dalloc(&b->d_out2[2], 16 * (uint64_t) p.queries.size() + 2 * p.total_peaks + p.groups.size() + 1,
                   b->stream) &&
            dalloc(&b->d_mrec, p.total_mrec, b->stream) &&
            dalloc(&b->d_tpq, 2 * p.total_mrec, b->stream) &&
            dalloc(&b->d_eslot, p.total_slots, b->stream) &&
            dalloc(&b->d_zblock, p.total_blocks, b->stream) &&
            dalloc(&b->d_zst, p.total_st, b->stream);
  static_assert(sizeof(QueryResult) == 64, "results sit in front of the 32-bit peak cells");
  if (ok) {
    b->peaks_off = p.queries.size() * sizeof(QueryResult);
    b->out_bytes = b->peaks_off + 2 * p.total_peaks * sizeof(uint32_t);
  }
  if (ok) ok = make_tensor_maps(b);
  if (ok) {
    std::vector<BlockList> lists(p.tracks.size());
    for (size_t i = 0; i < p.tracks.size(); ++i) {
      const Track& tr = p.tracks[i];
      lists[i] = BlockList{b->d_zblock + tr.block_base, b->d_zst + tr.st_base, tr.nblocks, tr.nst};
    }
    ok = upload(lists, &b->d_lists, b->stream);
  }
  if (ok) {
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < kMirrors && e == cudaSuccess; ++k) {
      e = cudaMallocHost((void**) &b->h_out[k], b->out_bytes ? b->out_bytes : 1);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_done[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { set_error("cudaMallocHost(results)", e); ok = false; }
  }
  if (ok) {
    // the table uploads read host vectors that die with this call's scope
    // only in `lists`; the plan's own vectors live as long as the batch
    const cudaError_t e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) { set_error("lgb_batch_create", e); ok = false; }
  }
  if (ok) {
    // high-priority side stream for the post-processing kernels; the batch
    // still works without it
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_post, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_fix, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_blocks, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      if (b->side) { cudaStreamDestroy(b->side); b->side = nullptr; }
    }
    // streams for the sweep launch groups after the first (only if there are any)
    if (p.groups.size() > 1 && cudaEventCreateWithFlags(&b->ev_gfork, cudaEventDisableTiming) == cudaSuccess) {
      const int want = (int) std::min<size_t>(p.groups.size() - 1, (size_t) lgb_batch::kGroupStreams);
      for (int j = 0; j < want; ++j) {
        if (cudaStreamCreateWithFlags(&b->gstream[j], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->ev_gjoin[j], cudaEventDisableTiming) != cudaSuccess) {
          cudaGetLastError();
          break;
        }
        b->ngstreams = j + 1;
      }
    }
  }
  if (!ok) { lgb_batch_destroy(b); return nullptr; }
  b->sms = (uint32_t) sms;
  {
    uint64_t most = 0;
    for (const Query& q : p.queries) {
      uint64_t blocks = 0;
      for (uint32_t m = 0; m < q.count; ++m) blocks += p.tracks[p.members[q.first + m]].nblocks;
      if (blocks > most) most = blocks;
    }
    b->query_cluster = query_cluster_size(most);
  }
  b->sweep_launches = (uint32_t) p.groups.size();
  uint32_t tp_launches = 0;
  for (const SweepGroup& g : p.groups) tp_launches += g.tpf ? ((g.params.packed && !g.run) ? 2u : 1u) : 0u;
  b->launches = b->sweep_launches + tp_launches + (p.total_slots ? 1 : 0) +
                ((p.total_blocks + p.total_st) ? 1 : 0) + (p.queries.empty() ? 0 : 1);
  return b;
}

// ---- the pieces of a step -------------------------------------------------------
// LOUDGAIN_B200_STEP_TRACE: a timing event behind a stage of the step, on the stream it runs on
static void step_mark(lgb_batch* b, const char* name, cudaStream_t s) {
  if (!b->trace) return;
  auto& marks = b->marks2[b->mark_set];
  if (b->nmarks == marks.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    marks.emplace_back(name, ev);
  }
  cudaEventRecord(marks[b->nmarks++].second, s);
}

// Peak cells zeroed, then every sweep launch group (the first on the batch's stream, the
// others round-robin on the group streams, joined again).
static int enqueue_sweeps(lgb_batch* b, const DeviceTables& t, int parity, bool cells_may_be_clean = false) {
  const Plan& p = b->plan;
  cudaError_t e = cudaSuccess;
  // peak cells, then per launch group a true-peak queue counter and a work-item ticket
  // (a pipelined run zeroes them behind its read-back, off the next sweep's path)
  if (!(cells_may_be_clean && b->cells_clean[parity]))
    e = cudaMemsetAsync(b->d_peaks(parity), 0, (2 * p.total_peaks + p.groups.size() + 1) * sizeof(uint32_t),
                        b->ms);
  b->cells_clean[parity] = false;
  if (e != cudaSuccess) { set_error("cudaMemsetAsync(peaks)", e); return 1; }
  step_mark(b, "memset", b->ms);
  if (b->timing) cudaEventRecord(b->ev0, b->ms);
  const bool gfork = p.groups.size() > 1 && b->ngstreams > 0;
  if (gfork) {
    e = cudaEventRecord(b->ev_gfork, b->ms);
    if (e != cudaSuccess) { set_error("fork(group streams)", e); return 1; }
  }
  size_t gidx = 0;
  for (const SweepGroup& g : p.groups) {
    SweepParams sp = g.params;
    sp.tracks = t.tracks; sp.work = t.work + g.first_warp; sp.recs = t.recs; sp.peaks = t.peaks;
    sp.mrec = b->d_mrec + g.mrec_base;
    sp.tmaps = b->d_tmaps;
    sp.tp_queue = b->d_tpq + 2 * g.mrec_base;
    cudaStream_t gs = b->ms;
    if (gfork && gidx > 0) {
      gs = b->gstream[(gidx - 1) % (size_t) b->ngstreams];
      if (gidx <= (size_t) b->ngstreams) {          // first use of this stream in the step
        e = cudaStreamWaitEvent(gs, b->ev_gfork, 0);
        if (e != cudaSuccess) { set_error("fork(group streams)", e); return 1; }
      }
    }
    sp.ctas_per_sm = b->pair_ctas;
    if (g.run) {
      sp.items = b->d_items + g.first_item;
      sp.tp_ticket = t.peaks + 2 * p.total_peaks + gidx;
      sp.run_counts = b->d_runcnt + (size_t) g.first_item * 32u;
      sp.run_queue = b->d_runq + 2 * g.queue_base;
    }
    ++gidx;
    e = g.run ? launch_sweep_run(sp, g.format, g.tpf, b->sms, gs)
        : sp.packed ? launch_sweep_pair(sp, g.format, g.tpf, gs)
                    : launch_sweep(sp, g.format, g.tpf, gs);
    if (e != cudaSuccess) { set_error("launch_sweep", e); return 1; }
  }
  if (gfork) {
    const size_t used = std::min<size_t>(p.groups.size() - 1, (size_t) b->ngstreams);
    for (size_t j = 0; j < used; ++j) {
      e = cudaEventRecord(b->ev_gjoin[j], b->gstream[j]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(b->ms, b->ev_gjoin[j], 0);
      if (e != cudaSuccess) { set_error("join(group streams)", e); return 1; }
    }
  }
  if (b->timing) { cudaEventRecord(b->ev1, b->ms); b->timed_run_pending = true; }
  step_mark(b, "sweep", b->ms);
  return 0;
}

// The true-peak pass of every group on the batch's stream: it needs the final sample peaks of
// every track of a group.  `fork`: the post-processing kernels run next to it on the side stream.
static int enqueue_truepeak(lgb_batch* b, const DeviceTables& t, bool fork, uint32_t cta_cap = 0) {
  const Plan& p = b->plan;
  uint32_t gi = 0;
  for (const SweepGroup& g : p.groups) {
    SweepParams sp = g.params;
    sp.tracks = t.tracks; sp.work = t.work + g.first_warp; sp.recs = t.recs; sp.peaks = t.peaks;
    sp.mrec = b->d_mrec + g.mrec_base;
    sp.tp_ticket = t.peaks + 2 * p.total_peaks + gi++;
    sp.npeak_words = (uint32_t) std::min<uint64_t>(2 * p.total_peaks, 0xffffffffu);
    sp.tp_queue = b->d_tpq + 2 * g.mrec_base;
    if (g.run) {
      sp.items = b->d_items + g.first_item;
      sp.run_counts = b->d_runcnt + (size_t) g.first_item * 32u;
      sp.run_queue = b->d_runq + 2 * g.queue_base;
    }
    // what the true-peak evaluation waits for (tuning, LOUDGAIN_B200_TP_HOLD): 0 nothing,
    // 1 the block kernel, 2 the fix-up kernel -- the evaluation's gathers slow the FP64 fix-up
    // down when the two run side by side
    static const int hold_mode = [] {
      const char* e = getenv("LOUDGAIN_B200_TP_HOLD");
      return e ? atoi(e) : 0;
    }();
    cudaEvent_t hold = !fork ? nullptr : hold_mode == 1 ? b->ev_post : hold_mode == 2 ? b->ev_fix : nullptr;
    cudaEvent_t hold_old = fork ? b->ev_post : nullptr;     // the round-1 kernels keep their arrangement
    const cudaError_t e = g.run ? launch_truepeak_run(sp, g.format, g.tpf, b->sms, b->ms, hold, cta_cap)
                          : sp.packed ? launch_truepeak_pair(sp, g.format, g.tpf, b->sms, b->ms, hold_old)
                                      : launch_truepeak(sp, g.format, g.tpf, b->sms, b->ms, hold_old);
    if (e != cudaSuccess) { set_error("launch_truepeak", e); return 1; }
  }
  step_mark(b, "true-peak pass", b->ms);
  if (b->timing) cudaEventRecord(b->ev2, b->ms);
  return 0;
}

// ev_blocks: the block lists of the run are complete (lgb_batch_wait_blocks).  While the step is
// being captured into a graph the record must be an external event node, so that streams
// outside the graph can wait for it.
static int record_blocks_event(lgb_batch* b, cudaStream_t ps) {
  if (!b->ev_blocks) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(ps, &cap);
  const cudaError_t e = cap == cudaStreamCaptureStatusActive
                            ? cudaEventRecordWithFlags(b->ev_blocks, ps, cudaEventRecordExternal)
                            : cudaEventRecord(b->ev_blocks, ps);
  if (e != cudaSuccess) { set_error("cudaEventRecord(blocks)", e); return 1; }
  return 0;
}

// Track and album queries on `ps` (with the second / third query stream next to it).
static int enqueue_queries(lgb_batch* b, const DeviceTables& t, cudaStream_t ps, int parity,
                           cudaStream_t qs = nullptr, cudaStream_t q2s = nullptr) {
  if (!qs) qs = b->qstream;
  if (!q2s) q2s = b->q2stream;
  const Plan& p = b->plan;
  cudaError_t e = cudaSuccess;
  if (b->xchg) {
    // albums span ranks: this rank's share goes out first, the track queries hide the
    // peers' latency, then the album totals are gated (lg_kernels.cu: xchg_*_kernel)
    lgb_exchange* x = b->xchg;
    XchgParams xp{};
    xp.world = x->world; xp.rank = x->rank; xp.nalbums = x->nalbums;
    xp.first_query = (uint32_t) p.tracks.size();
    xp.st_cap = x->st_cap;
    for (uint32_t r = 0; r < x->world; ++r) xp.peer[r] = x->peer[r];
    xp.st_off = b->d_xstoff;
    xp.ctl = x->d_ctl;
    // three chains side by side: publish -> gate -> finish | the tracks' ranges | the tracks'
    // loudness, then the albums' ranges (which need the peers' publish phase only)
    const uint32_t ntq = (uint32_t) p.tracks.size();
    e = cudaEventRecord(b->ev_q0, ps);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(qs, b->ev_q0, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(q2s, b->ev_q0, 0);
    if (e == cudaSuccess) e = launch_exchange_publish(t.lists, t.queries, t.members, b->abs_gate, xp, b->xcluster, ps);
    // Every kernel that WAITS for the ranks' flags must be ordered behind this rank's own
    // publish: with hundreds of albums the waiting CTAs of the range kernel can fill every
    // SM, and a publish kernel that cannot get an SM never raises the flag they wait for.
    if (e == cudaSuccess) e = cudaEventRecord(b->ev_pub, ps);
    step_mark(b, "x-publish", ps);
    if (e == cudaSuccess)
      e = launch_queries(t.lists, t.queries, t.members, ntq, b->abs_gate, t.results, qs, 1, 2);
    step_mark(b, "track ranges", qs);
    if (e == cudaSuccess) e = cudaEventRecord(b->ev_q1, qs);
    if (e == cudaSuccess)
      e = launch_queries(t.lists, t.queries, t.members, ntq, b->abs_gate, t.results, q2s, 1, 1);
    step_mark(b, "track loudness", q2s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(q2s, b->ev_pub, 0);
    if (e == cudaSuccess) e = launch_exchange_range(b->abs_gate, t.results, xp, b->xst_smem, q2s);
    step_mark(b, "x-range", q2s);
    if (e == cudaSuccess) e = cudaEventRecord(b->ev_q2, q2s);
    if (e == cudaSuccess) e = launch_exchange_gate(t.lists, t.queries, t.members, b->abs_gate, xp, b->xcluster, ps);
    step_mark(b, "x-gate", ps);
    if (e == cudaSuccess) e = launch_exchange_finish(t.results, xp, ps);
    step_mark(b, "x-finish", ps);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ps, b->ev_q1, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ps, b->ev_q2, 0);
    // behind both chains: the step counter every exchange kernel reads at its start
    if (e == cudaSuccess) e = launch_exchange_advance(xp, ps);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(x->h_ctl, x->d_ctl + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ps);
  } else if (qs && !b->timing) {
    // the loudness range (short-term lists, one CTA per query) next to the integrated
    // loudness (gating lists, a cluster per query): two launches on two streams
    e = cudaEventRecord(b->ev_q0, ps);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(qs, b->ev_q0, 0);
    if (e == cudaSuccess)
      e = launch_queries(t.lists, t.queries, t.members, (uint32_t) p.queries.size(), b->abs_gate, t.results,
                         qs, 1, 2);
    if (e == cudaSuccess) e = cudaEventRecord(b->ev_q1, qs);
    if (e == cudaSuccess)
      e = launch_queries(t.lists, t.queries, t.members, (uint32_t) p.queries.size(), b->abs_gate, t.results,
                         ps, b->query_cluster, 1);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ps, b->ev_q1, 0);
  } else {
    e = launch_queries(t.lists, t.queries, t.members, (uint32_t) p.queries.size(), b->abs_gate,
                       t.results, ps, b->query_cluster);
  }
  (void) parity;
  if (e != cudaSuccess) { set_error("launch_queries", e); return 1; }
  step_mark(b, "queries", ps);
  return 0;
}

// One complete step on the batch's stream: sweep, true-peak pass, FP64
// fix-up, blocks, queries, and the copy of the scalars into the pinned mirrors.
static int enqueue_step(lgb_batch* b, int parity) {
  const Plan& p = b->plan;
  const DeviceTables t = b->tables(parity);
  cudaError_t e = cudaSuccess;
  b->nmarks = 0;
  b->mark_set = parity;
  b->ms = b->stream;
  if (enqueue_sweeps(b, t, parity)) return 1;
  // Fork (not in timed runs: those keep everything on the main stream, between
  // the events): fix-up, slot and block kernels go to the high-priority side
  // stream first.  The true-peak evaluation fills every SM for its whole
  // duration, so it is held back until the block kernel is done: then the query
  // kernel (a handful of large CTAs, high priority) and the evaluation become
  // ready together and share the GPU, instead of the queries waiting for SMs.
  const bool fork = !b->timing && b->side != nullptr;
  cudaStream_t ps = fork ? b->side : b->stream;
  PostSizes z{(uint32_t) p.tracks.size(), p.total_recs, p.total_slots, p.total_blocks, p.total_st};
  auto post_kernels = [&]() -> int {
    e = launch_post(t, z, ps, fork ? b->ev_fix : nullptr);
    if (e != cudaSuccess) { set_error("launch_post", e); return 1; }
    step_mark(b, "fixslot+block", ps);
    if (fork) {
      e = cudaEventRecord(b->ev_post, ps);
      if (e != cudaSuccess) { set_error("cudaEventRecord(post)", e); return 1; }
    }
    return record_blocks_event(b, ps);
  };
  if (fork) {
    e = cudaEventRecord(b->ev_fork, b->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->side, b->ev_fork, 0);
    if (e != cudaSuccess) { set_error("fork(post-processing stream)", e); return 1; }
    if (post_kernels()) return 1;
  }
  // (next to the forked post-processing the evaluation leaves room for a fix-up CTA per SM)
  if (enqueue_truepeak(b, t, fork, fork ? 5u : 0u)) return 1;
  if (!fork && post_kernels()) return 1;
  if (enqueue_queries(b, t, ps, parity)) return 1;
  if (fork) {
    e = cudaEventRecord(b->ev_join, b->side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->stream, b->ev_join, 0);
    if (e != cudaSuccess) { set_error("join(post-processing stream)", e); return 1; }
  }
  if (b->out_bytes)
    e = cudaMemcpyAsync(b->h_out[parity], b->d_out2[parity], b->out_bytes, cudaMemcpyDeviceToHost, b->stream);
  if (e != cudaSuccess) { set_error("cudaMemcpyAsync(results)", e); return 1; }
  step_mark(b, "read-back", b->stream);
  return 0;
}

// Everything of a pipelined step behind the block kernel, on `ps` (direct or being captured):
// queries (with the album exchange), the results' read-back.
static int enqueue_post_tail(lgb_batch* b, const DeviceTables& t, cudaStream_t ps, int parity) {
  cudaError_t e = cudaSuccess;
  if (enqueue_queries(b, t, ps, parity, b->pq, b->pq2)) return 1;
  if (b->peaks_off)
    e = cudaMemcpyAsync(b->h_out[parity], b->d_out2[parity], b->peaks_off, cudaMemcpyDeviceToHost, ps);
  if (e != cudaSuccess) { set_error("cudaMemcpyAsync(results)", e); return 1; }
  return 0;
}

// A pipelined step (see lgb_batch: pstream).  Caller's stream: peak cells zeroed, sweep,
// true-peak evaluation -- nothing else, so the next run's sweep follows at once.  pstream:
// fix-up (direct launch), then blocks, queries and the results' read-back as a graph, then the
// peaks once the evaluation is through.
static int enqueue_step_pipelined(lgb_batch* b, int parity) {
  const Plan& p = b->plan;
  const DeviceTables t = b->tables(parity);
  b->nmarks = 0;
  b->mark_set = parity;
  b->ms = b->mstream;
  // behind everything the caller has enqueued so far (the producers of the PCM)
  cudaError_t e = cudaEventRecord(b->ev_in, b->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(b->mstream, b->ev_in, 0);
  if (e != cudaSuccess) { set_error("fork(pipelined sweep)", e); return 1; }
  if (b->mtrace) cudaEventRecord(b->tm0[parity], b->mstream);
  if (enqueue_sweeps(b, t, parity, true)) return 1;
  if (b->mtrace) cudaEventRecord(b->tm2[parity], b->mstream);
  // Fix-up and blocks start behind the sweep on a high-priority stream, next to the true-peak
  // evaluation, which leaves them room (5 CTAs per SM instead of 6: a fix-up CTA fits next to
  // them) -- they are through before the evaluation is.  On the SMs a sweep under way leaves
  // free the fix-up alone took 100 us, and with an album exchange behind it the post-processing
  // of a run then lasted as long as the step itself.  Everything else (queries, exchange,
  // read-back) is latency-bound and small: it starts behind the evaluation, together with the
  // next sweep -- which has the higher priority, so its persistent CTAs (a whole SM each) are
  // placed first and these kernels get the SMs it leaves.  (A kernel of theirs that sits on an
  // SM when the evaluation ends makes a sweep CTA wait, and the sweep lasts as long as its
  // last CTA.)
  e = cudaEventRecord(b->ev_swept, b->mstream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(b->fstream, b->ev_swept, 0);
  if (e != cudaSuccess) { set_error("fork(pipelined post-processing)", e); return 1; }
  if (enqueue_truepeak(b, t, false, 5u)) return 1;
  if (b->mtrace) cudaEventRecord(b->tm1[parity], b->mstream);
  e = cudaEventRecord(b->ev_mdone[parity], b->mstream);
  // (what the caller enqueues from here on comes behind the kernels that read the PCM)
  if (e == cudaSuccess) e = cudaStreamWaitEvent(b->stream, b->ev_mdone[parity], 0);
  if (e != cudaSuccess) { set_error("cudaEventRecord(main part)", e); return 1; }
  {
    // The block and short-term energies are still being read by the queries of the run before
    // this one (unless every mirror has its own): only the small block kernel waits for them,
    // the fix-up does not -- the slot energies have no other reader than the block kernel.
    PostSizes zf{(uint32_t) p.tracks.size(), p.total_recs, p.total_slots, 0, 0};
    PostSizes zb{(uint32_t) p.tracks.size(), p.total_recs, 0, p.total_blocks, p.total_st};
    e = launch_post(t, zf, b->fstream, nullptr);
    if (e == cudaSuccess && !b->blocks_per_mirror)
      e = cudaStreamWaitEvent(b->fstream, b->ev_done[(parity + kMirrors - 1) % kMirrors], 0);
    if (e == cudaSuccess) e = launch_post(t, zb, b->fstream, nullptr);
    step_mark(b, "fix-up+blocks", b->fstream);
    if (e != cudaSuccess) { set_error("launch_post", e); return 1; }
    if (record_blocks_event(b, b->fstream)) return 1;
    e = cudaEventRecord(b->ev_fix, b->fstream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->pstream, b->ev_fix, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->pstream, b->ev_mdone[parity], 0);
    // (with one set of chunk records the next run's sweep, which rewrites them, has to wait)
    if (e == cudaSuccess && !b->recs_per_mirror) e = cudaStreamWaitEvent(b->mstream, b->ev_fix, 0);
    if (e != cudaSuccess) { set_error("join(fix-up)", e); return 1; }
  }
  if (!b->pgraph[parity] && !b->pgraph_off) {
    cudaGraph_t g = nullptr;
    e = cudaStreamBeginCapture(b->pstream, cudaStreamCaptureModeRelaxed);
    if (e == cudaSuccess) {
      const int rc = enqueue_post_tail(b, t, b->pstream, parity);
      e = cudaStreamEndCapture(b->pstream, &g);
      if (rc != 0 && e == cudaSuccess) e = cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&b->pgraph[parity], g, 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess) {          // no capture on this stream: keep launching directly
      cudaGetLastError();
      b->pgraph[parity] = nullptr;
      b->pgraph_off = true;
    }
  }
  if (b->pgraph[parity]) {
    e = cudaGraphLaunch(b->pgraph[parity], b->pstream);
    if (e != cudaSuccess) { set_error("cudaGraphLaunch(post-processing)", e); return 1; }
  } else if (enqueue_post_tail(b, t, b->pstream, parity)) {
    return 1;
  }
  // the peak cells are final behind the evaluation
  e = cudaStreamWaitEvent(b->pstream, b->ev_mdone[parity], 0);
  if (e == cudaSuccess && b->out_bytes > b->peaks_off)
    e = cudaMemcpyAsync(b->h_out[parity] + b->peaks_off, b->d_peaks(parity), b->out_bytes - b->peaks_off,
                        cudaMemcpyDeviceToHost, b->pstream);
  step_mark(b, "read-back", b->pstream);
  // The cells for the next run on this mirror: it is enqueued only after this run has been
  // fetched (two runs in flight at most), i.e. after ev_done -- so its sweep needs no memset
  // in front of it.
  if (e == cudaSuccess)
    e = cudaMemsetAsync(b->d_peaks(parity), 0, (2 * p.total_peaks + p.groups.size() + 1) * sizeof(uint32_t),
                        b->pstream);
  if (e == cudaSuccess) e = cudaEventRecord(b->ev_done[parity], b->pstream);
  if (e != cudaSuccess) { set_error("cudaMemcpyAsync(peaks)", e); return 1; }
  b->cells_clean[parity] = true;
  b->post_in_flight = true;
  return 0;
}

extern "C" LG_EXPORT int lgb_batch_run(lgb_batch* b) {
  if (b->runs - b->fetched >= b->max_in_flight) {
    set_error(b->max_in_flight == 2u ? "lgb_batch_run: two runs are in flight already (fetch the older one first)"
                                     : "lgb_batch_run: too many runs in flight (fetch the oldest one first)");
    return 1;
  }
  const uint32_t k = b->runs++;
  const int parity = (int) (k % (uint32_t) kMirrors);
  if (k == 1 && !b->qstream && b->side) {
    // a batch that is run again gets a third stream: the two halves of its queries run
    // side by side from now on (a one-shot batch does not pay for the stream)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&b->qstream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_q0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_q1, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      if (b->qstream) { cudaStreamDestroy(b->qstream); b->qstream = nullptr; }
    }
  }
  if (k == 1 && b->pipeline && !b->pstream && b->side && b->qstream) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ok = cudaStreamCreateWithPriority(&b->pstream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->pq, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->pq2, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->mstream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->fstream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&b->ev_in, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&b->ev_swept, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&b->ev_pidle, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kMirrors && ok; ++i)
      ok = cudaEventCreateWithFlags(&b->ev_mdone[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      b->pipeline = false;
    } else {
      b->recs_per_mirror = true;
      for (int i = 1; i < kMirrors && b->recs_per_mirror; ++i)
        if (!dalloc(&b->d_recs_alt[i], b->plan.total_recs, b->stream)) {
          cudaGetLastError();        // no chunk records per mirror: the sweeps wait for the fix-up
          b->recs_per_mirror = false;
        }
    }
  }
  // timed runs (bench roofline leg), traced runs and the first run launch directly
  const bool direct = b->timing || b->graph_off || (b->trace && !b->trace_pipelined) || k < 1;
  // With an album exchange attached the runs are NOT pipelined by default: at two GPUs the
  // pipelined step ran hundreds of thousands of times (0.190 ms against 0.26 ms, profiles/
  // r02_bench_n2.json), but one run at two GPUs (with block buffers per mirror) and the one run
  // at eight ended in an exchange time-out -- a rank that did not arrive within 20 s -- during
  // the bench's single-step warm-up, and the cause was not found before the round's GPU time
  // ran out.  LOUDGAIN_B200_PIPELINE_EXCHANGE=1 pipelines them all the same.
  static const bool pipe_xchg = [] {
    const char* e = getenv("LOUDGAIN_B200_PIPELINE_EXCHANGE");
    return e && atoi(e) != 0;
  }();
  if (!direct && b->pipeline && b->pstream && b->mstream && b->fstream && (!b->xchg || pipe_xchg))
    return enqueue_step_pipelined(b, parity);
  if (b->post_in_flight) {
    // behind pipelined runs: this run's kernels rewrite what their post-processing reads
    if (cudaEventRecord(b->ev_pidle, b->pstream) != cudaSuccess ||
        cudaStreamWaitEvent(b->stream, b->ev_pidle, 0) != cudaSuccess) {
      set_error("lgb_batch_run: cannot order behind the pipelined runs");
      return 1;
    }
    b->post_in_flight = false;
  }
  auto done = [&](int rc) {
    if (rc == 0 && cudaEventRecord(b->ev_done[parity], b->stream) != cudaSuccess) {
      set_error("lgb_batch_run: cudaEventRecord failed");
      return 1;
    }
    return rc;
  };
  if (direct) return done(enqueue_step(b, parity));
  if (!b->graph[parity]) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeRelaxed);
    if (e == cudaSuccess) {
      const int rc = enqueue_step(b, parity);
      e = cudaStreamEndCapture(b->stream, &g);
      if (rc != 0 && e == cudaSuccess) e = cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&b->graph[parity], g, 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess) {          // capture not possible on this stream: keep launching directly
      cudaGetLastError();
      b->graph[parity] = nullptr;
      b->graph_off = true;
      return done(enqueue_step(b, parity));
    }
  }
  const cudaError_t e = cudaGraphLaunch(b->graph[parity], b->stream);
  if (e != cudaSuccess) { set_error("cudaGraphLaunch", e); return 1; }
  return done(0);
}

extern "C" LG_EXPORT int lgb_batch_set_max_in_flight(lgb_batch* b, uint32_t n) {
  if (n < 1u || n > (uint32_t) kMirrors) { set_error("lgb_batch_set_max_in_flight: 1 to 3 runs"); return 1; }
  b->max_in_flight = n;
  return 0;
}

extern "C" LG_EXPORT int lgb_batch_fetch(lgb_batch* b, lgb_result* track_results, lgb_result* album_results,
                               double* sample_peaks, double* true_peaks) {
  const Plan& p = b->plan;
  if (b->fetched == b->runs) { set_error("lgb_batch_fetch: no run to fetch"); return 1; }
  const int parity = (int) (b->fetched++ % (uint32_t) kMirrors);       // the oldest run that has not been fetched
  const cudaError_t e = cudaEventSynchronize(b->ev_done[parity]);
  if (e != cudaSuccess) { set_error("lgb_batch_fetch", e); return 1; }
  const QueryResult* h_results = reinterpret_cast<const QueryResult*>(b->h_out[parity]);
  const uint32_t* h_peaks = reinterpret_cast<const uint32_t*>(b->h_out[parity] + b->peaks_off);
  if (b->trace && b->marks2[parity].size() > 1 && b->runs > 3) {
    const auto& marks = b->marks2[parity];
    fprintf(stderr, "[lgb step%s%s]", getenv("RANK") ? " rank " : "", getenv("RANK") ? getenv("RANK") : "");
    float ms = 0.0f;
    // (two runs in flight: the other mirror's marks belong to the run after this one)
    if (!b->marks2[(parity + 1) % kMirrors].empty() &&
        cudaEventElapsedTime(&ms, marks[0].second, b->marks2[(parity + 1) % kMirrors][0].second) == cudaSuccess)
      fprintf(stderr, " next run's start %+.1f us;", 1e3f * ms);
    for (size_t i = 1; i < marks.size(); ++i) {
      if (cudaEventElapsedTime(&ms, marks[0].second, marks[i].second) != cudaSuccess) continue;
      fprintf(stderr, " %s +%.1f us;", marks[i].first, 1e3f * ms);
    }
    fprintf(stderr, "\n");
  }
  if (b->mtrace && b->fetched > 3 && b->post_in_flight) {
    float m = 0.0f, sw = 0.0f, gap = 0.0f;
    if (cudaEventElapsedTime(&m, b->tm0[parity], b->tm1[parity]) == cudaSuccess &&
        cudaEventElapsedTime(&sw, b->tm0[parity], b->tm2[parity]) == cudaSuccess &&
        cudaEventElapsedTime(&gap, b->tm1[parity], b->tm0[(parity + 1) % kMirrors]) == cudaSuccess)
      fprintf(stderr, "[lgb mtrace%s%s] sweep %.1f us, sweep + evaluation %.1f us, idle until the next run's %+.1f us\n",
              getenv("RANK") ? " rank " : "", getenv("RANK") ? getenv("RANK") : "", 1e3f * sw, 1e3f * m, 1e3f * gap);
  }
  if (b->xchg && *b->xchg->h_ctl) {
    set_error("lgb_batch_fetch: a rank of the album exchange did not arrive (timed out)");
    return 1;
  }
  if (b->timed_run_pending) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) {
      b->sweep_ms_total += ms;
      ++b->sweep_runs;
      if (cudaEventElapsedTime(&ms, b->ev1, b->ev2) == cudaSuccess) b->tp_ms_total += ms;
    }
    b->timed_run_pending = false;
  }
  const size_t nt = p.tracks.size();
  if (track_results) for (size_t i = 0; i < nt; ++i) to_result(h_results[i], track_results[i]);
  if (album_results) for (uint32_t a = 0; a < p.nalbums; ++a) to_result(h_results[nt + a], album_results[a]);
  if (sample_peaks || true_peaks) {
    for (size_t i = 0; i < nt; ++i) {
      const Track& tr = p.tracks[i];
      const double scale = tr.format == FMT_S16 ? 32768.0 : 1.0;
      for (uint32_t c = 0; c < tr.channels; ++c) {
        float sp, tp;
        memcpy(&sp, &h_peaks[2 * (tr.peak_base + c)], 4);
        memcpy(&tp, &h_peaks[2 * (tr.peak_base + c) + 1], 4);
        const double s = (double) sp / scale, t = (double) tp / scale;
        if (sample_peaks) sample_peaks[tr.peak_base + c] = s;
        if (true_peaks) true_peaks[tr.peak_base + c] = t > s ? t : s;
      }
    }
  }
  return 0;
}

extern "C" LG_EXPORT int lgb_batch_wait_blocks(lgb_batch* b, void* cuda_stream) {
  if (!b->ev_blocks || !b->side) {           // no fork: order after everything enqueued so far
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return 1;
    cudaEventRecord(ev, b->stream);
    const cudaError_t e = cudaStreamWaitEvent((cudaStream_t) cuda_stream, ev, 0);
    cudaEventDestroy(ev);
    return e == cudaSuccess ? 0 : 1;
  }
  const cudaError_t e = cudaStreamWaitEvent((cudaStream_t) cuda_stream, b->ev_blocks, 0);
  if (e != cudaSuccess) { set_error("lgb_batch_wait_blocks", e); return 1; }
  return 0;
}

#ifdef LG_PAIR_TRACE
// tuning builds only: the sweep's per-warp trace of group 0 (lg_pair.cu)
extern "C" LG_EXPORT uint64_t lgb_debug_trace(lgb_batch* b, uint64_t* out, uint64_t cap) {
  const SweepGroup& g = b->plan.groups[0];
  const uint64_t n = 2ull * g.nwarps;
  const uint64_t off = 2ull * g.nwarps * g.params.npairs * 32ull - n;
  cudaStreamSynchronize(b->stream);
  if (n <= cap) cudaMemcpy(out, b->d_tpq + 2 * g.mrec_base + off, n * 8, cudaMemcpyDeviceToHost);
  return n;
}
#endif

// True-peak candidates the last run's sweeps queued for evaluation (run-sweep
// groups; diagnostic: how much of the audio the screening could not rule out).
extern "C" LG_EXPORT uint64_t lgb_batch_truepeak_candidates(lgb_batch* b) {
  const Plan& p = b->plan;
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return 0;
  uint64_t total = 0;
  for (const SweepGroup& g : p.groups) {
    if (!g.run || !g.tpf) continue;
    const uint32_t grid = run_sweep_grid(g.params, b->sms);
    std::vector<uint32_t> h(grid);
    if (cudaMemcpy(h.data(), b->d_runcnt + (size_t) g.first_item * 32u, grid * sizeof(uint32_t),
                   cudaMemcpyDeviceToHost) != cudaSuccess)
      return 0;
    for (uint32_t c : h) total += c;
  }
  return total;
}

extern "C" LG_EXPORT uint64_t lgb_batch_total_samples(const lgb_batch* b) { return b->plan.total_samples; }
extern "C" LG_EXPORT uint64_t lgb_batch_peak_count(const lgb_batch* b) { return b->plan.total_peaks; }
extern "C" LG_EXPORT uint32_t lgb_batch_kernel_launches(const lgb_batch* b) {
  // a repeated batch launches the two halves of its queries separately (lgb_batch_run)
  return b->launches + ((b->qstream && !b->xchg && !b->plan.queries.empty()) ? 1u : 0u);
}
extern "C" LG_EXPORT uint32_t lgb_batch_sweep_launches(const lgb_batch* b) { return b->sweep_launches; }

extern "C" LG_EXPORT uint64_t lgb_batch_blocks(const lgb_batch* b, size_t track, int kind,
                                     const double** dev_ptr) {
  if (track >= b->plan.tracks.size()) { if (dev_ptr) *dev_ptr = nullptr; return 0; }
  const Track& tr = b->plan.tracks[track];
  // (buffers per mirror: those of the most recently enqueued run)
  const int m = b->blocks_per_mirror && b->runs ? (int) ((b->runs - 1u) % (uint32_t) kMirrors) : 0;
  const double* zb = m ? b->d_zblock_m[m] : b->d_zblock;
  const double* zs = m ? b->d_zst_m[m] : b->d_zst;
  const double* es = m ? b->d_eslot_m[m] : b->d_eslot;
  const double* p = kind == 0 ? zb + tr.block_base : kind == 1 ? zs + tr.st_base : es + tr.slot_base;
  if (dev_ptr) *dev_ptr = p;
  return kind == 0 ? tr.nblocks : kind == 1 ? tr.nst : tr.nslots;
}

extern "C" LG_EXPORT void lgb_batch_enable_timing(lgb_batch* b, int on) {
  if (on && !b->ev0) {
    cudaEventCreate(&b->ev0);
    cudaEventCreate(&b->ev1);
    cudaEventCreate(&b->ev2);
  }
  b->timing = on != 0;
  b->timed_run_pending = false;
  b->sweep_ms_total = 0.0;
  b->tp_ms_total = 0.0;
  b->sweep_runs = 0;
}

extern "C" LG_EXPORT double lgb_batch_sweep_ms(const lgb_batch* b) {
  return b->sweep_runs ? b->sweep_ms_total / (double) b->sweep_runs : 0.0;
}

extern "C" LG_EXPORT double lgb_batch_truepeak_ms(const lgb_batch* b) {
  return b->sweep_runs ? b->tp_ms_total / (double) b->sweep_runs : 0.0;
}

extern "C" LG_EXPORT int lgb_query_lists(const double* const* z, const uint32_t* nz,
                                         const double* const* st, const uint32_t* nst, size_t n,
                                         void* cuda_stream, lgb_result* out) {
  std::vector<BlockList> lists(n);
  for (size_t i = 0; i < n; ++i) lists[i] = BlockList{z[i], st[i], nz[i], nst[i]};
  QueryResult q;
  if (query_lists_sync(lists.data(), n, (cudaStream_t) cuda_stream, &q)) return 1;
  to_result(q, *out);
  return 0;
}

extern "C" LG_EXPORT int lgb_slots_query(const double* slots, uint64_t nslots, uint32_t s100,
                                         void* cuda_stream, lgb_result* out) {
  g_error.clear();
  if (!s100 || (nslots && !slots) || nslots > 0xfffffff0ull) {
    set_error("lgb_slots_query: invalid arguments");
    return 1;
  }
  cudaStream_t stream = (cudaStream_t) cuda_stream;
  const uint64_t nblocks = nslots >= 4 ? nslots - 3 : 0;
  const uint64_t nst = nslots >= 30 ? (nslots - 30) / 10 + 1 : 0;
  double *zblock = nullptr, *zst = nullptr;
  bool ok = dalloc(&zblock, nblocks, stream) && dalloc(&zst, nst, stream);
  QueryResult q;
  if (ok) {
    const cudaError_t e = launch_stream_blocks(slots, (int) s100, nblocks, nst, zblock, zst, stream);
    if (e != cudaSuccess) { set_error("launch_stream_blocks", e); ok = false; }
  }
  if (ok) {
    const BlockList bl{zblock, zst, (uint32_t) nblocks, (uint32_t) nst};
    ok = query_lists_sync(&bl, 1, stream, &q) == 0;
  }
  if (zblock) cudaFreeAsync(zblock, stream);
  if (zst) cudaFreeAsync(zst, stream);
  if (!ok) return 1;
  to_result(q, *out);
  return 0;
}

struct lgb_listquery {
  cudaStream_t stream = nullptr;
  BlockList* d_lists = nullptr;
  uint32_t* d_members = nullptr;
  Query* d_query = nullptr;
  QueryResult* d_result = nullptr;
  QueryResult* h_result = nullptr;   // pinned
  double abs_gate = 0.0;
  uint32_t cluster = 1;
};

extern "C" LG_EXPORT lgb_listquery* lgb_listquery_create(const double* const* z, const uint32_t* nz,
                                                         const double* const* st,
                                                         const uint32_t* nst, size_t n,
                                                         void* cuda_stream) {
  lgb_listquery* q = new lgb_listquery();
  q->stream = (cudaStream_t) cuda_stream;
  q->abs_gate = pow(10.0, (-70.0 + 0.691) / 10.0);
  std::vector<BlockList> hl(n);
  std::vector<uint32_t> hm(n);
  for (size_t i = 0; i < n; ++i) { hl[i] = BlockList{z[i], st[i], nz[i], nst[i]}; hm[i] = (uint32_t) i; }
  std::vector<Query> hq(1, Query{0, (uint32_t) n});
  {
    uint64_t blocks = 0;
    for (size_t i = 0; i < n; ++i) blocks += nz[i];
    q->cluster = query_cluster_size(blocks);
  }
  bool ok = upload(hl, &q->d_lists, q->stream) && upload(hm, &q->d_members, q->stream) &&
            upload(hq, &q->d_query, q->stream) && dalloc(&q->d_result, 1, q->stream);
  if (ok) {
    cudaError_t e = cudaMallocHost((void**) &q->h_result, sizeof(QueryResult));
    if (e == cudaSuccess) e = cudaStreamSynchronize(q->stream);
    if (e != cudaSuccess) { set_error("lgb_listquery_create", e); ok = false; }
  }
  if (!ok) { lgb_listquery_destroy(q); return nullptr; }
  return q;
}

extern "C" LG_EXPORT int lgb_listquery_run(lgb_listquery* q) {
  cudaError_t e = launch_queries(q->d_lists, q->d_query, q->d_members, 1, q->abs_gate, q->d_result,
                                 q->stream, q->cluster);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(q->h_result, q->d_result, sizeof(QueryResult), cudaMemcpyDeviceToHost,
                        q->stream);
  if (e != cudaSuccess) { set_error("lgb_listquery_run", e); return 1; }
  return 0;
}

extern "C" LG_EXPORT int lgb_listquery_fetch(lgb_listquery* q, lgb_result* out) {
  const cudaError_t e = cudaStreamSynchronize(q->stream);
  if (e != cudaSuccess) { set_error("lgb_listquery_fetch", e); return 1; }
  to_result(*q->h_result, *out);
  return 0;
}

extern "C" LG_EXPORT void lgb_listquery_destroy(lgb_listquery* q) {
  if (!q) return;
  void* const mem[] = {q->d_lists, q->d_members, q->d_query, q->d_result};
  for (void* m : mem) if (m) cudaFreeAsync(m, q->stream);
  if (q->h_result) cudaFreeHost(q->h_result);
  delete q;
}

// ---- album exchange across ranks ------------------------------------------------
extern "C" LG_EXPORT lgb_exchange* lgb_exchange_create(uint32_t world, uint32_t rank, uint32_t nalbums,
                                                       uint64_t st_capacity) {
  g_error.clear();
  if (world < 1 || world > kMaxWorld || rank >= world) {
    set_error("lgb_exchange_create: world must be 1..16 and rank < world");
    return nullptr;
  }
  lgb_exchange* x = new lgb_exchange();
  x->world = world; x->rank = rank; x->nalbums = nalbums;
  x->st_cap = st_capacity ? st_capacity : 1;
  const size_t bytes = xchg_region_bytes(world, nalbums, x->st_cap);
  const unsigned long long ctl0[8] = {1ull, 0, 0, 0, 0, 0, 0, 0};      // steps count from 1: flags start at 0
  cudaError_t e = cudaMalloc((void**) &x->region, bytes);
  if (e == cudaSuccess) e = cudaMemset(x->region, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**) &x->d_ctl, sizeof ctl0);
  if (e == cudaSuccess) e = cudaMemcpy(x->d_ctl, ctl0, sizeof ctl0, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMallocHost((void**) &x->h_ctl, sizeof(unsigned long long));
  if (e != cudaSuccess) {
    set_error("lgb_exchange_create", e);
    lgb_exchange_destroy(x);
    return nullptr;
  }
  *x->h_ctl = 0;
  x->peer[rank] = x->region;
  return x;
}

extern "C" LG_EXPORT int lgb_exchange_handle(lgb_exchange* x, void* out, size_t cap) {
  if (cap < sizeof(cudaIpcMemHandle_t)) { set_error("lgb_exchange_handle: 64 bytes are needed"); return 1; }
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, x->region);
  if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle", e); return 1; }
  memcpy(out, &h, sizeof h);
  return 0;
}

extern "C" LG_EXPORT int lgb_exchange_open(lgb_exchange* x, const void* handles) {
  for (uint32_t r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*) handles + (size_t) r * sizeof h, sizeof h);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle (peer access between the GPUs is required)", e); return 1; }
    x->peer[r] = (unsigned char*) p;
  }
  x->opened = true;
  return 0;
}

extern "C" LG_EXPORT void lgb_exchange_destroy(lgb_exchange* x) {
  if (!x) return;
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < x->world; ++r)
    if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->region) cudaFree(x->region);
  if (x->d_ctl) cudaFree(x->d_ctl);
  if (x->h_ctl) cudaFreeHost(x->h_ctl);
  delete x;
}

extern "C" LG_EXPORT uint64_t lgb_batch_album_shortterm_blocks(const lgb_batch* b) {
  const Plan& p = b->plan;
  uint64_t n = 0;
  for (const Track& tr : p.tracks) if (tr.album != LGB_NO_ALBUM) n += tr.nst;
  return n;
}

extern "C" LG_EXPORT int lgb_batch_attach_exchange(lgb_batch* b, lgb_exchange* x) {
  g_error.clear();
  const Plan& p = b->plan;
  if (x->nalbums != p.nalbums) { set_error("lgb_batch_attach_exchange: album counts differ"); return 1; }
  if (x->world > 1 && !x->opened) { set_error("lgb_batch_attach_exchange: lgb_exchange_open first"); return 1; }
  const size_t nt = p.tracks.size();
  std::vector<uint32_t> off(p.nalbums + 1, 0);
  for (uint32_t a = 0; a < p.nalbums; ++a) {
    const Query& q = p.queries[nt + a];
    uint32_t n = 0;
    for (uint32_t m = 0; m < q.count; ++m) n += p.tracks[p.members[q.first + m]].nst;
    off[a + 1] = off[a] + n;
  }
  if (off[p.nalbums] > x->st_cap) { set_error("lgb_batch_attach_exchange: st_capacity is too small"); return 1; }
  {
    // room for the union of an album's short-term energies: the largest local album, as
    // much again on every other rank and half of that on top; 26 000 doubles at most
    uint32_t largest = 0;
    for (uint32_t a = 0; a < p.nalbums; ++a) largest = std::max(largest, off[a + 1] - off[a]);
    const uint64_t want = (uint64_t) largest * x->world * 3u / 2u + 64u;
    b->xst_smem = (uint32_t) std::min<uint64_t>(want, 26000u);
    uint64_t most_blocks = 0;
    for (uint32_t a = 0; a < p.nalbums; ++a) {
      const Query& q = p.queries[nt + a];
      uint64_t nb = 0;
      for (uint32_t m = 0; m < q.count; ++m) nb += p.tracks[p.members[q.first + m]].nblocks;
      most_blocks = std::max(most_blocks, nb);
    }
    b->xcluster = query_cluster_size(most_blocks);
    if (const char* e = getenv("LOUDGAIN_B200_XCLUSTER")) b->xcluster = atoi(e) > 0 ? (uint32_t) atoi(e) : b->xcluster;   // tuning
  }
  if (cudaStreamSynchronize(b->stream) != cudaSuccess ||
      (b->pstream && cudaStreamSynchronize(b->pstream) != cudaSuccess)) {
    set_error("lgb_batch_attach_exchange: stream error");
    return 1;
  }
  if (b->d_xstoff) cudaFreeAsync(b->d_xstoff, b->stream);
  if (!upload(off, &b->d_xstoff, b->stream) || cudaStreamSynchronize(b->stream) != cudaSuccess) return 1;
  for (int k = 0; k < kMirrors; ++k) {
    if (b->graph[k]) { cudaGraphExecDestroy(b->graph[k]); b->graph[k] = nullptr; }
    if (b->pgraph[k]) { cudaGraphExecDestroy(b->pgraph[k]); b->pgraph[k] = nullptr; }
  }
  if (!b->qstream) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&b->qstream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_q0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_q1, cudaEventDisableTiming) != cudaSuccess) {
      set_error("lgb_batch_attach_exchange: stream creation failed");
      return 1;
    }
  }
  if (!b->q2stream) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&b->q2stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_pub, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_q2, cudaEventDisableTiming) != cudaSuccess) {
      set_error("lgb_batch_attach_exchange: stream creation failed");
      return 1;
    }
  }
  // (Off by default: at two GPUs the bench's long run of single steps then ended in an exchange
  // time-out -- a rank that did not arrive -- which was not tracked down this round;
  // LOUDGAIN_B200_BLOCKS_PER_MIRROR=1 turns it on.)
  static const bool want_blocks_per_mirror = [] {
    const char* e = getenv("LOUDGAIN_B200_BLOCKS_PER_MIRROR");
    return e && atoi(e) != 0;
  }();
  if (!b->blocks_per_mirror && want_blocks_per_mirror) {
    bool ok = true;
    for (int m = 1; m < kMirrors && ok; ++m) {
      ok = dalloc(&b->d_eslot_m[m], p.total_slots, b->stream) && dalloc(&b->d_zblock_m[m], p.total_blocks, b->stream) &&
           dalloc(&b->d_zst_m[m], p.total_st, b->stream);
      if (ok) {
        std::vector<BlockList> lists(p.tracks.size());
        for (size_t i = 0; i < p.tracks.size(); ++i) {
          const Track& tr = p.tracks[i];
          lists[i] = BlockList{b->d_zblock_m[m] + tr.block_base, b->d_zst_m[m] + tr.st_base, tr.nblocks, tr.nst};
        }
        ok = upload(lists, &b->d_lists_m[m], b->stream) && cudaStreamSynchronize(b->stream) == cudaSuccess;
      }
    }
    if (!ok) cudaGetLastError();        // (the fix-up then waits for the run before it)
    b->blocks_per_mirror = ok;
  }
  if (!b->xchg) b->launches += 4u + (nt ? 2u : 0u);       // publish, gate, range, finish, advance and the track
                                                          // queries as two launches, instead of the one query launch
  b->xchg = x;
  return 0;
}

extern "C" LG_EXPORT void lgb_batch_destroy(lgb_batch* b) {
  if (!b) return;
  if (b->ev0) { cudaEventDestroy(b->ev0); cudaEventDestroy(b->ev1); cudaEventDestroy(b->ev2); }
  for (auto& ms : b->marks2) for (auto& m : ms) cudaEventDestroy(m.second);
  void* const mem[] = {b->d_tracks, b->d_coefs, b->d_work, b->d_queries, b->d_members, b->d_lists,
                       b->d_recs, b->d_recs_alt[1], b->d_recs_alt[2], b->d_out2[0], b->d_out2[1], b->d_out2[2], b->d_mrec, b->d_tpq, b->d_tmaps, b->d_items, b->d_xi, b->d_runq, b->d_runcnt, b->d_eslot, b->d_zblock, b->d_zst, b->d_eslot_m[1], b->d_eslot_m[2], b->d_zblock_m[1], b->d_zblock_m[2],
                       b->d_zst_m[1], b->d_zst_m[2], b->d_lists_m[1], b->d_lists_m[2],
                       b->d_xstoff};
  cudaStreamSynchronize(b->stream);   // a run may still be writing the mirrors
  if (b->pstream) cudaStreamSynchronize(b->pstream);
  for (void* m : mem) if (m) cudaFreeAsync(m, b->stream);
  for (int k = 0; k < kMirrors; ++k) {
    if (b->graph[k]) cudaGraphExecDestroy(b->graph[k]);
    if (b->pgraph[k]) cudaGraphExecDestroy(b->pgraph[k]);
    if (b->ev_mdone[k]) cudaEventDestroy(b->ev_mdone[k]);
    if (b->ev_done[k]) cudaEventDestroy(b->ev_done[k]);
    if (b->h_out[k]) cudaFreeHost(b->h_out[k]);
  }
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_join) cudaEventDestroy(b->ev_join);
  if (b->ev_post) cudaEventDestroy(b->ev_post);
  if (b->ev_fix) cudaEventDestroy(b->ev_fix);
  if (b->ev_blocks) cudaEventDestroy(b->ev_blocks);
  if (b->side) cudaStreamDestroy(b->side);
  if (b->ev_swept) cudaEventDestroy(b->ev_swept);
  if (b->ev_pidle) cudaEventDestroy(b->ev_pidle);
  if (b->pstream) cudaStreamDestroy(b->pstream);
  if (b->pq) cudaStreamDestroy(b->pq);
  if (b->pq2) cudaStreamDestroy(b->pq2);
  if (b->mstream) cudaStreamDestroy(b->mstream);
  if (b->fstream) cudaStreamDestroy(b->fstream);
  if (b->ev_in) cudaEventDestroy(b->ev_in);
  if (b->ev_q0) cudaEventDestroy(b->ev_q0);
  if (b->ev_q1) cudaEventDestroy(b->ev_q1);
  if (b->qstream) cudaStreamDestroy(b->qstream);
  if (b->ev_q2) cudaEventDestroy(b->ev_q2);
  if (b->ev_pub) cudaEventDestroy(b->ev_pub);
  if (b->q2stream) cudaStreamDestroy(b->q2stream);
  if (b->ev_gfork) cudaEventDestroy(b->ev_gfork);
  for (int j = 0; j < lgb_batch::kGroupStreams; ++j) {
    if (b->ev_gjoin[j]) cudaEventDestroy(b->ev_gjoin[j]);
    if (b->gstream[j]) cudaStreamDestroy(b->gstream[j]);
  }
  delete b;
}
