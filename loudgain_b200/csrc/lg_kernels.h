// lg_kernels.h -- launch interface between the host layer (lg_batch.cu) and
// the kernels (lg_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lg_common.h"

namespace lg {

constexpr int kSweepThreads = 128;

// Device-resident tables of one batch.
struct DeviceTables {
  const Track* tracks;
  const CoefSet* coefs;
  const WarpWork* work;
  const Query* queries;
  const uint32_t* members;
  const BlockList* lists; // [ntracks]
  const cplx* xi_table;   // run sweep: mode-sum reference per chunk position (lg_design.h: make_run_coefs)
  ChunkRec* recs;       // [total_recs]
  uint32_t* peaks;      // [total_peaks][2]: sample peak, true peak (float bits)
  double* eslot;        // [total_slots]
  double* zblock;       // [total_blocks]
  double* zst;          // [total_st]
  QueryResult* results; // [nqueries]
  const double* hist_tab; // hist_table() when a track has Track::flags bit 0, else unused
};

struct PostSizes {
  uint32_t ntracks;
  uint64_t total_recs, total_slots, total_blocks, total_st;
};

// One sweep launch; `p` carries the group's constants and device pointers.
cudaError_t launch_sweep(const SweepParams& p, uint32_t format, int tpf, cudaStream_t stream);
// True-peak pass of the same group (no-op for rates without an interpolator);
// must follow the group's sweep on the stream.
// `hold` (optional): an event the heavy part of the pass waits for on `stream`
// (lg_batch.cu lets the small post-processing kernels go first).
cudaError_t launch_truepeak(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                            cudaStream_t stream, cudaEvent_t hold = nullptr);
// The same two passes for groups with an even channel count (lg_pair.cu:
// packed FP32, one lane per chunk and channel pair; SweepParams::packed).
cudaError_t launch_sweep_pair(const SweepParams& p, uint32_t format, int tpf, cudaStream_t stream);
cudaError_t launch_truepeak_pair(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                                 cudaStream_t stream, cudaEvent_t hold = nullptr);
// The run sweep and its true-peak pass (lg_run.cu: stereo tracks, one lane per
// run of chunks, persistent one-warp CTAs).
cudaError_t launch_sweep_run(const SweepParams& p, uint32_t format, int tpf, uint32_t sms, cudaStream_t stream);
uint32_t run_sweep_grid(const SweepParams& p, uint32_t sms);     // persistent CTAs launch_sweep_run starts
cudaError_t launch_truepeak_run(const SweepParams& p, uint32_t format, int tpf, uint32_t sms,
                                cudaStream_t stream, cudaEvent_t hold = nullptr, uint32_t cta_cap = 0);
// `fixed` (optional): recorded behind the fix-up kernel, before the block kernel.
cudaError_t launch_post(const DeviceTables& t, const PostSizes& z, cudaStream_t stream,
                        cudaEvent_t fixed = nullptr);
// 400 ms / 3 s blocks of one stream from its complete 100 ms slot list.
// hist_tab: hist_table() for EBUR128_MODE_HISTOGRAM semantics, nullptr for exact energies.
cudaError_t launch_stream_blocks(const double* eslot, int s100, uint64_t nblocks, uint64_t nst,
                                 double* zblock, double* zst, cudaStream_t stream,
                                 const double* hist_tab = nullptr);
// 1001 bin edges + 1000 bin centre energies of libebur128's histogram mode in the
// current device's memory (nullptr on failure).
const double* hist_table();
// Gated loudness + range for `nqueries` queries; all pointers are device memory.
// `cluster` = CTAs per query (thread-block cluster sharing the gating blocks of a
// query); query_cluster_size() picks it from the largest query's block count.
// `part`: 0 = the whole query, 1 = integrated loudness only, 2 = loudness range only (the
// halves touch different lists and different result fields and may run concurrently).
cudaError_t launch_queries(const BlockList* lists, const Query* queries, const uint32_t* members,
                           uint32_t nqueries, double abs_gate, QueryResult* results,
                           cudaStream_t stream, uint32_t cluster = 1, int part = 0);
uint32_t query_cluster_size(uint64_t max_gating_blocks);
// Album queries over tracks sharded across GPUs (lg_common.h: XchgParams).  publish: this
// rank's share to every peer; gate: relative threshold from all ranks' totals, this rank's
// sums behind it to every peer; range: results[first_query + album].range from the union of
// the short-term energies (needs only the publish phase: run it next to the gate on another
// stream; st_smem_doubles = energies it may stage in shared memory); finish: the loudness.
// `cluster` = CTAs that share an album's local gating blocks.
cudaError_t launch_exchange_publish(const BlockList* lists, const Query* queries, const uint32_t* members,
                                    double abs_gate, const XchgParams& x, uint32_t cluster, cudaStream_t stream);
cudaError_t launch_exchange_gate(const BlockList* lists, const Query* queries, const uint32_t* members,
                                 double abs_gate, const XchgParams& x, uint32_t cluster, cudaStream_t stream);
cudaError_t launch_exchange_range(double abs_gate, QueryResult* results, const XchgParams& x,
                                  uint32_t st_smem_doubles, cudaStream_t stream);
cudaError_t launch_exchange_finish(QueryResult* results, const XchgParams& x, cudaStream_t stream);
// ... and, behind every kernel of the step's exchange, the step counter moves on
cudaError_t launch_exchange_advance(const XchgParams& x, cudaStream_t stream);
// Sample peak and true peak (float bits, raw sample units) of track frames
// [first, first + count) per channel into out[2 * channels] (device memory).
cudaError_t launch_range_peaks(const void* pcm, uint32_t format, uint32_t channels, uint64_t first,
                               uint64_t count, int tpf, uint32_t* out, cudaStream_t stream);

// n 16-bit samples -> float, x / 32768 (device pointers).
cudaError_t launch_widen_s16(const void* in, void* out, size_t n, cudaStream_t stream);

}  // namespace lg
