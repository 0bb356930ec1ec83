/*
 * ebur128_b200.h -- batch extension of the B200 loudness library.
 *
 * The ebur128_* functions of ebur128.h are the drop-in surface that
 * /root/reference/src/scan.c binds to.  This header adds what the reference
 * API cannot express: measuring many tracks whose PCM is ALREADY resident in
 * HBM in one pass (album and library scans), which is
 * also what the ebur128_* layer itself calls at query time.  Plain C ABI: raw
 * device pointers and sizes, no CUDA or torch types in the signatures.
 *
 * One batch = a set of tracks + the queries over them: one per track
 * (ebur128_loudness_global / _range / peaks, scan.c:294-307) and one per
 * album (ebur128_loudness_global_multiple / _range_multiple, scan.c:383-391).
 */
#ifndef EBUR128_B200_H_
#define EBUR128_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGB_FORMAT_S16 0u      /* what scan.c:414 feeds: interleaved int16 */
#define LGB_FORMAT_F32 1u      /* ebur128_add_frames_float layout */
#define LGB_NO_ALBUM 0xffffffffu

typedef struct lgb_batch lgb_batch;

typedef struct {
  const void* pcm;             /* DEVICE pointer, interleaved [frames][channels], 16-byte aligned */
  uint64_t frames;
  uint32_t channels;           /* 1..64 */
  uint32_t samplerate;
  uint32_t format;             /* LGB_FORMAT_* */
  uint32_t album;              /* album index < nalbums, or LGB_NO_ALBUM */
  const uint8_t* weight_class; /* optional HOST array [channels]: 0 unused, 1 -> 1.0,
                                  2 -> 1.41, 3 -> 2.0; NULL = default channel map */
  uint64_t lead_in;            /* frames at the start that are context only: the tail of the
                                  preceding audio when the track is one time segment of a
                                  longer stream (lgb_slots_query).  They warm the filters and
                                  the interpolator; no true-peak output is taken inside them.
                                  Must be a whole number of 100 ms slots.  0 = a whole track. */
  uint32_t flags;              /* LGB_TRACK_* */
} lgb_track;

/* EBUR128_MODE_HISTOGRAM semantics for this track: its 400 ms and 3 s block energies
 * are replaced by the centre energies of libebur128's 0.1 LU histogram bins (blocks
 * below -70 LUFS are dropped) before any gate, sum or percentile sees them. */
#define LGB_TRACK_HISTOGRAM 1u

typedef struct {
  double loudness;             /* LUFS; -HUGE_VAL if nothing passed the gates */
  double range;                /* LU */
  double rel_threshold;        /* relative gate, as block energy */
  double sum_abs, sum_rel;     /* block-energy sums behind the two gates */
  uint64_t n_abs, n_rel;       /* block counts behind the two gates */
  uint64_t n_shortterm;        /* short-term blocks above the absolute gate */
} lgb_result;

/* Last error message of the calling thread ("" if none). */
const char* lgb_last_error(void);

/* Plans a batch on the current CUDA device and allocates its workspace.
 * `cuda_stream` is a cudaStream_t (NULL = default stream).  NULL on failure. */
lgb_batch* lgb_batch_create(const lgb_track* tracks, size_t ntracks, uint32_t nalbums,
                            void* cuda_stream);

/* Enqueues the whole measurement (sweep, fix-up, gating, range, read-back of the
 * scalars into pinned memory) on the batch's stream.  Asynchronous; 0 on success.
 * Up to two runs may be in flight (three: lgb_batch_set_max_in_flight): a caller that repeats
 * a batch can enqueue run k + 1 before it fetches run k, so that its own turn-around overlaps
 * the GPU.
 * From the second run of a batch on, the runs are pipelined: the sweep and the true-peak
 * evaluation (the kernels that read the PCM) run on a high-priority stream of the library,
 * ordered behind everything enqueued on the batch's stream so far; the fix-up, blocks,
 * queries and read-back of run k finish on other streams of the library while the sweep of
 * run k + 1 is under way.  Work enqueued on the batch's stream after this call is ordered
 * behind the PCM reads of the run, not behind its slot / block energies or results: those
 * are reached through lgb_batch_wait_blocks / lgb_batch_fetch.
 * (LOUDGAIN_B200_PIPELINE=0: every run completely on the batch's stream; that is also how a
 * batch with an album exchange attached runs unless LOUDGAIN_B200_PIPELINE_EXCHANGE=1.) */
int lgb_batch_run(lgb_batch* b);

/* How many runs may be in flight before the oldest has to be fetched: 1 to 3, 2 by default.
 * Three keep the GPU fed when the results of a run arrive late -- an album exchange between
 * several GPUs waits for the slowest rank -- and the host needs its time to turn around. */
int lgb_batch_set_max_in_flight(lgb_batch* b, uint32_t n);

/* Waits for the OLDEST run that has not been fetched yet and copies its results to
 * the host (an error if there is none).  Any pointer may be
 * NULL.  Peaks are linear amplitudes laid out track after track, one value
 * per channel; true_peaks already folds in the sample peak, as
 * ebur128_true_peak does. */
int lgb_batch_fetch(lgb_batch* b, lgb_result* track_results, lgb_result* album_results,
                    double* sample_peaks, double* true_peaks);

/* Plan facts. */
uint64_t lgb_batch_total_samples(const lgb_batch* b);   /* frames*channels over all tracks */
uint64_t lgb_batch_peak_count(const lgb_batch* b);      /* sum of channels */
uint32_t lgb_batch_kernel_launches(const lgb_batch* b); /* kernels one run launches */
uint32_t lgb_batch_sweep_launches(const lgb_batch* b);  /* of which sweep kernels */
/* True-peak candidates the last run's sweeps queued for the evaluation pass
 * (stereo tracks; diagnostic: the 24-frame pairs the screening against the
 * channel's running peak could not rule out, two per pair at most).  Waits for
 * the run. */
uint64_t lgb_batch_truepeak_candidates(lgb_batch* b);

/* Device views for diagnostics and multi-GPU merges (valid after run + sync):
 * kind 0 = 400 ms gating blocks, 1 = 3 s short-term blocks, 2 = 100 ms slot
 * energies of `track`.  Returns the element count and stores the DEVICE
 * pointer (doubles). */
uint64_t lgb_batch_blocks(const lgb_batch* b, size_t track, int kind, const double** dev_ptr);

/* Makes `cuda_stream` wait until the block lists (and slot energies) of the most
 * recently enqueued run are complete -- earlier than the end of the run, which
 * still has its true-peak pass and queries in flight.  For consumers of
 * lgb_batch_blocks on another stream (the multi-GPU album merge).  0 on success. */
int lgb_batch_wait_blocks(lgb_batch* b, void* cuda_stream);

/* Kernel timing for benchmarks: when enabled, every run brackets its sweep
 * launches with CUDA events on the batch's stream; fetch accumulates them.
 * lgb_batch_sweep_ms returns the mean sweep time per run in milliseconds over
 * the runs fetched since timing was (re-)enabled, 0 if none. */
void lgb_batch_enable_timing(lgb_batch* b, int on);
double lgb_batch_sweep_ms(const lgb_batch* b);
/* ... and the mean time of the true-peak pass that follows the sweep. */
double lgb_batch_truepeak_ms(const lgb_batch* b);

void lgb_batch_destroy(lgb_batch* b);

/* Gated loudness + range over the union of `n` block lists that live in
 * device memory (e.g. lists all-gathered from other GPUs over NCCL).
 * z[i] / st[i] are DEVICE pointers to doubles.  Synchronous.  0 on success. */
int lgb_query_lists(const double* const* z, const uint32_t* nz, const double* const* st,
                    const uint32_t* nst, size_t n, void* cuda_stream, lgb_result* out);

/* Gated loudness + range of ONE stream from its 100 ms slot energies (kind 2 of
 * lgb_batch_blocks), e.g. the slot lists of the time segments of a long stream
 * measured on several GPUs, concatenated in time order after dropping each
 * segment's lead-in slots: forms the 400 ms gating blocks (hop 100 ms) and the
 * 3 s short-term blocks (hop 1 s) over the whole list, then gates.
 * `slots` is a DEVICE pointer; s100 = frames per slot = (rate + 5) / 10.
 * Synchronous.  0 on success. */
int lgb_slots_query(const double* slots, uint64_t nslots, uint32_t s100, void* cuda_stream,
                    lgb_result* out);

/* Persistent form of lgb_query_lists for a fixed set of device buffers (a
 * pre-allocated all-gather target that is refilled every step): tables are
 * uploaded once, lgb_listquery_run only launches the kernel and an async
 * read-back on the stream, lgb_listquery_fetch waits for it. */
typedef struct lgb_listquery lgb_listquery;
lgb_listquery* lgb_listquery_create(const double* const* z, const uint32_t* nz,
                                    const double* const* st, const uint32_t* nst, size_t n,
                                    void* cuda_stream);
int lgb_listquery_run(lgb_listquery* q);
int lgb_listquery_fetch(lgb_listquery* q, lgb_result* out);
void lgb_listquery_destroy(lgb_listquery* q);

/* Device PCM the drop-in ebur128_* layer holds for all live states: current bytes,
 * high-water mark, and how many release passes have run.  States keep their PCM
 * in HBM until they are measured; when the total outgrows
 * LOUDGAIN_B200_PCM_BUDGET_MB (default 32768) the complete part of every state is
 * measured in one batch, its 100 ms energies and peaks are kept and the PCM is
 * freed (libebur128 itself keeps only block energies per state; scan.c:98-108
 * keeps every state alive until scan_deinit).  Any pointer may be NULL.  Reading
 * restarts the high-water mark at the current value. */
void lgb_dropin_pcm_bytes(uint64_t* now, uint64_t* peak, uint64_t* releases);

/* ---- albums whose tracks were measured on several GPUs --------------------
 * ebur128_loudness_global_multiple / _range_multiple (scan.c:383-391) over
 * states that live on different ranks (one process per GPU; bin/rgbpm2:150-175
 * is the reference's process-per-album model).  Every rank creates an exchange
 * region in its own HBM, the 64-byte handles are passed around by the host
 * (any transport; torch.distributed in engine.py), and every rank opens the
 * others' regions (CUDA IPC: the GPUs must have peer access, i.e. NVLink /
 * NVSwitch).  A batch with an exchange attached answers its ALBUM queries
 * together with the other ranks: album indices are global (the same nalbums
 * on every rank; a rank may hold no track of an album), per-track results stay
 * local.  Inside a run each rank reduces its own gating blocks; its kernels
 * store (sum, count) pairs and the short-term energies straight into the
 * peers' regions and wait on per-rank flags -- no host round trip, the whole
 * step stays one CUDA graph.  Every rank must run the batch the same number
 * of times; a rank that never arrives makes lgb_batch_fetch fail after 20 s.
 *
 * st_capacity: the largest lgb_batch_album_shortterm_blocks() of any rank. */
typedef struct lgb_exchange lgb_exchange;
lgb_exchange* lgb_exchange_create(uint32_t world, uint32_t rank, uint32_t nalbums,
                                  uint64_t st_capacity);
/* Writes this rank's 64-byte handle. 0 on success. */
int lgb_exchange_handle(lgb_exchange* x, void* out, size_t cap);
/* handles: world x 64 bytes in rank order (the own entry is ignored). */
int lgb_exchange_open(lgb_exchange* x, const void* handles);
void lgb_exchange_destroy(lgb_exchange* x);
/* 3 s short-term blocks of all tracks of the batch that belong to an album. */
uint64_t lgb_batch_album_shortterm_blocks(const lgb_batch* b);
/* Before the batch is run (or between runs, after a fetch). 0 on success. */
int lgb_batch_attach_exchange(lgb_batch* b, lgb_exchange* x);

/* Longest-processing-time-first assignment of `n` work items (tracks; cost =
 * frames x channels) to `world` ranks: rank_out[i] = rank of item i.  Host
 * arithmetic only.  Returns the largest rank load. */
uint64_t lgb_lpt_assign(const uint64_t* cost, size_t n, uint32_t world, uint32_t* rank_out);

/* ---- scan.c-shaped host driver -------------------------------------------
 * Replays the reference scanner's call sequence against the ebur128_* ABI of
 * this library for PCM held in HOST memory: scan_file's per-frame
 * ebur128_add_frames_short loop (scan.c:225-256,448), scan_get_track_result
 * (scan.c:275-330) and scan_set_album_result (scan.c:380-405), including the
 * reference's habit of re-querying the album once per track. */
typedef struct {
  const void* pcm;            /* HOST pointer, interleaved */
  uint64_t frames;
  uint32_t channels;
  uint32_t samplerate;
  uint32_t format;            /* LGB_FORMAT_*, optionally | LGB_HOST_CODEC_OPUS */
} lgb_host_track;

/* The file is an Opus stream: its gains are relative to -23 LUFS, i.e. the pre-gain is
 * lowered by 5 dB for the track (scan.c:309-311) and for an album that holds such a
 * track (scan.c:394-398), and loudness_reference follows (scan.c:328). */
#define LGB_HOST_CODEC_OPUS 0x100u
#define LGB_HOST_FORMAT_MASK 0xffu

typedef struct {              /* mirrors scan_result (scan.h:35-53) */
  double track_gain, track_peak, track_loudness, track_loudness_range;
  double album_gain, album_peak, album_loudness, album_loudness_range;
  double loudness_reference;
} lgb_scan_result;

/* chunk_frames: frames per ebur128_add_frames call (a decoded AVFrame is
 * about 1k-4k frames); 0 = one call per track.  0 on success. */
int lgb_scan_host(const lgb_host_track* tracks, size_t ntracks, size_t chunk_frames,
                  int do_album, double pre_gain, lgb_scan_result* out);

/* The same with up to `nthreads` scanner threads, one file per thread at a time
 * (the reference parallelises over files and albums with one process each,
 * bin/rgbpm2:170; ebur128 states are independent, so here the scanners share
 * one library instance and the whole album is measured in one GPU batch).
 * Queries run after all files are scanned, as in loudgain.c:299-340. */
int lgb_scan_host_mt(const lgb_host_track* tracks, size_t ntracks, size_t chunk_frames,
                     int do_album, double pre_gain, unsigned nthreads, lgb_scan_result* out);

/* ---- what loudgain does with a scan result --------------
 * Clipping prevention of loudgain.c:323-379 (-k / -K n): the peak a track /
 * album would have after its gain is compared with the limit
 * `max_true_peak_db` (dBTP; loudgain's default for -k is -1.0); with `prevent`
 * the gains in `r` are lowered so that the new peak just meets the limit.
 * Host arithmetic only (no GPU involved).  0 on success. */
typedef struct {
  int will_clip;               /* a peak is above the limit and was not corrected */
  int track_clipped;           /* track gain was lowered (loudgain's tclip) */
  int album_clipped;           /* album gain was lowered (aclip) */
  int album_would_clip;        /* album peak above the limit, not corrected (the album row's flag) */
  double track_new_peak;       /* linear peak after the (corrected) track gain */
  double album_new_peak;
} lgb_clip_info;

int lgb_clip_prevention(lgb_scan_result* r, int do_album, int prevent, double max_true_peak_db,
                        lgb_clip_info* info);

/* One row of loudgain's tab-separated `-O` output (loudgain.c:586-612) for a
 * track (album_row = 0) or for the album (album_row = 1, name "Album"); `unit`
 * is "dB" or "LU".  Returns the length written (without the terminating NUL),
 * or the length needed if `cap` is too small. */
size_t lgb_format_tab_row(const char* name, const lgb_scan_result* r, const lgb_clip_info* info,
                          int album_row, const char* unit, char* buf, size_t cap);

/* One row of the old mp3gain-compatible list, loudgain -o (loudgain.c:566-585):
 * "name<TAB>0<TAB>gain %.2f<TAB>peak * 32768 %.6f<TAB>0<TAB>0". */
size_t lgb_format_old_row(const char* name, const lgb_scan_result* r, int album_row, char* buf,
                          size_t cap);

/* loudgain's default human-readable block for a track or for the album
 * (loudgain.c:613-649); `opus` adds the Q7.8 number to the gain line. */
size_t lgb_format_human(const char* name, const lgb_scan_result* r, const lgb_clip_info* info,
                        int album_row, int opus, const char* unit, char* buf, size_t cap);

/* The tag values loudgain would write for a scan result, as text at the
 * reference's tag precision (tag.cc:178-203: gains "%.2f <unit>", peaks
 * "%.6f", reference "%.2f LUFS"; Opus files carry Q7.8 integers instead,
 * tag.cc:442-445,474-480).  One "NAME=value" pair per line; album lines only
 * with do_album, range / reference lines only with `extended` (loudgain's
 * -s e / -s l); `opus` selects the R128_* form.  Actual tag writing (TagLib)
 * stays with the host application.  Returns the length written, or the
 * length needed if `cap` is too small. */
size_t lgb_format_tags(const lgb_scan_result* r, int do_album, int extended, int opus,
                       const char* unit, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif

#endif /* EBUR128_B200_H_ */
