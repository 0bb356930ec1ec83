// lg_ebur128.cu -- the drop-in ebur128_* C ABI (include/ebur128.h) over the
// B200 batch engine.
//
// Call pattern being served (/root/reference/src/scan.c):
//   ebur128_init per file (:203-207); ebur128_add_frames_short per decoded
//   frame with a buffer the caller frees right away (:448-456); after ALL
//   files are scanned, per-track queries (:294-307) and per-album queries
//   (:383-391), the latter repeated once per track; ebur128_destroy (:102).
//
// So add_frames only stages: host memcpy into a pinned staging buffer, then
// cudaMemcpyAsync into the state's device PCM.  States are independent, as in
// libebur128: different threads may feed different states at the same time
// (the memcpy runs outside the library lock; the lock is only taken to claim
// or flush a staging buffer), which is how the ingest reaches PCIe rate.  The first query measures every
// state that has unmeasured audio in ONE batch (sweep + fix-up + gating on the
// GPU); later queries read cached scalars, and *_multiple queries run only the
// small gating/range kernel over block lists that are already in HBM.
//
// There is no CPU measurement path: without a usable CUDA device ebur128_init
// fails (NULL) with a message on stderr.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ebur128.h"
#include "../../include/ebur128_b200.h"
#include "lg_batch.h"
#include "lg_design.h"
#include "lg_kernels.h"

namespace {

constexpr size_t kStageBytes = 4u << 20;   // each pinned staging buffer
constexpr int kStageSlots = 32;            // pool of staging buffers (allocated on first use),
                                           // shared by all states: bounds the feeding threads

struct BatchHolder {
  lgb_batch* b = nullptr;
  ~BatchHolder() { if (b) lgb_batch_destroy(b); }
};

// One run of audio with fixed (rate, channels, sample format).
struct Segment {
  unsigned channels = 0;
  unsigned long rate = 0;
  uint32_t format = LGB_FORMAT_S16;
  bool hist = false;              // the state was created with EBUR128_MODE_HISTOGRAM
  char* d_pcm = nullptr;
  size_t cap = 0, fill = 0;       // bytes
  uint64_t frames = 0;
  uint8_t wclass[lg::kMaxChannels] = {0};
  // Incremental measurement (release_pass): frames [0, base) are measured and their PCM
  // is gone; what is kept of them is one energy per 100 ms slot and the peaks.  d_pcm
  // starts at track frame pcm_first <= base (lead-in for the next span, window queries).
  uint64_t base = 0, pcm_first = 0;
  double* d_slots = nullptr;      // device: slot energies of [0, base), then room for the tail's
  size_t slots_cap = 0;           // doubles
  std::vector<double> sp_done, tp_done;
  double *d_z = nullptr, *d_st = nullptr;   // block lists over the whole slot list (base > 0 only)
  uint32_t nz = 0, nst = 0;
  // measurement cache
  bool measured = false;
  std::shared_ptr<BatchHolder> batch;
  size_t track = 0;
  lgb_result res{};
  std::vector<double> sp, tp;
};

// One pinned staging buffer.  A state owns a slot while it is filling it; a
// flushed slot stays "in flight" until its H2D copy has completed.
struct StageSlot {
  char* buf = nullptr;             // allocated on first use
  cudaEvent_t ev = nullptr;
  bool in_flight = false;
  unsigned long long seq = 0;      // order of the flushes (oldest copy completes first)
  ebur128_state* owner = nullptr;
  size_t fill = 0;
};

struct Context {
  std::mutex mu;
  bool ready = false, failed = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  StageSlot slots[kStageSlots];
  unsigned long long flush_seq = 0;
  size_t pcm_bytes = 0;            // device PCM held by all live states (sum of capacities)
  size_t pcm_peak_bytes = 0;       // high-water mark (lgb_dropin_pcm_bytes)
  unsigned long long releases = 0; // release passes so far
  std::vector<ebur128_state*> live;
  // cache of the last *_multiple query (loudgain repeats it per track)
  std::vector<std::pair<const void*, size_t>> multi_key;
  lg::QueryResult multi_res{};
  bool multi_valid = false;
};

Context g_ctx;

// Makes the library's device current (only needed before CUDA calls, i.e. when
// a staging buffer is flushed or something is measured -- not per add_frames).
bool ctx_device() {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == g_ctx.device) return true;
  return cudaSetDevice(g_ctx.device) == cudaSuccess;
}

bool ctx_init() {
  if (g_ctx.ready) return ctx_device();
  if (g_ctx.failed) return false;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fprintf(stderr, "libebur128 (B200): no usable CUDA device (%s); this library has no CPU path\n",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    g_ctx.failed = true;
    return false;
  }
  const char* env = getenv("LOUDGAIN_B200_DEVICE");
  g_ctx.device = env ? atoi(env) : 0;
  if (g_ctx.device < 0 || g_ctx.device >= n) g_ctx.device = 0;
  e = cudaSetDevice(g_ctx.device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking);
  for (int i = 0; i < kStageSlots && e == cudaSuccess; ++i)
    e = cudaEventCreateWithFlags(&g_ctx.slots[i].ev, cudaEventDisableTiming);
  if (e == cudaSuccess) {
    // keep freed device memory in the stream-ordered pool: states come and go
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, g_ctx.device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (e != cudaSuccess) {
    fprintf(stderr, "libebur128 (B200): CUDA initialisation failed: %s\n", cudaGetErrorString(e));
    g_ctx.failed = true;
    return false;
  }
  g_ctx.ready = true;
  return true;
}

}  // namespace

struct ebur128_state_internal {
  std::vector<Segment> segs;
  int chmap[lg::kMaxChannels];
  int slot = -1;                // staging slot this state is filling, or -1
  // Held by the feeding thread for the whole of an add_frames call (uncontended: one
  // feeder per state, as with libebur128).  Whoever measures (library lock held) only
  // try_locks it: a state that is being fed right now is left alone, its staging slot
  // is not taken away.  Lock order: state, then library.
  std::mutex feed;
  unsigned long window_ms = 400, history_ms = ULONG_MAX;
  std::vector<float> convert;   // host scratch for int / double input
  std::vector<float> convert16; // ... and for 16-bit input to a state that has gone float
  bool poisoned = false;        // a staging failure lost frames: every later query fails
  // frames of the last add_frames call within the current segment, and their
  // peaks once asked for (ebur128_prev_sample_peak / _prev_true_peak)
  uint64_t prev_first = 0, prev_count = 0;
  bool prev_valid = false;
  std::vector<double> prev_sp, prev_tp;
};

namespace {

void default_map(int* map, unsigned channels) {
  static const int six[6] = {EBUR128_LEFT, EBUR128_RIGHT, EBUR128_CENTER, EBUR128_UNUSED,
                             EBUR128_LEFT_SURROUND, EBUR128_RIGHT_SURROUND};
  if (channels == 4) {
    map[0] = EBUR128_LEFT; map[1] = EBUR128_RIGHT;
    map[2] = EBUR128_LEFT_SURROUND; map[3] = EBUR128_RIGHT_SURROUND;
  } else if (channels == 5) {
    map[0] = EBUR128_LEFT; map[1] = EBUR128_RIGHT; map[2] = EBUR128_CENTER;
    map[3] = EBUR128_LEFT_SURROUND; map[4] = EBUR128_RIGHT_SURROUND;
  } else {
    for (unsigned i = 0; i < channels; ++i) map[i] = i < 6 ? six[i] : EBUR128_UNUSED;
  }
}

void sync_weights(ebur128_state* st) {
  Segment& s = st->d->segs.back();
  for (unsigned c = 0; c < st->channels; ++c) s.wclass[c] = lg::weight_class_of_role(st->d->chmap[c]);
}

void open_segment(ebur128_state* st, uint32_t format) {
  Segment s;
  s.channels = st->channels;
  s.rate = st->samplerate;
  s.format = format;
  s.hist = (st->mode & EBUR128_MODE_HISTOGRAM) == EBUR128_MODE_HISTOGRAM;
  st->d->segs.push_back(std::move(s));
  sync_weights(st);
}

size_t sample_bytes(uint32_t format) { return format == LGB_FORMAT_S16 ? 2 : 4; }

void invalidate(Segment& s) {
  s.measured = false;
  s.batch.reset();
  g_ctx.multi_valid = false;
}

size_t pcm_budget() {
  const char* e = getenv("LOUDGAIN_B200_PCM_BUDGET_MB");
  const long long mb = e ? atoll(e) : 0;
  return (size_t) (mb > 0 ? mb : 32768) << 20;        // default: 32 GB of the 180
}

size_t round_up(size_t v, size_t q) { return (v + q - 1) / q * q; }

void account_pcm(long long delta) {
  g_ctx.pcm_bytes = (size_t) ((long long) g_ctx.pcm_bytes + delta);
  if (g_ctx.pcm_bytes > g_ctx.pcm_peak_bytes) g_ctx.pcm_peak_bytes = g_ctx.pcm_bytes;
}

bool release_pass(ebur128_state* held);
bool widen_to_float(ebur128_state* st);
// the state whose add_frames the calling thread is inside of (it holds that state's feed lock)
thread_local ebur128_state* tl_feeding = nullptr;

// Pushes the filled part of the state's staging slot to the device and gives
// the slot back to the ring.  (Library lock held; `held`: the calling thread holds
// the state's feed lock.)
bool flush_stage(ebur128_state* st, bool held = false, bool may_release = true) {
  ebur128_state_internal* d = st->d;
  if (d->slot < 0) return true;
  StageSlot& sl = g_ctx.slots[d->slot];
  d->slot = -1;
  sl.owner = nullptr;
  if (!sl.fill) return true;
  if (!ctx_device()) return false;
  Segment& s = d->segs.back();
  const size_t need = s.fill + sl.fill;
  if (need > s.cap) {
    // a state that has been released before only keeps a tail: it grows by what it needs
    const size_t ncap = round_up(s.base ? need : std::max<size_t>(std::max<size_t>(s.cap * 2, need), 8u << 20),
                                 1u << 20);
    char* np = nullptr;
    if (cudaMallocAsync((void**) &np, ncap, g_ctx.stream) != cudaSuccess) return false;
    if (s.fill &&
        cudaMemcpyAsync(np, s.d_pcm, s.fill, cudaMemcpyDeviceToDevice, g_ctx.stream) != cudaSuccess)
      return false;
    if (s.d_pcm) cudaFreeAsync(s.d_pcm, g_ctx.stream);
    s.d_pcm = np;
    account_pcm((long long) ncap - (long long) s.cap);
    s.cap = ncap;
  }
  if (cudaMemcpyAsync(s.d_pcm + s.fill, sl.buf, sl.fill, cudaMemcpyHostToDevice, g_ctx.stream) !=
      cudaSuccess)
    return false;
  cudaEventRecord(sl.ev, g_ctx.stream);
  sl.in_flight = true;
  sl.seq = ++g_ctx.flush_seq;
  s.fill += sl.fill;
  sl.fill = 0;
  // the PCM of all live states has outgrown its budget: measure what is complete, keep
  // the 100 ms energies, give the PCM back
  if (may_release && g_ctx.pcm_bytes > pcm_budget()) return release_pass(held ? st : nullptr);
  return true;
}

// Finds a staging slot for `st` (library lock held): a free one; else one
// whose copy has completed; else one that an idle state left partly filled
// (its bytes are flushed first); else the oldest copy in flight is waited for.
bool acquire_slot(ebur128_state* st) {
  int pick = -1;
  for (int i = 0; i < kStageSlots && pick < 0; ++i)
    if (!g_ctx.slots[i].owner && !g_ctx.slots[i].in_flight && g_ctx.slots[i].buf) pick = i;
  for (int i = 0; i < kStageSlots && pick < 0; ++i)
    if (!g_ctx.slots[i].owner && !g_ctx.slots[i].in_flight) pick = i;        // not allocated yet
  if (pick < 0 && !ctx_device()) return false;
  for (int i = 0; i < kStageSlots && pick < 0; ++i) {
    StageSlot& sl = g_ctx.slots[i];
    if (!sl.owner && sl.in_flight && cudaEventQuery(sl.ev) == cudaSuccess) { sl.in_flight = false; pick = i; }
  }
  if (pick < 0) {
    for (int i = 0; i < kStageSlots; ++i) {
      StageSlot& sl = g_ctx.slots[i];
      ebur128_state* idle = sl.owner;              // (flush_stage clears sl.owner)
      if (idle && idle != st && idle->d->feed.try_lock()) {
        const bool ok = flush_stage(idle, true);
        idle->d->feed.unlock();
        if (!ok) return false;
        break;
      }
    }
    unsigned long long oldest = ~0ull;
    for (int i = 0; i < kStageSlots; ++i) {
      const StageSlot& sl = g_ctx.slots[i];
      if (!sl.owner && sl.in_flight && sl.seq < oldest) { oldest = sl.seq; pick = i; }
    }
    if (pick < 0) return true;        // every slot is being filled by another feeder: the caller waits
    if (cudaEventSynchronize(g_ctx.slots[pick].ev) != cudaSuccess) return false;
    g_ctx.slots[pick].in_flight = false;
  }
  StageSlot& sl = g_ctx.slots[pick];
  if (!sl.buf) {
    if (!ctx_device() || cudaMallocHost((void**) &sl.buf, kStageBytes) != cudaSuccess) {
      sl.buf = nullptr;
      return false;
    }
  }
  sl.owner = st;
  sl.fill = 0;
  st->d->slot = pick;
  return true;
}

// memcpy into a staging buffer with non-temporal stores: the buffer is written
// once and next read by the DMA engine, so pulling its lines into the cache
// first (what a plain memcpy of a few KB does) only costs memory bandwidth --
// with several scanner threads the host DRAM, not PCIe, is the ingest limit.
void stage_copy(char* dst, const char* src, size_t n) {
#if defined(__SSE2__)
  static const size_t nt_min = [] {
    const char* e = getenv("LOUDGAIN_B200_NT_MIN");      // tuning: smallest copy that uses streaming stores
    const long long v = e ? atoll(e) : 0;
    return (size_t) (v > 0 ? v : 256);
  }();
  if (n >= nt_min) {
    const size_t head = (16 - ((uintptr_t) dst & 15)) & 15;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t blocks = n / 64;
    while (blocks--) {
      const __m128i a = _mm_loadu_si128((const __m128i*) src);
      const __m128i b = _mm_loadu_si128((const __m128i*) (src + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*) (src + 32));
      const __m128i d = _mm_loadu_si128((const __m128i*) (src + 48));
      _mm_stream_si128((__m128i*) dst, a);
      _mm_stream_si128((__m128i*) (dst + 16), b);
      _mm_stream_si128((__m128i*) (dst + 32), c);
      _mm_stream_si128((__m128i*) (dst + 48), d);
      src += 64; dst += 64;
    }
    n &= 63;
    _mm_sfence();       // the stores must be globally visible before the DMA is enqueued
  }
#endif
  memcpy(dst, src, n);
}

// Copies `bytes` of caller PCM through the pinned staging pool.  The memcpy
// runs without the library lock; the lock is taken to claim or flush a slot.
bool stage_bytes(ebur128_state* st, const char* src, size_t bytes) {
  ebur128_state_internal* d = st->d;
  while (bytes) {
    while (d->slot < 0) {
      {
        std::lock_guard<std::mutex> lock(g_ctx.mu);
        if (!acquire_slot(st)) return false;
      }
      // all kStageSlots buffers are owned by other feeders that are inside add_frames
      // right now: wait for one of them to fill up and flush (no frames are dropped)
      if (d->slot < 0) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    StageSlot& sl = g_ctx.slots[d->slot];
    const size_t n = std::min(bytes, kStageBytes - sl.fill);
    stage_copy(sl.buf + sl.fill, src, n);
    sl.fill += n;
    src += n;
    bytes -= n;
    if (sl.fill == kStageBytes) {
      std::lock_guard<std::mutex> lock(g_ctx.mu);
      if (!flush_stage(st, true)) return false;
    }
  }
  return true;
}

int add_frames(ebur128_state* st, const void* src, size_t frames, uint32_t format) {
  if (!st || !st->d) return EBUR128_ERROR_NOMEM;
  ebur128_state_internal* d = st->d;
  // the common call takes no library-wide lock: scanner threads feeding different
  // states only meet when a 4 MB staging buffer is claimed or flushed
  std::lock_guard<std::mutex> feeding(d->feed);
  struct Mark { Mark(ebur128_state* s) { tl_feeding = s; } ~Mark() { tl_feeding = nullptr; } } mark(st);
  Segment* s = &d->segs.back();
  d->prev_first = s->frames;       // the "prev" peaks are those of this call's frames
  d->prev_count = frames;
  d->prev_valid = false;
  if (!frames) return EBUR128_SUCCESS;
  if (s->frames == 0) s->format = format;
  if (s->format != format) {
    // libebur128 converts every input type to double on the way in, so it can mix
    // them; here a state's PCM is kept in one format in HBM: a change of type
    // continues the state in float (the samples fed so far are widened once).
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    if (!widen_to_float(st)) return EBUR128_ERROR_NOMEM;
    s = &d->segs.back();
  }
  if (s->measured) {
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    invalidate(*s);
  }
  const void* in = src;
  if (format != s->format) {       // a float state fed 16-bit samples
    std::vector<float>& v = d->convert16;
    const size_t n = frames * st->channels;
    v.resize(n);
    const short* p = (const short*) src;
    for (size_t i = 0; i < n; ++i) v[i] = (float) p[i] / 32768.0f;
    in = v.data();
  }
  const size_t bytes = frames * st->channels * sample_bytes(s->format);
  if (!stage_bytes(st, (const char*) in, bytes)) {
    // frames of this call may be partly staged: the state cannot be trusted any more
    d->poisoned = true;
    return EBUR128_ERROR_NOMEM;
  }
  s->frames += frames;
  return EBUR128_SUCCESS;
}

// ---- incremental measurement -------------------------------------------------
// scan.c keeps every file's state alive until scan_deinit (scan.c:98-108), and
// libebur128 keeps only block energies per state.  Here a state holds PCM until
// it is measured; when all states together outgrow LOUDGAIN_B200_PCM_BUDGET_MB,
// release_pass measures what is complete of every state in ONE batch -- each
// state's new span [base, new_base) with the audio before it as lead-in (the
// filters and the interpolator warm up there: lgb_track.lead_in) -- keeps the
// span's 100 ms slot energies and peaks, and frees the PCM except for a tail
// (the next span's lead-in and the sliding windows).  A query later measures
// only the tail and forms the 400 ms / 3 s blocks over the whole slot list
// (lg_kernels.cu: stream_block_kernel), so every block is the same sum of slot
// energies as in a one-shot measurement.
//
// Spans start on multiples of `unit` = lcm(1 s, 16 bytes) frames: whole slots,
// whole short-term hops, and a 16-byte aligned first frame for the sweep.
uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }

struct SpanGeom {
  size_t fb;          // bytes per frame
  uint64_t s100, q, unit;
};

SpanGeom span_geom(const Segment& s) {
  SpanGeom g;
  g.fb = s.channels * sample_bytes(s.format);
  g.s100 = (s.rate + 5) / 10;
  g.q = 16 / gcd_u64(16, g.fb);
  const uint64_t sec = 10 * g.s100;
  g.unit = sec / gcd_u64(sec, g.q) * g.q;
  return g;
}

bool ensure_slots(Segment& s, size_t need) {
  if (need <= s.slots_cap) return true;
  const size_t ncap = std::max<size_t>(std::max<size_t>(2 * s.slots_cap, need), 1024);
  double* np = nullptr;
  if (cudaMallocAsync((void**) &np, ncap * sizeof(double), g_ctx.stream) != cudaSuccess) return false;
  const size_t have = (size_t) (s.base / ((s.rate + 5) / 10));
  if (have && cudaMemcpyAsync(np, s.d_slots, have * sizeof(double), cudaMemcpyDeviceToDevice, g_ctx.stream) !=
                  cudaSuccess)
    return false;
  if (s.d_slots) cudaFreeAsync(s.d_slots, g_ctx.stream);
  s.d_slots = np;
  s.slots_cap = ncap;
  return true;
}

// The track descriptor of the unmeasured part of `s` that is on the device up to
// frame `end`: frames [base, end) behind a lead-in of one unit (none at the start).
lgb_track span_track(const Segment& s, const SpanGeom& g, uint64_t end) {
  const uint64_t lead = s.base ? g.unit : 0;
  return lgb_track{s.d_pcm + (s.base - lead - s.pcm_first) * g.fb, end - s.base + lead, s.channels,
                   (uint32_t) s.rate, s.format, LGB_NO_ALBUM, s.wclass, lead,
                   s.hist ? LGB_TRACK_HISTOGRAM : 0u};
}

bool release_pass(ebur128_state* held) {
  struct Span { ebur128_state* st; Segment* s; SpanGeom g; uint64_t end, new_base; };
  std::vector<Span> spans;
  std::vector<ebur128_state*> locked;
  for (ebur128_state* st : g_ctx.live) {
    if (st != held && st != tl_feeding) {        // (those two feed locks are the calling thread's)
      if (!st->d->feed.try_lock()) continue;     // being fed right now: next time
      locked.push_back(st);
    }
    for (Segment& s : st->d->segs) {
      const SpanGeom g = span_geom(s);
      const uint64_t end = s.pcm_first + s.fill / g.fb;          // frames on the device
      const uint64_t new_base = end / g.unit * g.unit;
      if (new_base < s.base + 2 * g.unit) continue;              // too little to be worth a span
      spans.push_back(Span{st, &s, g, end, new_base});
    }
  }
  bool ok = true;
  if (!spans.empty()) {
    ++g_ctx.releases;
    std::vector<lgb_track> tracks(spans.size());
    for (size_t i = 0; i < spans.size(); ++i) tracks[i] = span_track(*spans[i].s, spans[i].g, spans[i].new_base);
    BatchHolder h;
    h.b = lgb_batch_create(tracks.data(), tracks.size(), 0, g_ctx.stream);
    std::vector<double> sp, tp;
    if (h.b) { sp.resize(lgb_batch_peak_count(h.b)); tp.resize(sp.size()); }
    ok = h.b && lgb_batch_run(h.b) == 0 && lgb_batch_fetch(h.b, nullptr, nullptr, sp.data(), tp.data()) == 0;
    if (!ok) fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
    size_t off = 0;
    for (size_t i = 0; ok && i < spans.size(); ++i) {
      Segment& s = *spans[i].s;
      const SpanGeom& g = spans[i].g;
      if (s.measured) invalidate(s);
      // slot energies of the span, behind the ones already kept
      const uint64_t lead = s.base ? g.unit : 0;
      const size_t have = (size_t) (s.base / g.s100), add = (size_t) ((spans[i].new_base - s.base) / g.s100);
      const double* dev = nullptr;
      const uint64_t n = lgb_batch_blocks(h.b, i, 2, &dev);
      ok = n == lead / g.s100 + add && ensure_slots(s, have + add + 64) &&
           cudaMemcpyAsync(s.d_slots + have, dev + lead / g.s100, add * sizeof(double), cudaMemcpyDeviceToDevice,
                           g_ctx.stream) == cudaSuccess;
      if (!ok) break;
      if (s.sp_done.empty()) { s.sp_done.assign(s.channels, 0.0); s.tp_done.assign(s.channels, 0.0); }
      for (unsigned c = 0; c < s.channels; ++c) {
        s.sp_done[c] = std::max(s.sp_done[c], sp[off + c]);
        s.tp_done[c] = std::max(s.tp_done[c], tp[off + c]);
      }
      off += s.channels;
      // keep the tail: the next span's lead-in and the longest sliding window
      const uint64_t win = ((uint64_t) spans[i].st->d->window_ms * s.rate + 999) / 1000;
      const uint64_t keep = g.unit + (win + g.s100 - 1) / g.s100 * g.s100 + g.s100;
      uint64_t first = spans[i].new_base > keep ? (spans[i].new_base - keep) / g.q * g.q : 0;
      if (first < s.pcm_first) first = s.pcm_first;
      const size_t tail = s.fill - (size_t) (first - s.pcm_first) * g.fb;   // may end inside a frame
      const size_t ncap = round_up(std::max<size_t>(tail, 1), 256u << 10);
      char* np = nullptr;
      ok = cudaMallocAsync((void**) &np, ncap, g_ctx.stream) == cudaSuccess &&
           cudaMemcpyAsync(np, s.d_pcm + (first - s.pcm_first) * g.fb, tail, cudaMemcpyDeviceToDevice,
                           g_ctx.stream) == cudaSuccess;
      if (!ok) break;
      cudaFreeAsync(s.d_pcm, g_ctx.stream);
      account_pcm((long long) ncap - (long long) s.cap);
      s.d_pcm = np; s.cap = ncap; s.fill = tail; s.pcm_first = first;
      s.base = spans[i].new_base;
    }
    // the batch's arrays are freed in stream order, behind the copies above
  }
  for (ebur128_state* st : locked) st->d->feed.unlock();
  return ok;
}

// Measures every live state that has audio the GPU has not looked at yet, in
// one batch.
bool measure_pending() {
  struct Item { Segment* s; SpanGeom g; long track; };
  std::vector<Item> todo;
  std::vector<lgb_track> tracks;
  for (ebur128_state* st : g_ctx.live) {
    // a state another thread is feeding right now is left alone (the caller
    // does not query a state while feeding it, as with libebur128)
    if (!st->d->feed.try_lock()) continue;
    const bool ok = flush_stage(st, true, false);     // (everything is about to be measured anyway)
    st->d->feed.unlock();
    if (!ok) return false;
    for (Segment& s : st->d->segs) {
      if (s.measured) continue;
      Item it{&s, span_geom(s), -1};
      if (s.frames > s.base || s.base == 0) {
        it.track = (long) tracks.size();
        tracks.push_back(span_track(s, it.g, s.frames));
      }
      todo.push_back(it);
    }
  }
  if (todo.empty()) return true;
  auto holder = std::make_shared<BatchHolder>();
  std::vector<lgb_result> res(tracks.size());
  std::vector<double> sp, tp;
  if (!tracks.empty()) {
    holder->b = lgb_batch_create(tracks.data(), tracks.size(), 0, g_ctx.stream);
    if (!holder->b) {
      fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
      return false;
    }
    sp.resize(lgb_batch_peak_count(holder->b));
    tp.resize(sp.size());
    if (lgb_batch_run(holder->b) || lgb_batch_fetch(holder->b, res.data(), nullptr, sp.data(), tp.data())) {
      fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
      return false;
    }
  }
  // states that were partly released: blocks over their whole slot list, one query each
  std::vector<lg::BlockList> lists;
  std::vector<Segment*> owners;
  size_t off = 0;
  for (Item& it : todo) {
    Segment& s = *it.s;
    s.sp.assign(s.channels, 0.0);
    s.tp.assign(s.channels, 0.0);
    if (it.track >= 0) {
      s.sp.assign(sp.begin() + off, sp.begin() + off + s.channels);
      s.tp.assign(tp.begin() + off, tp.begin() + off + s.channels);
      off += s.channels;
    }
    if (s.base == 0) {
      s.batch = holder;
      s.track = (size_t) it.track;
      s.res = res[it.track];
      s.measured = true;
      continue;
    }
    for (unsigned c = 0; c < s.channels; ++c) {
      s.sp[c] = std::max(s.sp[c], s.sp_done[c]);
      s.tp[c] = std::max(s.tp[c], s.tp_done[c]);
    }
    const SpanGeom& g = it.g;
    const size_t have = (size_t) (s.base / g.s100), add = (size_t) ((s.frames - s.base) / g.s100);
    bool ok = ensure_slots(s, have + add);
    if (ok && add) {
      const double* dev = nullptr;
      const uint64_t n = lgb_batch_blocks(holder->b, (size_t) it.track, 2, &dev);
      ok = n == g.unit / g.s100 + add &&
           cudaMemcpyAsync(s.d_slots + have, dev + g.unit / g.s100, add * sizeof(double), cudaMemcpyDeviceToDevice,
                           g_ctx.stream) == cudaSuccess;
    }
    const size_t nslots = have + add;
    s.nz = nslots >= 4 ? (uint32_t) (nslots - 3) : 0;
    s.nst = nslots >= 30 ? (uint32_t) ((nslots - 30) / 10 + 1) : 0;
    if (s.d_z) { cudaFreeAsync(s.d_z, g_ctx.stream); s.d_z = nullptr; }
    if (s.d_st) { cudaFreeAsync(s.d_st, g_ctx.stream); s.d_st = nullptr; }
    ok = ok && cudaMallocAsync((void**) &s.d_z, std::max<size_t>(s.nz, 1) * sizeof(double), g_ctx.stream) == cudaSuccess &&
         cudaMallocAsync((void**) &s.d_st, std::max<size_t>(s.nst, 1) * sizeof(double), g_ctx.stream) == cudaSuccess &&
         (!s.hist || lg::hist_table()) &&
         lg::launch_stream_blocks(s.d_slots, (int) g.s100, s.nz, s.nst, s.d_z, s.d_st, g_ctx.stream,
                                  s.hist ? lg::hist_table() : nullptr) == cudaSuccess;
    if (!ok) {
      fprintf(stderr, "libebur128 (B200): measuring a partly released state failed\n");
      return false;
    }
    lists.push_back(lg::BlockList{s.d_z, s.d_st, s.nz, s.nst});
    owners.push_back(&s);
  }
  if (!lists.empty()) {
    std::vector<lg::QueryResult> qr(lists.size());
    if (lg::query_each_sync(lists.data(), lists.size(), g_ctx.stream, qr.data())) {
      fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
      return false;
    }
    for (size_t i = 0; i < owners.size(); ++i) {
      Segment& s = *owners[i];
      const lg::QueryResult& q = qr[i];
      s.res = lgb_result{q.loudness, q.range, q.rel_thr, q.sum1, q.sum2, q.n1, q.n2, q.nst};
      s.batch.reset();
      s.measured = true;
    }
  }
  return true;
}

// The block lists of a measured segment: the measuring batch's, or the state's own
// when part of its audio was released before.
lg::BlockList segment_lists(const Segment& s) {
  lg::BlockList bl;
  if (s.base) return lg::BlockList{s.d_z, s.d_st, s.nz, s.nst};
  const double* p = nullptr;
  bl.nz = (uint32_t) lgb_batch_blocks(s.batch->b, s.track, 0, &p); bl.z = p;
  bl.nst = (uint32_t) lgb_batch_blocks(s.batch->b, s.track, 1, &p); bl.st = p;
  return bl;
}

// S16 -> float for a state that is fed a second sample type (library lock and the
// state's feed lock held).
bool widen_to_float(ebur128_state* st) {
  if (!ctx_device() || !flush_stage(st, true)) return false;
  Segment& s = st->d->segs.back();
  if (s.format == LGB_FORMAT_F32) return true;
  const size_t n = s.fill / 2;
  const size_t ncap = round_up(std::max<size_t>(n * 4, 8u << 20), 1u << 20);
  char* np = nullptr;
  if (cudaMallocAsync((void**) &np, ncap, g_ctx.stream) != cudaSuccess) return false;
  if (n && lg::launch_widen_s16(s.d_pcm, np, n, g_ctx.stream) != cudaSuccess) return false;
  if (s.d_pcm) cudaFreeAsync(s.d_pcm, g_ctx.stream);
  account_pcm((long long) ncap - (long long) s.cap);
  s.d_pcm = np; s.cap = ncap; s.fill = n * 4;
  s.format = LGB_FORMAT_F32;
  if (s.measured) invalidate(s);
  return true;
}

// ebur128_set_max_history: libebur128 stores only blocks above the absolute gate and
// keeps the newest `room` of them; `bl` is cut down to the stretch that holds them
// (the blocks below the gate inside it fail every gate anyway) and the rooms shrink
// by what was used.  A metering feature loudgain never touches: the lists are simply
// looked at on the host.
bool bound_lists(lg::BlockList& bl, uint64_t& room_z, uint64_t& room_st) {
  const double gate = pow(10.0, (-70.0 + 0.691) / 10.0);
  auto cut = [&](const double*& p, uint32_t& n, uint64_t& room) {
    std::vector<double> h(n);
    if (n && (cudaMemcpyAsync(h.data(), p, n * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream) != cudaSuccess ||
              cudaStreamSynchronize(g_ctx.stream) != cudaSuccess))
      return false;
    uint32_t first = n;
    while (first > 0 && room > 0) {
      --first;
      if (h[first] >= gate) --room;
    }
    p += first;
    n -= first;
    return true;
  };
  return cut(bl.z, bl.nz, room_z) && cut(bl.st, bl.nst, room_st);
}

// Gated loudness + range over the union of the given states' block lists.
int query_union(ebur128_state** sts, size_t n, lg::QueryResult* out) {
  if (!measure_pending()) return EBUR128_ERROR_NOMEM;
  std::vector<lg::BlockList> lists;
  std::vector<std::pair<const void*, size_t>> key;
  for (size_t i = 0; i < n; ++i) {
    if (!sts[i]) continue;
    if (sts[i]->d->poisoned) return EBUR128_ERROR_NOMEM;
    // ebur128_set_max_history: libebur128 then keeps only the newest history / 100 ms
    // gating blocks and history / 3000 ms short-term blocks of a state (never in
    // histogram mode, whose counters are unbounded)
    const unsigned long hist_ms = sts[i]->d->history_ms;
    uint64_t room_z = hist_ms == ULONG_MAX ? ~0ull : hist_ms / 100;
    uint64_t room_st = hist_ms == ULONG_MAX ? ~0ull : hist_ms / 3000;
    std::vector<Segment>& segs = sts[i]->d->segs;
    for (auto it = segs.rbegin(); it != segs.rend(); ++it) {       // newest audio first
      Segment& s = *it;
      if (!s.measured) {             // another thread is feeding it: nothing to report about it
        fprintf(stderr, "libebur128 (B200): a queried state is being fed by another thread\n");
        return EBUR128_ERROR_NOMEM;
      }
      lg::BlockList bl = segment_lists(s);
      if (!s.hist && hist_ms != ULONG_MAX && !bound_lists(bl, room_z, room_st)) return EBUR128_ERROR_NOMEM;
      lists.push_back(bl);
      key.emplace_back((const void*) bl.z, ((size_t) bl.nz << 32) ^ (size_t) bl.nst);
    }
  }
  if (g_ctx.multi_valid && key == g_ctx.multi_key) { *out = g_ctx.multi_res; return EBUR128_SUCCESS; }
  if (lg::query_lists_sync(lists.data(), lists.size(), g_ctx.stream, out)) {
    fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
    return EBUR128_ERROR_NOMEM;
  }
  g_ctx.multi_key = key;
  g_ctx.multi_res = *out;
  g_ctx.multi_valid = true;
  return EBUR128_SUCCESS;
}

int state_query(ebur128_state* st, lg::QueryResult* out) {
  if (st->d->poisoned) return EBUR128_ERROR_NOMEM;
  if (st->d->segs.size() == 1 && (st->d->history_ms == ULONG_MAX || st->d->segs[0].hist)) {
    if (!measure_pending() || !st->d->segs[0].measured) return EBUR128_ERROR_NOMEM;
    const lgb_result& r = st->d->segs[0].res;
    out->loudness = r.loudness; out->range = r.range; out->rel_thr = r.rel_threshold;
    out->sum1 = r.sum_abs; out->sum2 = r.sum_rel; out->n1 = r.n_abs; out->n2 = r.n_rel;
    out->nst = r.n_shortterm;
    return EBUR128_SUCCESS;
  }
  return query_union(&st, 1, out);
}

bool has_mode(const ebur128_state* st, int bits) { return (st->mode & bits) == bits; }

int peak_query(ebur128_state* st, unsigned ch, double* out, bool true_peak) {
  if (!has_mode(st, true_peak ? EBUR128_MODE_TRUE_PEAK : EBUR128_MODE_SAMPLE_PEAK))
    return EBUR128_ERROR_INVALID_MODE;
  if (ch >= st->channels) return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (st->d->poisoned || !ctx_init() || !measure_pending()) return EBUR128_ERROR_NOMEM;
  double m = 0.0;
  // Peaks survive a parameter change only while the channel count is kept.
  for (auto it = st->d->segs.rbegin(); it != st->d->segs.rend(); ++it) {
    if (it->channels != st->channels) break;
    if (!it->measured) return EBUR128_ERROR_NOMEM;
    m = std::max(m, true_peak ? it->tp[ch] : it->sp[ch]);
  }
  *out = m;
  return EBUR128_SUCCESS;
}

}  // namespace

// ------------------------------------------------------------------ the ABI

// Device PCM held by the drop-in layer (diagnostic): current bytes, high-water mark
// and the number of release passes so far.
extern "C" LG_EXPORT void lgb_dropin_pcm_bytes(uint64_t* now, uint64_t* peak, uint64_t* releases) {
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (now) *now = g_ctx.pcm_bytes;
  if (peak) *peak = g_ctx.pcm_peak_bytes;
  if (releases) *releases = g_ctx.releases;
  g_ctx.pcm_peak_bytes = g_ctx.pcm_bytes;          // the high-water mark restarts with every reading
}

extern "C" LG_EXPORT void ebur128_get_version(int* major, int* minor, int* patch) {
  *major = EBUR128_VERSION_MAJOR;
  *minor = EBUR128_VERSION_MINOR;
  *patch = EBUR128_VERSION_PATCH;
}

extern "C" LG_EXPORT ebur128_state* ebur128_init(unsigned int channels, unsigned long samplerate, int mode) {
  if (channels == 0 || channels > (unsigned) lg::kMaxChannels || samplerate < 16 ||
      samplerate > 2822400)
    return NULL;
  if (!(mode & EBUR128_MODE_M)) return NULL;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return NULL;
  ebur128_state* st = (ebur128_state*) calloc(1, sizeof(ebur128_state));
  if (!st) return NULL;
  st->d = new (std::nothrow) ebur128_state_internal();
  if (!st->d) { free(st); return NULL; }
  st->mode = mode;
  st->channels = channels;
  st->samplerate = samplerate;
  st->d->window_ms = has_mode(st, EBUR128_MODE_S) ? 3000 : 400;
  default_map(st->d->chmap, channels);
  open_segment(st, LGB_FORMAT_S16);
  g_ctx.live.push_back(st);
  return st;
}

extern "C" LG_EXPORT void ebur128_destroy(ebur128_state** stp) {
  if (!stp || !*stp) return;
  ebur128_state* st = *stp;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (g_ctx.ready) ctx_device();
  g_ctx.live.erase(std::remove(g_ctx.live.begin(), g_ctx.live.end(), st), g_ctx.live.end());
  g_ctx.multi_valid = false;
  if (st->d) {
    if (st->d->slot >= 0) {           // drop bytes that were staged but never flushed
      g_ctx.slots[st->d->slot].owner = nullptr;
      g_ctx.slots[st->d->slot].fill = 0;
    }
    for (Segment& s : st->d->segs) {
      // stream-ordered: anything still reading the PCM was enqueued before
      if (s.d_pcm) cudaFreeAsync(s.d_pcm, g_ctx.stream);
      if (s.d_slots) cudaFreeAsync(s.d_slots, g_ctx.stream);
      if (s.d_z) cudaFreeAsync(s.d_z, g_ctx.stream);
      if (s.d_st) cudaFreeAsync(s.d_st, g_ctx.stream);
      account_pcm(-(long long) s.cap);
      s.batch.reset();
    }
    delete st->d;
  }
  free(st);
  *stp = NULL;
}

extern "C" LG_EXPORT int ebur128_set_channel(ebur128_state* st, unsigned int ch, int value) {
  if (ch >= st->channels) return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  if (value == EBUR128_DUAL_MONO && (st->channels != 1 || ch != 0)) {
    fprintf(stderr, "EBUR128_DUAL_MONO only works with mono files!\n");
    return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  }
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  st->d->chmap[ch] = value;
  sync_weights(st);
  invalidate(st->d->segs.back());
  return EBUR128_SUCCESS;
}

extern "C" LG_EXPORT int ebur128_change_parameters(ebur128_state* st, unsigned int channels,
                                         unsigned long samplerate) {
  if (channels == 0 || channels > (unsigned) lg::kMaxChannels || samplerate < 16 ||
      samplerate > 2822400)
    return EBUR128_ERROR_NOMEM;
  if (channels == st->channels && samplerate == st->samplerate) return EBUR128_ERROR_NO_CHANGE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  if (!flush_stage(st)) return EBUR128_ERROR_NOMEM;
  const bool map_reset = channels != st->channels;
  st->channels = channels;
  st->samplerate = samplerate;
  if (map_reset) default_map(st->d->chmap, channels);
  // Filter, block schedule and interpolator restart; stored blocks are kept:
  // exactly a new track whose blocks join the same union.
  const uint32_t fmt = st->d->segs.back().format;
  if (st->d->segs.back().frames == 0) st->d->segs.pop_back();
  open_segment(st, fmt);
  st->d->prev_first = st->d->prev_count = 0;
  st->d->prev_valid = false;
  g_ctx.multi_valid = false;
  return EBUR128_SUCCESS;
}

extern "C" LG_EXPORT int ebur128_set_max_window(ebur128_state* st, unsigned long window) {
  if (has_mode(st, EBUR128_MODE_S) && window < 3000) window = 3000;
  else if (has_mode(st, EBUR128_MODE_M) && window < 400) window = 400;
  if (window == st->d->window_ms) return EBUR128_ERROR_NO_CHANGE;
  st->d->window_ms = window;
  return EBUR128_SUCCESS;
}

extern "C" LG_EXPORT int ebur128_set_max_history(ebur128_state* st, unsigned long history) {
  if (has_mode(st, EBUR128_MODE_LRA) && history < 3000) history = 3000;
  else if (has_mode(st, EBUR128_MODE_M) && history < 400) history = 400;
  if (history == st->d->history_ms) return EBUR128_ERROR_NO_CHANGE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  st->d->history_ms = history;
  g_ctx.multi_valid = false;
  return EBUR128_SUCCESS;
}

extern "C" LG_EXPORT int ebur128_add_frames_short(ebur128_state* st, const short* src, size_t frames) {
  return add_frames(st, src, frames, LGB_FORMAT_S16);
}

extern "C" LG_EXPORT int ebur128_add_frames_float(ebur128_state* st, const float* src, size_t frames) {
  return add_frames(st, src, frames, LGB_FORMAT_F32);
}

// int and double input are narrowed to float on the host (format shim only;
// all measurement stays on the GPU).
extern "C" LG_EXPORT int ebur128_add_frames_int(ebur128_state* st, const int* src, size_t frames) {
  if (!st || !st->d) return EBUR128_ERROR_NOMEM;
  std::vector<float>& v = st->d->convert;
  const size_t n = frames * st->channels;
  v.resize(n);
  for (size_t i = 0; i < n; ++i) v[i] = (float) ((double) src[i] / 2147483648.0);
  return add_frames(st, v.data(), frames, LGB_FORMAT_F32);
}

extern "C" LG_EXPORT int ebur128_add_frames_double(ebur128_state* st, const double* src, size_t frames) {
  if (!st || !st->d) return EBUR128_ERROR_NOMEM;
  std::vector<float>& v = st->d->convert;
  const size_t n = frames * st->channels;
  v.resize(n);
  for (size_t i = 0; i < n; ++i) v[i] = (float) src[i];
  return add_frames(st, v.data(), frames, LGB_FORMAT_F32);
}

extern "C" LG_EXPORT int ebur128_loudness_global(ebur128_state* st, double* out) {
  if (!has_mode(st, EBUR128_MODE_I)) return EBUR128_ERROR_INVALID_MODE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  lg::QueryResult r;
  const int rc = state_query(st, &r);
  if (rc == EBUR128_SUCCESS) *out = r.loudness;
  return rc;
}

extern "C" LG_EXPORT int ebur128_loudness_global_multiple(ebur128_state** sts, size_t size, double* out) {
  for (size_t i = 0; i < size; ++i)
    if (sts[i] && !has_mode(sts[i], EBUR128_MODE_I)) return EBUR128_ERROR_INVALID_MODE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  lg::QueryResult r;
  const int rc = query_union(sts, size, &r);
  if (rc == EBUR128_SUCCESS) *out = r.loudness;
  return rc;
}

extern "C" LG_EXPORT int ebur128_loudness_range(ebur128_state* st, double* out) {
  if (!has_mode(st, EBUR128_MODE_LRA)) return EBUR128_ERROR_INVALID_MODE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  lg::QueryResult r;
  const int rc = state_query(st, &r);
  if (rc == EBUR128_SUCCESS) *out = r.range;
  return rc;
}

extern "C" LG_EXPORT int ebur128_loudness_range_multiple(ebur128_state** sts, size_t size, double* out) {
  for (size_t i = 0; i < size; ++i)
    if (sts[i] && !has_mode(sts[i], EBUR128_MODE_LRA)) return EBUR128_ERROR_INVALID_MODE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  lg::QueryResult r;
  const int rc = query_union(sts, size, &r);
  if (rc == EBUR128_SUCCESS) *out = r.range;
  return rc;
}

extern "C" LG_EXPORT int ebur128_relative_threshold(ebur128_state* st, double* out) {
  if (!has_mode(st, EBUR128_MODE_I)) return EBUR128_ERROR_INVALID_MODE;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  lg::QueryResult r;
  const int rc = state_query(st, &r);
  if (rc != EBUR128_SUCCESS) return rc;
  *out = r.n1 ? 10.0 * log10(r.rel_thr) - 0.691 : -70.0;
  return EBUR128_SUCCESS;
}

extern "C" LG_EXPORT int ebur128_sample_peak(ebur128_state* st, unsigned int ch, double* out) {
  return peak_query(st, ch, out, false);
}

extern "C" LG_EXPORT int ebur128_true_peak(ebur128_state* st, unsigned int ch, double* out) {
  return peak_query(st, ch, out, true);
}

// ---- sliding-window queries -------------------------------------------------
// ebur128_loudness_momentary / _shortterm / _window: the loudness of the last
// 400 ms / 3 s / `window` ms fed so far (frames before the start of the audio
// count as silence).  loudgain never calls them; they are served from the same
// kernels: the tail of the state's PCM (the window plus one second of lead-in,
// right-aligned behind zeros when the audio is shorter) is measured as one
// track with `lead_in` set, and the window's 100 ms slot energies are summed.
// Windows must therefore be whole multiples of 100 ms (all three standard
// ones are).
namespace {

int window_query(ebur128_state* st, size_t nframes, double* out) {
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (st->d->poisoned || !ctx_init()) return EBUR128_ERROR_NOMEM;
  const size_t s100 = (st->samplerate + 5) / 10;
  const size_t ring = (st->d->window_ms * st->samplerate + 999) / 1000;
  if (!nframes || nframes % s100 || nframes > ((ring + s100 - 1) / s100) * s100)
    return EBUR128_ERROR_INVALID_MODE;
  if (!flush_stage(st)) return EBUR128_ERROR_NOMEM;
  Segment& s = st->d->segs.back();          // a parameter change restarts the window
  const size_t fb = s.channels * sample_bytes(s.format);
  const size_t lead = 10 * s100, total = lead + nframes;
  // (a partly released state keeps a tail that covers the window it was created with)
  const size_t have = std::min<size_t>(s.frames - s.pcm_first, total);
  char* tmp = nullptr;
  if (cudaMallocAsync((void**) &tmp, total * fb, g_ctx.stream) != cudaSuccess) return EBUR128_ERROR_NOMEM;
  bool ok = cudaMemsetAsync(tmp, 0, (total - have) * fb, g_ctx.stream) == cudaSuccess;
  if (ok && have)
    ok = cudaMemcpyAsync(tmp + (total - have) * fb, s.d_pcm + (s.frames - s.pcm_first - have) * fb, have * fb,
                         cudaMemcpyDeviceToDevice, g_ctx.stream) == cudaSuccess;
  double energy = 0.0;
  if (ok) {
    lgb_track t{tmp, total, s.channels, (uint32_t) s.rate, s.format, LGB_NO_ALBUM, s.wclass, lead};
    BatchHolder h;
    h.b = lgb_batch_create(&t, 1, 0, g_ctx.stream);
    ok = h.b && lgb_batch_run(h.b) == 0 && lgb_batch_fetch(h.b, nullptr, nullptr, nullptr, nullptr) == 0;
    if (ok) {
      const double* dev = nullptr;
      const uint64_t nslots = lgb_batch_blocks(h.b, 0, 2, &dev);
      std::vector<double> slots(nslots);
      ok = nslots == total / s100 &&
           cudaMemcpyAsync(slots.data(), dev, nslots * sizeof(double), cudaMemcpyDeviceToHost,
                           g_ctx.stream) == cudaSuccess &&
           cudaStreamSynchronize(g_ctx.stream) == cudaSuccess;
      for (size_t i = lead / s100; ok && i < nslots; ++i) energy += slots[i];
    }
    if (!ok) fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
  }
  cudaFreeAsync(tmp, g_ctx.stream);
  if (!ok) return EBUR128_ERROR_NOMEM;
  energy /= (double) nframes;
  *out = energy <= 0.0 ? -HUGE_VAL : 10.0 * log10(energy) - 0.691;
  return EBUR128_SUCCESS;
}

}  // namespace

extern "C" LG_EXPORT int ebur128_loudness_momentary(ebur128_state* st, double* out) {
  return window_query(st, 4 * ((st->samplerate + 5) / 10), out);
}

extern "C" LG_EXPORT int ebur128_loudness_shortterm(ebur128_state* st, double* out) {
  if (!has_mode(st, EBUR128_MODE_S)) return EBUR128_ERROR_INVALID_MODE;
  return window_query(st, 30 * ((st->samplerate + 5) / 10), out);
}

extern "C" LG_EXPORT int ebur128_loudness_window(ebur128_state* st, unsigned long window, double* out) {
  return window_query(st, (size_t) (st->samplerate * window / 1000), out);
}

// ---- peaks of the last add_frames call ---------------------------------------
// ebur128_prev_sample_peak / _prev_true_peak: a live-metering feature loudgain
// never uses.  The frames of the last call are still in
// the state's device PCM, so the query runs one small kernel over that range
// (lg_kernels.cu: range_peak_kernel; the interpolator's history is the audio
// before it) and caches the result until the next add_frames.
namespace {

int prev_peak_query(ebur128_state* st, unsigned ch, double* out, bool true_peak) {
  if (!has_mode(st, true_peak ? EBUR128_MODE_TRUE_PEAK : EBUR128_MODE_SAMPLE_PEAK))
    return EBUR128_ERROR_INVALID_MODE;
  if (ch >= st->channels) return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (st->d->poisoned || !ctx_init()) return EBUR128_ERROR_NOMEM;
  ebur128_state_internal* d = st->d;
  if (!d->prev_valid) {
    if (!flush_stage(st)) return EBUR128_ERROR_NOMEM;
    Segment& s = d->segs.back();
    const unsigned C = s.channels;
    d->prev_sp.assign(C, 0.0);
    d->prev_tp.assign(C, 0.0);
    // frames of the call that a release pass has already given back are not looked at
    // again (only a single call longer than the kept tail can lose any)
    uint64_t pfirst = d->prev_first, pcount = d->prev_count;
    if (pfirst < s.pcm_first) {
      const uint64_t drop = std::min<uint64_t>(s.pcm_first - pfirst, pcount);
      pfirst += drop; pcount -= drop;
    }
    if (pcount) {
      uint32_t* dev = nullptr;
      std::vector<uint32_t> host(2 * C);
      bool ok = cudaMallocAsync((void**) &dev, 2 * C * sizeof(uint32_t), g_ctx.stream) == cudaSuccess;
      if (ok)
        ok = lg::launch_range_peaks(s.d_pcm, s.format, C, pfirst - s.pcm_first, pcount,
                                    lg::true_peak_factor(s.rate), dev, g_ctx.stream) == cudaSuccess &&
             cudaMemcpyAsync(host.data(), dev, 2 * C * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             g_ctx.stream) == cudaSuccess &&
             cudaStreamSynchronize(g_ctx.stream) == cudaSuccess;
      if (dev) cudaFreeAsync(dev, g_ctx.stream);
      if (!ok) {
        fprintf(stderr, "libebur128 (B200): prev-peak query failed: %s\n",
                cudaGetErrorString(cudaGetLastError()));
        return EBUR128_ERROR_NOMEM;
      }
      const double scale = s.format == LGB_FORMAT_S16 ? 32768.0 : 1.0;
      for (unsigned c = 0; c < C; ++c) {
        float sp, tp;
        memcpy(&sp, &host[2 * c], 4);
        memcpy(&tp, &host[2 * c + 1], 4);
        d->prev_sp[c] = (double) sp / scale;
        d->prev_tp[c] = (double) tp / scale;
      }
    }
    d->prev_valid = true;
  }
  // as ebur128_true_peak: never below the sample peak of the same frames
  *out = true_peak ? std::max(d->prev_tp[ch], d->prev_sp[ch]) : d->prev_sp[ch];
  return EBUR128_SUCCESS;
}

}  // namespace

extern "C" LG_EXPORT int ebur128_prev_sample_peak(ebur128_state* st, unsigned int ch, double* out) {
  return prev_peak_query(st, ch, out, false);
}

extern "C" LG_EXPORT int ebur128_prev_true_peak(ebur128_state* st, unsigned int ch, double* out) {
  return prev_peak_query(st, ch, out, true);
}
