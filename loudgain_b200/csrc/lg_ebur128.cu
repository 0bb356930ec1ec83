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
#include <memory>
#include <mutex>
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
  char* d_pcm = nullptr;
  size_t cap = 0, fill = 0;       // bytes
  uint64_t frames = 0;
  uint8_t wclass[lg::kMaxChannels] = {0};
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
  std::atomic<bool> busy{false};  // inside add_frames (its slot must not be taken away)
  unsigned long window_ms = 400, history_ms = ULONG_MAX;
  std::vector<float> convert;   // host scratch for int / double input
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
  st->d->segs.push_back(std::move(s));
  sync_weights(st);
}

size_t sample_bytes(uint32_t format) { return format == LGB_FORMAT_S16 ? 2 : 4; }

void invalidate(Segment& s) {
  s.measured = false;
  s.batch.reset();
  g_ctx.multi_valid = false;
}

// Pushes the filled part of the state's staging slot to the device and gives
// the slot back to the ring.
bool flush_stage(ebur128_state* st) {
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
    size_t ncap = std::max<size_t>(std::max<size_t>(s.cap * 2, need), 32u << 20);
    char* np = nullptr;
    if (cudaMallocAsync((void**) &np, ncap, g_ctx.stream) != cudaSuccess) return false;
    if (s.fill &&
        cudaMemcpyAsync(np, s.d_pcm, s.fill, cudaMemcpyDeviceToDevice, g_ctx.stream) != cudaSuccess)
      return false;
    if (s.d_pcm) cudaFreeAsync(s.d_pcm, g_ctx.stream);
    s.d_pcm = np;
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
      if (sl.owner && sl.owner != st && !sl.owner->d->busy.load(std::memory_order_acquire)) {
        if (!flush_stage(sl.owner)) return false;
        break;
      }
    }
    unsigned long long oldest = ~0ull;
    for (int i = 0; i < kStageSlots; ++i) {
      const StageSlot& sl = g_ctx.slots[i];
      if (!sl.owner && sl.in_flight && sl.seq < oldest) { oldest = sl.seq; pick = i; }
    }
    if (pick < 0) {
      fprintf(stderr, "libebur128 (B200): more than %d states are being fed at the same time\n",
              kStageSlots);
      return false;
    }
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
  if (n >= 256) {
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
    if (d->slot < 0) {
      std::lock_guard<std::mutex> lock(g_ctx.mu);
      if (!acquire_slot(st)) return false;
    }
    StageSlot& sl = g_ctx.slots[d->slot];
    const size_t n = std::min(bytes, kStageBytes - sl.fill);
    stage_copy(sl.buf + sl.fill, src, n);
    sl.fill += n;
    src += n;
    bytes -= n;
    if (sl.fill == kStageBytes) {
      std::lock_guard<std::mutex> lock(g_ctx.mu);
      if (!flush_stage(st)) return false;
    }
  }
  return true;
}

int add_frames(ebur128_state* st, const void* src, size_t frames, uint32_t format) {
  if (!st || !st->d) return EBUR128_ERROR_NOMEM;
  ebur128_state_internal* d = st->d;
  Segment* s = &d->segs.back();
  d->prev_first = s->frames;       // the "prev" peaks are those of this call's frames
  d->prev_count = frames;
  d->prev_valid = false;
  if (!frames) return EBUR128_SUCCESS;
  {
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    if (!g_ctx.ready) return EBUR128_ERROR_NOMEM;
    if (s->frames == 0) s->format = format;
    if (s->format != format) {
      // Mixing sample types on one state is not supported: the PCM of a state
      // is kept in one format in HBM.
      fprintf(stderr, "libebur128 (B200): mixing sample formats on one state is not supported\n");
      return EBUR128_ERROR_INVALID_MODE;
    }
    if (s->measured) invalidate(*s);
    d->busy.store(true, std::memory_order_relaxed);
  }
  const size_t bytes = frames * st->channels * sample_bytes(format);
  const bool ok = stage_bytes(st, (const char*) src, bytes);
  if (ok) s->frames += frames;
  d->busy.store(false, std::memory_order_release);
  return ok ? EBUR128_SUCCESS : EBUR128_ERROR_NOMEM;
}

// Measures every live state that has audio the GPU has not looked at yet, in
// one batch.
bool measure_pending() {
  std::vector<Segment*> todo;
  for (ebur128_state* st : g_ctx.live) {
    // a state another thread is feeding right now is left alone (the caller
    // does not query a state while feeding it, as with libebur128)
    if (st->d->busy.load(std::memory_order_acquire)) continue;
    if (!flush_stage(st)) return false;
    for (Segment& s : st->d->segs)
      if (!s.measured) todo.push_back(&s);
  }
  if (todo.empty()) return true;
  std::vector<lgb_track> tracks(todo.size());
  for (size_t i = 0; i < todo.size(); ++i) {
    Segment& s = *todo[i];
    tracks[i] = lgb_track{s.d_pcm, s.frames, s.channels, (uint32_t) s.rate, s.format,
                          LGB_NO_ALBUM, s.wclass};
  }
  auto holder = std::make_shared<BatchHolder>();
  holder->b = lgb_batch_create(tracks.data(), tracks.size(), 0, g_ctx.stream);
  if (!holder->b) {
    fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
    return false;
  }
  std::vector<lgb_result> res(todo.size());
  std::vector<double> sp(lgb_batch_peak_count(holder->b)), tp(sp.size());
  if (lgb_batch_run(holder->b) || lgb_batch_fetch(holder->b, res.data(), nullptr, sp.data(), tp.data())) {
    fprintf(stderr, "libebur128 (B200): %s\n", lgb_last_error());
    return false;
  }
  size_t off = 0;
  for (size_t i = 0; i < todo.size(); ++i) {
    Segment& s = *todo[i];
    s.measured = true;
    s.batch = holder;
    s.track = i;
    s.res = res[i];
    s.sp.assign(sp.begin() + off, sp.begin() + off + s.channels);
    s.tp.assign(tp.begin() + off, tp.begin() + off + s.channels);
    off += s.channels;
  }
  return true;
}

// Gated loudness + range over the union of the given states' block lists.
int query_union(ebur128_state** sts, size_t n, lg::QueryResult* out) {
  if (!measure_pending()) return EBUR128_ERROR_NOMEM;
  std::vector<lg::BlockList> lists;
  std::vector<std::pair<const void*, size_t>> key;
  for (size_t i = 0; i < n; ++i) {
    if (!sts[i]) continue;
    for (Segment& s : sts[i]->d->segs) {
      lg::BlockList bl;
      const double* p = nullptr;
      bl.nz = (uint32_t) lgb_batch_blocks(s.batch->b, s.track, 0, &p); bl.z = p;
      bl.nst = (uint32_t) lgb_batch_blocks(s.batch->b, s.track, 1, &p); bl.st = p;
      lists.push_back(bl);
      key.emplace_back((const void*) s.batch->b, s.track);
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
  if (st->d->segs.size() == 1) {
    if (!measure_pending()) return EBUR128_ERROR_NOMEM;
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
  if (!ctx_init() || !measure_pending()) return EBUR128_ERROR_NOMEM;
  double m = 0.0;
  // Peaks survive a parameter change only while the channel count is kept.
  for (auto it = st->d->segs.rbegin(); it != st->d->segs.rend(); ++it) {
    if (it->channels != st->channels) break;
    m = std::max(m, true_peak ? it->tp[ch] : it->sp[ch]);
  }
  *out = m;
  return EBUR128_SUCCESS;
}

}  // namespace

// ------------------------------------------------------------------ the ABI

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
  st->d->history_ms = history;
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
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  const size_t s100 = (st->samplerate + 5) / 10;
  const size_t ring = (st->d->window_ms * st->samplerate + 999) / 1000;
  if (!nframes || nframes % s100 || nframes > ((ring + s100 - 1) / s100) * s100)
    return EBUR128_ERROR_INVALID_MODE;
  if (!flush_stage(st)) return EBUR128_ERROR_NOMEM;
  Segment& s = st->d->segs.back();          // a parameter change restarts the window
  const size_t fb = s.channels * sample_bytes(s.format);
  const size_t lead = 10 * s100, total = lead + nframes;
  const size_t have = std::min<size_t>(s.frames, total);
  char* tmp = nullptr;
  if (cudaMallocAsync((void**) &tmp, total * fb, g_ctx.stream) != cudaSuccess) return EBUR128_ERROR_NOMEM;
  bool ok = cudaMemsetAsync(tmp, 0, (total - have) * fb, g_ctx.stream) == cudaSuccess;
  if (ok && have)
    ok = cudaMemcpyAsync(tmp + (total - have) * fb, s.d_pcm + (s.frames - have) * fb, have * fb,
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
  if (!ctx_init()) return EBUR128_ERROR_NOMEM;
  ebur128_state_internal* d = st->d;
  if (!d->prev_valid) {
    if (!flush_stage(st)) return EBUR128_ERROR_NOMEM;
    Segment& s = d->segs.back();
    const unsigned C = s.channels;
    d->prev_sp.assign(C, 0.0);
    d->prev_tp.assign(C, 0.0);
    if (d->prev_count) {
      uint32_t* dev = nullptr;
      std::vector<uint32_t> host(2 * C);
      bool ok = cudaMallocAsync((void**) &dev, 2 * C * sizeof(uint32_t), g_ctx.stream) == cudaSuccess;
      if (ok)
        ok = lg::launch_range_peaks(s.d_pcm, s.format, C, d->prev_first, d->prev_count,
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
