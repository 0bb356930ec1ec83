// lg_scan.cu -- scan.c-shaped host driver over this library's own ebur128_*
// ABI (declared in include/ebur128_b200.h).  It exists so that the end-to-end
// path a loudgain user exercises -- host PCM in caller-owned buffers, one
// ebur128_add_frames_short per decoded frame, queries after all files -- can be
// run and timed without FFmpeg: /root/reference/src/scan.c:225-256,448 (feed
// loop), :275-330 (track result), :359-405 (album result).
#include <math.h>
#include <stdint.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <queue>
#include <thread>
#include <vector>

#include "../../include/ebur128.h"
#include "../../include/ebur128_b200.h"
#include "lg_batch.h"

namespace {

// scan.c:64
inline double lufs_to_rg(double l) { return -18.0 - l; }

double max_true_peak(ebur128_state* st) {
  double peak = 0.0;
  for (unsigned ch = 0; ch < st->channels; ++ch) {   // reads ->channels like scan.c:300
    double tmp;
    if (ebur128_true_peak(st, ch, &tmp) != EBUR128_SUCCESS) continue;
    if (tmp > peak) peak = tmp;
  }
  return peak;
}

}  // namespace

// scan_file (scan.c:110-273) for one track: init, then one add_frames call per
// `chunk_frames` frames.
static int scan_one(const lgb_host_track& t, size_t chunk_frames, ebur128_state** out) {
  *out = ebur128_init(t.channels, t.samplerate,
                      EBUR128_MODE_S | EBUR128_MODE_I | EBUR128_MODE_LRA |
                          EBUR128_MODE_SAMPLE_PEAK | EBUR128_MODE_TRUE_PEAK);
  if (!*out) return 1;
  const size_t step = chunk_frames ? chunk_frames : (t.frames ? t.frames : 1);
  const uint32_t format = t.format & LGB_HOST_FORMAT_MASK;
  const size_t fb = t.channels * (format == LGB_FORMAT_S16 ? 2u : 4u);
  for (uint64_t pos = 0; pos < t.frames; pos += step) {
    const size_t n = (size_t) (t.frames - pos < step ? t.frames - pos : step);
    const char* p = (const char*) t.pcm + pos * fb;
    const int e = format == LGB_FORMAT_S16
                      ? ebur128_add_frames_short(*out, (const short*) p, n)
                      : ebur128_add_frames_float(*out, (const float*) p, n);
    if (e != EBUR128_SUCCESS) return 2;
  }
  return 0;
}

extern "C" LG_EXPORT int lgb_scan_host(const lgb_host_track* tracks, size_t ntracks,
                                       size_t chunk_frames, int do_album, double pre_gain,
                                       lgb_scan_result* out) {
  return lgb_scan_host_mt(tracks, ntracks, chunk_frames, do_album, pre_gain, 1, out);
}

// The same with `nthreads` feeding threads, one file per thread at a time --
// the reference's own parallel model is one scanner per file/album
// (bin/rgbpm2:170, multiprocessing.Pool); here the scanners share one library
// instance, so that all files of the album are measured in one GPU batch.
extern "C" LG_EXPORT int lgb_scan_host_mt(const lgb_host_track* tracks, size_t ntracks,
                                          size_t chunk_frames, int do_album, double pre_gain,
                                          unsigned nthreads, lgb_scan_result* out) {
  std::vector<ebur128_state*> states(ntracks, nullptr);
  int rc = 0;
  const bool trace = getenv("LOUDGAIN_B200_TRACE") != nullptr;     // phase times on stderr
  const auto t_start = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  };
  // ---- scan_file for every file first (loudgain.c:299-305)
  if (nthreads <= 1 || ntracks <= 1) {
    for (size_t i = 0; i < ntracks && !rc; ++i) rc = scan_one(tracks[i], chunk_frames, &states[i]);
  } else {
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    auto worker = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= ntracks || err.load()) break;
        const int e = scan_one(tracks[i], chunk_frames, &states[i]);
        if (e) err.store(e);
      }
    };
    std::vector<std::thread> pool;
    const unsigned n = nthreads < ntracks ? nthreads : (unsigned) ntracks;
    for (unsigned k = 0; k < n; ++k) pool.emplace_back(worker);
    for (std::thread& th : pool) th.join();
    rc = err.load();
  }
  const double ms_feed = ms_since(t_start);
  const auto t_query = std::chrono::steady_clock::now();
  double ms_first = 0.0;
  // Opus gains are relative to -23 LUFS: scan.c:309-311 lowers the pre-gain by 5 dB for an
  // Opus track, scan.c:394-398 for an album that has one (loudgain refuses albums that
  // mix Opus with other codecs before it gets there)
  bool album_has_opus = false;
  for (size_t i = 0; i < ntracks; ++i) album_has_opus = album_has_opus || (tracks[i].format & LGB_HOST_CODEC_OPUS);
  const double user_pre_gain = pre_gain;
  // ---- results (loudgain.c:323-340): track, then album, per file
  for (size_t i = 0; i < ntracks && !rc; ++i) {
    pre_gain = user_pre_gain - ((tracks[i].format & LGB_HOST_CODEC_OPUS) ? 5.0 : 0.0);
    if (i == 1) ms_first = ms_since(t_query);
    lgb_scan_result& r = out[i];
    double global, range;
    if (ebur128_loudness_global(states[i], &global) != EBUR128_SUCCESS) global = 0.0;
    if (ebur128_loudness_range(states[i], &range) != EBUR128_SUCCESS) range = 0.0;
    r.track_gain = lufs_to_rg(global) + pre_gain;
    r.track_peak = max_true_peak(states[i]);
    r.track_loudness = global;
    r.track_loudness_range = range;
    r.album_gain = r.album_peak = r.album_loudness = r.album_loudness_range = 0.0;
    r.loudness_reference = lufs_to_rg(-pre_gain);
    if (do_album) {
      if (ebur128_loudness_global_multiple(states.data(), ntracks, &global) != EBUR128_SUCCESS)
        global = 0.0;
      if (ebur128_loudness_range_multiple(states.data(), ntracks, &range) != EBUR128_SUCCESS)
        range = 0.0;
      double apeak = 0.0;
      for (size_t j = 0; j < ntracks; ++j) {
        const double p = max_true_peak(states[j]);
        if (p > apeak) apeak = p;
      }
      r.album_gain = lufs_to_rg(global) + user_pre_gain - (album_has_opus ? 5.0 : 0.0);
      r.album_peak = apeak;
      r.album_loudness = global;
      r.album_loudness_range = range;
    }
  }
  const double ms_query = ms_since(t_query);
  const auto t_free = std::chrono::steady_clock::now();
  for (size_t i = 0; i < ntracks; ++i) ebur128_destroy(&states[i]);   // scan_deinit
  if (trace)
    fprintf(stderr, "lgb_scan_host: feed %.2f ms, queries %.2f ms (first track incl. the GPU batch "
            "%.2f ms), destroy %.2f ms\n", ms_feed, ms_query, ms_first, ms_since(t_free));
  return rc;
}

// ---- loudgain's use of a scan result ---------------------------------------

namespace {

// Gain (dB) applied to a linear peak; and the gain change that moves a peak
// from `from` to `to`.
inline double peak_after_gain(double gain_db, double peak) { return pow(10.0, gain_db / 20.0) * peak; }
inline double db_between(double from, double to) { return 20.0 * log10(from / to); }

}  // namespace

extern "C" LG_EXPORT int lgb_clip_prevention(lgb_scan_result* r, int do_album, int prevent,
                                             double max_true_peak_db, lgb_clip_info* info) {
  if (!r || !info) return 1;
  const double limit = pow(10.0, max_true_peak_db / 20.0);
  lgb_clip_info c = {0, 0, 0, 0, 0.0, 0.0};
  const double tpeak = peak_after_gain(r->track_gain, r->track_peak);
  const double apeak = do_album ? peak_after_gain(r->album_gain, r->album_peak) : 0.0;
  const bool t_over = tpeak > limit, a_over = do_album && apeak > limit;
  c.track_new_peak = tpeak;
  c.album_new_peak = apeak;
  c.will_clip = t_over || a_over;
  if (c.will_clip && prevent) {
    if (t_over) {
      r->track_gain -= db_between(tpeak, limit);
      c.track_new_peak = limit;
      c.track_clipped = 1;
    }
    if (a_over) {
      r->album_gain -= db_between(apeak, limit);
      c.album_new_peak = limit;
      c.album_clipped = 1;
    }
    c.will_clip = 0;
  }
  c.album_would_clip = a_over && !c.album_clipped;
  *info = c;
  return 0;
}

extern "C" LG_EXPORT size_t lgb_format_tab_row(const char* name, const lgb_scan_result* r,
                                               const lgb_clip_info* info, int album_row,
                                               const char* unit, char* buf, size_t cap) {
  const double loud = album_row ? r->album_loudness : r->track_loudness;
  const double range = album_row ? r->album_loudness_range : r->track_loudness_range;
  const double peak = album_row ? r->album_peak : r->track_peak;
  const double gain = album_row ? r->album_gain : r->track_gain;
  const double npeak = album_row ? info->album_new_peak : info->track_new_peak;
  const int clip = album_row ? info->album_would_clip : info->will_clip;
  const int fixed = album_row ? info->album_clipped : info->track_clipped;
  const int n = snprintf(buf, cap,
                         "%s\t%.2f LUFS\t%.2f %s\t%.6f\t%.2f dBTP\t%.2f LUFS\t%s\t%s\t%.2f %s\t%.6f\t%.2f dBTP\n",
                         name, loud, range, unit, peak, 20.0 * log10(peak), r->loudness_reference,
                         clip ? "Y" : "N", fixed ? "Y" : "N", gain, unit, npeak, 20.0 * log10(npeak));
  return n < 0 ? 0 : (size_t) n;
}

// loudgain.c:566-585: the old mp3gain-compatible list (-o).  Peaks in 16-bit sample
// units; the columns mp3gain used for MP3 gain steps and min / max are always 0.
extern "C" LG_EXPORT size_t lgb_format_old_row(const char* name, const lgb_scan_result* r, int album_row,
                                               char* buf, size_t cap) {
  const int n = snprintf(buf, cap, "%s\t%d\t%.2f\t%.6f\t%d\t%d\n", name, 0,
                         album_row ? r->album_gain : r->track_gain,
                         (album_row ? r->album_peak : r->track_peak) * 32768.0, 0, 0);
  return n < 0 ? 0 : (size_t) n;
}

// loudgain.c:613-649: the human-readable block of a track ("\nTrack: <name>\n ...") or of
// the album ("\nAlbum:\n ...").  With `opus` the gain line also shows the Q7.8 number that
// goes into R128_TRACK_GAIN / R128_ALBUM_GAIN (tag.cc:442-445).
extern "C" LG_EXPORT size_t lgb_format_human(const char* name, const lgb_scan_result* r,
                                             const lgb_clip_info* info, int album_row, int opus,
                                             const char* unit, char* buf, size_t cap) {
  const double loud = album_row ? r->album_loudness : r->track_loudness;
  const double range = album_row ? r->album_loudness_range : r->track_loudness_range;
  const double peak = album_row ? r->album_peak : r->track_peak;
  const double gain = album_row ? r->album_gain : r->track_gain;
  const int fixed = album_row ? info->album_clipped : info->track_clipped;
  const char* note = fixed ? " (corrected to prevent clipping)" : "";
  size_t len = 0;
  auto put = [&](const char* fmt, auto... args) {
    const int n = snprintf(len < cap ? buf + len : nullptr, len < cap ? cap - len : 0, fmt, args...);
    if (n > 0) len += (size_t) n;
  };
  if (album_row) put("%s", "\nAlbum:\n");
  else put("\nTrack: %s\n", name);
  put(" Loudness: %8.2f LUFS\n", loud);
  put(" Range:    %8.2f %s\n", range, unit);
  put(" Peak:     %8.6f (%.2f dBTP)\n", peak, 20.0 * log10(peak));
  if (opus) put(" Gain:     %8.2f %s (%d)%s\n", gain, unit, (int) round(gain * 256.0), note);
  else put(" Gain:     %8.2f %s%s\n", gain, unit, note);
  return len;
}

extern "C" LG_EXPORT size_t lgb_format_tags(const lgb_scan_result* r, int do_album, int extended,
                                            int opus, const char* unit, char* buf, size_t cap) {
  size_t len = 0;
  auto put = [&](const char* fmt, auto... args) {
    const int n = snprintf(len < cap ? buf + len : nullptr, len < cap ? cap - len : 0, fmt, args...);
    if (n > 0) len += (size_t) n;
  };
  if (opus) {
    // Q7.8 fixed point: round(gain * 2^8)
    put("R128_TRACK_GAIN=%d\n", (int) round(r->track_gain * 256.0));
    if (do_album) put("R128_ALBUM_GAIN=%d\n", (int) round(r->album_gain * 256.0));
    return len;
  }
  put("REPLAYGAIN_TRACK_GAIN=%.2f %s\n", r->track_gain, unit);
  put("REPLAYGAIN_TRACK_PEAK=%.6f\n", r->track_peak);
  if (do_album) {
    put("REPLAYGAIN_ALBUM_GAIN=%.2f %s\n", r->album_gain, unit);
    put("REPLAYGAIN_ALBUM_PEAK=%.6f\n", r->album_peak);
  }
  if (extended) {
    put("REPLAYGAIN_REFERENCE_LOUDNESS=%.2f LUFS\n", r->loudness_reference);
    put("REPLAYGAIN_TRACK_RANGE=%.2f %s\n", r->track_loudness_range, unit);
    if (do_album) put("REPLAYGAIN_ALBUM_RANGE=%.2f %s\n", r->album_loudness_range, unit);
  }
  return len;
}

// ---- library scans over several GPUs: which rank scans which track -------------
// The reference parallelises a library over files and albums with one process
// each and lets the OS balance them (bin/rgbpm2:150-175).  With one process per
// GPU the tracks are dealt out ahead of time: longest processing time first,
// always to the rank with the smallest load so far (cost = frames x channels,
// what the sweep's time is proportional to).  Ties go to the lower rank, equal
// costs keep their input order, so every rank computes the same assignment.
extern "C" LG_EXPORT uint64_t lgb_lpt_assign(const uint64_t* cost, size_t n, uint32_t world, uint32_t* rank_out) {
  if (!world) return 0;
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[a] > cost[b]; });
  // (load, rank) min-heap
  typedef std::pair<uint64_t, uint32_t> Slot;
  std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
  for (uint32_t r = 0; r < world; ++r) heap.push(Slot(0, r));
  uint64_t worst = 0;
  for (size_t i : order) {
    Slot s = heap.top();
    heap.pop();
    rank_out[i] = s.second;
    s.first += cost[i];
    if (s.first > worst) worst = s.first;
    heap.push(s);
  }
  return worst;
}
