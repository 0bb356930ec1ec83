// lg_scan.cu -- scan.c-shaped host driver over this library's own ebur128_*
// ABI (declared in include/ebur128_b200.h).  It exists so that the end-to-end
// path a loudgain user exercises -- host PCM in caller-owned buffers, one
// ebur128_add_frames_short per decoded frame, queries after all files -- can be
// run and timed without FFmpeg: /root/reference/src/scan.c:225-256,448 (feed
// loop), :275-330 (track result), :359-405 (album result).
#include <math.h>
#include <stdint.h>

#include <atomic>
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
  const size_t fb = t.channels * (t.format == LGB_FORMAT_S16 ? 2u : 4u);
  for (uint64_t pos = 0; pos < t.frames; pos += step) {
    const size_t n = (size_t) (t.frames - pos < step ? t.frames - pos : step);
    const char* p = (const char*) t.pcm + pos * fb;
    const int e = t.format == LGB_FORMAT_S16
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
  // ---- results (loudgain.c:323-340): track, then album, per file
  for (size_t i = 0; i < ntracks && !rc; ++i) {
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
      r.album_gain = lufs_to_rg(global) + pre_gain;
      r.album_peak = apeak;
      r.album_loudness = global;
      r.album_loudness_range = range;
    }
  }
  for (size_t i = 0; i < ntracks; ++i) ebur128_destroy(&states[i]);   // scan_deinit
  return rc;
}
