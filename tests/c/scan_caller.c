/* scan_caller.c -- a caller shaped like the reference scanner, in plain C,
 * compiled against include/ebur128.h and linked against a library that exports
 * the ebur128_* ABI (the product, or -- in tests -- the CPU oracle).
 *
 * It replays /root/reference/src/scan.c over synthetic PCM it generates itself
 * (no FFmpeg): ebur128_init with loudgain's mode word (scan.c:203-207), one
 * ebur128_add_frames_short call per 1024-frame "AVFrame" with a buffer that is
 * overwritten right after (scan.c:407-457), then per file the track queries of
 * scan_get_track_result (scan.c:275-330: loudness_global, loudness_range,
 * true_peak per channel read through st->channels), the album queries of
 * scan_set_album_result (scan.c:359-405: *_multiple over all states, re-issued
 * per track like loudgain.c:339-340), ebur128_destroy (scan.c:98-108) and
 * ebur128_get_version (loudgain.c:180-186).  Output: one line per track,
 * "index loudness range peak album_loudness album_range", %.10f.
 *
 * usage: scan_caller [ntracks] [seconds]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ebur128.h"

#define FRAME 1024

static unsigned int lcg(unsigned int* s) { *s = *s * 1664525u + 1013904223u; return *s; }

int main(int argc, char** argv) {
  const int ntracks = argc > 1 ? atoi(argv[1]) : 3;
  const double seconds = argc > 2 ? atof(argv[2]) : 8.0;
  const unsigned long rate = 44100;
  int major, minor, patch;
  ebur128_get_version(&major, &minor, &patch);
  if (major <= 1 && minor <= 2 && patch < 4) {              /* loudgain.c:183 */
    fprintf(stderr, "libebur128 >= 1.2.4 is needed\n");
    return 2;
  }
  ebur128_state** ebur128 = (ebur128_state**) calloc((size_t) ntracks, sizeof(ebur128_state*));
  short* buf = (short*) malloc(sizeof(short) * 2 * FRAME);
  if (!ebur128 || !buf) return 2;
  for (int i = 0; i < ntracks; ++i) {
    ebur128[i] = ebur128_init(2, rate,
                              EBUR128_MODE_S | EBUR128_MODE_I | EBUR128_MODE_LRA |
                                  EBUR128_MODE_SAMPLE_PEAK | EBUR128_MODE_TRUE_PEAK);
    if (ebur128[i] == NULL) {
      fprintf(stderr, "Could not initialize EBU R128 scanner\n");   /* scan.c:209 */
      return 3;
    }
    const long frames = (long) (seconds * (1.0 + 0.37 * i) * rate);
    unsigned int seed = 12345u + 77u * (unsigned) i;
    const double f1 = 220.0 * (i + 1), f2 = 3100.0 + 500.0 * i;
    for (long pos = 0; pos < frames; pos += FRAME) {
      const int n = frames - pos < FRAME ? (int) (frames - pos) : FRAME;
      for (int k = 0; k < n; ++k) {
        const double t = (double) (pos + k) / rate;
        /* level steps every 2 s so that the gates and the range have work to do */
        const double env = ((long) (t / 2.0) % 3 == 0) ? 0.08 : (((long) (t / 2.0) % 3 == 1) ? 0.5 : 0.9);
        const double noise = ((double) (lcg(&seed) >> 8) / 8388608.0 - 1.0) * 0.2;
        const double l = env * (0.6 * sin(6.283185307179586 * f1 * t) + 0.3 * sin(6.283185307179586 * f2 * t) + noise);
        const double r = env * (0.5 * sin(6.283185307179586 * f1 * t + 1.0) + noise * 0.7);
        double a = floor(l * 32767.0 + 0.5), b = floor(r * 32767.0 + 0.5);
        a = a > 32767.0 ? 32767.0 : (a < -32768.0 ? -32768.0 : a);
        b = b > 32767.0 ? 32767.0 : (b < -32768.0 ? -32768.0 : b);
        buf[2 * k] = (short) a;
        buf[2 * k + 1] = (short) b;
      }
      if (ebur128_add_frames_short(ebur128[i], buf, (size_t) n) != EBUR128_SUCCESS) {
        fprintf(stderr, "Error filtering\n");                          /* scan.c:451 */
        return 4;
      }
    }
  }
  for (int i = 0; i < ntracks; ++i) {
    double global, range, peak = 0.0, album_global, album_range;
    unsigned ch;
    if (ebur128_loudness_global(ebur128[i], &global) != EBUR128_SUCCESS) return 5;
    if (ebur128_loudness_range(ebur128[i], &range) != EBUR128_SUCCESS) return 5;
    for (ch = 0; ch < ebur128[i]->channels; ch++) {                   /* scan.c:300 */
      double tmp;
      if (ebur128_true_peak(ebur128[i], ch, &tmp) != EBUR128_SUCCESS) continue;
      if (tmp > peak) peak = tmp;
    }
    if (ebur128_loudness_global_multiple(ebur128, (size_t) ntracks, &album_global) != EBUR128_SUCCESS) return 6;
    if (ebur128_loudness_range_multiple(ebur128, (size_t) ntracks, &album_range) != EBUR128_SUCCESS) return 6;
    printf("%d %.10f %.10f %.10f %.10f %.10f\n", i, global, range, peak, album_global, album_range);
  }
  for (int i = 0; i < ntracks; ++i) {
    ebur128_destroy(&ebur128[i]);
    if (ebur128[i] != NULL) return 7;
  }
  free(ebur128);
  free(buf);
  return 0;
}
