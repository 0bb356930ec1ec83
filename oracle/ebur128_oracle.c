/*
 * ebur128_oracle.c -- CPU ORACLE (test infrastructure, NOT the product path).
 *
 * A plain-C, double-precision restatement of the loudness measurement that
 * loudgain's scanner drives through the ebur128_* API
 * (/root/reference/src/scan.c:203-207 init, :448-450 add_frames_short,
 * :294/:297/:303 track queries, :383-391 album queries, :102 destroy;
 * /root/reference/src/loudgain.c:180 version).
 *
 * PARITY UNPINNED: the arithmetic of this path lives in libebur128
 * (github.com/jiixyj/libebur128, floor v1.2.4: /root/reference/README.md:329,
 * /root/reference/debian/control:10), which is neither vendored in
 * /root/reference nor installed in this image, and the reference ships no
 * tests or golden vectors for it.  This file restates the published 1.2.x
 * algorithm: same block schedule, same direct-form-II
 * 4th-order K filter in double, same 49-tap Hann-windowed-sinc polyphase
 * interpolator with a float delay line and double accumulation, same
 * gating / percentile rules, same summation orders.  It is pinned only by
 * the ITU-R BS.1770 coefficient table, scipy.signal cross-checks and the
 * synthetic EBU Tech 3341/3342 cases (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product library
 * (loudgain_b200/csrc) never links or calls it.
 */
#include "../include/ebur128.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <xmmintrin.h>
#define FTZ_ENTER unsigned int mxcsr_saved_ = _mm_getcsr(); _mm_setcsr(mxcsr_saved_ | 0x8000u);
#define FTZ_LEAVE _mm_setcsr(mxcsr_saved_);
#else
#define FTZ_ENTER
#define FTZ_LEAVE
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------------------------------------------------------- types */

#define TP_TAPS 49
#define TP_MAX_FACTOR 4
#define HIST_BINS 1000

typedef struct {
  unsigned count;
  unsigned slot[TP_TAPS];   /* delay-line offset of each kept tap */
  double coef[TP_TAPS];
} subfilter;

typedef struct {
  unsigned factor, delay, channels, zi;
  subfilter phase[TP_MAX_FACTOR];
  float** line;             /* [channel][delay] circular delay lines */
} polyphase;

typedef struct {
  double* v;
  size_t n, cap;
} dlist;

struct ebur128_state_internal {
  double* ring;             /* [ring_frames][channels], K-weighted */
  size_t ring_frames;
  size_t ring_pos;          /* write position, in samples (frame*channels) */
  size_t need;              /* frames until the next block boundary */
  size_t s100;              /* frames per 100 ms */
  size_t st_count;          /* frames since the last short-term hop */
  int* chmap;
  double b[5], a[5];
  double (*w)[5];           /* DF-II delay elements per channel */
  dlist gate_blocks;        /* 400 ms energies >= absolute gate */
  dlist st_blocks;          /* 3 s energies >= absolute gate */
  int use_hist;
  unsigned long* gate_hist;
  unsigned long* st_hist;
  double* sample_peak;
  double* prev_sample_peak;
  double* true_peak;
  double* prev_true_peak;
  polyphase* os;
  float* os_in;
  float* os_out;
  size_t os_frames;
  unsigned long window_ms;
  unsigned long history_ms;
  size_t gate_max, st_max;
};

static double g_hist_energy[HIST_BINS];
static double g_hist_edge[HIST_BINS + 1];
static int g_tables_ready = 0;

static void tables_init(void) {
  int i;
  if (g_tables_ready) return;
  g_hist_edge[0] = pow(10.0, (-70.0 + 0.691) / 10.0);
  for (i = 0; i < HIST_BINS; ++i)
    g_hist_energy[i] = pow(10.0, ((double) i / 10.0 - 69.95 + 0.691) / 10.0);
  for (i = 1; i <= HIST_BINS; ++i)
    g_hist_edge[i] = pow(10.0, ((double) i / 10.0 - 70.0 + 0.691) / 10.0);
  g_tables_ready = 1;
}

static double abs_gate_energy(void) { return g_hist_edge[0]; }

static size_t hist_bin(double e) {
  size_t lo = 0, hi = HIST_BINS, mid;
  do {
    mid = (lo + hi) / 2;
    if (e >= g_hist_edge[mid]) lo = mid; else hi = mid;
  } while (hi - lo != 1);
  return lo;
}

static double to_lufs(double e) { return 10.0 * (log(e) / log(10.0)) - 0.691; }

static int dlist_push(dlist* l, double x, size_t max_len) {
  if (l->n == l->cap) {
    size_t nc = l->cap ? l->cap * 2 : 256;
    double* nv = (double*) realloc(l->v, nc * sizeof(double));
    if (!nv) return 1;
    l->v = nv; l->cap = nc;
  }
  l->v[l->n++] = x;
  if (l->n > max_len) {            /* bounded history: drop the oldest */
    size_t drop = l->n - max_len;
    memmove(l->v, l->v + drop, (l->n - drop) * sizeof(double));
    l->n -= drop;
  }
  return 0;
}

/* ------------------------------------------------------- K-weighting */

/* shelf (f0 1681.97 Hz, +4 dB) cascaded with the RLB
 * high-pass (f0 38.135 Hz), bilinear design, expanded to one 4th-order
 * section. */
static void kfilter_design(unsigned long rate, double b[5], double a[5]) {
  double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
  double K = tan(M_PI * f0 / (double) rate);
  double Vh = pow(10.0, G / 20.0);
  double Vb = pow(Vh, 0.4996667741545416);
  double den = 1.0 + K / Q + K * K;
  double sb[3], sa[3], hb[3] = {1.0, -2.0, 1.0}, ha[3];
  sb[0] = (Vh + Vb * K / Q + K * K) / den;
  sb[1] = 2.0 * (K * K - Vh) / den;
  sb[2] = (Vh - Vb * K / Q + K * K) / den;
  sa[0] = 1.0;
  sa[1] = 2.0 * (K * K - 1.0) / den;
  sa[2] = (1.0 - K / Q + K * K) / den;
  f0 = 38.13547087602444; Q = 0.5003270373238773;
  K = tan(M_PI * f0 / (double) rate);
  ha[0] = 1.0;
  ha[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K);
  ha[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);
  b[0] = sb[0] * hb[0];
  b[1] = sb[0] * hb[1] + sb[1] * hb[0];
  b[2] = sb[0] * hb[2] + sb[1] * hb[1] + sb[2] * hb[0];
  b[3] = sb[1] * hb[2] + sb[2] * hb[1];
  b[4] = sb[2] * hb[2];
  a[0] = sa[0] * ha[0];
  a[1] = sa[0] * ha[1] + sa[1] * ha[0];
  a[2] = sa[0] * ha[2] + sa[1] * ha[1] + sa[2] * ha[0];
  a[3] = sa[1] * ha[2] + sa[2] * ha[1];
  a[4] = sa[2] * ha[2];
}

/* Exposed for tests (coefficient table check). Not part of ebur128.h. */
void oracle_kfilter_coeffs(unsigned long rate, double* b5, double* a5) {
  kfilter_design(rate, b5, a5);
}

/* ------------------------------------------------ true-peak polyphase */

/* 49-tap Hann-windowed sinc split into `factor` phases. */
static polyphase* polyphase_new(unsigned factor, unsigned channels) {
  unsigned j, c;
  polyphase* p = (polyphase*) calloc(1, sizeof(*p));
  if (!p) return NULL;
  p->factor = factor;
  p->channels = channels;
  p->delay = (TP_TAPS + factor - 1) / factor;
  for (j = 0; j < TP_TAPS; ++j) {
    double m = (double) j - (double) (TP_TAPS - 1) / 2.0;
    double c0 = 1.0;
    if (fabs(m) > 0.000001) c0 = sin(m * M_PI / factor) / (m * M_PI / factor);
    c0 *= 0.5 * (1.0 - cos(2.0 * M_PI * j / (TP_TAPS - 1)));
    if (fabs(c0) > 0.000001) {
      subfilter* f = &p->phase[j % factor];
      f->coef[f->count] = c0;
      f->slot[f->count] = j / factor;
      f->count++;
    }
  }
  p->line = (float**) calloc(channels, sizeof(float*));
  if (!p->line) { free(p); return NULL; }
  for (c = 0; c < channels; ++c) {
    p->line[c] = (float*) calloc(p->delay, sizeof(float));
    if (!p->line[c]) return NULL;
  }
  return p;
}

static void polyphase_free(polyphase* p) {
  unsigned c;
  if (!p) return;
  for (c = 0; c < p->channels; ++c) free(p->line[c]);
  free(p->line);
  free(p);
}

/* Exposed for tests: taps of one phase. Returns tap count. */
unsigned oracle_tp_phase(unsigned factor, unsigned phase, double* coef, unsigned* slot) {
  unsigned t, n;
  polyphase* p = polyphase_new(factor, 1);
  if (!p) return 0;
  n = p->phase[phase].count;
  for (t = 0; t < n; ++t) { coef[t] = p->phase[phase].coef[t]; slot[t] = p->phase[phase].slot[t]; }
  polyphase_free(p);
  return n;
}

/* in: [frames][channels] float; out: [frames*factor][channels] float. */
static void polyphase_run(polyphase* p, size_t frames, const float* in, float* out) {
  size_t n;
  unsigned c, f, t;
  for (n = 0; n < frames; ++n) {
    for (c = 0; c < p->channels; ++c) {
      float* o = out + c;
      p->line[c][p->zi] = *in++;
      for (f = 0; f < p->factor; ++f) {
        const subfilter* sf = &p->phase[f];
        double acc = 0.0;
        for (t = 0; t < sf->count; ++t) {
          int i = (int) p->zi - (int) sf->slot[t];
          if (i < 0) i += (int) p->delay;
          acc += (double) p->line[c][i] * sf->coef[t];
        }
        *o = (float) acc;
        o += p->channels;
      }
    }
    out += p->channels * p->factor;
    if (++p->zi == p->delay) p->zi = 0;
  }
}

static int oversampler_setup(ebur128_state* st) {
  struct ebur128_state_internal* d = st->d;
  d->os = NULL; d->os_in = NULL; d->os_out = NULL; d->os_frames = 0;
  if ((st->mode & EBUR128_MODE_TRUE_PEAK) != EBUR128_MODE_TRUE_PEAK) return 0;
  if (st->samplerate < 96000) d->os = polyphase_new(4, st->channels);
  else if (st->samplerate < 192000) d->os = polyphase_new(2, st->channels);
  else return 0;                       /* >= 192 kHz: no oversampling */
  if (!d->os) return 1;
  d->os_frames = d->s100 * 4;
  d->os_in = (float*) malloc(d->os_frames * st->channels * sizeof(float));
  d->os_out = (float*) malloc(d->os_frames * st->channels * d->os->factor * sizeof(float));
  return (!d->os_in || !d->os_out);
}

static void oversampler_teardown(ebur128_state* st) {
  polyphase_free(st->d->os);
  free(st->d->os_in);
  free(st->d->os_out);
  st->d->os = NULL; st->d->os_in = NULL; st->d->os_out = NULL;
}

static void true_peak_scan(ebur128_state* st, size_t frames) {
  struct ebur128_state_internal* d = st->d;
  size_t i, n = frames * d->os->factor;
  unsigned c;
  polyphase_run(d->os, frames, d->os_in, d->os_out);
  for (i = 0; i < n; ++i)
    for (c = 0; c < st->channels; ++c) {
      double v = (double) d->os_out[i * st->channels + c];
      if (v < 0) v = -v;
      if (v > d->prev_true_peak[c]) d->prev_true_peak[c] = v;
    }
}

/* ------------------------------------------------------------ set-up */

static void default_channel_map(int* map, unsigned channels) {
  unsigned i;
  if (channels == 4) {
    map[0] = EBUR128_LEFT; map[1] = EBUR128_RIGHT;
    map[2] = EBUR128_LEFT_SURROUND; map[3] = EBUR128_RIGHT_SURROUND;
  } else if (channels == 5) {
    map[0] = EBUR128_LEFT; map[1] = EBUR128_RIGHT; map[2] = EBUR128_CENTER;
    map[3] = EBUR128_LEFT_SURROUND; map[4] = EBUR128_RIGHT_SURROUND;
  } else {
    static const int six[6] = {EBUR128_LEFT, EBUR128_RIGHT, EBUR128_CENTER,
                               EBUR128_UNUSED, EBUR128_LEFT_SURROUND,
                               EBUR128_RIGHT_SURROUND};
    for (i = 0; i < channels; ++i) map[i] = i < 6 ? six[i] : EBUR128_UNUSED;
  }
}

static size_t ring_frames_for(unsigned long rate, unsigned long window_ms, size_t s100) {
  size_t f = rate * window_ms / 1000;
  if (f % s100) f += s100 - f % s100;
  return f;
}

void ebur128_get_version(int* major, int* minor, int* patch) {
  *major = EBUR128_VERSION_MAJOR;
  *minor = EBUR128_VERSION_MINOR;
  *patch = EBUR128_VERSION_PATCH;
}

static int peaks_alloc(struct ebur128_state_internal* d, unsigned channels) {
  d->sample_peak = (double*) calloc(channels, sizeof(double));
  d->prev_sample_peak = (double*) calloc(channels, sizeof(double));
  d->true_peak = (double*) calloc(channels, sizeof(double));
  d->prev_true_peak = (double*) calloc(channels, sizeof(double));
  return !(d->sample_peak && d->prev_sample_peak && d->true_peak && d->prev_true_peak);
}

static void peaks_free(struct ebur128_state_internal* d) {
  free(d->sample_peak); free(d->prev_sample_peak);
  free(d->true_peak); free(d->prev_true_peak);
}

ebur128_state* ebur128_init(unsigned int channels, unsigned long samplerate, int mode) {
  ebur128_state* st;
  struct ebur128_state_internal* d;
  if (channels == 0 || channels > 64 || samplerate < 16 || samplerate > 2822400)
    return NULL;
  tables_init();
  st = (ebur128_state*) calloc(1, sizeof(*st));
  if (!st) return NULL;
  d = (struct ebur128_state_internal*) calloc(1, sizeof(*d));
  if (!d) { free(st); return NULL; }
  st->d = d;
  st->channels = channels;
  st->samplerate = samplerate;
  st->mode = mode;
  d->chmap = (int*) calloc(channels, sizeof(int));
  if (!d->chmap || peaks_alloc(d, channels)) goto fail;
  default_channel_map(d->chmap, channels);
  d->s100 = (samplerate + 5) / 10;
  d->history_ms = ULONG_MAX;
  d->gate_max = (size_t) -1;
  d->st_max = (size_t) -1;
  if ((mode & EBUR128_MODE_S) == EBUR128_MODE_S) d->window_ms = 3000;
  else if ((mode & EBUR128_MODE_M) == EBUR128_MODE_M) d->window_ms = 400;
  else goto fail;
  d->ring_frames = ring_frames_for(samplerate, d->window_ms, d->s100);
  d->ring = (double*) calloc(d->ring_frames * channels, sizeof(double));
  d->w = (double (*)[5]) calloc(channels, sizeof(double[5]));
  if (!d->ring || !d->w) goto fail;
  kfilter_design(samplerate, d->b, d->a);
  d->use_hist = (mode & EBUR128_MODE_HISTOGRAM) ? 1 : 0;
  if (d->use_hist) {
    d->gate_hist = (unsigned long*) calloc(HIST_BINS, sizeof(unsigned long));
    d->st_hist = (unsigned long*) calloc(HIST_BINS, sizeof(unsigned long));
    if (!d->gate_hist || !d->st_hist) goto fail;
  }
  if (oversampler_setup(st)) goto fail;
  d->need = d->s100 * 4;              /* the first block needs 400 ms */
  d->ring_pos = 0;
  d->st_count = 0;
  return st;
fail:
  ebur128_destroy(&st);
  return NULL;
}

void ebur128_destroy(ebur128_state** stp) {
  ebur128_state* st;
  if (!stp || !*stp) return;
  st = *stp;
  if (st->d) {
    oversampler_teardown(st);
    free(st->d->ring);
    free(st->d->w);
    free(st->d->chmap);
    free(st->d->gate_blocks.v);
    free(st->d->st_blocks.v);
    free(st->d->gate_hist);
    free(st->d->st_hist);
    peaks_free(st->d);
    free(st->d);
  }
  free(st);
  *stp = NULL;
}

int ebur128_set_channel(ebur128_state* st, unsigned int ch, int value) {
  if (ch >= st->channels) return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  if (value == EBUR128_DUAL_MONO && (st->channels != 1 || ch != 0)) {
    fprintf(stderr, "EBUR128_DUAL_MONO only works with mono files!\n");
    return EBUR128_ERROR_INVALID_CHANNEL_INDEX;
  }
  st->d->chmap[ch] = value;
  return EBUR128_SUCCESS;
}

int ebur128_change_parameters(ebur128_state* st, unsigned int channels,
                              unsigned long samplerate) {
  struct ebur128_state_internal* d = st->d;
  if (channels == 0 || channels > 64 || samplerate < 16 || samplerate > 2822400)
    return EBUR128_ERROR_NOMEM;
  if (channels == st->channels && samplerate == st->samplerate)
    return EBUR128_ERROR_NO_CHANGE;
  oversampler_teardown(st);
  free(d->ring); d->ring = NULL;
  if (channels != st->channels) {
    free(d->chmap); peaks_free(d); free(d->w);
    st->channels = channels;
    d->chmap = (int*) calloc(channels, sizeof(int));
    d->w = (double (*)[5]) calloc(channels, sizeof(double[5]));
    if (!d->chmap || !d->w || peaks_alloc(d, channels)) return EBUR128_ERROR_NOMEM;
    default_channel_map(d->chmap, channels);
  } else {
    memset(d->w, 0, channels * sizeof(double[5]));
  }
  if (samplerate != st->samplerate) {
    st->samplerate = samplerate;
    d->s100 = (samplerate + 5) / 10;
  }
  kfilter_design(st->samplerate, d->b, d->a);
  d->ring_frames = ring_frames_for(st->samplerate, d->window_ms, d->s100);
  d->ring = (double*) calloc(d->ring_frames * st->channels, sizeof(double));
  if (!d->ring) return EBUR128_ERROR_NOMEM;
  if (oversampler_setup(st)) return EBUR128_ERROR_NOMEM;
  d->need = d->s100 * 4;
  d->ring_pos = 0;
  d->st_count = 0;
  return EBUR128_SUCCESS;
}

int ebur128_set_max_window(ebur128_state* st, unsigned long window) {
  struct ebur128_state_internal* d = st->d;
  if ((st->mode & EBUR128_MODE_S) == EBUR128_MODE_S && window < 3000) window = 3000;
  else if ((st->mode & EBUR128_MODE_M) == EBUR128_MODE_M && window < 400) window = 400;
  if (window == d->window_ms) return EBUR128_ERROR_NO_CHANGE;
  d->window_ms = window;
  free(d->ring);
  d->ring_frames = ring_frames_for(st->samplerate, d->window_ms, d->s100);
  d->ring = (double*) calloc(d->ring_frames * st->channels, sizeof(double));
  if (!d->ring) return EBUR128_ERROR_NOMEM;
  d->need = d->s100 * 4;
  d->ring_pos = 0;
  d->st_count = 0;
  return EBUR128_SUCCESS;
}

int ebur128_set_max_history(ebur128_state* st, unsigned long history) {
  struct ebur128_state_internal* d = st->d;
  if ((st->mode & EBUR128_MODE_LRA) == EBUR128_MODE_LRA && history < 3000) history = 3000;
  else if ((st->mode & EBUR128_MODE_M) == EBUR128_MODE_M && history < 400) history = 400;
  if (history == d->history_ms) return EBUR128_ERROR_NO_CHANGE;
  d->history_ms = history;
  d->gate_max = history / 100;
  d->st_max = history / 3000;
  if (d->gate_blocks.n > d->gate_max) {
    size_t drop = d->gate_blocks.n - d->gate_max;
    memmove(d->gate_blocks.v, d->gate_blocks.v + drop, d->gate_max * sizeof(double));
    d->gate_blocks.n = d->gate_max;
  }
  if (d->st_blocks.n > d->st_max) {
    size_t drop = d->st_blocks.n - d->st_max;
    memmove(d->st_blocks.v, d->st_blocks.v + drop, d->st_max * sizeof(double));
    d->st_blocks.n = d->st_max;
  }
  return EBUR128_SUCCESS;
}

/* -------------------------------------------------------- the sweep */

static double channel_weight(int role) {
  switch (role) {
    case EBUR128_Mp110: case EBUR128_Mm110:
    case EBUR128_Mp060: case EBUR128_Mm060:
    case EBUR128_Mp090: case EBUR128_Mm090:
      return 1.41;
    case EBUR128_DUAL_MONO:
      return 2.0;
    default:
      return 1.0;
  }
}

/* Mean weighted energy of the last `frames` frames in the ring
 * (serial double sums, wrapped part first). */
static double ring_energy(const ebur128_state* st, size_t frames) {
  const struct ebur128_state_internal* d = st->d;
  const size_t C = st->channels;
  const size_t pos = d->ring_pos / C;
  double total = 0.0;
  size_t c, i;
  for (c = 0; c < C; ++c) {
    double s = 0.0;
    if (d->chmap[c] == EBUR128_UNUSED) continue;
    if (pos < frames) {
      for (i = 0; i < pos; ++i) {
        double x = d->ring[i * C + c]; s += x * x;
      }
      for (i = d->ring_frames - (frames - pos); i < d->ring_frames; ++i) {
        double x = d->ring[i * C + c]; s += x * x;
      }
    } else {
      for (i = pos - frames; i < pos; ++i) {
        double x = d->ring[i * C + c]; s += x * x;
      }
    }
    {
      double wgt = channel_weight(d->chmap[c]);
      if (wgt != 1.0) s *= wgt;
    }
    total += s;
  }
  return total / (double) frames;
}

static int record_block(struct ebur128_state_internal* d, double e, int short_term) {
  if (e < abs_gate_energy()) return 0;
  if (d->use_hist) {
    ++(short_term ? d->st_hist : d->gate_hist)[hist_bin(e)];
    return 0;
  }
  return short_term ? dlist_push(&d->st_blocks, e, d->st_max)
                    : dlist_push(&d->gate_blocks, e, d->gate_max);
}

#define DEFINE_FEED(NAME, T, SCALE)                                            \
  static void NAME(ebur128_state* st, const T* src, size_t frames) {           \
    struct ebur128_state_internal* d = st->d;                                  \
    const size_t C = st->channels;                                             \
    const double scale = (SCALE);                                              \
    double* dst = d->ring + d->ring_pos;                                       \
    size_t c, i;                                                               \
    FTZ_ENTER                                                                  \
    if ((st->mode & EBUR128_MODE_SAMPLE_PEAK) == EBUR128_MODE_SAMPLE_PEAK) {   \
      for (c = 0; c < C; ++c) {                                                \
        double m = 0.0;                                                        \
        for (i = 0; i < frames; ++i) {                                         \
          double x = (double) src[i * C + c];                                  \
          if (x > m) m = x; else if (-x > m) m = -x;                           \
        }                                                                      \
        m /= scale;                                                            \
        if (m > d->prev_sample_peak[c]) d->prev_sample_peak[c] = m;            \
      }                                                                        \
    }                                                                          \
    if (d->os) {                                                               \
      for (i = 0; i < frames * C; ++i)                                         \
        d->os_in[i] = (float) ((double) src[i] / scale);                       \
      true_peak_scan(st, frames);                                              \
    }                                                                          \
    for (c = 0; c < C; ++c) {                                                  \
      double* w = d->w[c];                                                     \
      if (d->chmap[c] == EBUR128_UNUSED) continue;                             \
      for (i = 0; i < frames; ++i) {                                           \
        w[0] = (double) src[i * C + c] / scale - d->a[1] * w[1] -              \
               d->a[2] * w[2] - d->a[3] * w[3] - d->a[4] * w[4];               \
        dst[i * C + c] = d->b[0] * w[0] + d->b[1] * w[1] + d->b[2] * w[2] +    \
                         d->b[3] * w[3] + d->b[4] * w[4];                      \
        w[4] = w[3]; w[3] = w[2]; w[2] = w[1]; w[1] = w[0];                    \
      }                                                                        \
      for (i = 1; i < 5; ++i)                                                  \
        if (fabs(w[i]) < DBL_MIN) w[i] = 0.0;                                  \
    }                                                                          \
    FTZ_LEAVE                                                                  \
  }

DEFINE_FEED(feed_short, short, 32768.0)
DEFINE_FEED(feed_int, int, 2147483648.0)
DEFINE_FEED(feed_float, float, 1.0)
DEFINE_FEED(feed_double, double, 1.0)

/* Block schedule of SURVEY.md A.2. */
#define DEFINE_ADD(NAME, T, FEED)                                              \
  int NAME(ebur128_state* st, const T* src, size_t frames) {                   \
    struct ebur128_state_internal* d = st->d;                                  \
    const size_t C = st->channels;                                             \
    size_t off = 0, c;                                                         \
    for (c = 0; c < C; ++c) {                                                  \
      d->prev_sample_peak[c] = 0.0;                                            \
      d->prev_true_peak[c] = 0.0;                                              \
    }                                                                          \
    while (frames > 0) {                                                       \
      if (frames >= d->need) {                                                 \
        FEED(st, src + off, d->need);                                          \
        off += d->need * C;                                                    \
        frames -= d->need;                                                     \
        d->ring_pos += d->need * C;                                            \
        if ((st->mode & EBUR128_MODE_I) == EBUR128_MODE_I)                     \
          if (record_block(d, ring_energy(st, d->s100 * 4), 0))                \
            return EBUR128_ERROR_NOMEM;                                        \
        if ((st->mode & EBUR128_MODE_LRA) == EBUR128_MODE_LRA) {               \
          d->st_count += d->need;                                              \
          if (d->st_count == d->s100 * 30) {                                   \
            if (record_block(d, ring_energy(st, d->s100 * 30), 1))             \
              return EBUR128_ERROR_NOMEM;                                      \
            d->st_count = d->s100 * 20;                                        \
          }                                                                    \
        }                                                                      \
        d->need = d->s100;                                                     \
        if (d->ring_pos == d->ring_frames * C) d->ring_pos = 0;                \
      } else {                                                                 \
        FEED(st, src + off, frames);                                           \
        d->ring_pos += frames * C;                                             \
        if ((st->mode & EBUR128_MODE_LRA) == EBUR128_MODE_LRA)                 \
          d->st_count += frames;                                               \
        d->need -= frames;                                                     \
        frames = 0;                                                            \
      }                                                                        \
    }                                                                          \
    for (c = 0; c < C; ++c) {                                                  \
      if (d->prev_sample_peak[c] > d->sample_peak[c])                          \
        d->sample_peak[c] = d->prev_sample_peak[c];                            \
      if (d->prev_true_peak[c] > d->true_peak[c])                              \
        d->true_peak[c] = d->prev_true_peak[c];                                \
    }                                                                          \
    return EBUR128_SUCCESS;                                                    \
  }

DEFINE_ADD(ebur128_add_frames_short, short, feed_short)
DEFINE_ADD(ebur128_add_frames_int, int, feed_int)
DEFINE_ADD(ebur128_add_frames_float, float, feed_float)
DEFINE_ADD(ebur128_add_frames_double, double, feed_double)

/* --------------------------------------------------------- queries */

/* Accumulate (sum, count) of stored gating blocks at or above `floor_e`
 * over a set of states; histogram states contribute bin centres. */
static void gate_accumulate(ebur128_state** sts, size_t n, int use_floor,
                            double floor_e, double* sum, size_t* count) {
  size_t i, j;
  for (i = 0; i < n; ++i) {
    struct ebur128_state_internal* d;
    if (!sts[i]) continue;
    d = sts[i]->d;
    if (d->use_hist) {
      size_t start = 0;
      if (use_floor && floor_e >= g_hist_edge[0]) {
        start = hist_bin(floor_e);
        if (floor_e > g_hist_energy[start]) ++start;
      }
      for (j = start; j < HIST_BINS; ++j) {
        *sum += d->gate_hist[j] * g_hist_energy[j];
        *count += d->gate_hist[j];
      }
    } else {
      for (j = 0; j < d->gate_blocks.n; ++j) {
        double z = d->gate_blocks.v[j];
        if (!use_floor || z >= floor_e) { *sum += z; ++*count; }
      }
    }
  }
}

/* SURVEY.md A.6. */
static int gated_loudness(ebur128_state** sts, size_t n, double* out) {
  double sum = 0.0, thr;
  size_t cnt = 0, i;
  for (i = 0; i < n; ++i)
    if (sts[i] && (sts[i]->mode & EBUR128_MODE_I) != EBUR128_MODE_I)
      return EBUR128_ERROR_INVALID_MODE;
  gate_accumulate(sts, n, 0, 0.0, &sum, &cnt);
  if (!cnt) { *out = -HUGE_VAL; return EBUR128_SUCCESS; }
  thr = sum / (double) cnt;
  thr *= pow(10.0, -10.0 / 10.0);
  sum = 0.0; cnt = 0;
  gate_accumulate(sts, n, 1, thr, &sum, &cnt);
  if (!cnt) { *out = -HUGE_VAL; return EBUR128_SUCCESS; }
  *out = to_lufs(sum / (double) cnt);
  return EBUR128_SUCCESS;
}

int ebur128_loudness_global(ebur128_state* st, double* out) {
  return gated_loudness(&st, 1, out);
}

int ebur128_loudness_global_multiple(ebur128_state** sts, size_t size, double* out) {
  return gated_loudness(sts, size, out);
}

int ebur128_relative_threshold(ebur128_state* st, double* out) {
  double sum = 0.0;
  size_t cnt = 0;
  if ((st->mode & EBUR128_MODE_I) != EBUR128_MODE_I) return EBUR128_ERROR_INVALID_MODE;
  gate_accumulate(&st, 1, 0, 0.0, &sum, &cnt);
  if (!cnt) { *out = -70.0; return EBUR128_SUCCESS; }
  *out = to_lufs(sum / (double) cnt * pow(10.0, -10.0 / 10.0));
  return EBUR128_SUCCESS;
}

static int window_loudness(ebur128_state* st, size_t frames, double* out) {
  double e;
  if (frames > st->d->ring_frames) return EBUR128_ERROR_INVALID_MODE;
  e = ring_energy(st, frames);
  *out = e <= 0.0 ? -HUGE_VAL : to_lufs(e);
  return EBUR128_SUCCESS;
}

int ebur128_loudness_momentary(ebur128_state* st, double* out) {
  return window_loudness(st, st->d->s100 * 4, out);
}

int ebur128_loudness_shortterm(ebur128_state* st, double* out) {
  if ((st->mode & EBUR128_MODE_S) != EBUR128_MODE_S) return EBUR128_ERROR_INVALID_MODE;
  return window_loudness(st, st->d->s100 * 30, out);
}

int ebur128_loudness_window(ebur128_state* st, unsigned long window, double* out) {
  size_t frames = st->samplerate * window / 1000;
  return window_loudness(st, frames, out);
}

static int cmp_double(const void* p, const void* q) {
  double a = *(const double*) p, b = *(const double*) q;
  return (a > b) - (a < b);
}

/* SURVEY.md A.7. */
int ebur128_loudness_range_multiple(ebur128_state** sts, size_t size, double* out) {
  size_t i, j, total = 0, nhist = 0, nlist = 0;
  for (i = 0; i < size; ++i) {
    if (!sts[i]) continue;
    if ((sts[i]->mode & EBUR128_MODE_LRA) != EBUR128_MODE_LRA)
      return EBUR128_ERROR_INVALID_MODE;
    if (sts[i]->d->use_hist) ++nhist; else ++nlist;
  }
  if (nhist && nlist) return EBUR128_ERROR_INVALID_MODE;
  if (nhist) {
    unsigned long hist[HIST_BINS];
    size_t n = 0, lo_i, hi_i, acc, k;
    double power = 0.0, floor_e;
    memset(hist, 0, sizeof(hist));
    for (i = 0; i < size; ++i) {
      if (!sts[i]) continue;
      for (j = 0; j < HIST_BINS; ++j) {
        hist[j] += sts[i]->d->st_hist[j];
        n += sts[i]->d->st_hist[j];
        power += sts[i]->d->st_hist[j] * g_hist_energy[j];
      }
    }
    if (!n) { *out = 0.0; return EBUR128_SUCCESS; }
    power /= (double) n;
    floor_e = pow(10.0, -20.0 / 10.0) * power;
    if (floor_e < g_hist_edge[0]) k = 0;
    else { k = hist_bin(floor_e); if (floor_e > g_hist_energy[k]) ++k; }
    n = 0;
    for (j = k; j < HIST_BINS; ++j) n += hist[j];
    if (!n) { *out = 0.0; return EBUR128_SUCCESS; }
    lo_i = (size_t) ((double) (n - 1) * 0.1 + 0.5);
    hi_i = (size_t) ((double) (n - 1) * 0.95 + 0.5);
    acc = 0; j = k;
    while (acc <= lo_i) acc += hist[j++];
    {
      double lo_e = g_hist_energy[j - 1];
      while (acc <= hi_i) acc += hist[j++];
      *out = to_lufs(g_hist_energy[j - 1]) - to_lufs(lo_e);
    }
    return EBUR128_SUCCESS;
  }
  for (i = 0; i < size; ++i)
    if (sts[i]) total += sts[i]->d->st_blocks.n;
  if (!total) { *out = 0.0; return EBUR128_SUCCESS; }
  {
    double* v = (double*) malloc(total * sizeof(double));
    double mean = 0.0, floor_e;
    size_t k = 0, first = 0, n;
    if (!v) return EBUR128_ERROR_NOMEM;
    for (i = 0; i < size; ++i) {
      if (!sts[i]) continue;
      memcpy(v + k, sts[i]->d->st_blocks.v, sts[i]->d->st_blocks.n * sizeof(double));
      k += sts[i]->d->st_blocks.n;
    }
    qsort(v, total, sizeof(double), cmp_double);
    for (j = 0; j < total; ++j) mean += v[j];
    mean /= (double) total;
    floor_e = pow(10.0, -20.0 / 10.0) * mean;
    while (first < total && v[first] < floor_e) ++first;
    n = total - first;
    if (n) {
      double hi = v[first + (size_t) ((double) (n - 1) * 0.95 + 0.5)];
      double lo = v[first + (size_t) ((double) (n - 1) * 0.1 + 0.5)];
      *out = to_lufs(hi) - to_lufs(lo);
    } else {
      *out = 0.0;
    }
    free(v);
  }
  return EBUR128_SUCCESS;
}

int ebur128_loudness_range(ebur128_state* st, double* out) {
  return ebur128_loudness_range_multiple(&st, 1, out);
}

#define PEAK_QUERY(NAME, MODEBITS, EXPR)                                       \
  int NAME(ebur128_state* st, unsigned int ch, double* out) {                  \
    if ((st->mode & (MODEBITS)) != (MODEBITS)) return EBUR128_ERROR_INVALID_MODE; \
    if (ch >= st->channels) return EBUR128_ERROR_INVALID_CHANNEL_INDEX;        \
    *out = (EXPR);                                                             \
    return EBUR128_SUCCESS;                                                    \
  }

PEAK_QUERY(ebur128_sample_peak, EBUR128_MODE_SAMPLE_PEAK, st->d->sample_peak[ch])
PEAK_QUERY(ebur128_prev_sample_peak, EBUR128_MODE_SAMPLE_PEAK, st->d->prev_sample_peak[ch])
PEAK_QUERY(ebur128_true_peak, EBUR128_MODE_TRUE_PEAK,
           st->d->true_peak[ch] > st->d->sample_peak[ch] ? st->d->true_peak[ch]
                                                         : st->d->sample_peak[ch])
PEAK_QUERY(ebur128_prev_true_peak, EBUR128_MODE_TRUE_PEAK,
           st->d->prev_true_peak[ch] > st->d->prev_sample_peak[ch]
               ? st->d->prev_true_peak[ch] : st->d->prev_sample_peak[ch])

/* ------------------------------------------------- test-only access */

/* Copies the stored block energies so tests can compare block lists, not
 * just the final scalars. kind 0 = 400 ms gating blocks, 1 = 3 s blocks. */
size_t oracle_block_count(const ebur128_state* st, int kind) {
  return kind ? st->d->st_blocks.n : st->d->gate_blocks.n;
}

size_t oracle_copy_blocks(const ebur128_state* st, int kind, double* dst, size_t cap) {
  const dlist* l = kind ? &st->d->st_blocks : &st->d->gate_blocks;
  size_t n = l->n < cap ? l->n : cap;
  memcpy(dst, l->v, n * sizeof(double));
  return n;
}
