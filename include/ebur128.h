/*
 * ebur128.h -- drop-in C ABI of the loudness-analysis library that loudgain's
 * scanner links against (`#include <ebur128.h>`: /root/reference/src/scan.c:42,
 * /root/reference/src/loudgain.c:72; found by cmake/FindEBUR128.cmake:9-26 as
 * header "ebur128.h" + library "ebur128").
 *
 * This header is written from scratch for the B200-native implementation
 * (libebur128.so built from loudgain_b200/csrc). It declares the libebur128
 * 1.2.x-compatible surface: same struct layout, same enum values, same
 * function signatures, so scan.c compiles and links unchanged.  The CPU oracle
 * (oracle/ebur128_oracle.c) exports the same symbols and is the checker.
 *
 * Each entry point cites the reference call site it serves.
 */
#ifndef EBUR128_H_
#define EBUR128_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EBUR128_VERSION_MAJOR 1
#define EBUR128_VERSION_MINOR 2
#define EBUR128_VERSION_PATCH 6

/* Channel roles for ebur128_set_channel(). loudgain never calls
 * ebur128_set_channel, so the default map applies (see ebur128_init). */
enum channel {
  EBUR128_UNUSED = 0,
  EBUR128_LEFT = 1,
  EBUR128_Mp030 = 1,
  EBUR128_RIGHT = 2,
  EBUR128_Mm030 = 2,
  EBUR128_CENTER = 3,
  EBUR128_Mp000 = 3,
  EBUR128_LEFT_SURROUND = 4,
  EBUR128_Mp110 = 4,
  EBUR128_RIGHT_SURROUND = 5,
  EBUR128_Mm110 = 5,
  EBUR128_DUAL_MONO,
  EBUR128_MpSC,
  EBUR128_MmSC,
  EBUR128_Mp060,
  EBUR128_Mm060,
  EBUR128_Mp090,
  EBUR128_Mm090,
  EBUR128_Mp135,
  EBUR128_Mm135,
  EBUR128_Mp180,
  EBUR128_Up000,
  EBUR128_Up030,
  EBUR128_Um030,
  EBUR128_Up045,
  EBUR128_Um045,
  EBUR128_Up090,
  EBUR128_Um090,
  EBUR128_Up110,
  EBUR128_Um110,
  EBUR128_Up135,
  EBUR128_Um135,
  EBUR128_Up180,
  EBUR128_Tp000,
  EBUR128_Bp000,
  EBUR128_Bp045,
  EBUR128_Bm045
};

enum error {
  EBUR128_SUCCESS = 0,
  EBUR128_ERROR_NOMEM,
  EBUR128_ERROR_INVALID_MODE,
  EBUR128_ERROR_INVALID_CHANNEL_INDEX,
  EBUR128_ERROR_NO_CHANGE
};

/* Mode bits; scan.c:205-206 passes S | I | LRA | SAMPLE_PEAK | TRUE_PEAK and
 * never HISTOGRAM. */
enum mode {
  EBUR128_MODE_M = (1 << 0),
  EBUR128_MODE_S = (1 << 1) | EBUR128_MODE_M,
  EBUR128_MODE_I = (1 << 2) | EBUR128_MODE_M,
  EBUR128_MODE_LRA = (1 << 3) | EBUR128_MODE_S,
  EBUR128_MODE_SAMPLE_PEAK = (1 << 4) | EBUR128_MODE_M,
  EBUR128_MODE_TRUE_PEAK = (1 << 5) | EBUR128_MODE_M | EBUR128_MODE_SAMPLE_PEAK,
  EBUR128_MODE_HISTOGRAM = (1 << 6)
};

struct ebur128_state_internal;

/* Public state. scan.c:300 and scan.c:368 read `channels` straight from this
 * struct, so the field order and types are part of the ABI. */
typedef struct {
  int mode;
  unsigned int channels;
  unsigned long samplerate;
  struct ebur128_state_internal* d;
} ebur128_state;

/* loudgain.c:180 -- version shown by `loudgain -h`; warns below 1.2.4. */
void ebur128_get_version(int* major, int* minor, int* patch);

/* scan.c:203-207 -- one state per scanned file. NULL on failure. */
ebur128_state* ebur128_init(unsigned int channels, unsigned long samplerate, int mode);

/* scan.c:102 -- frees the state and stores NULL into *st. */
void ebur128_destroy(ebur128_state** st);

int ebur128_set_channel(ebur128_state* st, unsigned int channel_number, int value);
int ebur128_change_parameters(ebur128_state* st, unsigned int channels,
                              unsigned long samplerate);
int ebur128_set_max_window(ebur128_state* st, unsigned long window);
int ebur128_set_max_history(ebur128_state* st, unsigned long history);

/* scan.c:448-450 -- the sweep entry. `src` is interleaved [frames][channels]
 * and belongs to the caller, who frees it right after the call (scan.c:456). */
int ebur128_add_frames_short(ebur128_state* st, const short* src, size_t frames);
int ebur128_add_frames_int(ebur128_state* st, const int* src, size_t frames);
int ebur128_add_frames_float(ebur128_state* st, const float* src, size_t frames);
int ebur128_add_frames_double(ebur128_state* st, const double* src, size_t frames);

/* scan.c:294 -- integrated loudness in LUFS; -HUGE_VAL when nothing passed
 * the gates. */
int ebur128_loudness_global(ebur128_state* st, double* out);
/* scan.c:383-386 -- integrated loudness over the union of several states. */
int ebur128_loudness_global_multiple(ebur128_state** sts, size_t size, double* out);

int ebur128_loudness_momentary(ebur128_state* st, double* out);
int ebur128_loudness_shortterm(ebur128_state* st, double* out);
int ebur128_loudness_window(ebur128_state* st, unsigned long window, double* out);

/* scan.c:297 -- loudness range in LU. */
int ebur128_loudness_range(ebur128_state* st, double* out);
/* scan.c:388-391 -- loudness range over the union of several states. */
int ebur128_loudness_range_multiple(ebur128_state** sts, size_t size, double* out);

/* Named by the north-star API list; linear amplitude per channel. */
int ebur128_sample_peak(ebur128_state* st, unsigned int channel_number, double* out);
/* Peak of the frames of the last add_frames call only (live metering; not used by loudgain). */
int ebur128_prev_sample_peak(ebur128_state* st, unsigned int channel_number, double* out);

/* scan.c:303 and scan.c:371 -- linear amplitude, max(true peak, sample peak). */
int ebur128_true_peak(ebur128_state* st, unsigned int channel_number, double* out);
int ebur128_prev_true_peak(ebur128_state* st, unsigned int channel_number, double* out);

int ebur128_relative_threshold(ebur128_state* st, double* out);

#ifdef __cplusplus
}
#endif

#endif /* EBUR128_H_ */
