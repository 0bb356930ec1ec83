// lg_packed.cuh -- the K-weighting step on a channel PAIR with Blackwell's
// packed FP32 instructions (FFMA2 / FADD2: two independent IEEE operations per
// issue slot, the filter constant broadcast from a uniform register), shared by
// the packed sweeps (lg_pair.cu, lg_run.cu).  Per channel this is the same
// operation sequence as lg_sweep.cuh: k_step, so results are bit-identical to
// the scalar sweep and to the host emulation (tests/emu).
#pragma once

#include <cuda_runtime.h>

#include "lg_common.h"

namespace lg {

__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

// S: any struct with float2 members xp, d1, w1, w2, v1, v2 (.x = channel 2j,
// .y = channel 2j + 1).  x - xp is written as fma(-1, xp, x): one rounding,
// the same value as the scalar subtraction.
template <class S>
__device__ __forceinline__ float2 k_step2(S& s, const float2 x, const SweepParams& k) {
  const float2 q = __ffma2_rn(bc2(-1.0f), s.xp, x);
  const float2 t = __ffma2_rn(bc2(k.ne2), s.w2, q);
  const float2 d = __ffma2_rn(bc2(k.c), s.d1, t);
  const float2 w = __fadd2_rn(s.w1, d);
  const float2 u = __ffma2_rn(bc2(k.np2), s.v2, d);
  const float2 v = __ffma2_rn(bc2(k.np1), s.v1, u);
  const float2 y = __ffma2_rn(bc2(k.r2), s.v2, __ffma2_rn(bc2(k.r1), s.v1, d));
  s.xp = x;
  s.w2 = s.w1; s.w1 = w; s.d1 = d;
  s.v2 = s.v1; s.v1 = v;
  return y;
}

// lg_sweep.cuh: mode_accumulate for both channels; fma(-yi, rot_im, sr) ==
// fma(yi, -rot_im, sr).
__device__ __forceinline__ void mode_accumulate2(float2& yr, float2& yi, const SweepParams& k,
                                                 const float2 sr, const float2 si) {
  const float2 nr = __ffma2_rn(yr, bc2(k.rot_re), __ffma2_rn(yi, bc2(-k.rot_im), sr));
  const float2 ni = __ffma2_rn(yr, bc2(k.rot_im), __ffma2_rn(yi, bc2(k.rot_re), si));
  yr = nr;
  yi = ni;
}

}  // namespace lg
