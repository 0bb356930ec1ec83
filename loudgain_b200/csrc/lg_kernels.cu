// lg_kernels.cu -- sm_100a kernels of the loudness path and their launchers.
//
//   sweep_kernel    fused K-weight + chunk energy + sample peak + true peak
//                   (what ebur128_add_frames_short does per call in the
//                   reference path, /root/reference/src/scan.c:448)
//   fixup_kernel    FP64 state carry + energy correction per chunk
//   slot_kernel     100 ms slot energies (channel-weighted)
//   block_kernel    400 ms gating blocks and 3 s short-term blocks
//   query_kernel    gated integrated loudness and loudness range over a set
//                   of tracks (ebur128_loudness_global[_multiple],
//                   ebur128_loudness_range[_multiple]: scan.c:294,297,383,388)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lg_common.h"
#include "lg_kernels.h"
#include "lg_post.cuh"
#include "lg_sweep.cuh"

namespace lg {

// ------------------------------------------------------------------ sweep

template <int FMT, int TPF>
__global__ void __launch_bounds__(kSweepThreads)
sweep_kernel(const Track* __restrict__ tracks, const CoefSet* __restrict__ coefs,
             const float* __restrict__ basis, const WarpWork* __restrict__ work,
             uint32_t nwarps, ChunkRec* __restrict__ recs, uint32_t* __restrict__ peaks) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (warp >= nwarps) return;
  const WarpWork ww = work[warp];
  const Track& tr = tracks[ww.track];
  const CoefSet& cs = coefs[tr.coef];
  const uint32_t C = tr.channels;
  const uint32_t ppc = (C + 1u) >> 1;
  const uint32_t cpw = 32u / ppc;
  const uint32_t slot = lane / ppc;
  const uint32_t pair = lane - slot * ppc;
  const uint32_t chunk = ww.first_chunk + slot;
  const bool active = slot < cpw && chunk < tr.nchunks;

  ChanOut out[2];
  out[0].sp = out[0].tp = out[1].sp = out[1].tp = 0.0f;
  const int ch0 = (int) (pair * 2u);
  const int nch = (ch0 + 1 < (int) C) ? 2 : 1;
  if (active) {
    GlobalSource<FMT> src;
    src.pcm = tr.pcm;
    src.frames = (long long) tr.frames;
    src.origin = (long long) chunk * cs.L - cs.W;
    src.channels = (int) C;
    src.ch0 = ch0;
    src.nch = nch;
    const long long left = (long long) tr.frames - (long long) chunk * cs.L;
    const int L_valid = left < cs.L ? (int) left : cs.L;
    sweep_chunk<TPF>(cs, basis + 2 * cs.basis_off, src, cs.L, L_valid, out);
    ChunkRec* r = recs + tr.rec_base + (uint64_t) chunk * C + ch0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < nch) {
        ChunkRec v;
        v.e0 = out[h].e0; v.xa = out[h].xa; v.xb = out[h].xb;
        v.pd = out[h].pd; v.pw = out[h].pw; v.qd = out[h].qd; v.qw = out[h].qw;
        r[h] = v;
      }
    }
  }
  // Peaks: non-negative floats order like their bit patterns.  Reduce over
  // the lanes of the warp that hold the same channel pair, one atomic each.
  const unsigned peers = __match_any_sync(0xffffffffu, active ? pair : 0xffffu);
  uint32_t v[4] = {__float_as_uint(out[0].sp), __float_as_uint(out[0].tp),
                   __float_as_uint(out[1].sp), __float_as_uint(out[1].tp)};
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __reduce_max_sync(peers, v[i]);
  if (active && lane == (uint32_t) (__ffs(peers) - 1)) {
    uint32_t* pk = peaks + 2 * (tr.peak_base + ch0);
    atomicMax(pk + 0, v[0]);
    atomicMax(pk + 1, v[1]);
    if (nch > 1) {
      atomicMax(pk + 2, v[2]);
      atomicMax(pk + 3, v[3]);
    }
  }
}

template <int FMT, int TPF>
static cudaError_t launch_sweep_t(const DeviceTables& t, uint32_t first_warp, uint32_t nwarps,
                                  cudaStream_t stream) {
  const uint32_t wpb = kSweepThreads / 32;
  const uint32_t blocks = (nwarps + wpb - 1) / wpb;
  sweep_kernel<FMT, TPF><<<blocks, kSweepThreads, 0, stream>>>(
      t.tracks, t.coefs, t.basis, t.work + first_warp, nwarps, t.recs, t.peaks);
  return cudaGetLastError();
}

cudaError_t launch_sweep(const DeviceTables& t, uint32_t format, int tpf, uint32_t first_warp,
                         uint32_t nwarps, cudaStream_t stream) {
  if (nwarps == 0) return cudaSuccess;
  if (format == FMT_S16) {
    if (tpf == 4) return launch_sweep_t<FMT_S16, 4>(t, first_warp, nwarps, stream);
    if (tpf == 2) return launch_sweep_t<FMT_S16, 2>(t, first_warp, nwarps, stream);
    return launch_sweep_t<FMT_S16, 0>(t, first_warp, nwarps, stream);
  }
  if (tpf == 4) return launch_sweep_t<FMT_F32, 4>(t, first_warp, nwarps, stream);
  if (tpf == 2) return launch_sweep_t<FMT_F32, 2>(t, first_warp, nwarps, stream);
  return launch_sweep_t<FMT_F32, 0>(t, first_warp, nwarps, stream);
}

// --------------------------------------------------------- post-processing

// Index of the last track whose `base` field is <= idx (tracks are laid out
// in increasing base order; empty tracks share a base with their successor,
// so step past them).
template <class GetBase>
__device__ uint32_t find_track(const Track* tracks, uint32_t ntracks, uint64_t idx, GetBase base) {
  uint32_t lo = 0, hi = ntracks;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (base(tracks[mid]) <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
fixup_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
             const CoefSet* __restrict__ coefs, const ChunkRec* __restrict__ recs,
             uint64_t total_recs, double* __restrict__ echunk) {
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total_recs) return;
  const uint32_t ti = find_track(tracks, ntracks, r, [](const Track& t) { return t.rec_base; });
  const Track& tr = tracks[ti];
  const CoefSet& cs = coefs[tr.coef];
  const uint64_t local = r - tr.rec_base;
  const uint64_t chunk = local / tr.channels;
  const uint32_t ch = (uint32_t) (local - chunk * tr.channels);
  if (chunk >= (uint64_t) tr.nslots * cs.k) return;   // tail chunks carry peaks only
  echunk[r] = chunk_true_energy(cs, recs + tr.rec_base + ch, tr.channels, (long long) chunk);
}

__global__ void __launch_bounds__(256)
slot_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
            const CoefSet* __restrict__ coefs, const double* __restrict__ echunk,
            uint64_t total_slots, double* __restrict__ eslot) {
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total_slots) return;
  const uint32_t ti = find_track(tracks, ntracks, s, [](const Track& t) { return t.slot_base; });
  const Track& tr = tracks[ti];
  eslot[s] = slot_energy(tr, coefs[tr.coef], echunk, (uint32_t) (s - tr.slot_base));
}

__global__ void __launch_bounds__(256)
block_kernel(const Track* __restrict__ tracks, uint32_t ntracks,
             const CoefSet* __restrict__ coefs, const double* __restrict__ eslot,
             uint64_t total_blocks, uint64_t total_st, double* __restrict__ zblock,
             double* __restrict__ zst) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total_blocks) {
    const uint32_t ti = find_track(tracks, ntracks, i, [](const Track& t) { return t.block_base; });
    const Track& tr = tracks[ti];
    zblock[i] = gating_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (i - tr.block_base));
  } else if (i < total_blocks + total_st) {
    const uint64_t j = i - total_blocks;
    const uint32_t ti = find_track(tracks, ntracks, j, [](const Track& t) { return t.st_base; });
    const Track& tr = tracks[ti];
    zst[j] = shortterm_block(eslot + tr.slot_base, coefs[tr.coef], (uint32_t) (j - tr.st_base));
  }
}

// ------------------------------------------------------------- reductions

constexpr int kQueryThreads = 512;

struct SumCount {
  double s;
  unsigned long long n;
};

// Deterministic CTA-wide (sum, count): fixed shuffle tree, then warp 0.
__device__ SumCount block_sum_count(double s, unsigned long long n, SumCount* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    n += __shfl_down_sync(0xffffffffu, n, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { scratch[wid].s = s; scratch[wid].n = n; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? scratch[lane].s : 0.0;
    n = lane < nw ? scratch[lane].n : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_down_sync(0xffffffffu, s, o);
      n += __shfl_down_sync(0xffffffffu, n, o);
    }
    if (lane == 0) { scratch[0].s = s; scratch[0].n = n; }
  }
  __syncthreads();
  const SumCount r = scratch[0];
  __syncthreads();
  return r;
}

// k-th smallest (0-based) among the member short-term energies >= floor_e.
// Positive doubles order like their 64-bit patterns: 8 passes of 8-bit radix
// select over the (L2-resident) values.
__device__ double select_kth(const BlockList* lists, const uint32_t* members, uint32_t count,
                             double floor_e, unsigned long long k,
                             unsigned int* hist, unsigned long long* shared_k,
                             unsigned long long* shared_prefix) {
  unsigned long long prefix = 0, mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (uint32_t m = 0; m < count; ++m) {
      const BlockList bl = lists[members[m]];
      for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x) {
        const double e = bl.st[i];
        if (!(e >= floor_e)) continue;
        const unsigned long long bits = (unsigned long long) __double_as_longlong(e);
        if ((bits & mask) != prefix) continue;
        atomicAdd(&hist[(bits >> shift) & 0xffull], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long acc = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (acc + hist[b] > k) break;
        acc += hist[b];
      }
      *shared_k = k - acc;
      *shared_prefix = prefix | ((unsigned long long) b << shift);
    }
    __syncthreads();
    k = *shared_k;
    prefix = *shared_prefix;
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return __longlong_as_double((long long) prefix);
}

__global__ void __launch_bounds__(kQueryThreads)
query_kernel(const BlockList* __restrict__ lists, const Query* __restrict__ queries,
             const uint32_t* __restrict__ members, double abs_gate,
             QueryResult* __restrict__ results) {
  __shared__ SumCount scratch[32];
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long sh_k, sh_prefix;
  const Query q = queries[blockIdx.x];
  const uint32_t* mem = members + q.first;
  QueryResult res;
  res.loudness = -HUGE_VAL; res.range = 0.0; res.rel_thr = 0.0;
  res.sum1 = res.sum2 = 0.0; res.n1 = res.n2 = res.nst = 0;

  // ---- integrated loudness: absolute gate, then relative gate at -10 LU
  double s = 0.0;
  unsigned long long n = 0;
  for (uint32_t m = 0; m < q.count; ++m) {
    const BlockList bl = lists[mem[m]];
    for (uint32_t i = threadIdx.x; i < bl.nz; i += blockDim.x) {
      const double e = bl.z[i];
      if (e >= abs_gate) { s += e; ++n; }
    }
  }
  SumCount a = block_sum_count(s, n, scratch);
  res.sum1 = a.s; res.n1 = a.n;
  if (a.n) {
    const double thr = a.s / (double) a.n * 0.1;
    res.rel_thr = thr;
    s = 0.0; n = 0;
    for (uint32_t m = 0; m < q.count; ++m) {
      const BlockList bl = lists[mem[m]];
      for (uint32_t i = threadIdx.x; i < bl.nz; i += blockDim.x) {
        const double e = bl.z[i];
        if (e >= abs_gate && e >= thr) { s += e; ++n; }
      }
    }
    SumCount b = block_sum_count(s, n, scratch);
    res.sum2 = b.s; res.n2 = b.n;
    if (b.n) res.loudness = energy_to_lufs(b.s / (double) b.n);
  }

  // ---- loudness range: -20 LU relative gate on short-term energies, then the
  //      10th / 95th percentile by rank
  s = 0.0; n = 0;
  for (uint32_t m = 0; m < q.count; ++m) {
    const BlockList bl = lists[mem[m]];
    for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x) {
      const double e = bl.st[i];
      if (e >= abs_gate) { s += e; ++n; }
    }
  }
  a = block_sum_count(s, n, scratch);
  res.nst = a.n;
  if (a.n) {
    double floor_e = a.s / (double) a.n * 0.01;
    if (floor_e < abs_gate) floor_e = abs_gate;
    s = 0.0; n = 0;
    for (uint32_t m = 0; m < q.count; ++m) {
      const BlockList bl = lists[mem[m]];
      for (uint32_t i = threadIdx.x; i < bl.nst; i += blockDim.x)
        if (bl.st[i] >= floor_e) ++n;
    }
    const SumCount c = block_sum_count(0.0, n, scratch);
    if (c.n) {
      const unsigned long long k_hi = (unsigned long long) ((double) (c.n - 1) * 0.95 + 0.5);
      const unsigned long long k_lo = (unsigned long long) ((double) (c.n - 1) * 0.1 + 0.5);
      const double hi = select_kth(lists, mem, q.count, floor_e, k_hi, hist, &sh_k, &sh_prefix);
      const double lo = select_kth(lists, mem, q.count, floor_e, k_lo, hist, &sh_k, &sh_prefix);
      res.range = energy_to_lufs(hi) - energy_to_lufs(lo);
    }
  }
  if (threadIdx.x == 0) results[blockIdx.x] = res;
}

cudaError_t launch_post(const DeviceTables& t, const PostSizes& z, cudaStream_t stream) {
  if (z.total_recs) {
    const unsigned blocks = (unsigned) ((z.total_recs + 255) / 256);
    fixup_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.recs, z.total_recs,
                                            t.echunk);
  }
  if (z.total_slots) {
    const unsigned blocks = (unsigned) ((z.total_slots + 255) / 256);
    slot_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.echunk, z.total_slots,
                                           t.eslot);
  }
  if (z.total_blocks + z.total_st) {
    const unsigned blocks = (unsigned) ((z.total_blocks + z.total_st + 255) / 256);
    block_kernel<<<blocks, 256, 0, stream>>>(t.tracks, z.ntracks, t.coefs, t.eslot, z.total_blocks,
                                            z.total_st, t.zblock, t.zst);
  }
  return cudaGetLastError();
}

cudaError_t launch_queries(const BlockList* lists, const Query* queries, const uint32_t* members,
                           uint32_t nqueries, double abs_gate, QueryResult* results,
                           cudaStream_t stream) {
  if (!nqueries) return cudaSuccess;
  query_kernel<<<nqueries, kQueryThreads, 0, stream>>>(lists, queries, members, abs_gate, results);
  return cudaGetLastError();
}

}  // namespace lg
