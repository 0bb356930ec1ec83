"""Batch measurement over PCM that is already resident in HBM
(include/ebur128_b200.h).  torch is used for device memory and streams only.

This is the path album / library scans take: all tracks of a
batch go through ONE fused sweep launch per kernel variant, then the FP64
fix-up and the gating / range reductions, all on the GPU; the host receives a
few scalars per track and per album -- the same quantities the reference
scanner reads back through ebur128_loudness_global[_multiple],
ebur128_loudness_range[_multiple] and ebur128_true_peak
(/root/reference/src/scan.c:294-307,383-391).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from collections.abc import Sequence

import numpy as np

from . import load_library

NO_ALBUM = 0xFFFFFFFF
FORMAT_S16, FORMAT_F32 = 0, 1


class _Track(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("frames", C.c_uint64), ("channels", C.c_uint32),
                ("samplerate", C.c_uint32), ("format", C.c_uint32), ("album", C.c_uint32),
                ("weight_class", C.c_void_p), ("lead_in", C.c_uint64), ("flags", C.c_uint32)]


class _Result(C.Structure):
    _fields_ = [("loudness", C.c_double), ("range", C.c_double), ("rel_threshold", C.c_double),
                ("sum_abs", C.c_double), ("sum_rel", C.c_double), ("n_abs", C.c_uint64),
                ("n_rel", C.c_uint64), ("n_shortterm", C.c_uint64)]


@dataclass
class Measurement:
    loudness: float
    range: float
    rel_threshold: float = 0.0
    sum_abs: float = 0.0
    sum_rel: float = 0.0
    n_abs: int = 0
    n_rel: int = 0
    n_shortterm: int = 0
    sample_peak: np.ndarray = field(default_factory=lambda: np.zeros(0))
    true_peak: np.ndarray = field(default_factory=lambda: np.zeros(0))


_RESULT_DTYPE = np.dtype([("loudness", "<f8"), ("range", "<f8"), ("rel_threshold", "<f8"),
                          ("sum_abs", "<f8"), ("sum_rel", "<f8"), ("n_abs", "<u8"), ("n_rel", "<u8"),
                          ("n_shortterm", "<u8")])
_NO_PEAKS = np.zeros(0)

class _Results(Sequence):
    """Per-track (or per-album) results of one fetch; Measurement objects are
    built on access."""

    def __init__(self, rows, sp, tp, offsets):
        self._rows, self._sp, self._tp, self._off = rows, sp, tp, offsets

    def __len__(self):
        return len(self._rows)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        row = self._rows[i].item()
        if self._off is None:
            return Measurement(*row, _NO_PEAKS, _NO_PEAKS)
        lo, hi = self._off[i], self._off[i + 1]
        return Measurement(*row, self._sp[lo:hi], self._tp[lo:hi])


_lib = None


def _bind():
    global _lib
    if _lib is not None:
        return _lib
    L = load_library().lib
    L.lgb_last_error.restype = C.c_char_p
    L.lgb_batch_create.argtypes = [C.POINTER(_Track), C.c_size_t, C.c_uint32, C.c_void_p]
    L.lgb_batch_create.restype = C.c_void_p
    L.lgb_batch_run.argtypes = [C.c_void_p]
    L.lgb_batch_fetch.argtypes = [C.c_void_p, C.POINTER(_Result), C.POINTER(_Result),
                                  C.c_void_p, C.c_void_p]
    L.lgb_batch_total_samples.argtypes = [C.c_void_p]
    L.lgb_batch_total_samples.restype = C.c_uint64
    L.lgb_batch_peak_count.argtypes = [C.c_void_p]
    L.lgb_batch_peak_count.restype = C.c_uint64
    L.lgb_batch_kernel_launches.argtypes = [C.c_void_p]
    L.lgb_batch_kernel_launches.restype = C.c_uint32
    L.lgb_batch_sweep_launches.argtypes = [C.c_void_p]
    L.lgb_batch_sweep_launches.restype = C.c_uint32
    L.lgb_batch_blocks.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.lgb_batch_blocks.restype = C.c_uint64
    L.lgb_batch_destroy.argtypes = [C.c_void_p]
    L.lgb_batch_destroy.restype = None
    _lib = L
    return L


def _err(L) -> str:
    return (L.lgb_last_error() or b"").decode()


class Batch:
    """A planned measurement of `tracks` = [(cuda tensor [frames, channels] of
    int16 or float32, sample rate)], optionally grouped into albums."""

    def __init__(self, tracks: Sequence, albums: Sequence[int] | None = None, stream=None,
                 lead_in: Sequence[int] | None = None, nalbums: int | None = None):
        """lead_in[i]: leading context frames of track i (a whole number of
        100 ms slots) when the track is one time segment of a longer stream."""
        import torch

        L = _bind()
        self._L = L
        self._keep = []
        n = len(tracks)
        arr = (_Track * max(n, 1))()
        nalb = int(nalbums) if nalbums is not None else 0
        self.channels = []
        for i, (pcm, rate) in enumerate(tracks):
            if not pcm.is_cuda:
                raise ValueError("Batch needs CUDA tensors (PCM resident in HBM)")
            if pcm.dim() == 1:
                pcm = pcm.view(-1, 1)
            pcm = pcm.contiguous()
            if pcm.data_ptr() % 16:
                # a row slice of a larger tensor: the sweep stages 16-byte units from frame 0
                pcm = pcm.clone()
            fmt = {torch.int16: FORMAT_S16, torch.float32: FORMAT_F32}[pcm.dtype]
            alb = NO_ALBUM if albums is None else int(albums[i])
            if alb != NO_ALBUM:
                nalb = max(nalb, alb + 1)
            self._keep.append(pcm)
            self.channels.append(pcm.shape[1])
            arr[i] = _Track(pcm.data_ptr(), pcm.shape[0], pcm.shape[1], int(rate), fmt, alb, None,
                            0 if lead_in is None else int(lead_in[i]), 0)
        self.ntracks, self.nalbums = n, nalb
        # result buffers are allocated once: fetch() sits between two steps of a
        # repeatedly run batch, where host time is GPU idle time
        self._tres = (_Result * max(n, 1))()
        self._ares = (_Result * max(nalb, 1))()
        self._npk = sum(self.channels)
        self._offsets = np.concatenate([[0], np.cumsum(self.channels)]).astype(np.int64)
        self._sp = np.zeros(max(self._npk, 1))
        self._tp = np.zeros(max(self._npk, 1))
        self._spp = self._sp.ctypes.data_as(C.c_void_p)
        self._tpp = self._tp.ctypes.data_as(C.c_void_p)
        self.stream = stream if stream is not None else torch.cuda.current_stream()
        self._h = L.lgb_batch_create(arr, n, nalb, C.c_void_p(self.stream.cuda_stream))
        if not self._h:
            raise RuntimeError("lgb_batch_create failed: " + _err(L))

    @property
    def total_samples(self) -> int:
        return self._L.lgb_batch_total_samples(self._h)

    @property
    def kernel_launches(self) -> int:
        return self._L.lgb_batch_kernel_launches(self._h)

    @property
    def sweep_launches(self) -> int:
        return self._L.lgb_batch_sweep_launches(self._h)

    def set_max_in_flight(self, n: int) -> None:
        """Runs that may be enqueued before the oldest is fetched (1..3, default 2)."""
        self._L.lgb_batch_set_max_in_flight.argtypes = [C.c_void_p, C.c_uint32]
        if self._L.lgb_batch_set_max_in_flight(self._h, n):
            raise RuntimeError("lgb_batch_set_max_in_flight failed: " + _err(self._L))

    def run(self) -> None:
        """Enqueue the whole measurement on the batch's stream (asynchronous)."""
        if self._L.lgb_batch_run(self._h):
            raise RuntimeError("lgb_batch_run failed: " + _err(self._L))

    def fetch(self) -> tuple[list[Measurement], list[Measurement]]:
        """Wait and read back per-track and per-album results."""
        L = self._L
        tres, ares, sp, tp = self._tres, self._ares, self._sp, self._tp
        if L.lgb_batch_fetch(self._h, tres, ares, self._spp, self._tpp):
            raise RuntimeError("lgb_batch_fetch failed: " + _err(L))

        # fetch() sits between two steps of a repeatedly run batch, where host
        # time is GPU idle time: snapshot the raw results (three small copies)
        # and build Measurement objects only for the entries that are looked at
        rows = np.frombuffer(tres, dtype=_RESULT_DTYPE, count=self.ntracks).copy()
        arows = np.frombuffer(ares, dtype=_RESULT_DTYPE, count=self.nalbums).copy()
        return (_Results(rows, sp.copy(), tp.copy(), self._offsets),
                _Results(arows, None, None, None))

    def blocks(self, track: int, kind: int = 0) -> np.ndarray:
        """Block energies of one track, copied to the host (0 = 400 ms gating
        blocks, 1 = 3 s short-term blocks, 2 = 100 ms slots).  Diagnostic."""
        # the block kernel of a repeatedly run batch is on the library's own stream
        self._L.lgb_batch_wait_blocks.argtypes = [C.c_void_p, C.c_void_p]
        if self._L.lgb_batch_wait_blocks(self._h, C.c_void_p(self.stream.cuda_stream)):
            raise RuntimeError("lgb_batch_wait_blocks failed: " + _err(self._L))
        self.stream.synchronize()
        return device_blocks(self, track, kind).cpu().numpy().copy()

    def close(self) -> None:
        if self._h:
            self._L.lgb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure(tracks, albums=None, stream=None):
    """One-shot convenience: plan, run, fetch."""
    b = Batch(tracks, albums, stream)
    try:
        b.run()
        return b.fetch()
    finally:
        b.close()


# ---------------------------------------------------------------- multi-GPU

class _DeviceView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


def device_blocks(batch: Batch, track: int, kind: int):
    """Block energies of one track as a CUDA tensor view (no copy)."""
    import torch

    p = C.c_void_p()
    n = batch._L.lgb_batch_blocks(batch._h, track, kind, C.byref(p))
    if not n:
        return torch.zeros(0, dtype=torch.float64, device="cuda")
    return torch.as_tensor(_DeviceView(p.value, n), device="cuda")


def gather_block_lists(dist, local: "torch.Tensor", world: int, group=None):
    """All-gathers variable-length float64 lists (one per rank) and returns
    them as a list of tensors.  Works on any backend / device (NCCL on GPUs,
    gloo on CPU in tests): sizes first, then one padded all_gather."""
    import torch

    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(max(sizes), 1)
    padded = torch.zeros(width, dtype=torch.float64, device=local.device)
    padded[:local.numel()] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def query_lists(z_lists, st_lists, stream=None) -> Measurement:
    """Gated loudness + range over the union of device-resident block lists
    (lgb_query_lists)."""
    import torch

    L = _bind()
    L.lgb_query_lists.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.c_size_t,
                                  C.c_void_p, C.POINTER(_Result)]
    n = len(z_lists)
    zp = (C.c_void_p * n)(*[t.data_ptr() if t.numel() else None for t in z_lists])
    sp = (C.c_void_p * n)(*[t.data_ptr() if t.numel() else None for t in st_lists])
    zn = (C.c_uint32 * n)(*[t.numel() for t in z_lists])
    sn = (C.c_uint32 * n)(*[t.numel() for t in st_lists])
    stream = stream if stream is not None else torch.cuda.current_stream()
    r = _Result()
    if L.lgb_query_lists(zp, zn, sp, sn, n, C.c_void_p(stream.cuda_stream), C.byref(r)):
        raise RuntimeError("lgb_query_lists failed: " + _err(L))
    return Measurement(r.loudness, r.range, r.rel_threshold, r.sum_abs, r.sum_rel, r.n_abs,
                       r.n_rel, r.n_shortterm)


class AlbumMerge:
    """Album result over tracks that are sharded across ranks.

    Every rank contributes the block lists (400 ms gating blocks and 3 s
    short-term blocks) of its local `tracks`; per step they are packed into one
    buffer, all-gathered over NCCL into a pre-allocated [world, width] tensor,
    and the library's gating / range kernel runs over the union on every rank
    (lgb_listquery_*).  Exact: the same block energies a single-GPU album query
    would see.  Sizes are exchanged once, at construction.
    """

    def __init__(self, batch: Batch, tracks, dist, world: int, group=None):
        import torch

        self.batch, self.dist, self.world, self.group = batch, dist, world, group
        self.tracks = list(tracks)
        # The merge runs on its own stream: it only waits for the batch's block
        # lists (lgb_batch_wait_blocks), not for its true-peak pass and queries.
        self.mstream = torch.cuda.Stream(device=torch.cuda.current_device())
        batch.stream.synchronize()
        self.z_views = [device_blocks(batch, t, 0) for t in self.tracks]
        self.st_views = [device_blocks(batch, t, 1) for t in self.tracks]
        nz = sum(v.numel() for v in self.z_views)
        nst = sum(v.numel() for v in self.st_views)
        dev = self.z_views[0].device if self.z_views else torch.device("cuda")
        mine = torch.tensor([nz, nst], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sizes, mine, group=group)
        self.sizes = [(int(s[0].item()), int(s[1].item())) for s in sizes]
        self.width = max(max(a + b for a, b in self.sizes), 1)
        self.send = torch.zeros(self.width, dtype=torch.float64, device=dev)
        self.recv = torch.zeros(world * self.width, dtype=torch.float64, device=dev)
        self.nz, self.nst = nz, nst
        # Pack plan: the library lays the block lists of consecutive tracks out
        # back to back, so adjacent views are merged into one device copy each.
        self._runs = []
        off = 0
        for v in self.z_views + self.st_views:
            n = v.numel()
            if n:
                if self._runs and self._runs[-1][0] + self._runs[-1][1] * 8 == v.data_ptr() \
                        and self._runs[-1][2] + self._runs[-1][1] == off:
                    self._runs[-1][1] += n
                else:
                    self._runs.append([v.data_ptr(), n, off])
            off += n
        self._run_views = [(torch.as_tensor(_DeviceView(p, n), device=dev), o) for p, n, o in self._runs]
        L = _bind()
        L.lgb_listquery_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                           C.c_size_t, C.c_void_p]
        L.lgb_listquery_create.restype = C.c_void_p
        L.lgb_listquery_run.argtypes = [C.c_void_p]
        L.lgb_listquery_fetch.argtypes = [C.c_void_p, C.POINTER(_Result)]
        L.lgb_listquery_destroy.argtypes = [C.c_void_p]
        L.lgb_listquery_destroy.restype = None
        base = self.recv.data_ptr()
        zp = (C.c_void_p * world)(*[base + r * self.width * 8 for r in range(world)])
        sp = (C.c_void_p * world)(*[base + (r * self.width + self.sizes[r][0]) * 8
                                    for r in range(world)])
        zn = (C.c_uint32 * world)(*[a for a, _ in self.sizes])
        sn = (C.c_uint32 * world)(*[b for _, b in self.sizes])
        self._L = L
        L.lgb_batch_wait_blocks.argtypes = [C.c_void_p, C.c_void_p]
        self._h = L.lgb_listquery_create(zp, zn, sp, sn, world,
                                         C.c_void_p(self.mstream.cuda_stream))
        if not self._h:
            raise RuntimeError("lgb_listquery_create failed: " + _err(L))

    def run(self) -> None:
        """Enqueue pack + all-gather + union query behind the batch's block
        kernels (call after batch.run(); fetch() before the next batch.run())."""
        import torch

        if self._L.lgb_batch_wait_blocks(self.batch._h, C.c_void_p(self.mstream.cuda_stream)):
            raise RuntimeError("lgb_batch_wait_blocks failed: " + _err(self._L))
        with torch.cuda.stream(self.mstream):
            for v, off in self._run_views:
                self.send[off:off + v.numel()].copy_(v, non_blocking=True)
            self.dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self._L.lgb_listquery_run(self._h):
            raise RuntimeError("lgb_listquery_run failed: " + _err(self._L))

    def fetch(self) -> Measurement:
        r = _Result()
        if self._L.lgb_listquery_fetch(self._h, C.byref(r)):
            raise RuntimeError("lgb_listquery_fetch failed: " + _err(self._L))
        return Measurement(r.loudness, r.range, r.rel_threshold, r.sum_abs, r.sum_rel, r.n_abs,
                           r.n_rel, r.n_shortterm)

    def close(self) -> None:
        if self._h:
            self._L.lgb_listquery_destroy(self._h)
            self._h = None


class AlbumExchange:
    """Album results over tracks that are sharded across ranks, computed inside
    the batch's own step (include/ebur128_b200.h: lgb_exchange_*).

    Album indices are global: every rank builds its Batch with the same
    `nalbums` (give Batch(..., nalbums=...) when a rank may hold no track of the
    last album).  Each rank reduces its own gating blocks; the kernels store the
    partial sums and the short-term energies straight into the peers' HBM over
    NVLink (CUDA IPC mappings of one region per rank) and wait on per-rank
    flags.  torch.distributed only carries the 64-byte handles and one size,
    once, here.  After this, batch.run() / batch.fetch() return the album
    results of ALL ranks' tracks, the same bits on every rank.
    """

    def __init__(self, batch: Batch, dist=None, world: int = 1, rank: int = 0, group=None):
        import torch

        L = _bind()
        L.lgb_exchange_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.lgb_exchange_create.restype = C.c_void_p
        L.lgb_exchange_handle.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.lgb_exchange_open.argtypes = [C.c_void_p, C.c_void_p]
        L.lgb_exchange_destroy.argtypes = [C.c_void_p]
        L.lgb_exchange_destroy.restype = None
        L.lgb_batch_album_shortterm_blocks.argtypes = [C.c_void_p]
        L.lgb_batch_album_shortterm_blocks.restype = C.c_uint64
        L.lgb_batch_attach_exchange.argtypes = [C.c_void_p, C.c_void_p]
        self._L, self.batch, self.world, self.rank = L, batch, world, rank
        dev = torch.device("cuda", torch.cuda.current_device())
        cap = torch.tensor([L.lgb_batch_album_shortterm_blocks(batch._h)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self._h = L.lgb_exchange_create(world, rank, batch.nalbums, int(cap.item()))
        if not self._h:
            raise RuntimeError("lgb_exchange_create failed: " + _err(L))
        if world > 1:
            mine = (C.c_ubyte * 64)()
            if L.lgb_exchange_handle(self._h, mine, 64):
                raise RuntimeError("lgb_exchange_handle failed: " + _err(L))
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            out = torch.empty(world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, t, group=group)
            raw = bytes(out.cpu().numpy().tobytes())
            if L.lgb_exchange_open(self._h, raw):
                raise RuntimeError("lgb_exchange_open failed: " + _err(L))
            dist.barrier(group=group)            # every region is mapped everywhere before anyone runs
        if L.lgb_batch_attach_exchange(batch._h, self._h):
            raise RuntimeError("lgb_batch_attach_exchange failed: " + _err(L))

    def close(self) -> None:
        if self._h:
            self._L.lgb_exchange_destroy(self._h)
            self._h = None


def lpt_assign(costs, world: int):
    """Rank of every work item, longest processing time first
    (lgb_lpt_assign; cost = frames x channels of a track)."""
    L = _bind()
    L.lgb_lpt_assign.argtypes = [C.POINTER(C.c_uint64), C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
    L.lgb_lpt_assign.restype = C.c_uint64
    n = len(costs)
    c = (C.c_uint64 * max(n, 1))(*[int(x) for x in costs])
    out = (C.c_uint32 * max(n, 1))()
    worst = L.lgb_lpt_assign(c, n, world, out)
    return [int(out[i]) for i in range(n)], int(worst)


def merge_album_across_ranks(batch: Batch, tracks, dist, world: int) -> Measurement:
    """One-shot form of AlbumMerge."""
    m = AlbumMerge(batch, tracks, dist, world)
    try:
        m.run()
        return m.fetch()
    finally:
        m.close()


# ------------------------------------------------- one stream, sharded by time

def segment_plan(frames: int, rate: int, parts: int, lead_in_slots: int = 10):
    """Cuts a stream of `frames` frames into `parts` contiguous time segments
    (BASELINE config 4).  Segments start on whole seconds (10
    slots of 100 ms), so that neither a slot nor a 1 s short-term hop straddles
    two segments; every segment but the first is preceded by `lead_in_slots`
    slots of the audio before it.  Returns [(first_frame_incl_lead_in,
    lead_in_frames, end_frame)] -- segment i measures frames
    [first + lead_in, end) and reads [first, end).

    The lead-in replaces an exchange of filter state between neighbours: a
    rank starts its filters at zero state one second early, and the K-filter's
    slowest mode (|pole| <= 0.9987 at 192 kHz) has decayed below 1e-18 of the
    state it started from long before the segment's own first frame."""
    s100 = (rate + 5) // 10
    sec = 10 * s100
    nsec = max(frames // sec, 1)
    cuts = [min((nsec * i // parts) * sec, frames) for i in range(parts)] + [frames]
    plan = []
    for i in range(parts):
        lead = min(lead_in_slots * s100, cuts[i]) if i else 0
        plan.append((cuts[i] - lead, lead, cuts[i + 1]))
    return plan


def slots_query(slots: "torch.Tensor", rate: int, stream=None) -> Measurement:
    """Gated loudness + range of one stream from its complete list of 100 ms
    slot energies on the device (lgb_slots_query)."""
    import torch

    L = _bind()
    L.lgb_slots_query.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(_Result)]
    stream = stream if stream is not None else torch.cuda.current_stream()
    slots = slots.contiguous()
    r = _Result()
    if L.lgb_slots_query(C.c_void_p(slots.data_ptr() if slots.numel() else None), slots.numel(),
                         (int(rate) + 5) // 10, C.c_void_p(stream.cuda_stream), C.byref(r)):
        raise RuntimeError("lgb_slots_query failed: " + _err(L))
    return Measurement(r.loudness, r.range, r.rel_threshold, r.sum_abs, r.sum_rel, r.n_abs,
                       r.n_rel, r.n_shortterm)


class StreamShard:
    """This rank's part of ONE stream that is sharded by time over `world`
    ranks (BASELINE config 4).

    `segments`: the rank's segments in time order, [(cuda tensor [frames,
    channels] holding lead-in + segment, lead_in_frames)], cut by segment_plan;
    rank r holds the segments that precede rank r+1's.  Planned once; every
    run() sweeps the local segments in one batch, packs their 100 ms slot
    energies (lead-in slots dropped), all-gathers them over NCCL/NVLink into a
    pre-allocated buffer in rank = time order, and fetch() forms blocks, gates
    and the range over the whole slot list (lgb_slots_query) and reduces the
    per-channel peaks with a MAX all-reduce.  The result is the same on every
    rank and equals the single-device measurement of the whole stream.
    """

    def __init__(self, segments, rate: int, dist=None, world: int = 1, group=None, stream=None):
        import torch

        self.rate, self.dist, self.world, self.group = int(rate), dist, world, group
        self.s100 = (self.rate + 5) // 10
        self.segments = list(segments)
        self.batch = Batch([(p, rate) for p, _ in self.segments], None, stream,
                           lead_in=[l for _, l in self.segments])
        self.stream = self.batch.stream
        self.dev = self.segments[0][0].device if self.segments else torch.device("cuda")
        self.nch = self.segments[0][0].shape[1] if self.segments else 0
        self.views = None
        nloc = sum((p.shape[0] - lead) // self.s100 for p, lead in self.segments)
        mine = torch.tensor([nloc], dtype=torch.int64, device=self.dev)
        sizes = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(sizes, mine, group=group)
            self.sizes = [int(x.item()) for x in sizes]
        else:
            self.sizes = [nloc]
        self.width = max(max(self.sizes), 1)
        self.send = torch.zeros(self.width, dtype=torch.float64, device=self.dev)
        self.recv = torch.zeros(world * self.width, dtype=torch.float64, device=self.dev)
        self.union = torch.zeros(max(sum(self.sizes), 1), dtype=torch.float64, device=self.dev)
        self.peaks = torch.zeros(2, max(self.nch, 1), dtype=torch.float64, device=self.dev)

    def run(self) -> None:
        import torch

        self.batch.run()
        if self.views is None:                   # device views of the slot lists, lead-in dropped
            self.views = [device_blocks(self.batch, i, 2)[lead // self.s100:]
                          for i, (_, lead) in enumerate(self.segments)]
        # the slot energies come from the fix-up kernel, which a repeatedly run batch keeps on
        # the library's own stream: order this stream behind the run's block kernel
        self.batch._L.lgb_batch_wait_blocks.argtypes = [C.c_void_p, C.c_void_p]
        if self.batch._L.lgb_batch_wait_blocks(self.batch._h, C.c_void_p(self.stream.cuda_stream)):
            raise RuntimeError("lgb_batch_wait_blocks failed: " + _err(self.batch._L))
        with torch.cuda.stream(self.stream):
            off = 0
            for v in self.views:
                self.send[off:off + v.numel()].copy_(v, non_blocking=True)
                off += v.numel()
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
                off = 0
                for r, n in enumerate(self.sizes):
                    self.union[off:off + n].copy_(self.recv[r * self.width:r * self.width + n],
                                                  non_blocking=True)
                    off += n
            else:
                self.union[:self.sizes[0]].copy_(self.send[:self.sizes[0]], non_blocking=True)

    def fetch(self) -> Measurement:
        import torch

        tres, _ = self.batch.fetch()
        nch = self.nch
        sp = np.max([m.sample_peak for m in tres], axis=0) if tres else np.zeros(nch)
        tp = np.max([m.true_peak for m in tres], axis=0) if tres else np.zeros(nch)
        with torch.cuda.stream(self.stream):
            self.peaks[:, :nch] = torch.as_tensor(np.stack([sp, tp]), device=self.dev)
            if self.world > 1:
                self.dist.all_reduce(self.peaks, op=self.dist.ReduceOp.MAX, group=self.group)
            out = slots_query(self.union[:sum(self.sizes)], self.rate, self.stream)
            pk = self.peaks.cpu().numpy()
        out.sample_peak, out.true_peak = pk[0, :nch].copy(), pk[1, :nch].copy()
        return out

    def close(self) -> None:
        self.views = None
        self.batch.close()


def measure_stream_segments(segments, rate: int, dist=None, world: int = 1, group=None,
                            stream=None) -> Measurement:
    """One-shot form of StreamShard."""
    s = StreamShard(segments, rate, dist, world, group, stream)
    try:
        s.run()
        return s.fetch()
    finally:
        s.close()


# ------------------------------------------------------------- scan.c driver

class _HostTrack(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("frames", C.c_uint64), ("channels", C.c_uint32),
                ("samplerate", C.c_uint32), ("format", C.c_uint32)]


class ScanResult(C.Structure):
    """Mirror of the reference's scan_result (scan.h:35-53)."""
    _fields_ = [(n, C.c_double) for n in (
        "track_gain", "track_peak", "track_loudness", "track_loudness_range", "album_gain",
        "album_peak", "album_loudness", "album_loudness_range", "loudness_reference")]


def scan_host(tracks, chunk_frames=4096, do_album=True, pre_gain=0.0, threads=1):
    """scan.c's call sequence over host PCM (numpy int16 / float32 arrays
    [frames, channels]) through the drop-in ebur128_* ABI: one
    ebur128_add_frames call per `chunk_frames` frames, `threads` scanner
    threads (one file each at a time), queries after all files.  Returns one
    ScanResult per track."""
    import numpy as np
    L = _bind()
    L.lgb_scan_host_mt.argtypes = [C.POINTER(_HostTrack), C.c_size_t, C.c_size_t, C.c_int,
                                   C.c_double, C.c_uint, C.POINTER(ScanResult)]
    keep = [np.ascontiguousarray(p) for p, _ in tracks]
    arr = (_HostTrack * len(tracks))()
    for i, (p, (_, rate)) in enumerate(zip(keep, tracks)):
        assert p.dtype in (np.int16, np.float32) and p.ndim == 2
        arr[i] = _HostTrack(p.ctypes.data, p.shape[0], p.shape[1], rate, 0 if p.dtype == np.int16 else 1)
    out = (ScanResult * len(tracks))()
    rc = L.lgb_scan_host_mt(arr, len(tracks), chunk_frames, 1 if do_album else 0, pre_gain, threads, out)
    if rc:
        raise RuntimeError(f"lgb_scan_host_mt failed ({rc}): {_err(L)}")
    return list(out)
