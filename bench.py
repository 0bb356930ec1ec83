#!/usr/bin/env python3
"""bench.py -- headline benchmark of the loudness hot path (BASELINE.json).

Metric: PCM Gsamples/s for the fused K-weight + gate + true-peak measurement,
whole job over N GPUs.  Workload at every N: BASELINE.json configs[1], a
synthetic 12-track 44.1 kHz stereo 16-bit album per GPU (album mode: per-track
and album results), generated on the device.  One "step" = one complete
measurement of the rank's album: sweep, FP64 fix-up, gating, range, peaks,
and the read-back of the per-track / per-album scalars.

  value     whole-job Gsamples/s with the PCM already resident in HBM
  e2e       the same album through the drop-in C ABI (ebur128_init /
            add_frames_short / queries, driven like scan.c) from pinned HOST
            buffers, copies inside the timed region
  roofline  the sweep kernel's algorithmic bytes (2 B per S16 sample, read
            once) / its mean launch time (CUDA events on the launching
            stream) against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (restatement of the reference's libebur128
            path) on one host core over a bounded sample of the same album

  configs   the other BASELINE.json shapes on the same GPUs, each with Gsamples/s
            and the sweep's fraction of the HBM roofline: cfg1 (one track), cfg3
            (96 kHz 5.1 float), cfg4 (one 10-hour stream, time-sharded over the
            ranks), cfg5 (library slice, tracks dealt out by LPT, all albums merged
            in one exchange), and cfg2 hard-clipped (worst case of the true-peak pass)

At N > 1 every rank scans its own 12 tracks and ONE album spans all 12 N tracks:
the album's gating runs inside the step's CUDA graph over NVLink peer memory
(engine.AlbumExchange); after the timed region the block lists are gathered to
rank 0 and the merged result is checked against a numpy gating of the union.

`--impl reference` times the reference's CPU path instead (the oracle port:
libebur128 itself is not in /root/reference), on the host cores, one worker
per track like bin/rgbpm2 -- the same album, the same 1024-frame calls.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pcm_gsamples_per_s_kweight_gate_truepeak"
UNIT = "Gsamples/s"
WORKLOAD = "cfg2: 12-track 44.1 kHz stereo S16 album per GPU, album mode (-a -k quantities)"


FEED_FRAMES = 1024          # frames per ebur128_add_frames_short call in BOTH arms (scan.c:448: one AVFrame)


def _ncu_traffic():
    """DRAM bytes per sweep launch from the committed ncu capture of this workload."""
    for name in ("r02_sweep_ncu_summary.json", "r01_sweep_ncu_summary.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            return float(json.load(open(path))["traffic_bytes_per_launch"]), "profiles/" + name
        except Exception:
            continue
    return None, None


def album_specs(rank: int):
    from loudgain_b200 import synth
    specs = synth.config2_specs(12)
    for s in specs:
        s.seed += 1000 * rank            # every rank scans a different album
    return specs


def scanner_threads(world: int, ntracks: int = 12) -> int:
    """One scanner thread per track, at most this rank's share of the host cores
    (both arms: the reference parallelises over files, bin/rgbpm2:170)."""
    return max(1, min(ntracks, usable_cores() // max(world, 1)))


def bench_config(world: int):
    """The `config` object -- identical in both arms."""
    samples = sum(2 * s.frames for s in album_specs(0))
    return {"workload": WORKLOAD, "samples_per_gpu": samples, "pcm_bytes_per_gpu": 2 * samples,
            "feed": f"ebur128_add_frames_short, {FEED_FRAMES}-frame calls from host PCM (scan.c:448), one scanner "
                    "thread per track up to the host cores per rank",
            "l2_policy": "input (495 MB per GPU) is larger than L2 (126 MB); no flush",
            "step_overlap": ("timed steps are enqueued three deep (runs k + 1 and k + 2 before fetch k) and "
                             "pipelined on the GPU (the post-processing of run k finishes under the sweep of "
                             "run k + 1); every step's results are read back and fetched inside the timed region"
                             if world == 1 or os.environ.get("LOUDGAIN_B200_PIPELINE_EXCHANGE", "0") != "0" else
                             "timed steps are enqueued three deep (runs k + 1 and k + 2 before fetch k), one CUDA "
                             "graph per step on one stream (with the album exchange attached the runs are not "
                             "pipelined on the GPU unless LOUDGAIN_B200_PIPELINE_EXCHANGE=1); every step's results "
                             "are read back and fetched inside the timed region"),
            "clock_sampling": f"{CLOCK_LOAD_STEPS[0]} + {CLOCK_LOAD_STEPS[1]} untimed steps of the same load "
                              "around the timed steps, nvidia-smi every 100 ms",
            "sharding": "by track; one album over all ranks' tracks, gated inside the step over NVLink peer "
                        "memory (lgb_exchange)" if world > 1 else "single GPU"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# untimed steps of the same load before / after the timed region (clock sampling)
CLOCK_LOAD_STEPS = (1500, 800)


# stdout carries the JSON line and nothing else: everything libraries print there
# (NCCL's version banner, for one) is sent to stderr, see main().
_JSON_OUT = None


def emit(obj) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def usable_cores() -> int:
    """Host cores this process may use: the affinity mask, capped by a cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def samples(self) -> int:
        return len(self.rows)

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_album(device, rank: int, fmt: str = "s16"):
    """cfg2 on `device`: list of (int16 tensor [frames, 2], rate)."""
    from loudgain_b200 import synth
    specs = album_specs(rank)
    if fmt == "f32":                     # tuning only: the float API's layout
        return [(synth.programme_float(s, device=device), s.rate) for s in specs]
    return [(synth.programme_s16(s, device=device), s.rate) for s in specs]


# ------------------------------------------------------------------ CPU arms

def _oracle_scan(lib, pcm_list, rate_list, chunk=FEED_FRAMES, threads=1):
    """scan.c's sequence on the oracle; returns seconds."""
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        st = lib.init(pcm_list[i].shape[1], rate_list[i])
        st.add_frames(pcm_list[i], chunk)
        r = (st.loudness_global(), st.loudness_range(), max(st.true_peaks()))
        return st, r

    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:       # ctypes drops the GIL inside the C calls
            done = list(ex.map(one, range(len(pcm_list))))
    else:
        done = [one(i) for i in range(len(pcm_list))]
    states = [d[0] for d in done]
    lib.loudness_global_multiple(states)
    lib.loudness_range_multiple(states)
    dt = time.perf_counter() - t0
    for st in states:
        st.destroy()
    return dt


def reference_arm(args):
    """The reference's CPU path on the host cores (rank 0 only): the whole album,
    the same call granularity and thread policy as the GPU arm's e2e leg."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from loudgain_b200 import synth
    from oracle import load_oracle
    lib = load_oracle()
    cores = usable_cores()
    # the same PCM as the GPU arm's rank 0: synthesised on the GPU when there is one (not
    # timed; bit-identical input for both arms), on the host otherwise
    try:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "cpu"
    pcm, rates = [], []
    for s in album_specs(0):
        pcm.append(synth.programme_s16(s, device=dev).cpu().numpy())
        rates.append(s.rate)
    samples = sum(p.size for p in pcm)
    threads = scanner_threads(1)            # rank 0 alone runs: it may use every core
    for _ in range(min(args.warmup, 1)):
        _oracle_scan(lib, pcm, rates, threads=threads)
    times = [_oracle_scan(lib, pcm, rates, threads=threads) for _ in range(args.steps)]
    total = sum(times)
    value = samples * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "the whole 12-track album per step (247.6 M samples), one worker thread per "
                                   f"track (rgbpm2 model), {FEED_FRAMES}-frame calls; warm-up capped at one pass"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": cores,
    }
    emit(line)


def cpu_baseline_leg(album_host, rates):
    """One core, bounded sample (about 10-20 s of CPU work)."""
    from oracle import load_oracle
    lib = load_oracle()
    budget = 120_000_000            # samples: ~12 s at ~10 Msamples/s
    pcm, rr, n = [], [], 0
    for p, r in zip(album_host, rates):
        if n >= budget:
            break
        take = min(p.shape[0], (budget - n) // p.shape[1])
        pcm.append(p[:take]); rr.append(r); n += take * p.shape[1]
    dt = _oracle_scan(lib, pcm, rr, threads=1)
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{len(pcm)} tracks / {n} samples of the same album, 1024-frame calls, "
                      "single thread (loudgain is single-threaded)"}


# ------------------------------------------------------------------- GPU arm

class HostTrack(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("frames", C.c_uint64), ("channels", C.c_uint32),
                ("samplerate", C.c_uint32), ("format", C.c_uint32)]


class ScanResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "track_gain", "track_peak", "track_loudness", "track_loudness_range", "album_gain",
        "album_peak", "album_loudness", "album_loudness_range", "loudness_reference")]


def _numpy_gating(z, st):
    """Gated loudness and range of block-energy lists (numpy; the oracle's rules,
    oracle/ebur128_oracle.c: gated_loudness / loudness_range)."""
    import numpy as np
    abs_gate = 10.0 ** ((-70.0 + 0.691) / 10.0)
    za = z[z >= abs_gate]
    loud = -np.inf
    if za.size:
        thr = za.sum() / za.size * 0.1
        zr = za[za >= thr]
        if zr.size:
            loud = 10.0 * np.log10(zr.sum() / zr.size) - 0.691
    rng = 0.0
    sa = st[st >= abs_gate]
    if sa.size:
        floor = max(sa.sum() / sa.size * 0.01, abs_gate)
        sr = np.sort(sa[sa >= floor])
        if sr.size:
            lo = sr[int((sr.size - 1) * 0.1 + 0.5)]
            hi = sr[int((sr.size - 1) * 0.95 + 0.5)]
            rng = 10.0 * np.log10(hi) - 10.0 * np.log10(lo)
    return float(loud), float(rng)


class Timer:
    """K steps between CUDA events on the launching stream, barrier + synchronize on
    both sides, max over ranks."""

    def __init__(self, torch, dist, world, dev, stream):
        self.torch, self.dist, self.world, self.dev, self.stream = torch, dist, world, dev, stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, step, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            out = step()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, out

    def run_deep(self, batch, steps, warmup, depth):
        """The same for a repeatedly run batch, `depth` runs in flight (every run's results are
        fetched inside the timed region)."""
        torch = self.torch
        batch.set_max_in_flight(depth)
        for _ in range(warmup):
            batch.run()
            out = batch.fetch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        inflight = 0
        for _ in range(steps):
            batch.run()
            inflight += 1
            if inflight == depth:
                out = batch.fetch()
                inflight -= 1
        while inflight:
            out = batch.fetch()
            inflight -= 1
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, out


def _timing_api(L):
    L.lgb_batch_enable_timing.argtypes = [C.c_void_p, C.c_int]
    L.lgb_batch_sweep_ms.argtypes = [C.c_void_p]
    L.lgb_batch_sweep_ms.restype = C.c_double
    L.lgb_batch_truepeak_ms.argtypes = [C.c_void_p]
    L.lgb_batch_truepeak_ms.restype = C.c_double
    L.lgb_batch_truepeak_candidates.argtypes = [C.c_void_p]
    L.lgb_batch_truepeak_candidates.restype = C.c_uint64


def _sweep_ms(L, batch, step, steps):
    """Mean duration of the batch's sweep launches (and of its true-peak pass), CUDA
    events on the launching stream inside lgb_batch_run."""
    L.lgb_batch_enable_timing(batch._h, 1)
    for _ in range(steps):
        step()
    out = L.lgb_batch_sweep_ms(batch._h), L.lgb_batch_truepeak_ms(batch._h)
    L.lgb_batch_enable_timing(batch._h, 0)
    return out


def extra_configs(args, torch, dist, timer, L, dev, rank, world, hbm):
    """The other BASELINE.json shapes on the same GPUs (parity of each is a test case
    in tests/test_gpu_parity.py; these are the throughput lines)."""
    from loudgain_b200 import engine, synth
    stream = torch.cuda.current_stream()
    steps, warmup = max(3, min(args.steps, 10)), 3
    out = {}

    def line(name, desc, batch, step, samples_all, bytes_local, sharding, more=None, deep=True):
        # (batches: three runs in flight, as in the headline; the time-sharded stream runs its
        # torch / NCCL tail between the steps, one step at a time)
        ms, res = timer.run_deep(batch, steps, max(warmup, 4), 3) if deep else timer.run(step, steps, warmup)
        sweep_ms, tp_ms = _sweep_ms(L, batch, step, steps)
        sweep_ms = timer.max_over_ranks(sweep_ms)
        o = {"workload": desc, "value": samples_all / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
             "samples": samples_all, "sweep_ms": sweep_ms, "truepeak_pass_ms": tp_ms,
             "sweep_launches": batch.sweep_launches, "kernel_launches_per_step": batch.kernel_launches,
             "sweep_hbm_gbs": bytes_local / (sweep_ms * 1e-3) / 1e9 if sweep_ms else None,
             "frac": bytes_local / (sweep_ms * 1e-3) / 1e9 / hbm if sweep_ms else None,
             "sharding": sharding}
        if more:
            o.update(more(res))
        out[name] = o

    def total(x: int) -> int:
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    # ---- cfg1: one 3-minute 44.1 kHz stereo S16 track per GPU
    spec = synth.config1_spec(180.0)
    spec.seed += rank
    pcm = synth.programme_s16(spec, device=dev)
    b = engine.Batch([(pcm, spec.rate)], None, stream)
    line("cfg1", "one 3-min 44.1 kHz stereo S16 track per GPU (15.9 M samples: 0.43 of one wave of the sweep)",
         b, lambda: (b.run(), b.fetch())[1], total(pcm.numel()), pcm.numel() * 2,
         "replicas: one independent track per GPU" if world > 1 else "single GPU",
         lambda r: {"loudness": r[0][0].loudness})
    b.close()
    del pcm, b

    # ---- cfg3: 96 kHz 5.1 float tracks (BS.1770 weights, 2x true peak, range)
    tracks = []
    for i in range(8):
        sp = synth.config3_spec(120.0)
        sp.seed += 10 * rank + i
        tracks.append((synth.programme_float(sp, device=dev), sp.rate))
    b = engine.Batch(tracks, [0] * len(tracks), stream)
    n = sum(t.numel() for t, _ in tracks)
    line("cfg3", "eight 2-min 96 kHz 5.1 float32 tracks per GPU, one album (channel weights, 2x true peak, range)",
         b, lambda: (b.run(), b.fetch())[1], total(n), n * 4,
         "replicas: one independent album per GPU" if world > 1 else "single GPU",
         lambda r: {"album_loudness": r[1][0].loudness, "album_range": r[1][0].range})
    b.close()
    del tracks, b

    # ---- cfg2, hard-clipped: every loud passage sits at full scale, so the true-peak
    # screening cannot rule anything out there (worst case of the candidate pass)
    album = make_album(dev, rank)
    clipped = [((t.to(torch.int32) * 6).clamp_(-32768, 32767).to(torch.int16), r) for t, r in album]
    del album
    b = engine.Batch(clipped, [0] * len(clipped), stream)
    n = sum(t.numel() for t, _ in clipped)
    line("cfg2_clipped", "the cfg2 album amplified by 15.6 dB and hard-clipped at full scale (true-peak worst case)",
         b, lambda: (b.run(), b.fetch())[1], total(n), n * 2,
         "replicas: one independent album per GPU" if world > 1 else "single GPU",
         lambda r: {"truepeak_candidate_frac": L.lgb_batch_truepeak_candidates(b._h) / (n / 24.0),
                    "max_true_peak": float(max(m.true_peak.max() for m in r[0]))})
    b.close()
    del clipped, b
    torch.cuda.empty_cache()

    # ---- cfg4: ONE 10-hour 48 kHz stereo stream, time-sharded over the ranks
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_cfg4
    seconds = 36000.0
    tot = int(seconds * bench_cfg4.RATE)
    first, lead, end = engine.segment_plan(tot, bench_cfg4.RATE, world)[rank]
    seg = bench_cfg4.stream_part(first, end, tot, dev)
    shard = engine.StreamShard([(seg, lead)], bench_cfg4.RATE, dist if world > 1 else None, world)
    line("cfg4", "one 10-hour 48 kHz stereo S16 stream (3.456 G samples in all)",
         shard.batch, lambda: (shard.run(), shard.fetch())[1], 2 * tot, seg.numel() * 2,
         "by time: whole-second segments, 1 s lead-in instead of an IIR state exchange (SURVEY 8e permits a "
         "warm-up halo); 100 ms slot energies all-gathered (NCCL), peaks MAX-all-reduced, blocks / gates / range "
         "over the whole slot list on every rank" if world > 1 else "single GPU",
         lambda m: {"loudness": m.loudness, "range": m.range}, deep=False)
    shard.close()
    del seg, shard
    torch.cuda.empty_cache()

    # ---- cfg5: library slice, 1250 tracks per GPU of the 10 000-track library's
    # generator (durations x 0.25 so that synthesis stays short), dealt out by LPT
    per_gpu, scale = 1250, 0.25
    specs, albums = synth.config5_specs(per_gpu * world, scale=scale)
    nalb = max(albums) + 1
    owner, worst = engine.lpt_assign([s.frames * s.channels for s in specs], world)
    mine = [i for i in range(len(specs)) if owner[i] == rank]
    tracks = [(synth.programme_s16(specs[i], device=dev), specs[i].rate) for i in mine]
    b = engine.Batch(tracks, [albums[i] for i in mine], stream, nalbums=nalb)
    x = engine.AlbumExchange(b, dist, world, rank) if world > 1 else None
    n = sum(t.numel() for t, _ in tracks)
    nall = sum(s.frames * s.channels for s in specs)
    line("cfg5", f"library slice: {per_gpu} tracks per GPU ({len(specs)} tracks, {nalb} albums, mixed "
                 f"22.05-192 kHz, 1/2/6 channels, S16; durations x {scale})",
         b, lambda: (b.run(), b.fetch())[1], nall, n * 2,
         f"by track, longest-processing-time-first (lgb_lpt_assign, worst rank load {worst / (nall / world):.4f} of "
         "the mean); all albums gated in one exchange over NVLink peer memory" if world > 1 else "single GPU",
         lambda r: {"albums": nalb, "album0_loudness": r[1][0].loudness})
    b.close()
    if x is not None:
        x.close()
    del tracks, b
    torch.cuda.empty_cache()
    return out


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from loudgain_b200 import build, engine

    build()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    album = make_album(dev, rank, args.format)
    stream = torch.cuda.current_stream()
    timer = Timer(torch, dist, world, dev, stream)
    barrier = timer.barrier
    # One album over ALL ranks' tracks (album index 0 everywhere): with more than one
    # rank its gating runs inside the step, over NVLink peer memory.
    batch = engine.Batch(album, [0] * len(album), stream)
    xchg = engine.AlbumExchange(batch, dist, world, rank) if world > 1 else None
    L = engine._bind()
    _timing_api(L)
    samples = batch.total_samples
    pcm_bytes = sum(t.numel() * t.element_size() for t, _ in album)

    def step():
        batch.run()
        return batch.fetch()

    # runs in flight in the timed region: run k + DEPTH - 1 is enqueued before run k is fetched
    DEPTH = 3
    batch.set_max_in_flight(DEPTH)

    # ---- resident-PCM throughput (steps replay the batch's CUDA graph)
    # The timed region is a few milliseconds, far shorter than nvidia-smi's
    # start-up and sampling period.  So the sampler starts first and the same
    # step keeps running untimed before and after the timed region (a fixed
    # count on every rank: the multi-GPU step waits for its peers); every clock
    # sample is then taken under the load the timed steps run in.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        for _ in range(args.warmup + CLOCK_LOAD_STEPS[0]):
            step()
        if world == 1:
            # nvidia-smi's query (every 100 ms) holds up kernel launches for about a millisecond, a
            # third of the timed region: keep the load running until a sample has just arrived,
            # so that the next one is 100 ms away.  (With several ranks the steps are collective
            # -- the album exchange -- and every rank runs the fixed count above.)
            seen, t_end = clk.samples(), time.perf_counter() + 0.4
            while clk.samples() == seen and time.perf_counter() < t_end:
                step()
        barrier()
        # The K timed steps run three deep: steps k + 1 and k + 2 are enqueued before the
        # results of step k are fetched (lgb_batch_run keeps three result mirrors), so the
        # host's turn-around between steps -- and, with several GPUs, the wait for the slowest
        # rank of the album exchange -- overlaps the GPU; every step's results are read.
        e0.record(stream)
        inflight = 0
        for _ in range(args.steps):
            batch.run()
            inflight += 1
            if inflight == DEPTH:
                tres, ares = batch.fetch()
                inflight -= 1
        while inflight:
            tres, ares = batch.fetch()
            inflight -= 1
        e1.record(stream)
        barrier()
        for _ in range(CLOCK_LOAD_STEPS[1]):
            step()
        barrier()
    ms_total = timer.max_over_ranks(e0.elapsed_time(e1))
    # ---- roofline leg: the same steps launched directly, with CUDA events
    # around the sweep kernel on the launching stream
    sweep_ms, tp_ms = _sweep_ms(L, batch, step, args.steps)
    value = samples * world * args.steps / (ms_total * 1e-3) / 1e9

    if args.quick:
        # diagnostic: the same steps enqueued back to back, one read-back at the
        # end -- GPU time per step without the host turn-around between steps
        barrier()
        e0.record(stream)
        inflight = 0
        for _ in range(args.steps):
            batch.run()
            inflight += 1
            if inflight == DEPTH:
                batch.fetch()
                inflight -= 1
        while inflight:
            batch.fetch()
            inflight -= 1
        e1.record(stream)
        barrier()
        piped_ms = e0.elapsed_time(e1) / args.steps
        cand = L.lgb_batch_truepeak_candidates(batch._h)
        if rank == 0:
            emit({"quick": True, "n_gpus": world, "value": value, "tp_candidates": cand,
                  "tp_candidate_frac": cand / (samples / 24.0), "ms_per_step": ms_total / args.steps,
                  "ms_per_step_back_to_back": piped_ms,
                  "sweep_ms": sweep_ms, "truepeak_ms": tp_ms,
                  "sweep_gsamples": samples / (sweep_ms * 1e-3) / 1e9 if sweep_ms else None,
                  "album_loudness": ares[0].loudness})
        batch.close()
        if xchg is not None:
            xchg.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the merged album against a numpy gating of the union of all ranks' block
    # lists (outside the timed region; every rank must also hold the same bits)
    merged_check = None
    if world > 1:
        z = torch.cat([engine.device_blocks(batch, t, 0) for t in range(len(album))])
        st = torch.cat([engine.device_blocks(batch, t, 1) for t in range(len(album))])
        zs = engine.gather_block_lists(dist, z, world)
        sts = engine.gather_block_lists(dist, st, world)
        mine = torch.tensor([ares[0].loudness, ares[0].range], dtype=torch.float64, device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        if rank == 0:
            want_l, want_r = _numpy_gating(torch.cat(zs).cpu().numpy(), torch.cat(sts).cpu().numpy())
            merged_check = {"numpy_union_loudness": want_l, "numpy_union_range": want_r,
                            "loudness_diff": abs(want_l - ares[0].loudness),
                            "range_diff": abs(want_r - ares[0].range),
                            "same_bits_on_every_rank": all(torch.equal(every[0], e) for e in every),
                            "tolerance": 1e-9}
            merged_check["ok"] = bool(merged_check["loudness_diff"] <= 1e-9 and
                                      merged_check["range_diff"] <= 1e-9 and
                                      merged_check["same_bits_on_every_rank"])

    # ---- end to end through the drop-in ABI from pinned host memory
    host = [t.cpu().pin_memory() for t, _ in album]
    arr = (HostTrack * len(host))()
    for i, (h, (_, rate)) in enumerate(zip(host, album)):
        arr[i] = HostTrack(h.data_ptr(), h.shape[0], h.shape[1], rate, 0 if args.format == "s16" else 1)
    out = (ScanResult * len(host))()
    L.lgb_scan_host_mt.argtypes = [C.POINTER(HostTrack), C.c_size_t, C.c_size_t, C.c_int, C.c_double,
                                   C.c_uint, C.POINTER(ScanResult)]
    os.environ.setdefault("LOUDGAIN_B200_DEVICE", str(local))
    e2e_steps = max(1, min(args.steps, 5))
    threads = scanner_threads(world, len(host))

    def e2e_leg(nthreads):
        for _ in range(min(args.warmup, 2)):
            assert L.lgb_scan_host_mt(arr, len(host), FEED_FRAMES, 1, 0.0, nthreads, out) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            assert L.lgb_scan_host_mt(arr, len(host), FEED_FRAMES, 1, 0.0, nthreads, out) == 0
        torch.cuda.synchronize()
        dt = timer.max_over_ranks(time.perf_counter() - t0)
        return samples * world * e2e_steps / dt / 1e9

    e2e_value = e2e_leg(threads)
    e2e_single = e2e_leg(1) if threads > 1 else e2e_value
    d2h = (len(host) + 1) * 64 + 2 * 4 * sum(h.shape[1] for h in host)

    # ---- consistency: both paths measured the same tracks (the drop-in path's album
    # is the rank's own 12 tracks; the resident path's album spans all ranks)
    for i in range(len(host)):
        assert abs(out[i].track_loudness - tres[i].loudness) < 1e-9, "e2e and resident paths disagree"
    if world == 1:
        assert abs(out[0].album_loudness - ares[0].loudness) < 1e-9

    launches = batch.kernel_launches
    rates = [r for _, r in album]
    album_loudness, album_range = ares[0].loudness, ares[0].range
    batch.close()
    if xchg is not None:
        xchg.close()
    del album, batch
    torch.cuda.empty_cache()

    # ---- the same album as float32 (the ebur128_add_frames_float layout: 4 B per
    # sample, where the sweep is bound by HBM rather than by instruction dispatch)
    f32_ms = None
    if world == 1:
        falbum = make_album(dev, rank, "f32")
        fbatch = engine.Batch(falbum, [0] * len(falbum), stream)
        for _ in range(args.warmup):
            fbatch.run(); fbatch.fetch()
        f32_ms, _ = _sweep_ms(L, fbatch, lambda: (fbatch.run(), fbatch.fetch()), args.steps)
        f32_bytes = sum(t.numel() * t.element_size() for t, _ in falbum)
        fbatch.close()
        del falbum, fbatch
        torch.cuda.empty_cache()

    hbm, src = _peaks()
    configs = None
    if not args.no_configs:
        configs = extra_configs(args, torch, dist, timer, L, dev, rank, world, hbm)

    if rank == 0:
        achieved = samples * 2 / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        cpu = cpu_baseline_leg([h.numpy() for h in host], rates) if world == 1 else None
        traffic, traffic_file = _ncu_traffic()
        cfg = bench_config(world)          # (identical to the reference arm's)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm if achieved else None, "traffic": traffic,
                         "traffic_source": f"dram__bytes_read+write per launch, ncu --set full, {traffic_file}",
                         "kernel": "run_sweep_kernel<S16 stereo, true-peak screening> (lg_run.cu)",
                         "kernel_ms": sweep_ms,
                         # the candidates the sweep's screening leaves are evaluated by a
                         # second pass (tp_filter_run_kernel + tp_eval_run_kernel); with it:
                         "truepeak_pass_ms": tp_ms,
                         "achieved_incl_truepeak_pass": samples * 2 / ((sweep_ms + tp_ms) * 1e-3) / 1e9
                         if sweep_ms > 0 else None,
                         "peak_source": src,
                         "algorithmic_bytes": "2 B per S16 sample, read once (SURVEY 8d)"},
            "roofline_float_input": None if not f32_ms else {
                "bound": "hbm", "achieved": f32_bytes / (f32_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": f32_bytes / (f32_ms * 1e-3) / 1e9 / hbm, "kernel": "run_sweep_kernel<F32 stereo>",
                "kernel_ms": f32_ms, "algorithmic_bytes": "4 B per float sample, read once (SURVEY 8d)",
                "note": "same album as float32 PCM; informational, the metric's workload is the S16 album"},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pcm_bytes,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "scanner_threads": threads,
                    "single_thread_value": e2e_single},
            "gpu_launches": launches * args.steps,
            "album_loudness": album_loudness, "album_range": album_range,
            "merged_album_loudness": album_loudness if world > 1 else None,
            "merged_album_check": merged_check,
            "configs": configs,
            "host_cores": usable_cores(),
        }
        emit(line)
        if merged_check is not None and not merged_check["ok"]:
            raise SystemExit("merged album disagrees with the numpy gating of the union: " + json.dumps(merged_check))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--format", default="s16", choices=["s16", "f32"],
                    help="PCM sample format of the synthetic album (f32: tuning only)")
    ap.add_argument("--quick", action="store_true",
                    help="tuning runs: skip the CPU baseline and the end-to-end leg")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the other BASELINE.json shapes (cfg1, cfg3, cfg4, cfg5, clipped album)")
    args = ap.parse_args()
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
