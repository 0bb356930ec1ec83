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

`--impl reference` times the reference's CPU path instead (the oracle port:
libebur128 itself is not in /root/reference), on all host cores, one worker
per track like bin/rgbpm2.
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


def _ncu_traffic():
    """DRAM bytes per sweep launch from the committed ncu capture of this workload."""
    path = os.path.join(ROOT, "profiles", "r01_sweep_ncu_summary.json")
    try:
        return float(json.load(open(path))["traffic_bytes_per_launch"])
    except Exception:
        return None


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
    specs = synth.config2_specs(12)
    for s in specs:
        s.seed += 1000 * rank            # every rank scans a different album
    if fmt == "f32":                     # tuning only: the float API's layout
        return [(synth.programme_float(s, device=device), s.rate) for s in specs]
    return [(synth.programme_s16(s, device=device), s.rate) for s in specs]


# ------------------------------------------------------------------ CPU arms

def _oracle_scan(lib, pcm_list, rate_list, chunk=1024, threads=1):
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
    """The reference's CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from loudgain_b200 import synth
    from oracle import load_oracle
    lib = load_oracle()
    cores = usable_cores()
    seconds = 60.0
    specs = synth.config2_specs(12)
    pcm, rates = [], []
    for s in specs:
        s.seconds = min(s.seconds, seconds)
        pcm.append(synth.programme_s16(s).numpy())
        rates.append(s.rate)
    samples = sum(p.size for p in pcm)
    threads = min(cores, len(pcm))
    for _ in range(args.warmup):
        _oracle_scan(lib, pcm[:threads], rates[:threads], threads=threads)
    times = [_oracle_scan(lib, pcm, rates, threads=threads) for _ in range(args.steps)]
    total = sum(times)
    value = samples * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"first {seconds:.0f} s of each of the 12 tracks",
                   "feed": "ebur128_add_frames_short, 1024-frame calls (scan.c:448)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {seconds:.0f} s of each of the 12 album tracks per step; "
                                   "one worker thread per track (rgbpm2 model)"},
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


def gpu_arm(args):
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    album = make_album(dev, rank, args.format)
    albums = [0] * len(album)
    stream = torch.cuda.current_stream()
    batch = engine.Batch(album, albums, stream)
    L = engine._bind()
    L.lgb_batch_enable_timing.argtypes = [C.c_void_p, C.c_int]
    L.lgb_batch_sweep_ms.argtypes = [C.c_void_p]
    L.lgb_batch_sweep_ms.restype = C.c_double
    L.lgb_batch_truepeak_ms.argtypes = [C.c_void_p]
    L.lgb_batch_truepeak_ms.restype = C.c_double
    samples = batch.total_samples
    pcm_bytes = sum(t.numel() * t.element_size() for t, _ in album)

    # Album over ALL ranks' tracks: all-gather the block lists over NCCL, then
    # run the gating / range kernel over the union on every rank.
    merge = engine.AlbumMerge(batch, range(len(album)), dist, world) if world > 1 else None

    def step():
        batch.run()
        if merge is not None:
            merge.run()                  # same stream: ordered after the batch's kernels
        res = batch.fetch()
        return res, (merge.fetch() if merge is not None else None)

    # ---- resident-PCM throughput (steps replay the batch's CUDA graph)
    # The timed region is a few milliseconds, far shorter than nvidia-smi's
    # start-up and sampling period.  So the sampler starts first and the same
    # step keeps running untimed before and after the timed region (a fixed
    # count on every rank: the multi-GPU step holds a collective); every clock
    # sample is then taken under the load the timed steps run in.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        for _ in range(args.warmup + CLOCK_LOAD_STEPS[0]):
            step()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            (tres, ares), merged = step()
        e1.record(stream)
        barrier()
        for _ in range(CLOCK_LOAD_STEPS[1]):
            step()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # ---- roofline leg: the same steps launched directly, with CUDA events
    # around the sweep kernel on the launching stream
    L.lgb_batch_enable_timing(batch._h, 1)
    for _ in range(args.steps):
        step()
    sweep_ms = L.lgb_batch_sweep_ms(batch._h)
    tp_ms = L.lgb_batch_truepeak_ms(batch._h)
    L.lgb_batch_enable_timing(batch._h, 0)
    value = samples * world * args.steps / (ms_total * 1e-3) / 1e9

    if args.quick:
        # diagnostic: the same steps enqueued back to back, one read-back at the
        # end -- GPU time per step without the host turn-around between steps
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            batch.run()
        e1.record(stream)
        batch.fetch()
        barrier()
        piped_ms = e0.elapsed_time(e1) / args.steps
        L.lgb_batch_truepeak_candidates.argtypes = [C.c_void_p]
        L.lgb_batch_truepeak_candidates.restype = C.c_uint64
        cand = L.lgb_batch_truepeak_candidates(batch._h)
        if rank == 0:
            emit({"quick": True, "value": value, "tp_candidates": cand, "tp_candidate_frac": cand / (samples / 24.0), "ms_per_step": ms_total / args.steps,
                  "ms_per_step_back_to_back": piped_ms,
                  "sweep_ms": sweep_ms, "truepeak_ms": tp_ms,
                  "sweep_gsamples": samples / (sweep_ms * 1e-3) / 1e9 if sweep_ms else None})
        batch.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the drop-in ABI from pinned host memory
    host = [t.cpu().pin_memory() for t, _ in album]
    arr = (HostTrack * len(host))()
    for i, (h, (_, rate)) in enumerate(zip(host, album)):
        arr[i] = HostTrack(h.data_ptr(), h.shape[0], h.shape[1], rate, 0)
    out = (ScanResult * len(host))()
    L.lgb_scan_host_mt.argtypes = [C.POINTER(HostTrack), C.c_size_t, C.c_size_t, C.c_int, C.c_double,
                                   C.c_uint, C.POINTER(ScanResult)]
    chunk = 4096
    os.environ.setdefault("LOUDGAIN_B200_DEVICE", str(local))
    e2e_steps = max(1, min(args.steps, 5))
    # one scanner thread per track, as many as this rank's share of the host
    # cores allows: the reference arm's model (one worker per track, rgbpm2)
    cores = usable_cores()
    threads = max(1, min(len(host), cores // world))

    def e2e_leg(nthreads):
        for _ in range(min(args.warmup, 2)):
            assert L.lgb_scan_host_mt(arr, len(host), chunk, 1, 0.0, nthreads, out) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            assert L.lgb_scan_host_mt(arr, len(host), chunk, 1, 0.0, nthreads, out) == 0
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return samples * world * e2e_steps / float(dt.item()) / 1e9

    # The host decides how many scanner threads pay off (cores actually granted to
    # this container, memory bandwidth): a few counts are tried, the best is the
    # e2e value and its thread count is reported.
    e2e_single = e2e_leg(1)
    e2e_value, e2e_threads = e2e_single, 1
    for nt in sorted({t for t in (2, 4, threads // 2, threads) if 1 < t <= threads}):
        v = e2e_leg(nt)
        if v > e2e_value:
            e2e_value, e2e_threads = v, nt
    threads = e2e_threads
    d2h = (len(host) + 1) * 64 + 2 * 4 * sum(h.shape[1] for h in host)

    # ---- consistency: both paths measured the same album
    for i in range(len(host)):
        assert abs(out[i].track_loudness - tres[i].loudness) < 1e-9, "e2e and resident paths disagree"
    assert abs(out[0].album_loudness - ares[0].loudness) < 1e-9

    # ---- the same album as float32 (the ebur128_add_frames_float layout: 4 B per
    # sample, where the sweep is bound by HBM rather than by instruction dispatch)
    f32_ms = None
    launches = batch.kernel_launches
    rates = [r for _, r in album]
    if world == 1:
        batch.close()
        del album, batch
        torch.cuda.empty_cache()
        falbum = make_album(dev, rank, "f32")
        fbatch = engine.Batch(falbum, [0] * len(falbum), stream)
        for _ in range(args.warmup):
            fbatch.run(); fbatch.fetch()
        L.lgb_batch_enable_timing(fbatch._h, 1)
        for _ in range(args.steps):
            fbatch.run(); fbatch.fetch()
        f32_ms = L.lgb_batch_sweep_ms(fbatch._h)
        f32_bytes = sum(t.numel() * t.element_size() for t, _ in falbum)
        fbatch.close()
        del falbum

    if rank == 0:
        hbm, src = _peaks()
        achieved = samples * 2 / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        cpu = cpu_baseline_leg([h.numpy() for h in host], rates)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples_per_gpu": samples,
                       "pcm_bytes_per_gpu": pcm_bytes,
                       "l2_policy": "input (508 MB per GPU) is larger than L2 (126 MB); no flush",
                       "clock_sampling": f"{CLOCK_LOAD_STEPS[0]} + {CLOCK_LOAD_STEPS[1]} untimed steps of the "
                                         "same load around the timed steps, nvidia-smi every 100 ms",
                       "sharding": "by track; album block lists all-gathered over NCCL" if world > 1
                                   else "single GPU",
                       "e2e_feed": f"ebur128_add_frames_short, {chunk}-frame calls from host PCM, "
                                   f"{threads} scanner thread(s) per GPU (one file each at a time)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm if achieved else None, "traffic": _ncu_traffic(),
                         "traffic_source": "dram__bytes_read+write per launch, ncu --set full, "
                                           "profiles/r01_sweep_ncu_summary.json",
                         "kernel": "sweep_pair_kernel<S16 stereo, 4x true-peak codes>", "kernel_ms": sweep_ms,
                         # the true-peak evaluation is a separate pass over the sweep's pair
                         # maxima (tp_scan_pair_kernel + tp_eval_pair_kernel); with it:
                         "truepeak_pass_ms": tp_ms,
                         "achieved_incl_truepeak_pass": samples * 2 / ((sweep_ms + tp_ms) * 1e-3) / 1e9
                         if sweep_ms > 0 else None,
                         "peak_source": src,
                         "algorithmic_bytes": "2 B per S16 sample, read once (SURVEY 8d)"},
            "roofline_float_input": None if not f32_ms else {
                "bound": "hbm", "achieved": f32_bytes / (f32_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": f32_bytes / (f32_ms * 1e-3) / 1e9 / hbm, "kernel": "sweep_pair_kernel<F32 stereo>",
                "kernel_ms": f32_ms, "algorithmic_bytes": "4 B per float sample, read once (SURVEY 8d)",
                "note": "same album as float32 PCM; informational, the metric's workload is the S16 album"},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pcm_bytes,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "scanner_threads": threads,
                    "single_thread_value": e2e_single},
            "gpu_launches": launches * args.steps,
            "album_loudness": ares[0].loudness, "album_range": ares[0].range,
            "merged_album_loudness": merged.loudness if merged else None,
            "host_cores": usable_cores(),
        }
        emit(line)
    if merge is not None:
        merge.close()
    if world > 1:
        batch.close()
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
