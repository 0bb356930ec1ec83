#!/usr/bin/env python3
"""Tuning: end-to-end ingest rate of the drop-in path (lgb_scan_host_mt over the cfg2 album
in pinned host memory) against call size and scanner threads."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from loudgain_b200 import engine, synth  # noqa: E402
import bench  # noqa: E402

L = engine._bind()
host = [synth.programme_s16(s, device="cuda").cpu().pin_memory() for s in synth.config2_specs(12)]
arr = (bench.HostTrack * len(host))()
for i, h in enumerate(host):
    arr[i] = bench.HostTrack(h.data_ptr(), h.shape[0], h.shape[1], 44100, 0)
out = (bench.ScanResult * len(host))()
L.lgb_scan_host_mt.argtypes = [C.POINTER(bench.HostTrack), C.c_size_t, C.c_size_t, C.c_int, C.c_double,
                               C.c_uint, C.POINTER(bench.ScanResult)]
samples = sum(h.numel() for h in host)
print("NT_MIN", os.environ.get("LOUDGAIN_B200_NT_MIN"), "cores", bench.usable_cores())
for threads in (1, 4, 8, 12):
    row = []
    for chunk in (256, 1024, 4096, 16384):
        L.lgb_scan_host_mt(arr, len(host), chunk, 1, 0.0, threads, out)
        t0 = time.perf_counter()
        for _ in range(3):
            assert L.lgb_scan_host_mt(arr, len(host), chunk, 1, 0.0, threads, out) == 0
        torch.cuda.synchronize()
        row.append(samples * 3 / (time.perf_counter() - t0) / 1e9)
    print(f"threads {threads:2d}: " + "  ".join(f"chunk {c}: {v:6.2f}" for c, v in zip((256, 1024, 4096, 16384), row)))
