#!/usr/bin/env python3
"""Tuning: per-warp timeline of the packed sweep (library built with
-DLG_PAIR_TRACE as variant 'tr').  Runs the sweep alone (timed mode keeps the
true-peak pass from overwriting the trace) and prints when warps start and end."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LG_LIB_SUFFIX"] = "tr"
import torch  # noqa: E402

import bench  # noqa: E402
from loudgain_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
album = bench.make_album(dev, 0)
batch = engine.Batch(album, [0] * len(album))
L = engine._bind()
for _ in range(3):
    batch.run(); batch.fetch()
L.lgb_debug_trace.restype = C.c_uint64
L.lgb_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
buf = np.zeros(1 << 16, dtype=np.uint64)
# the scan/eval kernels of the same step only touch the head of the queue
batch.run(); batch.fetch()
n = L.lgb_debug_trace(batch._h, buf.ctypes.data, buf.size)
t = buf[:n].reshape(-1, 2)
start = t[:, 0].astype(np.float64)
dur = (t[:, 1] & ((1 << 48) - 1)).astype(np.float64)
sm = (t[:, 1] >> 48).astype(np.int64)
t0 = start.min()
start = (start - t0) / 1e3
end = start + dur / 1e3
print(f"warps {len(start)}, kernel span {end.max():.1f} us; warp duration mean {dur.mean() / 1e3:.1f} us, "
      f"min {dur.min() / 1e3:.1f}, max {dur.max() / 1e3:.1f}")
print("start-time percentiles (us):", np.percentile(start, [0, 10, 50, 53, 55, 60, 90, 100]).round(1))
print("end-time percentiles (us):  ", np.percentile(end, [0, 10, 40, 50, 60, 90, 99, 100]).round(1))
per_sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
per_sm_n = np.array([(sm == s).sum() for s in np.unique(sm)])
print("per-SM last end (us): min %.1f mean %.1f max %.1f; warps per SM min %d max %d" %
      (per_sm_end.min(), per_sm_end.mean(), per_sm_end.max(), per_sm_n.min(), per_sm_n.max()))
first = start < np.percentile(start, 50)
print("first-wave warps: duration mean %.1f us; later warps: %.1f us" %
      (dur[first].mean() / 1e3, dur[~first].mean() / 1e3))
batch.close()
