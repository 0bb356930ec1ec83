"""Shared test helpers: the oracle, the test-only host emulation of the
device math (tests/emu), and comparison utilities."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# Tolerances of BASELINE.json's north_star.
TOL_LU = 0.01            # integrated loudness and range, LU
TOL_TP_REL = 1e-6        # true peak, relative
# What this implementation is expected to hold in practice (tighter).
GOAL_LU = 2e-4


class LgbTrack(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("frames", C.c_uint64), ("channels", C.c_uint32),
                ("samplerate", C.c_uint32), ("format", C.c_uint32), ("album", C.c_uint32),
                ("weight_class", C.c_void_p), ("lead_in", C.c_uint64), ("flags", C.c_uint32)]


class LgbResult(C.Structure):
    _fields_ = [("loudness", C.c_double), ("range", C.c_double), ("rel_threshold", C.c_double),
                ("sum_abs", C.c_double), ("sum_rel", C.c_double), ("n_abs", C.c_uint64),
                ("n_rel", C.c_uint64), ("n_shortterm", C.c_uint64)]


NO_ALBUM = 0xFFFFFFFF


def build_emu() -> str:
    src = os.path.join(HERE, "emu", "emu.cpp")
    so = os.path.join(HERE, "emu", "libemu.so")
    deps = [src] + [os.path.join(ROOT, "loudgain_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "loudgain_b200", "csrc"))
                    if f.endswith((".h", ".cuh"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fPIC",
                               "-shared", "-o", so, src])
    return so


def emu_measure(tracks, albums=None, target_tasks=0, lead_in=None):
    """tracks: list of (pcm ndarray [frames, ch] int16/float32, rate).
    Returns dict with per-track / per-album results, peaks and block lists."""
    lib = C.CDLL(build_emu())
    n = len(tracks)
    arr = (LgbTrack * n)()
    keep = []
    nalb = 0
    for i, (pcm, rate) in enumerate(tracks):
        pcm = np.ascontiguousarray(pcm)
        keep.append(pcm)
        fmt = {np.dtype(np.int16): 0, np.dtype(np.float32): 1}[pcm.dtype]
        alb = NO_ALBUM if albums is None else albums[i]
        if alb != NO_ALBUM:
            nalb = max(nalb, alb + 1)
        arr[i] = LgbTrack(pcm.ctypes.data, pcm.shape[0], pcm.shape[1], rate, fmt, alb, None,
                          0 if lead_in is None else lead_in[i])
    tb, ts = C.c_uint64(), C.c_uint64()
    lib.emu_plan_sizes(arr, C.c_size_t(n), C.c_uint64(target_tasks), C.byref(tb), C.byref(ts))
    tres = (LgbResult * n)()
    ares = (LgbResult * max(nalb, 1))()
    npk = sum(t[0].shape[1] for t in tracks)
    sp = np.zeros(npk); tp = np.zeros(npk); tps = np.zeros(npk)
    blocks = np.zeros(max(tb.value, 1)); st = np.zeros(max(ts.value, 1))
    clen = np.zeros(n, dtype=np.int32)
    nslots = [t[0].shape[0] // ((t[1] + 5) // 10) for t in tracks]
    slots = np.zeros(max(sum(nslots), 1))
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_measure(arr, C.c_size_t(n), C.c_uint32(nalb), C.c_uint64(target_tasks), tres, ares,
                         dp(sp), dp(tp), dp(blocks), dp(st), dp(clen), dp(tps), dp(slots))
    assert rc == 0
    out = {"tracks": [], "albums": [], "blocks": blocks[:tb.value], "st": st[:ts.value],
           "chunk_len": clen,
           "slots": np.split(slots[:sum(nslots)], np.cumsum(nslots)[:-1]) if n else []}
    off = 0
    for i, (pcm, _) in enumerate(tracks):
        ch = pcm.shape[1]
        r = tres[i]
        out["tracks"].append({"loudness": r.loudness, "range": r.range, "n_abs": r.n_abs,
                              "n_rel": r.n_rel, "n_st": r.n_shortterm,
                              "sample_peak": sp[off:off + ch].copy(),
                              "true_peak": tp[off:off + ch].copy(),
                              # the device's two-pass (screened) evaluation
                              "true_peak_screened": tps[off:off + ch].copy()})
        off += ch
    for a in range(nalb):
        r = ares[a]
        out["albums"].append({"loudness": r.loudness, "range": r.range, "n_abs": r.n_abs,
                              "n_rel": r.n_rel, "n_st": r.n_shortterm})
    return out


def oracle_measure(lib, tracks, albums=None, chunk_frames=1024):
    """Drives the oracle the way scan.c drives libebur128."""
    from oracle import blocks as oracle_blocks
    states, res = [], {"tracks": [], "albums": []}
    for pcm, rate in tracks:
        st = lib.init(pcm.shape[1], rate)
        st.add_frames(pcm, chunk_frames)
        states.append(st)
        res["tracks"].append({"loudness": st.loudness_global(), "range": st.loudness_range(),
                              "sample_peak": np.array(st.sample_peaks()),
                              "true_peak": np.array(st.true_peaks()),
                              "blocks": oracle_blocks(lib, st, 0), "st": oracle_blocks(lib, st, 1)})
    if albums is not None:
        nalb = max(a for a in albums if a != NO_ALBUM) + 1 if any(a != NO_ALBUM for a in albums) else 0
        for a in range(nalb):
            mem = [s for s, al in zip(states, albums) if al == a]
            res["albums"].append({"loudness": lib.loudness_global_multiple(mem),
                                  "range": lib.loudness_range_multiple(mem)})
    for st in states:
        st.destroy()
    return res


def lu_diff(a: float, b: float) -> float:
    if np.isinf(a) or np.isinf(b):
        return 0.0 if a == b else float("inf")
    return abs(a - b)


def rel_diff(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    m = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.where(d == 0, 0.0, d / m))) if a.size else 0.0


def gate_slots(slots, rate):
    """Integrated loudness and range of one stream from its 100 ms slot
    energies (numpy restatement of lgb_slots_query / SURVEY.md A.2, A.6, A.7)."""
    s100 = (rate + 5) // 10
    slots = np.asarray(slots, dtype=np.float64)
    n = len(slots)
    z = np.array([slots[b:b + 4].sum() for b in range(max(n - 3, 0))]) / (4.0 * s100)
    st = np.array([slots[10 * j:10 * j + 30].sum() for j in range((n - 30) // 10 + 1 if n >= 30 else 0)])
    st = st / (30.0 * s100)
    gate = 10 ** ((-70 + 0.691) / 10)
    z = z[z >= gate]
    loud = -np.inf
    if len(z):
        z = z[z >= 0.1 * z.mean()]
        loud = 10 * np.log10(z.mean()) - 0.691
    st = np.sort(st[st >= gate])
    rng = 0.0
    if len(st):
        st = st[st >= 0.01 * st.mean()]
        if len(st):
            hi = st[int((len(st) - 1) * 0.95 + 0.5)]
            lo = st[int((len(st) - 1) * 0.1 + 0.5)]
            rng = 10 * np.log10(hi / lo)
    return loud, rng
