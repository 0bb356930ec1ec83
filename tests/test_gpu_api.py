"""GPU parity of the rest of the libebur128 1.2.x surface
against the CPU oracle on the same PCM, through the C ABI: channel roles,
parameter changes, the int / double entry points, mixed sample types on one
state, EBUR128_MODE_HISTOGRAM, bounded history -- and of the host layer's
bookkeeping: incremental measurement under a PCM budget and more feeding
threads than staging buffers.  loudgain itself uses none of the former
(scan.c:203-207,448) and all of the latter."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from loudgain_b200 import capi, synth
from tests.helpers import TOL_LU, TOL_TP_REL, lu_diff, rel_diff

pytestmark = pytest.mark.gpu


def _both(product, oracle, fn):
    return fn(product), fn(oracle)


def _summary(st):
    return {"L": st.loudness_global(), "R": st.loudness_range(),
            "sp": np.array(st.sample_peaks()), "tp": np.array(st.true_peaks())}


def _same(g, o, peaks=True):
    assert lu_diff(g["L"], o["L"]) <= TOL_LU
    assert lu_diff(g["R"], o["R"]) <= TOL_LU
    if peaks:
        np.testing.assert_array_equal(g["sp"], o["sp"])
        assert rel_diff(g["tp"], o["tp"]) <= TOL_TP_REL


def test_set_channel_roles(product, oracle):
    """DUAL_MONO counts a mono file twice, side positions weigh 1.41, UNUSED drops a
    channel (libebur128's channel weights; the reference never calls set_channel)."""
    mono = synth.programme_s16(synth.TrackSpec(seed=901, rate=48000, channels=1, seconds=20.0)).numpy()
    quad = synth.programme_s16(synth.TrackSpec(seed=902, rate=44100, channels=4, seconds=20.0)).numpy()

    def dual(lib):
        st = lib.init(1, 48000)
        assert st.set_channel(0, capi.DUAL_MONO) == capi.SUCCESS
        st.add_frames(mono, 4096)
        out = _summary(st)
        st.destroy()
        return out

    def plain(lib):
        st = lib.init(1, 48000)
        st.add_frames(mono, 4096)
        out = _summary(st)
        st.destroy()
        return out

    g, o = _both(product, oracle, dual)
    _same(g, o)
    assert abs((g["L"] - plain(product)["L"]) - 10.0 * np.log10(2.0)) < 1e-6

    def roles(lib):
        st = lib.init(4, 44100)
        assert st.set_channel(0, capi.Mp060) == capi.SUCCESS
        assert st.set_channel(1, capi.Mp090) == capi.SUCCESS
        assert st.set_channel(2, capi.UNUSED) == capi.SUCCESS
        assert st.set_channel(3, capi.CENTER) == capi.SUCCESS
        assert st.set_channel(4, capi.LEFT) == capi.ERROR_INVALID_CHANNEL_INDEX
        assert st.set_channel(0, capi.DUAL_MONO) == capi.ERROR_INVALID_CHANNEL_INDEX
        st.add_frames(quad, 1024)
        out = _summary(st)
        st.destroy()
        return out

    g, o = _both(product, oracle, roles)
    _same(g, o)


def test_change_parameters(product, oracle):
    """A parameter change restarts filter, block schedule and interpolator; the blocks
    stored so far stay in the state's union."""
    a = synth.programme_s16(synth.TrackSpec(seed=903, rate=44100, channels=2, seconds=12.0)).numpy()
    b = synth.programme_s16(synth.TrackSpec(seed=904, rate=48000, channels=2, seconds=9.0)).numpy()
    c = synth.programme_s16(synth.TrackSpec(seed=905, rate=48000, channels=1, seconds=7.0)).numpy()

    def run(lib):
        st = lib.init(2, 44100)
        st.add_frames(a, 1024)
        assert st.change_parameters(2, 44100) == capi.ERROR_NO_CHANGE
        assert st.change_parameters(2, 48000) == capi.SUCCESS
        st.add_frames(b, 1024)
        mid = _summary(st)
        assert st.change_parameters(1, 48000) == capi.SUCCESS
        st.add_frames(c, 1024)
        out = {"L": st.loudness_global(), "R": st.loudness_range(), "mid": mid,
               "sp": np.array(st.sample_peaks()), "tp": np.array(st.true_peaks())}
        st.destroy()
        return out

    g, o = _both(product, oracle, run)
    _same(g, o)
    _same(g["mid"], o["mid"])


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_int_and_double_input(product, oracle, dtype):
    spec = synth.TrackSpec(seed=906, rate=44100, channels=2, seconds=15.0)
    x = synth.programme_float(spec).numpy().astype(np.float64)
    pcm = np.round(x * 2147483647.0).clip(-2147483648, 2147483647).astype(np.int32) if dtype == np.int32 else x

    def run(lib):
        st = lib.init(2, 44100)
        st.add_frames(pcm, 2048)
        out = _summary(st)
        st.destroy()
        return out

    g, o = _both(product, oracle, run)
    # the product narrows both types to float32 on the way in: peaks agree
    # to float precision, not bit for bit
    _same(g, o, peaks=False)
    assert rel_diff(g["sp"], o["sp"]) <= 1e-6 and rel_diff(g["tp"], o["tp"]) <= 2e-6


@pytest.mark.parametrize("order", ["short_then_float", "float_then_short"])
def test_mixed_sample_types_on_one_state(product, oracle, order):
    """libebur128 accepts any add_frames_* on the same state (scan.c only uses _short)."""
    spec = synth.TrackSpec(seed=907, rate=44100, channels=2, seconds=16.0)
    s16 = synth.programme_s16(spec).numpy()
    half = (len(s16) // 2 // 7) * 7 + 3               # not a slot boundary
    as_float = (s16.astype(np.float32) / 32768.0)
    parts = [s16[:half], as_float[half:]] if order == "short_then_float" else [as_float[:half], s16[half:]]

    def run(lib):
        st = lib.init(2, 44100)
        for p in parts:
            st.add_frames(p, 1024)
        out = _summary(st)
        st.destroy()
        return out

    def pure(lib):
        st = lib.init(2, 44100)
        st.add_frames(s16, 1024)
        out = _summary(st)
        st.destroy()
        return out

    g, o = _both(product, oracle, run)
    _same(g, o)
    _same(g, pure(product))              # x / 32768 is exact in float: the same samples either way


def test_histogram_mode(product, oracle):
    """EBUR128_MODE_HISTOGRAM: block energies become 0.1 LU bin centres before any gate
    or percentile; states of both kinds may meet in one *_multiple query."""
    specs = synth.config2_specs(ntracks=3, scale=0.12)
    tracks = [synth.programme_s16(s).numpy() for s in specs]
    mode = capi.MODE_LOUDGAIN | capi.MODE_HISTOGRAM

    def run(lib):
        sts = []
        for i, pcm in enumerate(tracks):
            st = lib.init(2, 44100, mode if i < 2 else capi.MODE_LOUDGAIN)
            st.add_frames(pcm, 4096)
            sts.append(st)
        out = {"tracks": [_summary(st) for st in sts],
               "rel": [st.relative_threshold() for st in sts],
               "album_L": lib.loudness_global_multiple(sts[:2]),
               "album_R": lib.loudness_range_multiple(sts[:2]),
               "mixed_L": lib.loudness_global_multiple(sts)}
        for st in sts:
            st.destroy()
        return out

    g, o = _both(product, oracle, run)
    for gt, ot in zip(g["tracks"], o["tracks"]):
        _same(gt, ot)
    for a, b in zip(g["rel"], o["rel"]):
        assert abs(a - b) <= TOL_LU
    assert lu_diff(g["album_L"], o["album_L"]) <= TOL_LU
    assert lu_diff(g["album_R"], o["album_R"]) <= TOL_LU
    assert lu_diff(g["mixed_L"], o["mixed_L"]) <= TOL_LU
    # the bins are visible: exact and histogram results of the same audio differ, by < 0.1 LU
    exact = product.init(2, 44100)
    exact.add_frames(tracks[0], 4096)
    d = abs(exact.loudness_global() - g["tracks"][0]["L"])
    exact.destroy()
    assert 0.0 < d < 0.1


def test_bounded_history(product, oracle):
    """ebur128_set_max_history: only the newest blocks of a state are gated."""
    spec = synth.TrackSpec(seed=908, rate=44100, channels=2, seconds=40.0)
    pcm = synth.programme_s16(spec).numpy()

    def run(lib):
        st = lib.init(2, 44100)
        assert st.set_max_history(10000) == capi.SUCCESS
        assert st.set_max_history(10000) == capi.ERROR_NO_CHANGE
        st.add_frames(pcm, 4096)
        out = _summary(st)
        st.destroy()
        return out

    g, o = _both(product, oracle, run)
    _same(g, o)
    full = product.init(2, 44100)
    full.add_frames(pcm, 4096)
    assert abs(full.loudness_global() - g["L"]) > 1e-3      # the bound matters on this material
    full.destroy()


def _pcm_stats(product):
    fn = product.lib.lgb_dropin_pcm_bytes
    fn.argtypes = [C.POINTER(C.c_uint64)] * 3
    fn.restype = None
    now, peak, rel = C.c_uint64(), C.c_uint64(), C.c_uint64()
    fn(C.byref(now), C.byref(peak), C.byref(rel))
    return now.value, peak.value, rel.value


@pytest.mark.parametrize("rate,channels,dtype,n,sec0,dsec", [
    (44100, 2, np.int16, 8, 31.3, 7.7),       # 5.5 - 15 MB per state: several staging buffers each
    (44100, 1, np.int16, 5, 61.0, 9.1),       # mono: spans start on 2 s multiples (16-byte alignment)
    (96000, 6, np.float32, 4, 11.2, 4.3)])    # 26 - 55 MB per state
def test_incremental_measurement_under_a_budget(product, oracle, rate, channels, dtype, n, sec0, dsec,
                                                monkeypatch):
    """scan.c keeps every state alive until the end (scan.c:98-108); the PCM of all live
    states must not pile up in HBM.  With a small budget the complete part of every state
    is measured and released while the scan goes on; the results stay those of a one-shot
    measurement (same blocks: sums of the same 100 ms energies), album union included."""
    secs = [sec0 + dsec * i for i in range(n)]
    pcms = []
    for i, sec in enumerate(secs):
        spec = synth.TrackSpec(seed=950 + i, rate=rate, channels=channels, seconds=sec,
                               lfe_channel=3 if channels == 6 else None)
        x = synth.programme_float(spec)
        pcms.append(synth.quantise_s16(x).numpy() if dtype == np.int16 else x.numpy())
    total = sum(p.nbytes for p in pcms)

    def run(lib):
        sts = []
        for p in pcms:
            st = lib.init(channels, rate)
            st.add_frames(p, 1024)
            sts.append(st)
        out = {"tracks": [_summary(st) for st in sts],
               "album_L": lib.loudness_global_multiple(sts), "album_R": lib.loudness_range_multiple(sts)}
        out["window"] = [sts[-1].loudness_momentary(), sts[-1].loudness_shortterm()]
        # feeding goes on after a query
        sts[0].add_frames(pcms[1][:rate * 4], 1024)
        out["more"] = _summary(sts[0])
        for st in sts:
            st.destroy()
        return out

    o = run(oracle)
    one_shot = run(product)
    _, _, rel0 = _pcm_stats(product)
    monkeypatch.setenv("LOUDGAIN_B200_PCM_BUDGET_MB", "24")
    before = _pcm_stats(product)[0]                    # (also restarts the high-water mark)
    g = run(product)
    now, peak, rel = _pcm_stats(product)
    assert rel > rel0                                   # release passes did run
    assert now == before                                # everything was given back
    for gt, ot, st1 in zip(g["tracks"], o["tracks"], one_shot["tracks"]):
        _same(gt, ot)
        assert lu_diff(gt["L"], st1["L"]) <= 1e-5 and lu_diff(gt["R"], st1["R"]) <= 1e-5
        np.testing.assert_array_equal(gt["sp"], st1["sp"])
        np.testing.assert_array_equal(gt["tp"], st1["tp"])
    assert lu_diff(g["album_L"], o["album_L"]) <= TOL_LU and lu_diff(g["album_R"], o["album_R"]) <= TOL_LU
    for a, b in zip(g["window"], o["window"]):
        assert lu_diff(a, b) <= TOL_LU
    _same(g["more"], o["more"])
    assert total > 24 * 2 ** 20                          # the scan as a whole does not fit the budget
    # never more than the budget plus what one state adds before the next check (its
    # buffer doubles as it grows) -- far below the whole scan
    assert peak - before < 24 * 2 ** 20 + 2 * max(p.nbytes for p in pcms) + 8 * 2 ** 20
    assert peak - before < total


def test_more_feeding_threads_than_staging_buffers(product):
    """40 scanner threads feed 40 states at once through a pool of 32 pinned staging
    buffers: a thread that finds every buffer taken waits, no frame is dropped
    (scan.c:448-453 would only print 'Error filtering' and go on)."""
    n = 40
    specs = [synth.TrackSpec(seed=980 + i, rate=44100, channels=2, seconds=3.0 + 0.1 * i) for i in range(n)]
    pcms = [synth.programme_s16(s).numpy() for s in specs]
    want = []
    for p in pcms:
        st = product.init(2, 44100)
        st.add_frames(p, 4096)
        want.append((st.loudness_global(), st.true_peaks()))
        st.destroy()
    states = [product.init(2, 44100) for _ in range(n)]
    gate = threading.Barrier(n)
    errors = []

    def feed(i):
        try:
            gate.wait()
            states[i].add_frames(pcms[i], 64)       # many small calls: the threads overlap for long
        except Exception as e:                       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=feed, args=(i,)) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    got = [(st.loudness_global(), st.true_peaks()) for st in states]
    for st in states:
        st.destroy()
    assert got == want


def test_histogram_state_released_under_a_budget(product, oracle, monkeypatch):
    """A histogram-mode state whose PCM is released in spans: its blocks are formed over the
    kept slot list and quantised there (stream_block_kernel with the histogram table)."""
    spec = synth.TrackSpec(seed=991, rate=48000, channels=2, seconds=70.0)
    pcm = synth.programme_s16(spec).numpy()
    mode = capi.MODE_LOUDGAIN | capi.MODE_HISTOGRAM

    def run(lib):
        sts = [lib.init(2, 48000, mode) for _ in range(3)]
        for st in sts:
            st.add_frames(pcm, 2048)
        out = _summary(sts[1])
        out["album_L"] = lib.loudness_global_multiple(sts)
        for st in sts:
            st.destroy()
        return out

    o = run(oracle)
    rel0 = _pcm_stats(product)[2]
    monkeypatch.setenv("LOUDGAIN_B200_PCM_BUDGET_MB", "16")
    g = run(product)
    assert _pcm_stats(product)[2] > rel0
    _same(g, o)
    assert lu_diff(g["album_L"], o["album_L"]) <= TOL_LU
