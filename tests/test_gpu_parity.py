"""GPU parity: the CUDA path, called through the C ABI (ebur128_* as scan.c
calls it, and the lgb_* batch extension), against the CPU oracle on the same
PCM.  Tolerances are the north star's: loudness / range 0.01 LU, gains equal
at 2 decimals, sample peak bit-exact, true peak 1e-6 relative -- plus the
tighter figure this implementation is expected to hold."""
import numpy as np
import pytest

from loudgain_b200 import synth
from tests import cases
from tests.helpers import GOAL_LU, NO_ALBUM, TOL_LU, TOL_TP_REL, lu_diff, oracle_measure, rel_diff

pytestmark = pytest.mark.gpu


def _drive(lib, tracks, albums, chunk_frames):
    states, res = [], {"tracks": [], "albums": []}
    for pcm, rate in tracks:
        st = lib.init(pcm.shape[1], rate)
        st.add_frames(pcm, chunk_frames)
        states.append(st)
    for st in states:      # queries only after every file is scanned (loudgain.c:323-340)
        res["tracks"].append({"loudness": st.loudness_global(), "range": st.loudness_range(),
                              "sample_peak": np.array(st.sample_peaks()),
                              "true_peak": np.array(st.true_peaks())})
    if albums is not None:
        for a in range(max(albums) + 1):
            mem = [s for s, al in zip(states, albums) if al == a]
            for _ in range(2):          # loudgain repeats the album query per track
                g = lib.loudness_global_multiple(mem)
                r = lib.loudness_range_multiple(mem)
            res["albums"].append({"loudness": g, "range": r})
    for st in states:
        st.destroy()
    return res


def _gain2(loudness):
    """ReplayGain at tag precision (scan.c:64, tag.cc:178)."""
    return "%.2f" % (-18.0 - loudness)


def _check(o, g, tol=GOAL_LU):
    assert lu_diff(g["loudness"], o["loudness"]) <= TOL_LU
    assert lu_diff(g["range"], o["range"]) <= TOL_LU
    assert lu_diff(g["loudness"], o["loudness"]) <= tol
    assert lu_diff(g["range"], o["range"]) <= tol
    if np.isfinite(o["loudness"]):
        # identical at 2 decimals unless the oracle value sits on a rounding edge
        frac = abs(((-18.0 - o["loudness"]) * 100.0) % 1.0 - 0.5)
        if frac > 0.05:
            assert _gain2(g["loudness"]) == _gain2(o["loudness"])
    if "sample_peak" in o:
        np.testing.assert_array_equal(g["sample_peak"], o["sample_peak"])
        assert rel_diff(g["true_peak"], o["true_peak"]) <= TOL_TP_REL


def test_config1_track(product, oracle):
    """cfg1: 44.1 kHz stereo S16 track through the drop-in ABI, 1024-frame calls."""
    spec = synth.config1_spec(60.0)
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, spec.rate)])
    g = _drive(product, [(pcm, spec.rate)], None, 1024)
    _check(o["tracks"][0], g["tracks"][0])


def test_config2_album(product, oracle):
    """cfg2: album mode -- per-track and *_multiple results."""
    specs = synth.config2_specs(ntracks=12, scale=0.1)
    tracks = [(synth.programme_s16(s).numpy(), s.rate) for s in specs]
    albums = [0] * len(tracks)
    o = oracle_measure(oracle, tracks, albums)
    g = _drive(product, tracks, albums, 4096)
    for ot, gt in zip(o["tracks"], g["tracks"]):
        _check(ot, gt)
    _check(o["albums"][0], g["albums"][0])
    assert max(t["true_peak"].max() for t in o["tracks"]) > 1.0   # -k has work to do


def test_config3_multichannel_96k(product, oracle):
    """cfg3: 96 kHz 5.1, S16 reduction (what the reference measures) and float."""
    spec = synth.config3_spec(20.0)
    x = synth.programme_float(spec)
    for pcm in (synth.quantise_s16(x).numpy(), x.numpy()):
        o = oracle_measure(oracle, [(pcm, spec.rate)])
        g = _drive(product, [(pcm, spec.rate)], None, 4096)
        _check(o["tracks"][0], g["tracks"][0])
        assert o["tracks"][0]["true_peak"][3] > 0.5      # loud LFE shows in the peaks


def test_config4_long_stream_slice(product, oracle):
    """cfg4 shape at a size the oracle finishes quickly: 48 kHz stereo, 10 min."""
    spec = synth.config4_spec(600.0)
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, spec.rate)], chunk_frames=1 << 16)
    g = _drive(product, [(pcm, spec.rate)], None, 1 << 16)
    _check(o["tracks"][0], g["tracks"][0])


def test_config5_library_batch(product, oracle):
    """cfg5 shape, scaled: mixed rate / channel tracks in albums, batch API with
    PCM resident in HBM."""
    import torch
    from loudgain_b200 import engine
    specs, albums = synth.config5_specs(ntracks=40, scale=0.05)
    host = [synth.programme_s16(s).numpy() for s in specs]
    tracks = [(h, s.rate) for h, s in zip(host, specs)]
    o = oracle_measure(oracle, tracks, albums)
    dev = [(torch.from_numpy(h).cuda(), s.rate) for h, s in zip(host, specs)]
    tres, ares = engine.measure(dev, albums)
    for ot, m in zip(o["tracks"], tres):
        _check(ot, {"loudness": m.loudness, "range": m.range, "sample_peak": m.sample_peak,
                    "true_peak": m.true_peak})
    for oa, m in zip(o["albums"], ares):
        _check(oa, {"loudness": m.loudness, "range": m.range})


def test_batch_blocks_match_oracle(product, oracle):
    """Block lists, not just the scalars derived from them."""
    import torch
    from loudgain_b200 import engine
    spec = synth.config1_spec(45.0)
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, spec.rate)])["tracks"][0]
    b = engine.Batch([(torch.from_numpy(pcm).cuda(), spec.rate)])
    b.run()
    b.fetch()
    gate = 10 ** ((-70 + 0.691) / 10)
    z = b.blocks(0, 0)
    st = b.blocks(0, 1)
    b.close()
    z, st = z[z >= gate], st[st >= gate]
    assert len(z) == len(o["blocks"]) and rel_diff(z, o["blocks"]) < 2e-5
    assert len(st) == len(o["st"]) and rel_diff(st, o["st"]) < 2e-5


@pytest.mark.parametrize("rate,channels,fmt", [(44100, 2, "s16"), (44100, 2, "f32"), (48000, 2, "s16"),
                                               (96000, 2, "f32"), (96000, 6, "s16"), (48000, 1, "f32")])
@pytest.mark.parametrize("kind", ["tone", "noise"])
def test_dc_offset_quiet_programme(product, oracle, rate, channels, fmt, kind):
    """a -60 dBFS programme on an offset of 0.1 .. 0.9 FS
    (tests/cases.py).  Through the batch API (block lists within 1e-5 relative of
    the oracle's double-precision filter) and through ebur128_add_frames_*."""
    import torch
    from loudgain_b200 import engine
    gate = 10 ** ((-70 + 0.691) / 10)
    for dc in (0.1, 0.3, 0.6, 0.9):
        x = cases.dc_offset_quiet_programme(rate, dc, kind, channels)
        pcm = cases.to_s16(x) if fmt == "s16" else x.astype(np.float32)
        o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
        b = engine.Batch([(torch.from_numpy(pcm).cuda(), rate)])
        b.run()
        tres, _ = b.fetch()
        z, st = b.blocks(0, 0), b.blocks(0, 1)
        b.close()
        m = tres[0]
        _check(o, {"loudness": m.loudness, "range": m.range, "sample_peak": m.sample_peak,
                   "true_peak": m.true_peak})
        z, st = z[z >= gate], st[st >= gate]
        assert len(z) == len(o["blocks"]) and rel_diff(z, o["blocks"]) <= 1e-5, dc
        assert len(st) == len(o["st"]) and rel_diff(st, o["st"]) <= 1e-5, dc
        if dc == 0.6:
            g = _drive(product, [(pcm, rate)], None, 1024)
            _check(o, g["tracks"][0])


def test_chunking_invariance(product):
    """Any split of add_frames calls gives identical results."""
    rng = np.random.default_rng(5)
    pcm = (rng.standard_normal((44100 * 8, 2)) * 4000).astype(np.int16)
    ref = None
    for split in (None, 1024, 4410, 4409, 100000, 1):
        st = product.init(2, 44100)
        if split == 1:
            st.add_frames(pcm[:700], 1)
            st.add_frames(pcm[700:], 7777)
        else:
            st.add_frames(pcm, split)
        got = (st.loudness_global(), st.loudness_range(), tuple(st.sample_peaks()),
               tuple(st.true_peaks()))
        st.destroy()
        ref = ref or got
        assert got == ref


def test_config4_time_segments(product, oracle):
    """cfg4 shape, scaled: one long stereo 48 kHz stream measured as time
    segments with one second of lead-in each (the per-rank work of the
    time-sharded scan, here all on one GPU), merged through lgb_slots_query."""
    import torch
    from loudgain_b200 import engine
    spec = synth.config1_spec(64.2)
    spec.rate = 48000
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, 48000)])["tracks"][0]
    whole, _ = engine.measure([(torch.from_numpy(pcm).cuda(), 48000)])
    for parts in (1, 2, 5):
        plan = engine.segment_plan(len(pcm), 48000, parts)
        segs = [(torch.from_numpy(pcm[a:e]).cuda(), lead) for a, lead, e in plan]
        m = engine.measure_stream_segments(segs, 48000)
        _check(o, {"loudness": m.loudness, "range": m.range, "sample_peak": m.sample_peak,
                   "true_peak": m.true_peak})
        # against the unsegmented device measurement: same blocks up to the
        # FP64 rounding of the state carry, identical peaks
        assert abs(m.loudness - whole[0].loudness) < 1e-9 and abs(m.range - whole[0].range) < 1e-9
        np.testing.assert_array_equal(m.true_peak, whole[0].true_peak)
        assert m.n_abs == whole[0].n_abs and m.n_shortterm == whole[0].n_shortterm


@pytest.mark.parametrize("seconds", [0.25, 2.0, 7.3])
def test_sliding_window_queries(product, oracle, seconds):
    """ebur128_loudness_momentary / _shortterm / _window (not used by loudgain):
    the last 400 ms / 3 s / n x 100 ms of what was fed, silence before the start."""
    rng = np.random.default_rng(int(seconds * 100))
    n = int(seconds * 44100)
    env = np.linspace(0.2, 1.0, n)[:, None]
    pcm = (rng.standard_normal((n, 2)) * 6000 * env).astype(np.int16)

    def drive(L):
        st = L.init(2, 44100)
        st.add_frames(pcm[: n // 3], 1000)
        first = st.loudness_momentary()
        st.add_frames(pcm[n // 3:], 4096)
        out = (first, st.loudness_momentary(), st.loudness_shortterm(), st.loudness_window(1000),
               st.loudness_global())
        st.destroy()
        return out

    got, want = drive(product), drive(oracle)
    for g, w in zip(got, want):
        assert lu_diff(g, w) <= GOAL_LU


@pytest.mark.parametrize("nslots", [29, 3000, 70000, 400000])
def test_slots_query_sizes(product, nslots):
    """lgb_slots_query against a numpy restatement, from a list too short for a
    short-term block up to eleven hours of slots (the gating blocks of a large
    query are shared out over a thread-block cluster; short-term energies beyond
    the shared-memory cache come from global memory)."""
    import torch
    from loudgain_b200 import engine
    from tests.helpers import gate_slots
    rng = np.random.default_rng(nslots)
    # 100 ms slot energy sums at 48 kHz: programme around -20 LUFS with quiet and silent stretches
    level = rng.choice([0.0, 1e-9, 1e-3, 1e-2, 3e-2], size=nslots, p=[0.05, 0.1, 0.25, 0.4, 0.2])
    slots = level * rng.uniform(0.5, 1.5, nslots) * 4800
    m = engine.slots_query(torch.from_numpy(slots).cuda(), 48000)
    loud, rng_lu = gate_slots(slots, 48000)
    assert lu_diff(m.loudness, loud) < 1e-9 and abs(m.range - rng_lu) < 1e-9


def test_threaded_scanners_match_single_thread(product):
    """Different threads feed different states at the same time (states are
    independent, as in libebur128): the results do not depend on it, also with
    more files than scanner threads and more states than staging buffers."""
    from loudgain_b200 import engine
    rng = np.random.default_rng(11)
    tracks = []
    for i in range(40):
        n = int(rng.integers(30000, 400000))
        ch = 2 if i % 5 else 1
        tracks.append(((rng.standard_normal((n, ch)) * (1500 + 300 * i)).astype(np.int16),
                       44100 if i % 3 else 48000))
    fields = [n for n, _ in engine.ScanResult._fields_]
    one = engine.scan_host(tracks, chunk_frames=1024, threads=1)
    for threads in (4, 12):
        many = engine.scan_host(tracks, chunk_frames=1024, threads=threads)
        for a, b in zip(one, many):
            assert [getattr(a, f) for f in fields] == [getattr(b, f) for f in fields]


@pytest.mark.parametrize("frames", [0, 1, 11, 4409, 17639, 17640, 17641, 132300])
def test_ragged_lengths(product, oracle, frames):
    rng = np.random.default_rng(frames)
    pcm = (rng.standard_normal((frames, 2)) * 5000).astype(np.int16)
    o = oracle_measure(oracle, [(pcm, 44100)])
    g = _drive(product, [(pcm, 44100)], None, 1024)
    _check(o["tracks"][0], g["tracks"][0])


def test_silence_and_errors(product):
    from loudgain_b200 import capi
    st = product.init(2, 44100)
    st.add_frames(np.zeros((44100 * 4, 2), dtype=np.int16))
    assert st.loudness_global() == -np.inf and st.loudness_range() == 0.0
    assert st.true_peaks() == [0.0, 0.0]
    rc, _ = st._scalar("ebur128_true_peak", 2)
    assert rc == capi.ERROR_INVALID_CHANNEL_INDEX
    st.destroy()
    assert st.ptr is None
    assert product.try_init(0, 44100) is None
    st = product.init(2, 44100, capi.MODE_I)
    rc, _ = st._scalar("ebur128_loudness_range")
    assert rc == capi.ERROR_INVALID_MODE
    st.destroy()


def test_sample_peak_full_scale(product):
    pcm = np.zeros((48000, 2), dtype=np.int16)
    pcm[100, 0] = -32768
    pcm[47999, 1] = 32767
    st = product.init(2, 48000)
    st.add_frames(pcm)
    assert st.sample_peaks() == [1.0, 32767 / 32768.0]
    st.destroy()


@pytest.mark.parametrize("rate,channels,dtype", [(44100, 2, np.int16), (96000, 6, np.int16),
                                                 (192000, 2, np.int16), (48000, 1, np.float32)])
def test_prev_peaks_of_last_call(product, oracle, rate, channels, dtype):
    """ebur128_prev_sample_peak / _prev_true_peak: peaks of the frames of the last
    add_frames call only (interpolator history = the audio before it), after
    every call of a ragged sequence; sample peak exact, true peak within 1e-6."""
    rng = np.random.default_rng(1234 + rate + channels)
    n = rate // 2
    x = rng.uniform(-0.9, 0.9, size=(n, channels)) * np.linspace(0.2, 1.0, n)[:, None]
    pcm = np.round(x * 32767).astype(np.int16) if dtype == np.int16 else x.astype(np.float32)
    a, b = product.init(channels, rate), oracle.init(channels, rate)
    pos = 0
    for count in (1, 7, 1024, 4097, 333, n):
        count = min(count, n - pos)
        if count == 0:
            break
        for st in (a, b):
            st.add_frames(pcm[pos:pos + count])
        pos += count
        for ch in range(channels):
            assert a.prev_sample_peak(ch) == b.prev_sample_peak(ch)
            assert a.prev_true_peak(ch) == pytest.approx(b.prev_true_peak(ch), rel=1e-6, abs=0)
    # the running peaks are unaffected by the queries
    assert a.sample_peaks() == b.sample_peaks()
    assert np.allclose(a.true_peaks(), b.true_peaks(), rtol=1e-6, atol=0)
    a.destroy(); b.destroy()


def test_known_answers(product):
    """EBU Tech 3341 / 3342 / BS.1770 cases straight through the CUDA path."""
    for name, (pcm, rate, want, tol) in cases.loudness_cases().items():
        with product.init(pcm.shape[1], rate) as st:
            st.add_frames(pcm, 4800)
            assert abs(st.loudness_global() - want) <= tol, name
    for name, (pcm, rate, want, tol) in cases.range_cases().items():
        with product.init(pcm.shape[1], rate) as st:
            st.add_frames(pcm, 4800)
            assert abs(st.loudness_range() - want) <= tol, name
    for name, (pcm, rate, want, up, down) in cases.true_peak_cases().items():
        with product.init(pcm.shape[1], rate) as st:
            st.add_frames(pcm, 4800)
            db = 20 * np.log10(max(st.true_peaks()))
            assert want - down <= db <= want + up, name


def test_true_peak_at_chunk_offsets(product, oracle):
    rate = 44100
    for off in (0, 1, 5, 11, 12, 13, 200, 293, 294, 440, 441):
        pcm = np.zeros((44100, 1), dtype=np.int16)
        k = 22050 + off
        pcm[k - 1:k + 1, 0] = 30000
        pcm[k - 3:k - 1, 0] = -9000
        pcm[k + 1:k + 3, 0] = -9000
        o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
        g = _drive(product, [(pcm, rate)], None, 4096)["tracks"][0]
        assert rel_diff(g["true_peak"], o["true_peak"]) <= TOL_TP_REL, off


def test_gpu_matches_host_emulation(product):
    """The host compile of the device math (tests/emu) predicts the GPU: the
    FP32 sweep bit for bit (true peaks identical), the FP64 post-processing to
    rounding (nvcc contracts a*b+c into DFMA, the host build does not)."""
    import torch
    from loudgain_b200 import engine
    from tests.helpers import emu_measure
    spec = synth.config1_spec(20.0)
    pcm = synth.programme_s16(spec).numpy()
    b = engine.Batch([(torch.from_numpy(pcm).cuda(), spec.rate)])
    b.run()
    tres, _ = b.fetch()
    z = b.blocks(0, 0)
    b.close()
    # same chunk length as the GPU plan: ask the emulation for the same task target
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    e = emu_measure([(pcm, spec.rate)], target_tasks=sms * 2048)
    np.testing.assert_allclose(z, e["blocks"], rtol=1e-13)
    np.testing.assert_array_equal(tres[0].true_peak, e["tracks"][0]["true_peak"])
    np.testing.assert_array_equal(tres[0].sample_peak, e["tracks"][0]["sample_peak"])


def test_album_exchange_single_rank(product):
    """The cross-rank album path (lgb_exchange_*: publish / gate / finish kernels over
    the exchange region) with one rank must give what the plain album query gives, on
    every repeat (the steps alternate between the region's two halves and the second
    run on replays the step as a CUDA graph)."""
    import torch
    from loudgain_b200 import engine

    specs = synth.config2_specs(ntracks=6, scale=0.08)
    tracks = [(synth.programme_s16(s, device="cuda"), s.rate) for s in specs]
    albums = [0, 1, 0, 1, 0, 2]
    want_t, want_a = engine.measure(tracks, albums)
    b = engine.Batch(tracks, albums, nalbums=4)            # album 3 has no track at all
    x = engine.AlbumExchange(b)
    try:
        for _ in range(4):
            b.run()
            got_t, got_a = b.fetch()
            for w, g in zip(want_t, got_t):
                assert g.loudness == w.loudness and g.range == w.range
            for a in range(3):
                assert lu_diff(got_a[a].loudness, want_a[a].loudness) <= 1e-10
                assert got_a[a].range == want_a[a].range
                assert (got_a[a].n_abs, got_a[a].n_rel, got_a[a].n_shortterm) == \
                       (want_a[a].n_abs, want_a[a].n_rel, want_a[a].n_shortterm)
            assert got_a[3].loudness == -np.inf and got_a[3].range == 0.0
    finally:
        b.close()
        x.close()
    torch.cuda.synchronize()


def _oracle_parallel(oracle, tracks):
    """The oracle on full-size tracks, one thread per track (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(t):
        st = oracle.init(t[0].shape[1], t[1])
        st.add_frames(t[0], 1024)
        return st

    with ThreadPoolExecutor(min(len(tracks), 12)) as ex:
        return list(ex.map(one, tracks))


def test_config1_full_size(product, oracle):
    """cfg1 at BASELINE.json's size: the 3-minute 44.1 kHz stereo S16 track, 1024-frame calls."""
    spec = synth.config1_spec(180.0)
    pcm = synth.programme_s16(spec, device="cuda").cpu().numpy()
    (st,) = _oracle_parallel(oracle, [(pcm, spec.rate)])
    o = {"loudness": st.loudness_global(), "range": st.loudness_range(),
         "sample_peak": np.array(st.sample_peaks()), "true_peak": np.array(st.true_peaks())}
    st.destroy()
    g = _drive(product, [(pcm, spec.rate)], None, 1024)
    _check(o, g["tracks"][0])


def test_config2_full_size(product, oracle):
    """cfg2 at BASELINE.json's size: the whole 12-track album (247.6 M samples), per-track and
    album results, gains at tag precision, clipping prevention inputs (true peaks)."""
    specs = synth.config2_specs(12)
    tracks = [(synth.programme_s16(s, device="cuda").cpu().numpy(), s.rate) for s in specs]
    sts = _oracle_parallel(oracle, tracks)
    o_tracks = [{"loudness": st.loudness_global(), "range": st.loudness_range(),
                 "sample_peak": np.array(st.sample_peaks()), "true_peak": np.array(st.true_peaks())}
                for st in sts]
    o_album = {"loudness": oracle.loudness_global_multiple(sts), "range": oracle.loudness_range_multiple(sts)}
    for st in sts:
        st.destroy()
    g = _drive(product, tracks, [0] * len(tracks), 1024)
    for ot, gt in zip(o_tracks, g["tracks"]):
        _check(ot, gt)
    _check(o_album, g["albums"][0])


def test_two_runs_in_flight(product):
    """lgb_batch_run keeps two result mirrors: run k + 1 may be enqueued before run k is
    fetched; fetch returns the oldest unfetched run; a third run without a fetch is refused."""
    import torch
    from loudgain_b200 import engine

    specs = synth.config2_specs(ntracks=3, scale=0.05)
    tracks = [(synth.programme_s16(s, device="cuda"), s.rate) for s in specs]
    want_t, want_a = engine.measure(tracks, [0, 0, 0])
    b = engine.Batch(tracks, [0, 0, 0])
    try:
        with pytest.raises(RuntimeError):
            b.fetch()                                  # nothing has run yet
        b.run()
        b.run()
        with pytest.raises(RuntimeError):
            b.run()                                    # two are in flight
        for _ in range(6):                             # direct launches first, then the two graphs
            t, a = b.fetch()
            for w, g in zip(want_t, t):
                assert g.loudness == w.loudness and g.range == w.range
                np.testing.assert_array_equal(g.true_peak, w.true_peak)
            assert a[0].loudness == want_a[0].loudness and a[0].range == want_a[0].range
            b.run()
        b.fetch()
        b.fetch()
    finally:
        b.close()
    torch.cuda.synchronize()


@pytest.mark.parametrize("depth", [2, 3])
def test_pipelined_runs_see_their_own_pcm(product, depth):
    """From its second run on a batch is pipelined (lg_batch.cu: the post-processing of run k
    finishes on the library's own stream while the sweep of run k + 1 is under way).  The
    audio is REPLACED between runs here (a copy on the batch's stream), two or three runs in
    flight (lgb_batch_set_max_in_flight):
    every fetch must return the results of the audio its own run swept, bit for bit what a
    one-shot measurement of that audio gives -- no stale mirror, no result of the wrong run,
    no peak cell shared between two runs."""
    import torch
    from loudgain_b200 import engine

    specs = synth.config2_specs(ntracks=4, scale=0.06)
    v0 = [synth.programme_s16(s, device="cuda") for s in specs]
    v1 = [torch.flip(t, dims=[0]) // 2 if i % 2 else (t // 3) for i, t in enumerate(v0)]
    rates = [s.rate for s in specs]
    albums = [0, 0, 1, 1]
    want = [engine.measure(list(zip(v, rates)), albums) for v in (v0, v1)]
    bufs = [t.clone() for t in v0]
    b = engine.Batch(list(zip(bufs, rates)), albums)

    def load(which):
        for dst, src in zip(bufs, (v0, v1)[which]):
            dst.copy_(src, non_blocking=True)

    def check(which, got):
        want_t, want_a = want[which]
        got_t, got_a = got
        for w, g in zip(want_t, got_t):
            assert g.loudness == w.loudness and g.range == w.range
            np.testing.assert_array_equal(g.true_peak, w.true_peak)
            np.testing.assert_array_equal(g.sample_peak, w.sample_peak)
        for w, g in zip(want_a, got_a):
            assert g.loudness == w.loudness and g.range == w.range

    try:
        order = [0, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0]
        b.set_max_in_flight(depth)
        pending = []
        for which in order:
            load(which)
            b.run()
            pending.append(which)
            if len(pending) == depth:
                with pytest.raises(RuntimeError):
                    b.run()               # one too many in flight
                check(pending.pop(0), b.fetch())
        while pending:
            check(pending.pop(0), b.fetch())
        assert want[0][0][0].loudness != want[1][0][0].loudness      # the two variants do differ
    finally:
        b.close()
    torch.cuda.synchronize()
