"""CPU check of the kernel MATH before any GPU time is spent.

tests/emu compiles the per-thread device functions of loudgain_b200/csrc
(lg_sweep.cuh, lg_post.cuh, lg_plan.h) for the host -- explicit fmaf, so bit
for bit what the GPU computes -- and runs them sequentially.  These tests
compare that against the oracle.  This emulation is test infrastructure; the
product library contains no CPU path (see tests/test_abi.py).
"""
import numpy as np
import pytest

from loudgain_b200 import synth
from tests import cases
from tests.helpers import (GOAL_LU, NO_ALBUM, TOL_TP_REL, emu_measure, lu_diff, oracle_measure,
                           rel_diff)

ABS_GATE = 10 ** ((-70 + 0.691) / 10)


def _check(o, e, goal=GOAL_LU):
    assert lu_diff(e["loudness"], o["loudness"]) <= goal
    assert lu_diff(e["range"], o["range"]) <= goal
    np.testing.assert_array_equal(e["sample_peak"], o["sample_peak"])   # bit-exact
    assert rel_diff(e["true_peak"], o["true_peak"]) <= TOL_TP_REL
    # the device's two-pass scheme (sweep records iteration maxima, the true-peak
    # pass evaluates only what can still raise the peak) loses nothing
    np.testing.assert_array_equal(e["true_peak_screened"], e["true_peak"])


@pytest.mark.parametrize("target_tasks", [0, 500, 20])
def test_programme_stereo_s16(oracle, target_tasks):
    """cfg1 shape (44.1 kHz stereo S16), three chunk lengths."""
    spec = synth.config1_spec(35.0)
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, spec.rate)])["tracks"][0]
    e = emu_measure([(pcm, spec.rate)], target_tasks=target_tasks)
    _check(o, e["tracks"][0])
    got = e["blocks"][e["blocks"] >= ABS_GATE]
    assert len(got) == len(o["blocks"])
    assert rel_diff(got, o["blocks"]) < 2e-5
    got = e["st"][e["st"] >= ABS_GATE]
    assert len(got) == len(o["st"]) and rel_diff(got, o["st"]) < 2e-5


@pytest.mark.parametrize("rate,channels,fmt", [
    (48000, 2, "s16"), (48000, 2, "f32"), (44100, 1, "s16"), (22050, 2, "s16"),
    (32000, 1, "f32"), (96000, 6, "s16"), (88200, 2, "s16"), (192000, 2, "s16"),
    (48000, 5, "s16"), (48000, 4, "f32"), (44100, 3, "s16"), (11025, 2, "s16"),
])
def test_rates_channels_formats(oracle, rate, channels, fmt):
    spec = synth.TrackSpec(seed=rate + channels, rate=rate, channels=channels, seconds=8.0,
                           lfe_channel=3 if channels == 6 else None, surround_db=1.5)
    x = synth.programme_float(spec)
    pcm = synth.quantise_s16(x).numpy() if fmt == "s16" else x.numpy()
    o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
    e = emu_measure([(pcm, rate)])["tracks"][0]
    _check(o, e)


def test_album_union(oracle):
    """cfg2 shape, scaled down: per-track and album results."""
    specs = synth.config2_specs(ntracks=4, scale=0.06)
    tracks = [(synth.programme_s16(s).numpy(), s.rate) for s in specs]
    albums = [0, 0, 0, 0]
    o = oracle_measure(oracle, tracks, albums)
    e = emu_measure(tracks, albums)
    for ot, et in zip(o["tracks"], e["tracks"]):
        _check(ot, et)
    assert lu_diff(e["albums"][0]["loudness"], o["albums"][0]["loudness"]) <= GOAL_LU
    assert lu_diff(e["albums"][0]["range"], o["albums"][0]["range"]) <= GOAL_LU


def test_mixed_rate_album_and_loose_track(oracle):
    rng = np.random.default_rng(3)
    a = (rng.standard_normal((48000 * 5, 2)) * 2000).astype(np.int16)
    b = (rng.standard_normal((96000 * 4, 1)) * 0.05).astype(np.float32)
    c = (rng.standard_normal((44100 * 6, 2)) * 6000).astype(np.int16)
    tracks = [(a, 48000), (b, 96000), (c, 44100)]
    albums = [0, 0, NO_ALBUM]
    o = oracle_measure(oracle, tracks, albums)
    e = emu_measure(tracks, albums)
    for ot, et in zip(o["tracks"], e["tracks"]):
        _check(ot, et)
    assert lu_diff(e["albums"][0]["loudness"], o["albums"][0]["loudness"]) <= GOAL_LU


@pytest.mark.parametrize("frames", [0, 1, 11, 73, 4409, 4410, 17639, 17640, 17641, 132299, 132300])
def test_ragged_lengths(oracle, frames):
    """Empty, sub-block, exactly-one-block and exactly-3-s inputs."""
    rng = np.random.default_rng(frames)
    pcm = (rng.standard_normal((frames, 2)) * 5000).astype(np.int16)
    o = oracle_measure(oracle, [(pcm, 44100)])["tracks"][0]
    e = emu_measure([(pcm, 44100)])["tracks"][0]
    _check(o, e)
    assert e["n_abs"] == len(o["blocks"]) and e["n_st"] == len(o["st"])


def test_dc_offset_and_bass(oracle):
    """Zero-start chunks see a large high-pass transient; the correction must
    cancel it (DC offset) and carry slow state exactly (41 Hz tone)."""
    rate = 44100
    t = np.arange(rate * 10) / rate
    rng = np.random.default_rng(9)
    dc = 0.05 + 0.003 * rng.standard_normal(len(t))
    bass = 0.4 * np.sin(2 * np.pi * 41.0 * t) + 0.001 * rng.standard_normal(len(t))
    for sig in (dc, bass):
        pcm = cases.to_s16(np.stack([sig, -sig], axis=1))
        o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
        e = emu_measure([(pcm, rate)])["tracks"][0]
        _check(o, e)


@pytest.mark.parametrize("rate,channels,fmt", [(44100, 2, "s16"), (44100, 2, "f32"), (48000, 2, "s16"),
                                               (96000, 2, "f32"), (96000, 6, "s16"), (48000, 1, "f32"),
                                               (44100, 1, "s16")])
@pytest.mark.parametrize("kind", ["tone", "noise"])
def test_dc_offset_quiet_programme(oracle, rate, channels, fmt, kind):
    """a -60 dBFS programme on an offset of 0.1 .. 0.9 FS.
    Block energies within 1e-5 relative of the reference's double-precision
    direct form II (oracle/ebur128_oracle.c: filter loop; scan.c:448), loudness
    and range within 2e-4 LU."""
    for dc in (0.1, 0.3, 0.6, 0.9):
        x = cases.dc_offset_quiet_programme(rate, dc, kind, channels)
        pcm = cases.to_s16(x) if fmt == "s16" else x.astype(np.float32)
        o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
        e = emu_measure([(pcm, rate)])
        _check(o, e["tracks"][0])
        got = e["blocks"][e["blocks"] >= ABS_GATE]
        assert len(got) == len(o["blocks"])
        assert rel_diff(got, o["blocks"]) <= 1e-5, dc
        got = e["st"][e["st"] >= ABS_GATE]
        assert len(got) == len(o["st"]) and rel_diff(got, o["st"]) <= 1e-5, dc


def test_full_scale_square_and_impulses(oracle):
    rate = 48000
    n = rate * 5
    sq = np.where((np.arange(n) // 37) % 2 == 0, 32767, -32768).astype(np.int16)
    imp = np.zeros(n, dtype=np.int16)
    imp[::1000] = 32767
    imp[500::1000] = -32768
    pcm = np.stack([sq, imp], axis=1)
    o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
    e = emu_measure([(pcm, rate)])["tracks"][0]
    _check(o, e)
    assert e["true_peak"][0] > 1.2      # inter-sample overs of a clipped square


def test_ebu_cases_through_emulation(oracle):
    for name, (pcm, rate, want, tol) in cases.loudness_cases().items():
        e = emu_measure([(pcm, rate)])["tracks"][0]
        assert abs(e["loudness"] - want) <= tol, name
    for name, (pcm, rate, want, tol) in cases.range_cases().items():
        e = emu_measure([(pcm, rate)])["tracks"][0]
        assert abs(e["range"] - want) <= tol, name
    for name, (pcm, rate, want, up, down) in cases.true_peak_cases().items():
        e = emu_measure([(pcm, rate)])["tracks"][0]
        db = 20 * np.log10(e["true_peak"].max())
        assert want - down <= db <= want + up, name


@pytest.mark.parametrize("rate", [44100, 96000])
def test_true_peak_at_every_chunk_offset(oracle, rate):
    """An inter-sample over placed at each offset inside a chunk (and across
    chunk and warm-up boundaries) must be seen with full history."""
    L = emu_measure([(np.zeros((rate, 1), dtype=np.int16), rate)])["chunk_len"][0]
    base = 20 * L
    for off in list(range(0, 30)) + [L // 2, L - 13, L - 12, L - 2, L - 1]:
        pcm = np.zeros((40 * L, 1), dtype=np.int16)
        k = base + off
        pcm[k - 1:k + 1, 0] = 30000           # two equal samples: peak lies between them
        pcm[k - 3:k - 1, 0] = -9000
        pcm[k + 1:k + 3, 0] = -9000
        o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
        e = emu_measure([(pcm, rate)])["tracks"][0]
        assert o["true_peak"][0] > o["sample_peak"][0] * 1.05
        assert rel_diff(e["true_peak"], o["true_peak"]) <= TOL_TP_REL, off


def test_time_segments_with_lead_in(oracle):
    """cfg4's time sharding, on the host emulation: a stream cut into segments
    that start one second early (lead-in instead of a filter-state exchange);
    the concatenated slot energies give the whole stream's loudness and range,
    and the maxima of the segments' peaks the whole stream's peaks."""
    from loudgain_b200.engine import segment_plan
    from tests.helpers import gate_slots
    rate = 48000
    spec = synth.config1_spec(47.3)
    spec.rate = rate
    pcm = synth.programme_s16(spec).numpy()
    o = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
    s100 = (rate + 5) // 10
    for parts in (2, 3):
        plan = segment_plan(len(pcm), rate, parts)
        assert plan[0][:2] == (0, 0) and plan[-1][2] == len(pcm)
        segs = [(pcm[a:e], rate) for a, _, e in plan]
        e = emu_measure(segs, lead_in=[l for _, l, _ in plan])
        slots = np.concatenate([s[l // s100:] for s, (_, l, _) in zip(e["slots"], plan)])
        assert len(slots) == len(pcm) // s100
        loud, rng = gate_slots(slots, rate)
        assert lu_diff(loud, o["loudness"]) <= GOAL_LU and lu_diff(rng, o["range"]) <= GOAL_LU
        sp = np.max([t["sample_peak"] for t in e["tracks"]], axis=0)
        tp = np.max([t["true_peak"] for t in e["tracks"]], axis=0)
        tps = np.max([t["true_peak_screened"] for t in e["tracks"]], axis=0)
        np.testing.assert_array_equal(sp, o["sample_peak"])
        assert rel_diff(tp, o["true_peak"]) <= TOL_TP_REL
        np.testing.assert_array_equal(tps, tp)


@pytest.mark.parametrize("rate,seconds,dtype", [(44100, 61.3, np.int16), (48000, 33.0, np.int16),
                                              (96000, 20.0, np.float32), (22050, 45.0, np.float32),
                                              (44100, 0.9, np.int16), (32000, 12.0, np.int16),
                                              (11025, 30.0, np.float32), (192000, 6.0, np.int16)])
def test_run_view_of_a_plan(rate, seconds, dtype, monkeypatch):
    """The run sweep's view of a stereo track (lg_run.cu stages rows of a 2-D tensor,
    row = run, through TMA): 16-byte pitched rows, stages that tile warm-up + run,
    complete / partial rows classified right, every chunk owned by exactly one
    lane -- on small and machine-filling plans."""
    import ctypes as C
    from tests.helpers import LgbTrack, build_emu
    lib = C.CDLL(build_emu())
    lib.emu_check_run_view.restype = C.c_longlong
    n = int(rate * seconds)
    pcm = np.zeros((n, 2), dtype=dtype)
    arr = (LgbTrack * 2)(LgbTrack(pcm.ctypes.data, n, 2, rate, 0 if dtype == np.int16 else 1,
                                  0xffffffff, None, 0),
                         LgbTrack(pcm.ctypes.data, n // 3, 2, rate, 0 if dtype == np.int16 else 1,
                                  0xffffffff, None, 0))
    full = C.c_longlong()
    for sms in ("148", "4", "1"):
        monkeypatch.setenv("LOUDGAIN_B200_SMS", sms)
        got = lib.emu_check_run_view(arr, C.c_size_t(2), C.c_uint64(0), C.byref(full))
        assert got >= 2, f"item {-1 - got} has an inconsistent view"
        assert full.value <= got


def test_run_sweep_grid():
    """CTAs of the run sweep (lg_common.h: run_grid_ctas).  One round of work items: the SMs that
    are used are filled up to the fullest sub-partition's level (warp w runs on sub-partition
    w % 4) and the rest are left to the kernels of the run before (pipelined runs); several
    rounds, or a tuning switch: every SM."""
    import ctypes as C
    from tests.helpers import build_emu
    lib = C.CDLL(build_emu())
    lib.emu_run_grid.restype = C.c_uint32
    grid = lambda n, sms=148, warps=16, spare=1: lib.emu_run_grid(n, sms, warps, spare)
    assert grid(2200) == 138                 # the 12-track album: 15 items per SM -> 16, ten SMs free
    assert grid(2200, spare=0) == 148
    assert grid(2368) == 148 and grid(2369) == 148 and grid(5000) == 148
    assert grid(100) == 100 and grid(148) == 148 and grid(0) == 0
    assert grid(1018) == 128                 # 7 per SM -> 8: two per sub-partition either way
    assert grid(1200) == 100                 # 9 per SM -> 12
    for n in range(149, 2369, 37):
        g = grid(n)
        per_sm = -(-n // 148)
        fill = min(16, 4 * (-(-per_sm // 4)))
        assert g <= 148 and g * fill >= n                     # every item has a warp in one round
        assert -(-(-(-n // g)) // 4) == -(-per_sm // 4)       # no sub-partition fuller than on 148 SMs
    assert grid(600, sms=4) == 4 and grid(3, sms=4) == 3


def test_tail_filler_plan(oracle, monkeypatch):
    """Planner option behind LOUDGAIN_B200_TAIL_FRAC / _TAIL_DIV: the tracks that hold the
    last part of the batch get shorter chunks (their own launch group); results
    stay within the parity goals, track by track and for the album."""
    specs = [synth.TrackSpec(seed=900 + i, rate=44100, channels=2, seconds=12.0 + 3 * i) for i in range(3)]
    tracks = [(synth.programme_s16(s).numpy(), s.rate) for s in specs]
    monkeypatch.setenv("LOUDGAIN_B200_RUN", "0")       # the option belongs to the one-chunk-per-lane sweeps
    o = oracle_measure(oracle, tracks, albums=[0, 0, 0])
    base = emu_measure(tracks, albums=[0, 0, 0], target_tasks=2500)
    monkeypatch.setenv("LOUDGAIN_B200_TAIL_FRAC", "0.4")
    monkeypatch.setenv("LOUDGAIN_B200_TAIL_DIV", "3")
    e = emu_measure(tracks, albums=[0, 0, 0], target_tasks=2500)
    assert list(base["chunk_len"]) == [base["chunk_len"][0]] * 3
    assert e["chunk_len"][0] == base["chunk_len"][0] and e["chunk_len"][2] < base["chunk_len"][2]
    for i in range(3):
        _check(o["tracks"][i], e["tracks"][i])
    assert lu_diff(e["albums"][0]["loudness"], o["albums"][0]["loudness"]) <= GOAL_LU
    assert lu_diff(e["albums"][0]["range"], o["albums"][0]["range"]) <= GOAL_LU
