"""Known-answer and behavioural tests of the CPU oracle (CPU only)."""
import numpy as np
import pytest

from loudgain_b200 import capi
from tests import cases

LOUD = cases.loudness_cases()
RANGE = cases.range_cases()
PEAK = cases.true_peak_cases()


@pytest.mark.parametrize("name", sorted(LOUD))
def test_integrated_loudness(oracle, name):
    pcm, rate, want, tol = LOUD[name]
    with oracle.init(pcm.shape[1], rate) as st:
        st.add_frames(pcm, 1024)
        assert abs(st.loudness_global() - want) <= tol


@pytest.mark.parametrize("name", sorted(RANGE))
def test_loudness_range(oracle, name):
    pcm, rate, want, tol = RANGE[name]
    with oracle.init(pcm.shape[1], rate) as st:
        st.add_frames(pcm, 1024)
        assert abs(st.loudness_range() - want) <= tol


@pytest.mark.parametrize("name", sorted(PEAK))
def test_true_peak(oracle, name):
    pcm, rate, want, up, down = PEAK[name]
    with oracle.init(pcm.shape[1], rate) as st:
        st.add_frames(pcm, 1024)
        db = 20 * np.log10(max(st.true_peaks()))
        assert want - down <= db <= want + up


def test_short_input_matches_float_within_quantisation(oracle):
    pcm, rate, want, tol = LOUD["3341-1"]
    with oracle.init(2, rate) as st:
        st.add_frames(cases.to_s16(pcm), 1024)
        assert abs(st.loudness_global() - want) <= tol


def test_chunking_invariance(oracle):
    """Results do not depend on how frames are split across calls (A.2)."""
    rng = np.random.default_rng(5)
    pcm = (rng.standard_normal((44100 * 8, 2)) * 4000).astype(np.int16)
    ref = None
    for split in (None, 1, 1024, 4410, 4409, 17640, 100000):
        with oracle.init(2, 44100) as st:
            if split == 1:
                st.add_frames(pcm[:3000], 1)
                st.add_frames(pcm[3000:], 7777)
            else:
                st.add_frames(pcm, split)
            got = (st.loudness_global(), st.loudness_range(), tuple(st.sample_peaks()),
                   tuple(st.true_peaks()))
        if ref is None:
            ref = got
        assert got == ref


def test_empty_short_and_silent(oracle):
    with oracle.init(2, 44100) as st:
        assert st.loudness_global() == -np.inf and st.loudness_range() == 0.0
        assert st.true_peaks() == [0.0, 0.0]
    with oracle.init(2, 44100) as st:          # shorter than one 400 ms block
        st.add_frames(np.full((17000, 2), 1000, dtype=np.int16))
        assert st.loudness_global() == -np.inf and st.loudness_range() == 0.0
        assert st.sample_peaks() == [1000 / 32768.0] * 2
    with oracle.init(2, 44100) as st:          # digital silence
        st.add_frames(np.zeros((44100 * 4, 2), dtype=np.int16))
        assert st.loudness_global() == -np.inf and st.loudness_range() == 0.0


def test_sample_peak_exact_and_negative_full_scale(oracle):
    pcm = np.zeros((48000, 2), dtype=np.int16)
    pcm[100, 0] = -32768
    pcm[200, 1] = 32767
    with oracle.init(2, 48000) as st:
        st.add_frames(pcm)
        assert st.sample_peaks() == [1.0, 32767 / 32768.0]


def test_channel_map_and_weights(oracle):
    """5.1 default map: index 3 (LFE) is ignored for loudness, counted for
    peaks; surrounds weigh 1.41."""
    rate = 48000
    t = np.arange(rate * 5) / rate
    s = (0.1 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    z = np.zeros_like(s)

    def loud(cols):
        with oracle.init(6, rate) as st:
            st.add_frames(np.stack(cols, axis=1))
            return st.loudness_global(), st.true_peaks()

    l_only, _ = loud([s, z, z, z, z, z])
    lfe_only, pk = loud([z, z, z, s, z, z])
    ls_only, _ = loud([z, z, z, z, s, z])
    assert lfe_only == -np.inf and pk[3] > 0.099
    assert abs((ls_only - l_only) - 10 * np.log10(1.41)) < 1e-9


def test_album_union_equals_concatenated_blocks(oracle):
    from oracle import blocks
    rng = np.random.default_rng(11)
    a = (rng.standard_normal((44100 * 6, 2)) * 3000).astype(np.int16)
    b = (rng.standard_normal((44100 * 9, 2)) * 300).astype(np.int16)
    sa, sb = oracle.init(2, 44100), oracle.init(2, 44100)
    sa.add_frames(a, 1024); sb.add_frames(b, 1024)
    z = np.concatenate([blocks(oracle, sa, 0), blocks(oracle, sb, 0)])
    thr = 0.1 * z.mean()
    want = 10 * np.log10(z[z >= thr].mean()) - 0.691
    assert abs(oracle.loudness_global_multiple([sa, sb]) - want) < 1e-9
    assert oracle.loudness_global_multiple([sa, None, sb]) == oracle.loudness_global_multiple([sa, sb])
    sa.destroy(); sb.destroy()


def test_mode_and_argument_errors(oracle):
    assert oracle.try_init(0, 44100) is None
    assert oracle.try_init(2, 3) is None
    st = oracle.init(2, 44100, capi.MODE_I)
    rc, _ = st._scalar("ebur128_loudness_range")
    assert rc == capi.ERROR_INVALID_MODE
    rc, _ = st._scalar("ebur128_true_peak", 0)
    assert rc == capi.ERROR_INVALID_MODE
    st.destroy()
    st = oracle.init(2, 44100)
    rc, _ = st._scalar("ebur128_true_peak", 2)
    assert rc == capi.ERROR_INVALID_CHANNEL_INDEX
    st.destroy()
    assert st.ptr is None


def test_192k_has_no_oversampling(oracle):
    rate = 192000
    t = np.arange(rate) / rate
    x = (0.5 * np.sin(2 * np.pi * 40000 * t + 0.3)).astype(np.float32).reshape(-1, 1)
    with oracle.init(1, rate) as st:
        st.add_frames(x)
        assert st.true_peak(0) == st.sample_peak(0)


def _meter_trace(oracle, pcm, rate, hop_s, query):
    """Feeds `pcm` in hops of `hop_s` seconds and reads the sliding-window meter after each."""
    hop = int(round(hop_s * rate))
    out = []
    with oracle.init(pcm.shape[1], rate) as st:
        for a in range(0, len(pcm) - hop + 1, hop):
            st.add_frames(pcm[a:a + hop], 1024)
            out.append(((a + hop) / rate, query(st)))
    return out


def test_tech3341_case9_shortterm_constant(oracle):
    """EBU Tech 3341 case 9: 1.34 s at -20 dBFS then 1.66 s at -30 dBFS, five times -- every 3 s
    window holds the same mix, so the short-term meter reads -23.0 +-0.1 LUFS from 3 s on."""
    pcm = cases._stereo([(-20.0, 1.34), (-30.0, 1.66)] * 5)
    trace = _meter_trace(oracle, pcm, cases.RATE, 0.1, lambda st: st.loudness_shortterm())
    late = [v for t, v in trace if t >= 3.0]
    assert len(late) > 100
    assert max(abs(v + 23.0) for v in late) <= 0.1


def test_tech3341_case12_momentary_constant(oracle):
    """EBU Tech 3341 case 12: 0.18 s at -20 dBFS then 0.22 s at -30 dBFS, repeated -- every
    400 ms window holds the same mix: the momentary meter reads -23.0 +-0.1 LUFS from 1 s on."""
    pcm = cases._stereo([(-20.0, 0.18), (-30.0, 0.22)] * 25)
    trace = _meter_trace(oracle, pcm, cases.RATE, 0.02, lambda st: st.loudness_momentary())
    late = [v for t, v in trace if t >= 1.0]
    assert len(late) > 400
    assert max(abs(v + 23.0) for v in late) <= 0.1


@pytest.mark.parametrize("lead_ms", [0, 150, 1000, 2850])
def test_tech3341_cases10_13_window_maxima(oracle, lead_ms):
    """EBU Tech 3341 cases 10 and 13: a -23 dBFS tone of exactly one window's length (3 s /
    0.4 s) behind silences of different lengths and followed by silence -- the maximum of the
    short-term / momentary meter is -23.0 +-0.1 LUFS wherever the tone starts."""
    for window_s, query, hop in ((3.0, lambda st: st.loudness_shortterm(), 0.05),
                                 (0.4, lambda st: st.loudness_momentary(), 0.01)):
        pcm = cases._stereo([(-200.0, lead_ms / 1000.0 + 1e-9), (-23.0, window_s), (-200.0, 1.0)])
        trace = _meter_trace(oracle, pcm, cases.RATE, hop, query)
        best = max(v for _, v in trace)
        assert abs(best + 23.0) <= 0.1, (window_s, lead_ms, best)


def test_tech3341_cases11_14_rising_maxima(oracle):
    """EBU Tech 3341 cases 11 and 14 (in their structure): tones of one window's length at
    -38, -36 ... -20 dBFS, each between silences longer than the window -- the meter's maxima
    within the successive segments step up by the same amounts, +-0.1 LU."""
    levels = list(range(-38, -18, 2))
    for window_s, query, hop in ((3.0, lambda st: st.loudness_shortterm(), 0.1),
                                 (0.4, lambda st: st.loudness_momentary(), 0.02)):
        gap = window_s + 0.6
        parts = []
        for lv in levels:
            parts += [(-200.0, gap), (float(lv), window_s)]
        pcm = cases._stereo(parts + [(-200.0, gap)])
        trace = _meter_trace(oracle, pcm, cases.RATE, hop, query)
        seg = gap + window_s
        for i, lv in enumerate(levels):
            lo, hi = i * seg + gap, (i + 1) * seg + gap - 0.5
            best = max(v for t, v in trace if lo <= t <= hi)
            assert abs(best - lv) <= 0.1, (window_s, lv, best)
