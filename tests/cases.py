"""Synthesised known-answer cases: EBU Tech 3341 / 3342 and ITU-R BS.1770
(SURVEY.md Appendix B).  The reference ships no fixtures for the ebur128
boundary, so these bound gross errors of any implementation of the path
(oracle and CUDA alike); fine parity is CUDA-vs-oracle."""
from __future__ import annotations

import numpy as np

RATE = 48000


def _sine(db, seconds, freq=1000.0, rate=RATE, phase=0.0, start=0):
    n = int(round(seconds * rate))
    t = (np.arange(n) + start) / rate
    return 10.0 ** (db / 20.0) * np.sin(2 * np.pi * freq * t + phase)


def _stereo(parts, rate=RATE):
    """parts = [(dBFS, seconds)]: identical in-phase 1 kHz sine in L and R."""
    segs, pos = [], 0
    for db, sec in parts:
        s = _sine(db, sec, rate=rate, start=pos)
        pos += len(s)
        segs.append(s)
    x = np.concatenate(segs)
    return np.stack([x, x], axis=1).astype(np.float32)


def loudness_cases():
    """name -> (pcm float32 [frames, ch], rate, expected LUFS, tolerance)."""
    c = {}
    c["3341-1"] = (_stereo([(-23.0, 20)]), RATE, -23.0, 0.1)
    c["3341-2"] = (_stereo([(-33.0, 20)]), RATE, -33.0, 0.1)
    c["3341-3"] = (_stereo([(-36.0, 10), (-23.0, 60), (-36.0, 10)]), RATE, -23.0, 0.1)
    c["3341-4"] = (_stereo([(-72.0, 10), (-36.0, 10), (-23.0, 60), (-36.0, 10), (-72.0, 10)]),
                   RATE, -23.0, 0.1)
    c["3341-5"] = (_stereo([(-26.0, 20), (-20.0, 20.1), (-26.0, 20)]), RATE, -23.0, 0.1)
    # 3341-6: 5.0 channels L R C Ls Rs
    n = 20 * RATE
    t = np.arange(n) / RATE
    s = np.sin(2 * np.pi * 1000.0 * t)
    g = lambda db: 10.0 ** (db / 20.0)
    five = np.stack([g(-28) * s, g(-28) * s, g(-24) * s, g(-30) * s, g(-30) * s], axis=1)
    c["3341-6"] = (five.astype(np.float32), RATE, -23.0, 0.1)
    # BS.1770: 0 dBFS 997 Hz in the left channel only
    x = np.zeros((n, 2), dtype=np.float32)
    x[:, 0] = np.sin(2 * np.pi * 997.0 * t)
    c["bs1770-997"] = (x, RATE, -3.01, 0.02)
    return c


def range_cases():
    """name -> (pcm, rate, expected LRA, tolerance)."""
    c = {}
    c["3342-1"] = (_stereo([(-20.0, 20), (-30.0, 20)]), RATE, 10.0, 1.0)
    c["3342-2"] = (_stereo([(-20.0, 20), (-15.0, 20)]), RATE, 5.0, 1.0)
    c["3342-3"] = (_stereo([(-40.0, 20), (-20.0, 20)]), RATE, 20.0, 1.0)
    c["3342-4"] = (_stereo([(-50.0, 20), (-35.0, 20), (-20.0, 20), (-35.0, 20), (-50.0, 20)]),
                   RATE, 15.0, 1.0)
    return c


def true_peak_cases():
    """name -> (pcm, rate, expected dBTP, +tol, -tol)."""
    c = {}
    n = RATE
    t = np.arange(n)

    # 20 ms raised-cosine fades keep the band-limited onset/offset transient
    # (Gibbs overshoot of an abruptly gated tone) out of the peak reading.
    ramp = 0.5 * (1 - np.cos(np.pi * np.arange(960) / 960.0))
    fade = np.ones(n)
    fade[:960] = ramp
    fade[-960:] = ramp[::-1]

    def tone(div, amp, phase_deg):
        x = amp * np.sin(2 * np.pi * t / div + np.deg2rad(phase_deg)) * fade
        return np.stack([x, x], axis=1).astype(np.float32)

    c["3341-15"] = (tone(4, 0.5, 0.0), RATE, -6.0, 0.2, 0.4)
    c["3341-16"] = (tone(4, 0.5, 45.0), RATE, -6.0, 0.2, 0.4)
    c["3341-17"] = (tone(6, 0.5, 60.0), RATE, -6.0, 0.2, 0.4)
    c["3341-18"] = (tone(8, 0.5, 67.5), RATE, -6.0, 0.2, 0.4)
    c["3341-19"] = (tone(4, 1.41, 45.0), RATE, 3.0, 0.2, 0.4)
    return c


def to_s16(x: np.ndarray) -> np.ndarray:
    return np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)


def dc_offset_quiet_programme(rate, dc, kind, channels=2, seconds=9.0, seed=5):
    """A quiet programme riding on a large constant offset (float64 [frames, ch]).

    kind "tone": 1 kHz at -60 dBFS swept +/-10 dB; "noise": white noise around
    -50 dBFS swept +/-8 dB.  Channel 0 carries offset `dc`, the others 0.7 * dc
    with the programme inverted.  The offset starts at frame 0 (a real step the
    reference's filter sees too); afterwards the K-weighted output is 1e4..1e5
    times smaller than the offset, which is what breaks an implementation that
    lets the offset into its filter state in single precision."""
    t = np.arange(int(rate * seconds)) / rate
    rng = np.random.default_rng(seed)
    if kind == "tone":
        sig = 10 ** ((-60 + 10 * np.sin(2 * np.pi * 0.13 * t)) / 20) * np.sin(2 * np.pi * 1000 * t)
    else:
        sig = 10 ** ((-50 + 8 * np.sin(2 * np.pi * 0.13 * t)) / 20) * 0.5 * rng.standard_normal(len(t))
    cols = [dc + sig] + [0.7 * dc - sig] * (channels - 1)
    return np.stack(cols, axis=1)
