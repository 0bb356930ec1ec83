"""Deterministic synthetic PCM for the five BASELINE.json configs (SURVEY.md
section 8(d)).  torch is used as an array library only (CPU in tests, CUDA in
bench.py so that large inputs are born in HBM).

A "programme-like" track = a sum of sines (50 Hz - 12 kHz, 1/f amplitudes)
plus tilted noise, under a slow piecewise-constant level envelope spanning
about 25 dB, with two digital-silence gaps and one -60 dBFS bed, so that the
-70 LUFS absolute gate, the -10 LU relative gate and the loudness range all
have something to do.  Peaks are normalised close to full scale so that
inter-sample overs occur.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class TrackSpec:
    seed: int
    rate: int
    channels: int
    seconds: float
    level_db: float = 0.0          # overall level offset applied before quantising
    peak_dbfs: float = -0.3        # normalised sample peak (before level_db)
    lfe_channel: int | None = None  # this channel carries 30-80 Hz only, loud
    surround_db: float = 0.0       # boost of channels 4,5 (Ls, Rs)
    dc: float = 0.0                # DC offset in full-scale units

    @property
    def frames(self) -> int:
        return int(round(self.seconds * self.rate))


def _envelope_segments(rng: np.random.Generator, seconds: float):
    """Piecewise-constant gains (linear) with silence gaps and a quiet bed."""
    edges, gains = [0.0], []
    t = 0.0
    while t < seconds:
        dur = rng.uniform(2.0, 8.0)
        gains.append(10.0 ** (rng.uniform(-25.0, 0.0) / 20.0))
        t += dur
        edges.append(min(t, seconds))
    nseg = len(gains)
    if nseg >= 6:
        for idx in rng.choice(np.arange(1, nseg - 1), size=3, replace=False)[:2]:
            gains[idx] = 0.0                    # digital silence
        bed = int(rng.integers(1, nseg - 1))
        if gains[bed] != 0.0:
            gains[bed] = 10.0 ** (-60.0 / 20.0)  # -60 dBFS bed
    return np.asarray(edges), np.asarray(gains)


def _tilt_fir(device) -> torch.Tensor:
    """Short FIR giving white noise a gentle low-pass (pink-ish) tilt."""
    taps = np.array([0.049922035, -0.095993537, 0.050612699, -0.004408786])
    # impulse response of a 3-pole pink approximation, truncated
    b = np.array([0.049922035, -0.095993537, 0.050612699, -0.004408786])
    a = np.array([1.0, -2.494956002, 2.017265875, -0.522189400])
    h = np.zeros(64)
    x = np.zeros(64); x[0] = 1.0
    for n in range(64):
        acc = sum(b[k] * x[n - k] for k in range(4) if n - k >= 0)
        acc -= sum(a[k] * h[n - k] for k in range(1, 4) if n - k >= 0)
        h[n] = acc
    del taps
    return torch.tensor(h[::-1].copy(), dtype=torch.float32, device=device).view(1, 1, -1)


def programme_float(spec: TrackSpec, device="cpu") -> torch.Tensor:
    """Float32 [frames, channels] in full-scale units, before quantisation."""
    dev = torch.device(device)
    n, C, rate = spec.frames, spec.channels, spec.rate
    rng = np.random.default_rng(spec.seed)
    edges, gains = _envelope_segments(rng, spec.seconds)
    nsin = 40
    fmax = min(12000.0, 0.45 * rate)
    freqs = np.exp(rng.uniform(math.log(50.0), math.log(fmax), size=(C, nsin)))
    amps = (1.0 / np.sqrt(freqs / 50.0)) * rng.uniform(0.5, 1.0, size=(C, nsin))
    phases = rng.uniform(0, 2 * math.pi, size=(C, nsin))
    out = torch.empty((n, C), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    fir = _tilt_fir(dev)
    piece = rate * 10
    e_t = torch.tensor(edges[1:-1], dtype=torch.float64, device=dev)
    g_t = torch.tensor(gains, dtype=torch.float32, device=dev)
    for p0 in range(0, n, piece):
        m = min(piece, n - p0)
        gen.manual_seed(spec.seed * 1000003 + p0 // piece)
        t = (torch.arange(m, device=dev, dtype=torch.float64) + p0) / rate
        env = g_t[torch.bucketize(t, e_t, right=True)]
        for c in range(C):
            if spec.lfe_channel is not None and c == spec.lfe_channel:
                x = torch.zeros(m, dtype=torch.float64, device=dev)
                for f, ph in zip((31.0, 47.0, 63.0, 79.0), phases[c, :4]):
                    x += torch.sin(2 * math.pi * f * t + ph)
                out[p0:p0 + m, c] = (0.9 * x / 4.0).float()   # loud, not enveloped
                continue
            x = torch.zeros(m, dtype=torch.float64, device=dev)
            for f, a, ph in zip(freqs[c], amps[c], phases[c]):
                x += a * torch.sin(2 * math.pi * f * t + ph)
            x = x.float() / float(np.sqrt((amps[c] ** 2).sum() / 2.0)) * 0.25
            w = torch.randn(m + 63, generator=gen, device=dev, dtype=torch.float32)
            noise = torch.nn.functional.conv1d(w.view(1, 1, -1), fir).view(-1)
            x = x + noise * 1.5
            if c in (4, 5) and spec.surround_db:
                x = x * (10.0 ** (spec.surround_db / 20.0))
            out[p0:p0 + m, c] = x * env
    peak = out.abs().max().clamp_min(1e-9)
    out *= (10.0 ** (spec.peak_dbfs / 20.0)) / peak
    if spec.level_db:
        out *= 10.0 ** (spec.level_db / 20.0)
    if spec.dc:
        out += spec.dc
    return out


def quantise_s16(x: torch.Tensor) -> torch.Tensor:
    """round(x * 32767) clipped to int16, as SURVEY.md 8(d) specifies."""
    return torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)


def programme_s16(spec: TrackSpec, device="cpu") -> torch.Tensor:
    return quantise_s16(programme_float(spec, device))


# ------------------------------------------------------------------ configs

def config1_spec(seconds: float = 180.0) -> TrackSpec:
    """cfg1: one 44.1 kHz stereo 16-bit track, seed 17701."""
    return TrackSpec(seed=17701, rate=44100, channels=2, seconds=seconds)


def config2_specs(ntracks: int = 12, scale: float = 1.0) -> list[TrackSpec]:
    """cfg2: 12-track 44.1 kHz stereo album, durations 150-330 s, per-track
    level offsets -8..+4 dB; hot tracks exceed 0 dBTP so that -k matters."""
    rng = np.random.default_rng(17702)
    durs = rng.uniform(150.0, 330.0, size=ntracks) * scale
    levels = rng.uniform(-8.0, 4.0, size=ntracks)
    levels[:2] = (3.0, 4.0)
    return [TrackSpec(seed=17702 + i, rate=44100, channels=2, seconds=float(durs[i]),
                      level_db=float(levels[i])) for i in range(ntracks)]


def config3_spec(seconds: float = 120.0) -> TrackSpec:
    """cfg3: 96 kHz 5.1 (L R C LFE Ls Rs), loud LFE, surrounds +1.5 dB."""
    return TrackSpec(seed=17703, rate=96000, channels=6, seconds=seconds, lfe_channel=3,
                     surround_db=1.5)


def config4_spec(seconds: float = 36000.0) -> TrackSpec:
    """cfg4: one long 48 kHz stereo stream (10 h at full size)."""
    return TrackSpec(seed=17704, rate=48000, channels=2, seconds=seconds)


def config5_specs(ntracks: int = 10000, scale: float = 1.0) -> tuple[list[TrackSpec], list[int]]:
    """cfg5: library of mixed rate / channel-count tracks grouped into albums.
    Returns (specs, album index per track)."""
    rng = np.random.default_rng(17705)
    rates = np.array([22050, 32000, 44100, 48000, 88200, 96000, 192000])
    rate_p = np.array([2, 2, 60, 25, 3, 6, 2]) / 100.0
    chans = np.array([1, 2, 6])
    chan_p = np.array([10, 85, 5]) / 100.0
    specs, albums = [], []
    album = 0
    left = 0
    for i in range(ntracks):
        if left == 0:
            left = int(rng.integers(5, 21))
            if i:
                album += 1
        left -= 1
        rate = int(rng.choice(rates, p=rate_p))
        ch = int(rng.choice(chans, p=chan_p))
        dur = float(np.exp(rng.uniform(math.log(30.0), math.log(600.0)))) * scale
        specs.append(TrackSpec(seed=17705 + i, rate=rate, channels=ch, seconds=dur,
                               level_db=float(rng.uniform(-10.0, 2.0)),
                               lfe_channel=3 if ch == 6 else None))
        albums.append(album)
    return specs, albums
