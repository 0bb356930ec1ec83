#!/usr/bin/env python3
"""Writes tests/golden/r128_vectors.json: results of the CPU oracle
(oracle/ebur128_oracle.c, the restatement of the reference's libebur128 path)
on small seeded inputs, driven like scan.c (1024-frame add_frames_short calls).

The reference ships no fixtures for this boundary and libebur128 is not
available in the build image, so these vectors do NOT
come from the reference: they freeze the oracle's behaviour (any later change
of the oracle, the synthesiser or the product shows up as a diff against
them) and travel to the GPU box, where the product is checked against them.

    python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from loudgain_b200 import synth  # noqa: E402
from oracle import load_oracle  # noqa: E402


def cases():
    """name -> (int16 pcm [frames, ch], rate)."""
    out = {}
    s = synth.config1_spec(12.0)
    out["cfg1_12s_44k1_stereo"] = (synth.programme_s16(s).numpy(), s.rate)
    s = synth.config3_spec(6.0)
    out["cfg3_6s_96k_5p1"] = (synth.programme_s16(s).numpy(), s.rate)
    s = synth.config4_spec(9.0)
    out["cfg4_9s_48k_stereo"] = (synth.programme_s16(s).numpy(), s.rate)
    rng = np.random.default_rng(20261018)
    out["noise_5s_22k05_mono"] = ((rng.standard_normal((22050 * 5, 1)) * 6000).astype(np.int16), 22050)
    out["noise_4s_192k_stereo"] = ((rng.standard_normal((192000 * 4, 2)) * 3000).astype(np.int16), 192000)
    t = np.arange(48000 * 8) / 48000.0
    sine = np.round(10 ** (-23 / 20) * 32767 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16)
    out["sine_1k_-23dbfs_8s_48k_stereo"] = (np.stack([sine, sine], axis=1), 48000)
    clip = np.clip(rng.standard_normal((44100 * 4, 2)) * 30000, -32768, 32767).astype(np.int16)
    out["clipped_4s_44k1_stereo"] = (clip, 44100)
    out["short_0s3_44k1_stereo"] = ((rng.standard_normal((13230, 2)) * 5000).astype(np.int16), 44100)
    return out


def main():
    lib = load_oracle()
    vec = {}
    for name, (pcm, rate) in cases().items():
        st = lib.init(pcm.shape[1], rate)
        st.add_frames(pcm, 1024)
        loud = st.loudness_global()
        vec[name] = {"rate": rate, "channels": int(pcm.shape[1]), "frames": int(pcm.shape[0]),
                     "pcm_crc32": int(__import__("zlib").crc32(np.ascontiguousarray(pcm).tobytes())),
                     "loudness": None if np.isinf(loud) else loud, "range": st.loudness_range(),
                     "sample_peak": [float(x) for x in st.sample_peaks()],
                     "true_peak": [float(x) for x in st.true_peaks()]}
        st.destroy()
    path = os.path.join(ROOT, "tests", "golden", "r128_vectors.json")
    with open(path, "w") as f:
        json.dump({"source": "oracle/ebur128_oracle.c via tools/make_golden.py (not libebur128: "
                             "parity unpinned, DESIGN.md section 2)", "vectors": vec}, f, indent=1)
    print(path, len(vec), "vectors")


if __name__ == "__main__":
    main()
