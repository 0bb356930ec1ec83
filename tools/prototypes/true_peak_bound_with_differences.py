"""What a tighter screening bound would buy the true-peak pass ("next").

Today the sweep rules a 24-frame pair out when ||c||_1 * max|x| over its window (the pair and the
12 frames of history) cannot beat the channel's running peak.  A polyphase output is
    y = sum_k c_k x[n-k] = s * x[a] + sum_k c_k (x[n-k] - x[a]),   s = sum_k c_k,
and |x[n-k] - x[a]| <= |k - a'| * D with D the largest first difference in the window, so
    |y| <= |s| * M + G * D,     G = sum_k |c_k| |k - a'|   (a' = the anchor tap that minimises it)
is a bound too -- and the sweep computes the first difference anyway (the high-pass works on
it).  Near its peaks programme material moves slowly from sample to sample, so this bound is
much tighter than 1.8645 * M.  This script (CPU, numpy): the constants from the library's own tap
table, a brute-force check that the bound holds, and the share of pairs each bound leaves as
candidates on the bench's album.

    python tools/prototypes/true_peak_bound_with_differences.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loudgain_b200 import synth  # noqa: E402


def taps4():
    src = open(os.path.join(ROOT, "loudgain_b200", "csrc", "lg_tp_coefs.h")).read()
    body = src[src.index("kTp4d[3][12]"):]
    body = body[:body.index("};")]
    rows = re.findall(r"\{([^{}]*)\}", body)
    return np.array([[float.fromhex(v.strip()) for v in r.split(",") if v.strip()] for r in rows])


C = taps4()                                   # [3 phases][12 taps], tap t multiplies x[n - t]
L1 = np.abs(C).sum(1)
S = C.sum(1)
K = np.arange(12)
G = np.array([min((np.abs(c) * np.abs(K - a)).sum() for a in range(12)) for c in C])
print("phase  ||c||_1   sum c     G (anchor at the best tap)")
for p in range(3):
    print(f"  {p + 1}    {L1[p]:.4f}   {S[p]:+.4f}   {G[p]:.4f}")
L1M, SM, GM = L1.max() * 1.0002, np.abs(S).max() * 1.0002, G.max() * 1.0002


def outputs(x):
    """max over the three phases of |y| for every frame n >= 11 of a mono float64 signal."""
    w = np.lib.stride_tricks.sliding_window_view(x, 12)[:, ::-1]      # w[n-11][t] = x[n - t]
    return np.abs(w @ C.T).max(1)


def pair_maxima(v, width):
    """max of v over [24 p - width + 24, 24 p + 24) for every pair p (zeros before the start)."""
    n = (len(v) // 24) * 24
    m24 = v[:n].reshape(-1, 24).max(1)
    if width == 24:
        return m24
    prev = np.concatenate([[0.0], m24[:-1]])
    return np.maximum(m24, prev)              # (the sweep's window: this pair and the one before)


def check_bound(rng):
    worst = 0.0
    for kind in range(6):
        n = 24 * 4000
        if kind == 0: x = rng.uniform(-1, 1, n)
        elif kind == 1: x = np.sign(rng.standard_normal(n))
        elif kind == 2: x = np.clip(np.cumsum(rng.standard_normal(n)) * 0.05, -1, 1)
        elif kind == 3: x = np.sin(np.arange(n) * rng.uniform(0.01, 3.1))
        elif kind == 4: x = np.clip(rng.standard_normal(n) * 2, -1, 1)
        else: x = rng.standard_normal(n) * np.exp(-np.arange(n) / 3000.0)
        y = np.concatenate([np.zeros(11), outputs(x)])
        M = pair_maxima(np.abs(x), 48)
        D = pair_maxima(np.abs(np.diff(x, prepend=0.0)), 48)
        Y = pair_maxima(y, 24)
        bound = np.minimum(L1M * M, SM * M + GM * D)
        worst = max(worst, float((Y - bound).max()))
    return worst


def shares(pcm, peak):
    x = pcm.astype(np.float64) / 32768.0
    M = pair_maxima(np.abs(x), 48)
    D = pair_maxima(np.abs(np.diff(x, prepend=0.0)), 48)
    old = L1M * M > peak
    new = np.minimum(L1M * M, SM * M + GM * D) > peak
    return old.mean(), new.mean()


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    print(f"largest (true maximum - bound) over six adversarial signal classes: {check_bound(rng):+.3e} (must be <= 0)")
    specs = synth.config2_specs(12, scale=0.25)
    tot_old = tot_new = 0.0
    print("track  channel  candidates today   with the difference term   (against the final sample peak)")
    for i, s in enumerate(specs):
        pcm = synth.programme_s16(s).numpy()
        for ch in range(2):
            peak = np.abs(pcm[:, ch].astype(np.float64)).max() / 32768.0
            o, n = shares(pcm[:, ch], peak)
            tot_old += o / 24
            tot_new += n / 24
            if i < 4:
                print(f"  {i:2d}      {ch}        {o:7.4f}            {n:7.4f}")
    print(f"album mean: {tot_old:.4f} -> {tot_new:.4f}  ({tot_old / max(tot_new, 1e-12):.1f}x fewer)")
