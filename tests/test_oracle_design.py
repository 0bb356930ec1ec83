"""Pins the oracle's filter design against published constants: the ITU-R
BS.1770 48 kHz coefficient table, an independent scipy realisation, and the
structure of the 49-tap true-peak interpolator."""
import numpy as np
import pytest
import scipy.signal as sg

from oracle import kfilter_coeffs, tp_phase


def test_bs1770_48k_table(oracle):
    b, a = kfilter_coeffs(oracle, 48000)
    # ITU-R BS.1770-4 Table 1/2: stage 1 (shelf) and stage 2 (RLB high-pass).
    pb = np.array([1.53512485958697, -2.69169618940638, 1.19839281085285])
    pa = np.array([1.0, -1.69065929318241, 0.73248077421585])
    rb = np.array([1.0, -2.0, 1.0])
    ra = np.array([1.0, -1.99004745483398, 0.99007225036621])
    np.testing.assert_allclose(b, np.convolve(pb, rb), rtol=0, atol=2e-13)
    np.testing.assert_allclose(a, np.convolve(pa, ra), rtol=0, atol=2e-13)


def test_44k1_check_values(oracle):
    b, a = kfilter_coeffs(oracle, 44100)
    np.testing.assert_allclose(b, [1.53084123, -5.71266246, 8.0018803, -4.98913815, 1.16907908],
                               atol=5e-9)
    np.testing.assert_allclose(a, [1, -3.65282479, 5.01108676, -3.06315925, 0.70489871], atol=5e-9)
    radii = np.sort(np.abs(np.roots(a)))
    np.testing.assert_allclose(radii, [0.84415, 0.84415, 0.99458, 0.99458], atol=2e-5)


@pytest.mark.parametrize("rate", [22050, 32000, 44100, 48000, 88200, 96000, 192000])
def test_filter_matches_scipy(oracle, rate):
    """The oracle's block energies equal an independent float64 lfilter run."""
    from oracle import blocks
    rng = np.random.default_rng(rate)
    n = rate * 2
    x = (rng.standard_normal((n, 1)) * 0.1).astype(np.float32)
    b, a = kfilter_coeffs(oracle, rate)
    y = sg.lfilter(b, a, x[:, 0].astype(np.float64))
    s100 = (rate + 5) // 10
    nb = (n - 4 * s100) // s100 + 1
    ref = np.array([np.mean(y[k * s100:k * s100 + 4 * s100] ** 2) for k in range(nb)])
    with oracle.init(1, rate) as st:
        st.add_frames(x, 1000)
        got = blocks(oracle, st, 0)
    np.testing.assert_allclose(got, ref, rtol=1e-10)


def test_true_peak_phases(oracle):
    c0, s0 = tp_phase(oracle, 4, 0)
    assert list(s0) == [6] and c0[0] == 1.0
    c1, s1 = tp_phase(oracle, 4, 1)
    c2, s2 = tp_phase(oracle, 4, 2)
    c3, s3 = tp_phase(oracle, 4, 3)
    for s in (s1, s2, s3):
        assert list(s) == list(range(12))
    np.testing.assert_allclose(c3, c1[::-1], rtol=1e-12)      # mirror pair
    np.testing.assert_allclose(c2, c2[::-1], rtol=1e-12)      # symmetric
    np.testing.assert_allclose([c1.sum(), c2.sum(), c3.sum()], [1.00048, 1.00090, 1.00048],
                               atol=1e-5)
    d0, t0 = tp_phase(oracle, 2, 0)
    d1, t1 = tp_phase(oracle, 2, 1)
    assert list(t0) == [12] and d0[0] == 1.0
    assert list(t1) == list(range(24))
    np.testing.assert_allclose(d1, d1[::-1], rtol=1e-12)


def test_generated_tap_header_matches_oracle(oracle):
    """tools/gen_tp_coefs.py (what the CUDA sweep multiplies by) restates the
    same prototype as the oracle."""
    import os
    import re
    path = os.path.join(os.path.dirname(__file__), "..", "loudgain_b200", "csrc", "lg_tp_coefs.h")
    text = open(path).read()

    def rows(name):
        body = text[text.index(name):]
        body = body[body.index("{") + 1:body.index("};")]
        return [[float.fromhex(v.rstrip("f")) for v in re.findall(r"-?0x[0-9a-f.]+p[-+]?\d+f?", r)]
                for r in body.strip().split("\n")]

    d4 = rows("kTp4d")
    for p in range(3):
        c, _ = tp_phase(oracle, 4, p + 1)
        np.testing.assert_allclose(d4[p], c, rtol=1e-14)
    f4 = rows("kTp4f")
    for p in range(3):
        c, _ = tp_phase(oracle, 4, p + 1)
        np.testing.assert_array_equal(np.array(f4[p], dtype=np.float32), c.astype(np.float32))
    c, _ = tp_phase(oracle, 2, 1)
    np.testing.assert_allclose(rows("kTp2d")[0], c, rtol=1e-14)
