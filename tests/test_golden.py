"""Golden vectors (tests/golden/r128_vectors.json, written by
tools/make_golden.py from the CPU oracle on small seeded inputs): the oracle
still reproduces them (CPU), and the product matches them on the B200 through
the drop-in C ABI.  They are NOT reference outputs -- libebur128 is not
available here -- they freeze what parity is measured
against."""
import json
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import cases  # noqa: E402

from tests.helpers import TOL_LU, TOL_TP_REL, lu_diff, rel_diff  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "r128_vectors.json")) as f:
    GOLDEN = json.load(f)["vectors"]


@pytest.fixture(scope="module")
def inputs():
    c = cases()
    assert set(c) == set(GOLDEN)
    for name, (pcm, rate) in c.items():
        g = GOLDEN[name]
        assert (rate, pcm.shape[1], pcm.shape[0]) == (g["rate"], g["channels"], g["frames"])
        assert zlib.crc32(np.ascontiguousarray(pcm).tobytes()) == g["pcm_crc32"], \
            f"{name}: the synthesiser no longer produces the PCM the vector was made from"
    return c


def _measure(lib, pcm, rate):
    st = lib.init(pcm.shape[1], rate)
    st.add_frames(pcm, 1024)
    out = (st.loudness_global(), st.loudness_range(), np.array(st.sample_peaks()),
           np.array(st.true_peaks()))
    st.destroy()
    return out


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_reproduces_golden(oracle, inputs, name):
    g = GOLDEN[name]
    loud, rng, sp, tp = _measure(oracle, *inputs[name])
    want = -np.inf if g["loudness"] is None else g["loudness"]
    assert lu_diff(loud, want) < 1e-9 and abs(rng - g["range"]) < 1e-9
    np.testing.assert_array_equal(sp, g["sample_peak"])
    assert rel_diff(tp, g["true_peak"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_product_matches_golden(product, inputs, name):
    g = GOLDEN[name]
    loud, rng, sp, tp = _measure(product, *inputs[name])
    want = -np.inf if g["loudness"] is None else g["loudness"]
    assert lu_diff(loud, want) <= TOL_LU and abs(rng - g["range"]) <= TOL_LU     # 0.01 LU
    if g["loudness"] is not None:
        assert "%.2f" % (-18.0 - loud) == "%.2f" % (-18.0 - want) or \
            abs(((-18.0 - want) * 100.0) % 1.0 - 0.5) < 0.05                       # tag precision
    np.testing.assert_array_equal(sp, g["sample_peak"])                           # bit-exact
    assert rel_diff(tp, g["true_peak"]) <= TOL_TP_REL                             # 1e-6
