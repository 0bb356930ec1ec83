"""loudgain's use of a scan result: clipping prevention
(/root/reference/src/loudgain.c:323-379) and the `-O` row
(loudgain.c:586-612).  Host arithmetic of the product library; no GPU."""
import ctypes as C
import math

import pytest

from loudgain_b200 import load_library


class ScanResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "track_gain", "track_peak", "track_loudness", "track_loudness_range", "album_gain",
        "album_peak", "album_loudness", "album_loudness_range", "loudness_reference")]


class ClipInfo(C.Structure):
    _fields_ = [("will_clip", C.c_int), ("track_clipped", C.c_int), ("album_clipped", C.c_int),
                ("album_would_clip", C.c_int), ("track_new_peak", C.c_double),
                ("album_new_peak", C.c_double)]


@pytest.fixture(scope="module")
def L():
    lib = load_library().lib
    lib.lgb_clip_prevention.argtypes = [C.POINTER(ScanResult), C.c_int, C.c_int, C.c_double,
                                        C.POINTER(ClipInfo)]
    lib.lgb_format_tab_row.argtypes = [C.c_char_p, C.POINTER(ScanResult), C.POINTER(ClipInfo), C.c_int,
                                       C.c_char_p, C.c_char_p, C.c_size_t]
    lib.lgb_format_tab_row.restype = C.c_size_t
    lib.lgb_format_tags.argtypes = [C.POINTER(ScanResult), C.c_int, C.c_int, C.c_int, C.c_char_p,
                                    C.c_char_p, C.c_size_t]
    lib.lgb_format_tags.restype = C.c_size_t
    return lib


def _result(tl, tp, al, ap):
    # scan.c:64: gain = -18 - loudness (reference -18 LUFS)
    return ScanResult(-18.0 - tl, tp, tl, 7.5, -18.0 - al, ap, al, 9.25, -18.0)


def test_no_clipping_needed(L):
    r, c = _result(-9.0, 0.98, -10.0, 1.02), ClipInfo()
    assert L.lgb_clip_prevention(r, 1, 1, -1.0, c) == 0
    assert (c.will_clip, c.track_clipped, c.album_clipped) == (0, 0, 0)
    assert r.track_gain == -9.0 and r.album_gain == -8.0
    assert c.track_new_peak == pytest.approx(10 ** (-9 / 20) * 0.98)


def test_quiet_track_is_limited(L):
    # a quiet track with a full-scale peak: +5 dB would push the peak to 1.78
    r, c = _result(-23.0, 1.0, -20.0, 1.0), ClipInfo()
    L.lgb_clip_prevention(r, 1, 0, -1.0, c)
    assert c.will_clip == 1 and c.track_clipped == 0 and c.album_would_clip == 1
    assert r.track_gain == 5.0
    L.lgb_clip_prevention(r, 1, 1, -1.0, c)
    limit = 10 ** (-1 / 20)
    assert (c.will_clip, c.track_clipped, c.album_clipped, c.album_would_clip) == (0, 1, 1, 0)
    assert c.track_new_peak == pytest.approx(limit) and c.album_new_peak == pytest.approx(limit)
    # the corrected gains put the peaks exactly on the limit
    assert 10 ** (r.track_gain / 20) * r.track_peak == pytest.approx(limit, rel=1e-12)
    assert 10 ** (r.album_gain / 20) * r.album_peak == pytest.approx(limit, rel=1e-12)
    assert r.track_gain == pytest.approx(-1.0) and r.album_gain == pytest.approx(-1.0)


def test_track_only_mode_ignores_album(L):
    r, c = _result(-23.0, 0.5, -30.0, 1.0), ClipInfo()
    L.lgb_clip_prevention(r, 0, 1, -1.0, c)
    assert (c.will_clip, c.track_clipped, c.album_clipped) == (0, 0, 0) and r.album_gain == 12.0


def test_tab_row_format(L):
    r, c = _result(-23.0, 1.0, -20.0, 1.0), ClipInfo()
    L.lgb_clip_prevention(r, 1, 1, -1.0, c)
    buf = C.create_string_buffer(512)
    n = L.lgb_format_tab_row(b"a.flac", r, c, 0, b"dB", buf, 512)
    row = buf.value.decode()
    assert n == len(row)
    assert row == ("a.flac\t-23.00 LUFS\t7.50 dB\t1.000000\t0.00 dBTP\t-18.00 LUFS\tN\tY\t-1.00 dB\t"
                   "%.6f\t-1.00 dBTP\n" % 10 ** (-1 / 20))
    L.lgb_format_tab_row(b"Album", r, c, 1, b"LU", buf, 512)
    cols = buf.value.decode().rstrip("\n").split("\t")
    assert cols[0] == "Album" and cols[1] == "-20.00 LUFS" and cols[2] == "9.25 LU"
    assert cols[6:9] == ["N", "Y", "-1.00 LU"] and len(cols) == 11
    assert math.isclose(float(cols[9]), 10 ** (-1 / 20), abs_tol=1e-6)


def test_tag_values_at_reference_precision(L):
    """tag.cc:178-203 (ReplayGain text tags) and tag.cc:442-445 (Opus Q7.8)."""
    r = _result(-9.456, 0.9876543, -10.004, 1.0234567)
    buf = C.create_string_buffer(512)
    n = L.lgb_format_tags(r, 1, 1, 0, b"dB", buf, 512)
    text = buf.value.decode()
    assert n == len(text)
    assert text.splitlines() == [
        "REPLAYGAIN_TRACK_GAIN=-8.54 dB", "REPLAYGAIN_TRACK_PEAK=0.987654",
        "REPLAYGAIN_ALBUM_GAIN=-8.00 dB", "REPLAYGAIN_ALBUM_PEAK=1.023457",
        "REPLAYGAIN_REFERENCE_LOUDNESS=-18.00 LUFS", "REPLAYGAIN_TRACK_RANGE=7.50 dB",
        "REPLAYGAIN_ALBUM_RANGE=9.25 dB"]
    L.lgb_format_tags(r, 0, 0, 0, b"LU", buf, 512)
    assert buf.value.decode().splitlines() == ["REPLAYGAIN_TRACK_GAIN=-8.54 LU",
                                               "REPLAYGAIN_TRACK_PEAK=0.987654"]
    L.lgb_format_tags(r, 1, 0, 1, b"dB", buf, 512)
    assert buf.value.decode().splitlines() == ["R128_TRACK_GAIN=%d" % round(-8.544 * 256),
                                               "R128_ALBUM_GAIN=%d" % round(-7.996 * 256)]
    assert L.lgb_format_tags(r, 1, 1, 0, b"dB", None, 0) == n      # length query


# ---- the reference's own numbers: docs/images/test-{1,2,3}.csv.png -------------------------

def _golden_rows():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loudgain_O_rows.json")
    return json.load(open(path))


def test_reference_screenshot_tables(L):
    """The three `-O` tables of the reference's README (transcribed to
    tests/golden/loudgain_O_rows.json).  The tables print loudness with two
    decimals only, so the unrounded loudness of a row is recovered from test-1
    (no clipping prevention: New_Peak = 10^(gain/20) * True_Peak with gain = -18 - L,
    scan.c:64,317) and must round back to the printed value; lgb_clip_prevention +
    lgb_format_tab_row then have to reproduce every column of all three tables
    (loudgain.c:323-379,586-612): text columns exactly, the 6-decimal peaks to the
    last digit the 6-digit inputs determine."""
    g = _golden_rows()
    n = len(g["names"]) - 1                       # tracks; the last row is the album
    t1 = g["tables"]["test-1"]
    loud = [-18.0 - 20.0 * math.log10(t1["new_peak"][i] / g["true_peak"][i]) for i in range(n + 1)]
    for i in range(n + 1):
        assert "%.2f" % loud[i] == "%.2f" % g["loudness"][i]
    buf = C.create_string_buffer(1024)
    for name, t in g["tables"].items():
        pre = t["pre_gain"]
        for i in range(n):
            # scan_get_track_result / scan_set_album_result (scan.c:317,327,399-403)
            r = ScanResult(-18.0 - loud[i] + pre, g["true_peak"][i], loud[i], g["range"][i],
                           -18.0 - loud[n] + pre, g["true_peak"][n], loud[n], g["range"][n], -18.0 + pre)
            c = ClipInfo()
            assert L.lgb_clip_prevention(r, 1, 1 if t["prevent"] else 0, g["max_true_peak_db"], c) == 0
            rows = []
            L.lgb_format_tab_row(g["names"][i].encode(), r, c, 0, b"dB", buf, 1024)
            rows.append((i, buf.value.decode()))
            if i == n - 1:                        # loudgain prints the album after the last track
                L.lgb_format_tab_row(b"Album", r, c, 1, b"dB", buf, 1024)
                rows.append((n, buf.value.decode()))
            for k, row in rows:
                cols = row.rstrip("\n").split("\t")
                assert len(cols) == len(g["columns"])
                want = [g["names"][k], "%.2f LUFS" % g["loudness"][k], "%.2f dB" % g["range"][k],
                        "%.6f" % g["true_peak"][k], "%.2f dBTP" % g["true_peak_db"][k],
                        "%.2f LUFS" % t["reference"], t["will_clip"][k], t["clip_prevent"][k],
                        "%.2f dB" % t["gain"][k], None, "%.2f dBTP" % t["new_peak_db"][k]]
                for col, (got, exp) in enumerate(zip(cols, want)):
                    if exp is not None:
                        assert got == exp, (name, k, g["columns"][col], got, exp)
                assert abs(float(cols[9]) - t["new_peak"][k]) <= 2.5e-6, (name, k, cols[9])


def test_old_style_and_human_output(L):
    """loudgain -o (loudgain.c:566-585) and the default human-readable blocks
    (loudgain.c:613-649), byte for byte."""
    L.lgb_format_old_row.argtypes = [C.c_char_p, C.POINTER(ScanResult), C.c_int, C.c_char_p, C.c_size_t]
    L.lgb_format_old_row.restype = C.c_size_t
    L.lgb_format_human.argtypes = [C.c_char_p, C.POINTER(ScanResult), C.POINTER(ClipInfo), C.c_int, C.c_int,
                                   C.c_char_p, C.c_char_p, C.c_size_t]
    L.lgb_format_human.restype = C.c_size_t
    r, c = _result(-23.0, 0.98, -22.0, 1.02), ClipInfo()
    assert L.lgb_clip_prevention(r, 1, 1, -1.0, c) == 0
    buf = C.create_string_buffer(1024)
    n = L.lgb_format_old_row(b"a.flac", r, 0, buf, 1024)
    assert buf.value.decode() == "a.flac\t0\t%.2f\t%.6f\t0\t0\n" % (r.track_gain, 0.98 * 32768.0) and n == len(buf.value)
    L.lgb_format_old_row(b"Album", r, 1, buf, 1024)
    assert buf.value.decode() == "Album\t0\t%.2f\t%.6f\t0\t0\n" % (r.album_gain, 1.02 * 32768.0)
    n = L.lgb_format_human(b"a.flac", r, c, 0, 0, b"dB", buf, 1024)
    note = " (corrected to prevent clipping)" if c.track_clipped else ""
    want = ("\nTrack: a.flac\n Loudness: %8.2f LUFS\n Range:    %8.2f dB\n Peak:     %8.6f (%.2f dBTP)\n"
            " Gain:     %8.2f dB%s\n") % (-23.0, 7.5, 0.98, 20.0 * math.log10(0.98), r.track_gain, note)
    assert buf.value.decode() == want and n == len(want)
    assert c.track_clipped == 1 and c.album_clipped == 1   # +5 dB on a full-scale peak needs the correction
    L.lgb_format_human(b"", r, c, 1, 1, b"LU", buf, 1024)
    note = " (corrected to prevent clipping)" if c.album_clipped else ""
    want = ("\nAlbum:\n Loudness: %8.2f LUFS\n Range:    %8.2f LU\n Peak:     %8.6f (%.2f dBTP)\n"
            " Gain:     %8.2f LU (%d)%s\n") % (-22.0, 9.25, 1.02, 20.0 * math.log10(1.02), r.album_gain,
                                              round(r.album_gain * 256.0), note)
    assert buf.value.decode() == want
    assert L.lgb_format_human(b"x", r, c, 0, 0, b"dB", None, 0) > 0          # length query
