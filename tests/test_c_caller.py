"""The drop-in boundary seen from C: (1) what the reference's SHIPPED binary imports
from libebur128 (fixture made from bin/loudgain_0.5.3-1ubuntu1_amd64.deb by
tools/make_deb_imports.py) is exported by the product library under the SONAME the
binary NEEDs; (2) a plain C caller that replays scan.c's call sequence
(tests/c/scan_caller.c) compiles against include/ebur128.h as C99, links against
the product library, and -- on the GPU -- prints the numbers the oracle build of the
same program prints."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "scan_caller.c")
LIBDIR = os.path.join(ROOT, "loudgain_b200", "lib")
DEB = "/root/reference/bin/loudgain_0.5.3-1ubuntu1_amd64.deb"


def _fixture():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "loudgain_deb_imports.json")))


def _exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_shipped_binary_imports_are_exported(product_path):
    fx = _fixture()
    assert fx["undefined_ebur128_symbols"], "empty fixture"
    missing = set(fx["undefined_ebur128_symbols"]) - _exported(product_path)
    assert not missing, f"the shipped loudgain binary would not resolve: {sorted(missing)}"
    dyn = subprocess.check_output(["readelf", "-d", product_path], text=True)
    soname = [line.split("[")[1].split("]")[0] for line in dyn.splitlines() if "(SONAME)" in line]
    assert soname and soname[0] in fx["needed"], (soname, fx["needed"])


@pytest.mark.skipif(not os.path.exists(DEB) or not shutil.which("ar"), reason="reference tree not present")
def test_fixture_matches_the_reference_deb():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_deb_imports import deb_imports
    got, fx = deb_imports(), _fixture()
    assert got["undefined_ebur128_symbols"] == fx["undefined_ebur128_symbols"]
    assert got["needed"] == fx["needed"]


def _build(tmp_path, name, libdir, lib):
    """Links the caller against `lib`; the loader will look for the library's SONAME, so
    an install-style link of that name is put next to the program (rpath)."""
    exe = str(tmp_path / name)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1",
                           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, f"-l:{lib}", f"-Wl,-rpath,{tmp_path}", "-lm"])
    dyn = subprocess.check_output(["readelf", "-d", os.path.join(libdir, lib)], text=True)
    soname = [line.split("[")[1].split("]")[0] for line in dyn.splitlines() if "(SONAME)" in line]
    link = tmp_path / (soname[0] if soname else lib)
    if not link.exists():
        os.symlink(os.path.join(libdir, lib), link)
    return exe


def _oracle_exe(tmp_path):
    from oracle import build_oracle
    path = build_oracle()
    return _build(tmp_path, "scan_caller_oracle", os.path.dirname(path), os.path.basename(path))


def _rows(text):
    return [[float(v) for v in line.split()] for line in text.strip().splitlines()]


def test_c_caller_compiles_links_and_runs_on_the_oracle(tmp_path, product_path):
    exe = _build(tmp_path, "scan_caller_product", LIBDIR, os.path.basename(product_path))
    needed = subprocess.check_output(["readelf", "-d", exe], text=True)
    assert "libebur128.so.1" in needed            # bound by SONAME, like the shipped binary
    rows = _rows(subprocess.check_output([_oracle_exe(tmp_path), "3", "6"], text=True))
    assert len(rows) == 3 and all(len(r) == 6 for r in rows)
    assert all(-40.0 < r[1] < 0.0 and r[2] > 1.0 and 0.1 < r[3] < 1.2 for r in rows)
    assert len({r[4] for r in rows}) == 1         # the album is the same whichever track asks


@pytest.mark.gpu
def test_c_caller_on_the_gpu_matches_the_oracle_build(tmp_path, product_path):
    prod = _build(tmp_path, "scan_caller_product", LIBDIR, os.path.basename(product_path))
    got = _rows(subprocess.check_output([prod, "4", "9"], text=True))
    want = _rows(subprocess.check_output([_oracle_exe(tmp_path), "4", "9"], text=True))
    assert len(got) == len(want) == 4
    for g, w in zip(got, want):
        assert g[0] == w[0]
        for k in (1, 2, 4, 5):
            assert abs(g[k] - w[k]) <= 2e-4, (g, w)          # LU
        assert abs(g[3] - w[3]) <= 1e-6 * w[3], (g, w)        # true peak, relative
