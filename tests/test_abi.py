"""The drop-in boundary: the CUDA library builds, loads and exports every
symbol include/*.h declares (no compute calls here: CPU only)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ebur128|lgb)_[a-z0-9_]+)\s*\(", text)))


def _exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


@pytest.fixture(scope="module")
def libpath():
    from loudgain_b200 import build
    return build()


def test_exports_every_declared_symbol(libpath):
    have = _exported(libpath)
    for header in ("ebur128.h", "ebur128_b200.h"):
        names = _declared(header)
        assert len(names) >= 10
        missing = [n for n in names if n not in have]
        assert not missing, f"{header}: not exported: {missing}"


def test_exports_nothing_else(libpath):
    """-fvisibility=hidden: only the C ABI leaves the library."""
    extra = [n for n in _exported(libpath) if not n.startswith(("ebur128_", "lgb_"))]
    assert not extra, extra


def test_oracle_exports_the_same_abi():
    from oracle import build_oracle
    have = _exported(build_oracle())
    assert not [n for n in _declared("ebur128.h") if n not in have]


def test_binding_covers_header():
    from loudgain_b200.capi import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == _declared("ebur128.h")


def test_soname_and_arch(libpath):
    dyn = subprocess.check_output(["readelf", "-d", libpath], text=True)
    assert "libebur128.so.1" in dyn        # what the shipped loudgain binary NEEDs
    sass = subprocess.check_output(["cuobjdump", "-lelf", libpath], text=True)
    assert "sm_100a" in sass


def test_loads_and_reports_version(libpath):
    from loudgain_b200 import load_library
    lib = load_library()
    assert lib.version() >= (1, 2, 4)      # loudgain.c:183 warns below 1.2.4


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call the oracle / emulation."""
    pkg = os.path.join(ROOT, "loudgain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ebur128_oracle" not in text and "libemu" not in text, f


def test_no_gpu_means_loud_failure(libpath):
    """Without a CUDA device the library refuses to initialise (no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from loudgain_b200 import load_library
    assert load_library().try_init(2, 44100) is None
