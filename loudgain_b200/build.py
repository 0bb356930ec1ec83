"""Builds libebur128.so (CUDA, sm_100a) in-tree with nvcc.

The library is the product: it exports the ebur128_* drop-in ABI
(include/ebur128.h) and the lgb_* batch extension (include/ebur128_b200.h).
nvcc cross-compiles without a GPU, so this runs anywhere the toolkit is.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# LG_LIB_SUFFIX / LG_NVCC_EXTRA build tuning variants next to the product library
_SUFFIX = os.environ.get("LG_LIB_SUFFIX", "")
LIB_PATH = os.path.join(LIB_DIR, f"libebur128{'_' + _SUFFIX if _SUFFIX else ''}.so")
SOURCES = ("lg_kernels.cu", "lg_pair.cu", "lg_run.cu", "lg_batch.cu", "lg_ebur128.cu", "lg_scan.cu")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--cudart", "shared", "-Wno-deprecated-gpu-targets",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("ebur128.h", "ebur128_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every CUDA source for sm_100a and links libebur128.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", f"{_SUFFIX}.o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("LG_NVCC_EXTRA", "").split(), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, "-shared", "--cudart", "shared", "-Wno-deprecated-gpu-targets", "-o", LIB_PATH, *objs,
           "-Xlinker", "-soname=libebur128.so.1", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
