"""Loader for the CPU oracle (oracle/ebur128_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from the
loudgain_b200 package.  PARITY UNPINNED: see the header of ebur128_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libebur128_oracle.so")
_SRC = os.path.join(_HERE, "ebur128_oracle.c")


def build_oracle(force: bool = False) -> str:
    """Compile the oracle with gcc if missing or older than its source."""
    stale = (not os.path.exists(ORACLE_SO)
             or os.path.getmtime(ORACLE_SO) < os.path.getmtime(_SRC))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libebur128_oracle.so"])
    return ORACLE_SO


def load_oracle():
    """Returns the oracle wrapped in the same ctypes binding the product uses."""
    from loudgain_b200.capi import Ebur128Lib, StateP

    lib = Ebur128Lib(build_oracle())
    L = lib.lib
    L.oracle_kfilter_coeffs.argtypes = [C.c_ulong, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.oracle_kfilter_coeffs.restype = None
    L.oracle_tp_phase.argtypes = [C.c_uint, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint)]
    L.oracle_tp_phase.restype = C.c_uint
    L.oracle_block_count.argtypes = [StateP, C.c_int]
    L.oracle_block_count.restype = C.c_size_t
    L.oracle_copy_blocks.argtypes = [StateP, C.c_int, C.POINTER(C.c_double), C.c_size_t]
    L.oracle_copy_blocks.restype = C.c_size_t
    return lib


def kfilter_coeffs(lib, rate: int) -> tuple[np.ndarray, np.ndarray]:
    b = (C.c_double * 5)()
    a = (C.c_double * 5)()
    lib.lib.oracle_kfilter_coeffs(rate, b, a)
    return np.array(b[:]), np.array(a[:])


def tp_phase(lib, factor: int, phase: int) -> tuple[np.ndarray, np.ndarray]:
    coef = (C.c_double * 49)()
    slot = (C.c_uint * 49)()
    n = lib.lib.oracle_tp_phase(factor, phase, coef, slot)
    return np.array(coef[:n]), np.array(slot[:n], dtype=np.int64)


def blocks(lib, state, kind: int) -> np.ndarray:
    """Stored block energies of an oracle state (0 = 400 ms, 1 = 3 s)."""
    n = lib.lib.oracle_block_count(state.ptr, kind)
    out = np.empty(n, dtype=np.float64)
    if n:
        lib.lib.oracle_copy_blocks(state.ptr, kind, out.ctypes.data_as(C.POINTER(C.c_double)), n)
    return out
