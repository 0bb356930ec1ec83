"""loudgain_b200 -- B200-native loudness analysis behind the libebur128 C ABI.

The product is loudgain_b200/lib/libebur128.so (CUDA, sm_100a), built by
loudgain_b200.build from loudgain_b200/csrc.  This package only loads it and
mirrors the reference scanner's host-side interface:

  capi      ctypes binding of include/ebur128.h (what scan.c calls)
  engine    batch API over HBM-resident PCM (include/ebur128_b200.h): Batch, the
            multi-GPU album exchange and time sharding, and scan_host, the binding of
            the library's scan.c-shaped driver (lgb_scan_host_mt in csrc/lg_scan.cu)
  wavio     WAV / RIFF reader with swr-style narrowing to 16 bit
  synth     synthetic PCM for the BASELINE.json configs

There is no CPU measurement path in this package; loading fails loudly if the
CUDA library has not been built.
"""
from .build import LIB_PATH, build  # noqa: F401


def load_library():
    """The product library through the ebur128_* binding.  Raises if missing."""
    import os

    from .capi import Ebur128Lib

    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built; run `python -m loudgain_b200.build` "
            "(there is no CPU fallback)")
    return Ebur128Lib(LIB_PATH)
