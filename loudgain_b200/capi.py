"""ctypes binding of the ebur128_* C ABI declared in include/ebur128.h.

The same binding class drives any library exporting that ABI: the product
library (loudgain_b200/csrc -> libebur128.so, CUDA inside) and, in tests only,
the CPU oracle.  It mirrors the call sequence of the reference scanner
(/root/reference/src/scan.c:203-207 init, :448-450 add_frames_short,
:294-307 track queries, :383-391 album queries, :102 destroy).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Sequence

import numpy as np

# enum mode (include/ebur128.h)
MODE_M = 1 << 0
MODE_S = (1 << 1) | MODE_M
MODE_I = (1 << 2) | MODE_M
MODE_LRA = (1 << 3) | MODE_S
MODE_SAMPLE_PEAK = (1 << 4) | MODE_M
MODE_TRUE_PEAK = (1 << 5) | MODE_M | MODE_SAMPLE_PEAK
MODE_HISTOGRAM = 1 << 6
#: the mode word loudgain passes (scan.c:205-206)
MODE_LOUDGAIN = MODE_S | MODE_I | MODE_LRA | MODE_SAMPLE_PEAK | MODE_TRUE_PEAK

SUCCESS, ERROR_NOMEM, ERROR_INVALID_MODE, ERROR_INVALID_CHANNEL_INDEX, ERROR_NO_CHANGE = range(5)

# enum channel (subset used by tests)
UNUSED, LEFT, RIGHT, CENTER, LEFT_SURROUND, RIGHT_SURROUND, DUAL_MONO = range(7)
Mp060, Mm060, Mp090, Mm090 = 9, 10, 11, 12        # include/ebur128.h: side positions, weight 1.41


class StateStruct(C.Structure):
    """Public part of ebur128_state; scan.c:300 reads `channels` directly."""
    _fields_ = [("mode", C.c_int), ("channels", C.c_uint),
                ("samplerate", C.c_ulong), ("d", C.c_void_p)]


StateP = C.POINTER(StateStruct)

#: every symbol include/ebur128.h declares
ABI_SYMBOLS = (
    "ebur128_get_version", "ebur128_init", "ebur128_destroy", "ebur128_set_channel",
    "ebur128_change_parameters", "ebur128_set_max_window", "ebur128_set_max_history",
    "ebur128_add_frames_short", "ebur128_add_frames_int", "ebur128_add_frames_float",
    "ebur128_add_frames_double", "ebur128_loudness_global",
    "ebur128_loudness_global_multiple", "ebur128_loudness_momentary",
    "ebur128_loudness_shortterm", "ebur128_loudness_window", "ebur128_loudness_range",
    "ebur128_loudness_range_multiple", "ebur128_sample_peak", "ebur128_prev_sample_peak",
    "ebur128_true_peak", "ebur128_prev_true_peak", "ebur128_relative_threshold",
)


class Ebur128Error(RuntimeError):
    def __init__(self, fn: str, code: int):
        super().__init__(f"{fn} failed with error {code}")
        self.code = code


class Ebur128Lib:
    """A loaded library exporting the ebur128_* ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"ebur128 library not built: {path}")
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        L.ebur128_get_version.argtypes = [C.POINTER(C.c_int)] * 3
        L.ebur128_get_version.restype = None
        L.ebur128_init.argtypes = [C.c_uint, C.c_ulong, C.c_int]
        L.ebur128_init.restype = StateP
        L.ebur128_destroy.argtypes = [C.POINTER(StateP)]
        L.ebur128_destroy.restype = None
        L.ebur128_set_channel.argtypes = [StateP, C.c_uint, C.c_int]
        L.ebur128_change_parameters.argtypes = [StateP, C.c_uint, C.c_ulong]
        L.ebur128_set_max_window.argtypes = [StateP, C.c_ulong]
        L.ebur128_set_max_history.argtypes = [StateP, C.c_ulong]
        for name in ("short", "int", "float", "double"):
            fn = getattr(L, f"ebur128_add_frames_{name}")
            fn.argtypes = [StateP, C.c_void_p, C.c_size_t]
            fn.restype = C.c_int
        for name in ("loudness_global", "loudness_momentary", "loudness_shortterm",
                     "loudness_range", "relative_threshold"):
            fn = getattr(L, f"ebur128_{name}")
            fn.argtypes = [StateP, C.POINTER(C.c_double)]
            fn.restype = C.c_int
        L.ebur128_loudness_window.argtypes = [StateP, C.c_ulong, C.POINTER(C.c_double)]
        for name in ("loudness_global_multiple", "loudness_range_multiple"):
            fn = getattr(L, f"ebur128_{name}")
            fn.argtypes = [C.POINTER(StateP), C.c_size_t, C.POINTER(C.c_double)]
            fn.restype = C.c_int
        for name in ("sample_peak", "prev_sample_peak", "true_peak", "prev_true_peak"):
            fn = getattr(L, f"ebur128_{name}")
            fn.argtypes = [StateP, C.c_uint, C.POINTER(C.c_double)]
            fn.restype = C.c_int

    def version(self) -> tuple[int, int, int]:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.ebur128_get_version(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def init(self, channels: int, samplerate: int, mode: int = MODE_LOUDGAIN) -> "State":
        p = self.lib.ebur128_init(channels, samplerate, mode)
        if not p:
            raise Ebur128Error("ebur128_init", -1)
        return State(self, p)

    def try_init(self, channels: int, samplerate: int, mode: int = MODE_LOUDGAIN):
        p = self.lib.ebur128_init(channels, samplerate, mode)
        return State(self, p) if p else None

    def _multi(self, fn, states: Sequence["State | None"]) -> tuple[int, float]:
        arr = (StateP * len(states))(*[(s.ptr if s is not None else StateP()) for s in states])
        out = C.c_double()
        rc = fn(arr, len(states), C.byref(out))
        return rc, out.value

    def loudness_global_multiple(self, states: Sequence["State | None"]) -> float:
        rc, v = self._multi(self.lib.ebur128_loudness_global_multiple, states)
        if rc != SUCCESS:
            raise Ebur128Error("ebur128_loudness_global_multiple", rc)
        return v

    def loudness_range_multiple(self, states: Sequence["State | None"]) -> float:
        rc, v = self._multi(self.lib.ebur128_loudness_range_multiple, states)
        if rc != SUCCESS:
            raise Ebur128Error("ebur128_loudness_range_multiple", rc)
        return v


_ADDERS = {
    np.dtype(np.int16): "ebur128_add_frames_short",
    np.dtype(np.int32): "ebur128_add_frames_int",
    np.dtype(np.float32): "ebur128_add_frames_float",
    np.dtype(np.float64): "ebur128_add_frames_double",
}


class State:
    """One ebur128_state*, owned until destroy()."""

    def __init__(self, lib: Ebur128Lib, ptr):
        self._lib = lib
        self.ptr = ptr

    # --- struct fields the caller may read (scan.c:300,368)
    @property
    def channels(self) -> int:
        return self.ptr.contents.channels

    @property
    def samplerate(self) -> int:
        return self.ptr.contents.samplerate

    @property
    def mode(self) -> int:
        return self.ptr.contents.mode

    def destroy(self) -> None:
        if self.ptr:
            p = StateP(self.ptr.contents)
            self._lib.lib.ebur128_destroy(C.byref(p))
            assert not p, "ebur128_destroy must store NULL (scan.c:102 contract)"
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def add_frames(self, pcm: np.ndarray, chunk_frames: int | None = None) -> None:
        """Feed interleaved PCM [frames, channels] (or flat), optionally split
        into chunk_frames-sized calls like the reference's per-AVFrame loop."""
        pcm = np.ascontiguousarray(pcm)
        fn_name = _ADDERS[pcm.dtype]
        fn = getattr(self._lib.lib, fn_name)
        ch = self.channels
        flat = pcm.reshape(-1)
        assert flat.size % ch == 0
        frames = flat.size // ch
        step = frames if not chunk_frames else chunk_frames
        base = flat.ctypes.data
        isz = flat.itemsize
        pos = 0
        while pos < frames:
            n = min(step, frames - pos)
            rc = fn(self.ptr, C.c_void_p(base + pos * ch * isz), n)
            if rc != SUCCESS:
                raise Ebur128Error(fn_name, rc)
            pos += n
        if frames == 0:
            rc = fn(self.ptr, C.c_void_p(base), 0)
            if rc != SUCCESS:
                raise Ebur128Error(fn_name, rc)

    def add_frames_calls(self, pcm: np.ndarray, splits: Iterable[int]) -> None:
        """Feed PCM with an explicit list of per-call frame counts."""
        pcm = np.ascontiguousarray(pcm).reshape(-1, self.channels)
        pos = 0
        for n in splits:
            self.add_frames(pcm[pos:pos + n])
            pos += n
        if pos < len(pcm):
            self.add_frames(pcm[pos:])

    def _scalar(self, name: str, *args) -> tuple[int, float]:
        out = C.c_double()
        rc = getattr(self._lib.lib, name)(self.ptr, *args, C.byref(out))
        return rc, out.value

    def _checked(self, name: str, *args) -> float:
        rc, v = self._scalar(name, *args)
        if rc != SUCCESS:
            raise Ebur128Error(name, rc)
        return v

    def loudness_global(self) -> float:
        return self._checked("ebur128_loudness_global")

    def loudness_range(self) -> float:
        return self._checked("ebur128_loudness_range")

    def loudness_momentary(self) -> float:
        return self._checked("ebur128_loudness_momentary")

    def loudness_shortterm(self) -> float:
        return self._checked("ebur128_loudness_shortterm")

    def loudness_window(self, window_ms: int) -> float:
        return self._checked("ebur128_loudness_window", window_ms)

    def relative_threshold(self) -> float:
        return self._checked("ebur128_relative_threshold")

    def sample_peak(self, ch: int) -> float:
        return self._checked("ebur128_sample_peak", ch)

    def true_peak(self, ch: int) -> float:
        return self._checked("ebur128_true_peak", ch)

    def prev_sample_peak(self, ch: int) -> float:
        return self._checked("ebur128_prev_sample_peak", ch)

    def prev_true_peak(self, ch: int) -> float:
        return self._checked("ebur128_prev_true_peak", ch)

    def set_channel(self, ch: int, role: int) -> int:
        return self._lib.lib.ebur128_set_channel(self.ptr, ch, role)

    def change_parameters(self, channels: int, samplerate: int) -> int:
        return self._lib.lib.ebur128_change_parameters(self.ptr, channels, samplerate)

    def set_max_history(self, history_ms: int) -> int:
        return self._lib.lib.ebur128_set_max_history(self.ptr, history_ms)

    def set_max_window(self, window_ms: int) -> int:
        return self._lib.lib.ebur128_set_max_window(self.ptr, window_ms)

    def sample_peaks(self) -> list[float]:
        return [self.sample_peak(c) for c in range(self.channels)]

    def true_peaks(self) -> list[float]:
        return [self.true_peak(c) for c in range(self.channels)]
