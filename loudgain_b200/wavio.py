"""WAV/RIFF reader for the scan driver: PCM files can
be scanned without FFmpeg.  What the reference feeds libebur128 is ALWAYS
interleaved S16 (/root/reference/src/scan.c:414-450: every decoded frame goes
through swr_convert to AV_SAMPLE_FMT_S16, without dither), so the reader
returns int16 and narrows wider formats the way that conversion does:

  8-bit unsigned  -> (x - 128) << 8
  24-bit          -> x >> 8          (decoded as x << 8 in S32, then S32 >> 16)
  32-bit          -> x >> 16
  32-bit float    -> clip(round_half_even(x * 32768)) to int16

Supports WAVE_FORMAT_PCM (1), IEEE_FLOAT (3) and WAVE_FORMAT_EXTENSIBLE
(0xFFFE, sub-format taken from the GUID).  Host code; numpy only.
"""
from __future__ import annotations

import struct

import numpy as np

WAVE_FORMAT_PCM = 1
WAVE_FORMAT_IEEE_FLOAT = 3
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class WavError(ValueError):
    pass


def _chunks(buf: bytes):
    """(id, payload offset, size) of the RIFF sub-chunks."""
    if len(buf) < 12 or buf[0:4] not in (b"RIFF", b"RF64") or buf[8:12] != b"WAVE":
        raise WavError("not a RIFF/WAVE file")
    pos = 12
    while pos + 8 <= len(buf):
        cid = buf[pos:pos + 4]
        size = struct.unpack_from("<I", buf, pos + 4)[0]
        yield cid, pos + 8, size
        pos += 8 + size + (size & 1)          # chunks are word aligned


def to_s16(raw: np.ndarray, fmt: int, bits: int) -> np.ndarray:
    """Narrows decoded samples to int16 like swr_convert(..., AV_SAMPLE_FMT_S16)."""
    if fmt == WAVE_FORMAT_IEEE_FLOAT:
        x = np.rint(raw.astype(np.float32) * np.float32(32768.0))
        return np.clip(x, -32768, 32767).astype(np.int16)
    if bits == 8:
        return ((raw.astype(np.int16) - 128) << 8).astype(np.int16)
    if bits == 16:
        return raw.astype(np.int16, copy=False)
    if bits == 24:
        return (raw >> 8).astype(np.int16)
    if bits == 32:
        return (raw >> 16).astype(np.int16)
    raise WavError(f"unsupported sample size: {bits} bits")


def read_wav(path: str):
    """Returns (int16 ndarray [frames, channels], sample rate)."""
    with open(path, "rb") as f:
        buf = f.read()
    fmt = channels = rate = bits = align = None
    data = None
    for cid, off, size in _chunks(buf):
        if cid == b"fmt ":
            if size < 16:
                raise WavError("short fmt chunk")
            fmt, channels, rate, _, align, bits = struct.unpack_from("<HHIIHH", buf, off)
            if fmt == WAVE_FORMAT_EXTENSIBLE:
                if size < 40:
                    raise WavError("short WAVE_FORMAT_EXTENSIBLE chunk")
                fmt = struct.unpack_from("<H", buf, off + 24)[0]      # first field of the sub-format GUID
        elif cid == b"data":
            end = min(off + size, len(buf)) if size != 0xFFFFFFFF else len(buf)
            data = memoryview(buf)[off:end]
    if fmt is None or data is None:
        raise WavError("missing fmt or data chunk")
    if fmt not in (WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT) or not channels or not rate:
        raise WavError(f"unsupported WAV format tag {fmt}")
    width = (bits + 7) // 8
    if align != width * channels:
        raise WavError("inconsistent block alignment")
    frames = len(data) // align
    data = data[:frames * align]
    if fmt == WAVE_FORMAT_IEEE_FLOAT:
        if bits != 32:
            raise WavError("only 32-bit float WAV is supported")
        raw = np.frombuffer(data, dtype="<f4")
    elif width == 1:
        raw = np.frombuffer(data, dtype=np.uint8)
    elif width == 2:
        raw = np.frombuffer(data, dtype="<i2")
    elif width == 3:
        b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        raw = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        raw = (raw ^ 0x800000) - 0x800000                                # sign-extend 24 bits
    elif width == 4:
        raw = np.frombuffer(data, dtype="<i4")
    else:
        raise WavError(f"unsupported sample size: {bits} bits")
    pcm = to_s16(raw, fmt, bits).reshape(frames, channels)
    return np.ascontiguousarray(pcm), int(rate)
