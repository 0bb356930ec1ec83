"""WAV reader of the scan driver (loudgain_b200/wavio.py): formats and the
narrowing to S16 that scan.c:414-450 applies through swr_convert."""
import struct
import wave

import numpy as np
import pytest

from loudgain_b200.wavio import WavError, read_wav


def _riff(fmt_chunk: bytes, data: bytes, extra: bytes = b"") -> bytes:
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt_chunk)) + fmt_chunk + extra + \
           b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b"")
    return b"RIFF" + struct.pack("<I", len(body)) + body


def _fmt(tag, ch, rate, bits, extensible_sub=None):
    align = ch * ((bits + 7) // 8)
    base = struct.pack("<HHIIHH", 0xFFFE if extensible_sub else tag, ch, rate, rate * align, align, bits)
    if extensible_sub:
        guid = struct.pack("<H", extensible_sub) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
        base += struct.pack("<HHI", 22, bits, 0x3F) + guid
    return base


def test_pcm16_matches_wave_module(tmp_path):
    rng = np.random.default_rng(1)
    pcm = rng.integers(-32768, 32768, size=(1000, 2), dtype=np.int16)
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    got, rate = read_wav(str(p))
    assert rate == 44100
    np.testing.assert_array_equal(got, pcm)


def test_pcm24_extensible_six_channels(tmp_path):
    rng = np.random.default_rng(2)
    v = rng.integers(-(1 << 23), 1 << 23, size=(500, 6)).astype(np.int32)
    raw = np.stack([(v & 0xFF), (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=-1).astype(np.uint8).tobytes()
    p = tmp_path / "b.wav"
    # an unknown chunk with odd size in front of the data: must be skipped, word aligned
    p.write_bytes(_riff(_fmt(1, 6, 96000, 24, extensible_sub=1), raw, extra=b"LIST" + struct.pack("<I", 3) + b"abc\x00"))
    got, rate = read_wav(str(p))
    assert rate == 96000 and got.shape == (500, 6)
    np.testing.assert_array_equal(got, (v >> 8).astype(np.int16))      # S32 >> 16 of (x << 8)


def test_pcm32_pcm8_and_float(tmp_path):
    v32 = np.array([[-2147483648, 2147483647], [65535, -65536], [1 << 16, -1]], dtype="<i4")
    p = tmp_path / "c.wav"
    p.write_bytes(_riff(_fmt(1, 2, 48000, 32), v32.tobytes()))
    np.testing.assert_array_equal(read_wav(str(p))[0], np.array([[-32768, 32767], [0, -1], [1, -1]], dtype=np.int16))
    v8 = np.array([[0], [128], [255]], dtype=np.uint8)
    p = tmp_path / "d.wav"
    p.write_bytes(_riff(_fmt(1, 1, 8000, 8), v8.tobytes()))
    np.testing.assert_array_equal(read_wav(str(p))[0], np.array([[-32768], [0], [127 << 8]], dtype=np.int16))
    vf = np.array([[1.0, -1.0], [0.5, 0.25 / 32768], [2.0, -3.0]], dtype="<f4")
    p = tmp_path / "e.wav"
    p.write_bytes(_riff(_fmt(3, 2, 44100, 32), vf.tobytes()))
    np.testing.assert_array_equal(read_wav(str(p))[0], np.array([[32767, -32768], [16384, 0], [32767, -32768]], dtype=np.int16))


def test_rejects_garbage(tmp_path):
    p = tmp_path / "f.wav"
    p.write_bytes(b"not a wav file at all")
    with pytest.raises(WavError):
        read_wav(str(p))
    p.write_bytes(_riff(_fmt(0x55, 2, 44100, 16), b"\x00" * 8))          # MP3-in-WAV
    with pytest.raises(WavError):
        read_wav(str(p))
