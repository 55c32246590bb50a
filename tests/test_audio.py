"""Audio ingestion, CPU side: the RIFF/WAVE walk and the length rule of libpfasr against the oracle
restatement of AudioHelper.GetFileSample / Resample (oracle/audio.py), plus hand-checked cases of the C# arithmetic.
The device conversion itself is covered in tests/test_gpu_audio.py."""
import ctypes as C
import struct

import numpy as np
import pytest

from aliparaformerasr_b200 import _lib, audio
from oracle import audio as oaudio


def make_wav(payload: bytes, tag: int, bits: int, channels: int, rate: int, extensible: bool = False,
             junk_before: bool = False, overstate: bool = False) -> bytes:
    align = channels * bits // 8
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, rate, rate * align, align, bits, 22, bits, 0, tag) + guid_tail
    else:
        fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * align, align, bits)
    chunks = b""
    if junk_before:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"          # odd-sized chunk: padded to a word
    chunks += b"fmt " + struct.pack("<I", len(fmt)) + fmt
    size = len(payload) + (1000 if overstate else 0)
    chunks += b"data" + struct.pack("<I", size) + payload
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def test_wav_parse_formats():
    rng = np.random.default_rng(0)
    s16 = rng.integers(-32768, 32767, 200, dtype=np.int16)
    a = audio.parse_wav(make_wav(s16.tobytes(), 1, 16, 2, 44100, junk_before=True))
    assert (a.format, a.channels, a.sample_rate, a.n_values) == (_lib.PF_AUDIO_S16, 2, 44100, 200)
    assert np.array_equal(a.data, s16)
    f32 = rng.standard_normal(77).astype(np.float32)
    a = audio.parse_wav(make_wav(f32.tobytes(), 3, 32, 1, 16000, extensible=True))
    assert (a.format, a.channels, a.sample_rate, a.n_values) == (_lib.PF_AUDIO_F32, 1, 16000, 77)
    assert np.array_equal(a.data, f32)
    s24 = rng.integers(0, 256, 3 * 50, dtype=np.uint8)
    a = audio.parse_wav(make_wav(s24.tobytes(), 1, 24, 1, 8000))
    assert (a.format, a.n_values) == (_lib.PF_AUDIO_S24, 50) and np.array_equal(a.data, s24)
    a = audio.parse_wav(make_wav(bytes(range(100)), 1, 8, 1, 8000))
    assert (a.format, a.n_values) == (_lib.PF_AUDIO_U8, 100)
    s32 = rng.integers(-2**31, 2**31 - 1, 40, dtype=np.int32)
    a = audio.parse_wav(make_wav(s32.tobytes(), 1, 32, 2, 48000))
    assert (a.format, a.n_values) == (_lib.PF_AUDIO_S32, 40) and np.array_equal(a.data, s32)
    # a partial trailing frame is dropped; an overstated data length is clamped to the file
    a = audio.parse_wav(make_wav(s16.tobytes()[:-1], 1, 16, 2, 44100))
    assert a.n_values == 198
    a = audio.parse_wav(make_wav(s16.tobytes(), 1, 16, 1, 16000, overstate=True))
    assert a.n_values == 200


def test_wav_parse_rejects():
    lib = _lib.load()
    out = _lib.PfAudio()
    for blob in (b"", b"RIFF\0\0\0\0WAVX", b"OggS" + b"\0" * 40, make_wav(b"\0" * 8, 2, 4, 1, 8000), make_wav(b"\0" * 8, 1, 16, 1, 8000)[:30]):
        buf = C.create_string_buffer(blob, max(1, len(blob)))
        assert lib.pf_wav_parse(buf, len(blob), C.byref(out)) == _lib.PF_ERR_UNSUPPORTED
    assert lib.pf_wav_parse(None, 0, C.byref(out)) == _lib.PF_ERR_BAD_ARG


def test_num_samples_follows_math_round():
    def n(values, rate, ch):
        return audio.from_samples(np.zeros(values, np.int16), rate, ch).num_samples()
    assert n(1000, 16000, 1) == 1000
    assert n(1000, 16000, 2) == 1000            # no down-mix at 16 kHz (AudioHelper.cs:27-30)
    assert n(1000, 8000, 1) == 2000
    assert n(1001, 8000, 2) == 1000             # odd stereo length: the last value is ignored
    assert n(44100, 44100, 1) == 16000
    assert n(3, 32000, 1) == 2                  # 1.5 -> 2 (half to even)
    assert n(5, 32000, 1) == 2                  # 2.5 -> 2 (half to even, not away from zero)
    assert n(0, 8000, 1) == 0
    for values, rate, ch in ((777, 22050, 1), (9999, 48000, 2), (12345, 11025, 1), (31, 96000, 2)):
        assert n(values, rate, ch) == oaudio.resample(np.zeros(values, np.float32), rate, 16000, ch).size
    with pytest.raises(ValueError):
        n(100, 8000, 3)                          # ArgumentException: only 1 or 2 channels
    with pytest.raises(ValueError):
        audio.from_samples(np.zeros(4, np.int16), 0, 1).num_samples()


def test_oracle_resample_hand_cases():
    # 8 kHz -> 16 kHz: ratio 0.5, every second output is the midpoint, the tail repeats the last sample
    x = np.asarray([0.0, 1.0, 3.0], np.float32)
    assert oaudio.resample(x, 8000, 16000).tolist() == [0.0, 0.5, 1.0, 2.0, 3.0, 3.0]
    # stereo: (L + R) * 0.5f first
    st = np.asarray([0.0, 2.0, 1.0, 3.0, 5.0, 7.0], np.float32)
    assert oaudio.resample(st, 8000, 16000, 2).tolist() == [1.0, 1.5, 2.0, 4.0, 6.0, 6.0]
    # 32 kHz -> 16 kHz: plain decimation by two
    x = np.arange(10, dtype=np.float32)
    assert oaudio.resample(x, 32000, 16000).tolist() == [0.0, 2.0, 4.0, 6.0, 8.0]
    # sample-format table
    assert oaudio.to_float(np.asarray([-32768, 0, 16384], np.int16), oaudio.S16).tolist() == [-1.0, 0.0, 0.5]
    assert oaudio.to_float(np.asarray([0, 128, 255], np.uint8), oaudio.U8).tolist() == [-1.0, 0.0, 127 / 128]
    assert oaudio.to_float(np.asarray([0, 0, 0x80, 0, 0, 0x40], np.uint8), oaudio.S24).tolist() == [-1.0, 0.5]
    assert oaudio.to_float(np.asarray([-2**31, 2**30], np.int32), oaudio.S32).tolist() == [-1.0, 0.5]


def test_wav_parse_survives_corrupted_images():
    """Truncations and random byte flips of valid files: the parser may accept or refuse, but what it returns always lies
    inside the image (it reads untrusted files)."""
    lib = _lib.load()
    rng = np.random.default_rng(11)
    base = [make_wav(rng.integers(-3000, 3000, 400, dtype=np.int16).tobytes(), 1, 16, 2, 44100, junk_before=True),
            make_wav(rng.standard_normal(100).astype(np.float32).tobytes(), 3, 32, 1, 8000, extensible=True),
            make_wav(bytes(range(90)), 1, 24, 1, 16000)]
    width = {_lib.PF_AUDIO_U8: 1, _lib.PF_AUDIO_S16: 2, _lib.PF_AUDIO_S24: 3, _lib.PF_AUDIO_S32: 4, _lib.PF_AUDIO_F32: 4}
    for _ in range(3000):
        blob = bytearray(base[int(rng.integers(0, len(base)))])
        if rng.random() < 0.5:
            blob = blob[: int(rng.integers(0, len(blob) + 1))]
        for _ in range(int(rng.integers(0, 6))):
            if blob:
                blob[int(rng.integers(0, len(blob)))] = int(rng.integers(0, 256))
        buf = np.frombuffer(bytes(blob) + b"\0", dtype=np.uint8)[: len(blob)]
        out = _lib.PfAudio()
        st = lib.pf_wav_parse(buf.ctypes.data if len(blob) else None, len(blob), C.byref(out))
        if st == _lib.PF_OK:
            off = out.data - buf.ctypes.data
            assert out.format in width and out.channels >= 1 and out.sample_rate >= 1 and out.n_values >= 0
            assert 0 <= off and off + out.n_values * width[out.format] <= len(blob)
        else:
            assert st in (_lib.PF_ERR_UNSUPPORTED, _lib.PF_ERR_BAD_ARG)
