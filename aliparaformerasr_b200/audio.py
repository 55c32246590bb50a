"""Audio ingestion through libpfasr (include/pf_abi.h "audio ingestion"; csrc/audio.cu): the caller-side step in front of
``AddSamples``.  Mirrors AliParaformerAsr.Examples/Utils/AudioHelper.cs:12-32 (``GetFileSample``) and :223-279
(``Resample``); the sample conversion, down-mix and resampling run on the device."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

_DTYPES = {_lib.PF_AUDIO_U8: np.uint8, _lib.PF_AUDIO_S16: np.dtype("<i2"), _lib.PF_AUDIO_S24: np.uint8,
           _lib.PF_AUDIO_S32: np.dtype("<i4"), _lib.PF_AUDIO_F32: np.dtype("<f4")}


@dataclass
class Audio:
    """Raw interleaved samples of one file.  ``data`` is a contiguous array: uint8 / int16 / int32 / float32 values, or
    the 3-byte little-endian groups of 24-bit audio as uint8."""
    data: np.ndarray
    format: int
    channels: int
    sample_rate: int

    @property
    def n_values(self) -> int:
        return self.data.size // 3 if self.format == _lib.PF_AUDIO_S24 else self.data.size

    def as_pf_audio(self) -> _lib.PfAudio:
        return _lib.PfAudio(self.data.ctypes.data, self.n_values, self.format, self.channels, self.sample_rate, 0)

    def num_samples(self) -> int:
        """Length of the float[] ``GetFileSample`` returns for this audio."""
        a = self.as_pf_audio()
        n = _lib.load().pf_audio_num_samples(C.byref(a))
        if n < 0:
            raise ValueError(_lib.load().pf_last_error().decode(errors="replace"))   # ArgumentException in the C#
        return int(n)

    def to_pcm(self) -> np.ndarray:
        """The converted samples (device round trip through the test hook; product runs use ``Engine.run_audio``)."""
        a = self.as_pf_audio()
        n = C.c_int64(0)
        out = np.zeros(max(1, self.num_samples()), np.float32)
        _lib.check(_lib.load().pf_dbg_audio_convert(C.byref(a), _lib.fptr(out), out.size, C.byref(n)))
        return out[: n.value]


def from_samples(samples: np.ndarray, sample_rate: int, channels: int = 1, format: Optional[int] = None) -> Audio:
    x = np.ascontiguousarray(samples)
    if format is None:
        format = {np.dtype("uint8"): _lib.PF_AUDIO_U8, np.dtype("int16"): _lib.PF_AUDIO_S16, np.dtype("int32"): _lib.PF_AUDIO_S32,
                  np.dtype("float32"): _lib.PF_AUDIO_F32}[x.dtype]
    return Audio(x.reshape(-1), format, channels, sample_rate)


def parse_wav(blob: bytes) -> Audio:
    """``pf_wav_parse``: RIFF/WAVE image -> ``Audio`` viewing the data chunk (the bytes are copied once into numpy)."""
    buf = np.frombuffer(blob, dtype=np.uint8)
    a = _lib.PfAudio()
    _lib.check(_lib.load().pf_wav_parse(buf.ctypes.data, buf.size, C.byref(a)))
    off = a.data - buf.ctypes.data
    width = {_lib.PF_AUDIO_U8: 1, _lib.PF_AUDIO_S16: 2, _lib.PF_AUDIO_S24: 3, _lib.PF_AUDIO_S32: 4, _lib.PF_AUDIO_F32: 4}[a.format]
    raw = buf[off: off + a.n_values * width].copy()
    data = raw if a.format in (_lib.PF_AUDIO_U8, _lib.PF_AUDIO_S24) else raw.view(_DTYPES[a.format])
    return Audio(data, a.format, a.channels, a.sample_rate)


def read_wav(path: str) -> Audio:
    with open(path, "rb") as f:
        return parse_wav(f.read())
