"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the caller-side audio step.

* ``to_float``: the PCM sample providers AudioFileReader chains in the un-vendored NuGet **NAudio 2.2.1**
  (AliParaformerAsr.Examples.csproj:21): Pcm8BitToSampleProvider ``b / 128f - 1.0f``, Pcm16 ``s / 32768f``, Pcm24
  ``s / 8388608f``, Pcm32 ``s / (Int32.MaxValue + 1f)``, IEEE float unchanged.  Restated from the published NAudio
  sources; parity unpinned for this table (no NAudio binary offline), the arithmetic is exact in every case but the
  int32 -> float rounding, which IEEE fixes.
* ``resample`` / ``get_file_sample``: AliParaformerAsr.Examples/Utils/AudioHelper.cs:223-279 and :12-32.
"""
from __future__ import annotations

import numpy as np

U8, S16, S24, S32, F32 = range(5)


def to_float(data: np.ndarray, fmt: int) -> np.ndarray:
    f32 = np.float32
    if fmt == U8:
        return (data.astype(f32) / f32(128.0) - f32(1.0)).astype(f32)
    if fmt == S16:
        return (data.astype(f32) / f32(32768.0)).astype(f32)
    if fmt == S24:
        b = data.reshape(-1, 3).astype(np.int32)
        v = (b[:, 2].astype(np.int8).astype(np.int32) << 16) | (b[:, 1] << 8) | b[:, 0]
        return (v.astype(f32) / f32(8388608.0)).astype(f32)
    if fmt == S32:
        return (data.astype(f32) / f32(2147483648.0)).astype(f32)
    return data.astype(f32)


def resample(source: np.ndarray, source_rate: int, target_rate: int, source_channels: int = 1) -> np.ndarray:
    """AudioHelper.Resample (AudioHelper.cs:223-279)."""
    if source_rate <= 0 or target_rate <= 0:
        raise ValueError("sample rates must be positive")
    if source_channels not in (1, 2):
        raise ValueError("only mono or stereo input")
    source = np.asarray(source, np.float32)
    if source.size == 0:
        return np.zeros(0, np.float32)
    mono = source
    if source_channels == 2:
        n = source.size // 2
        mono = ((source[0:2 * n:2] + source[1:2 * n:2]) * np.float32(0.5)).astype(np.float32)
    ratio = float(source_rate) / float(target_rate)
    target_len = int(np.rint(mono.size / ratio))                 # Math.Round: half to even
    i = np.arange(target_len, dtype=np.float64)
    pos = i * ratio
    idx = pos.astype(np.int64)
    frac = pos - idx
    edge = idx >= mono.size - 1
    i0 = np.minimum(idx, mono.size - 1)
    i1 = np.minimum(idx + 1, mono.size - 1)
    val = ((1.0 - frac) * mono[i0].astype(np.float64) + frac * mono[i1].astype(np.float64)).astype(np.float32)
    val[edge] = mono[-1]
    return val


def get_file_sample(data: np.ndarray, fmt: int, channels: int, sample_rate: int) -> np.ndarray:
    """AudioHelper.GetFileSample (AudioHelper.cs:12-32) from the decoded data chunk on: float conversion, and only for a
    rate other than 16 kHz the down-mix + resample (a 16 kHz stereo file stays interleaved)."""
    x = to_float(data, fmt)
    if sample_rate != 16000:
        x = resample(x, sample_rate, 16000, source_channels=channels)
    return x
