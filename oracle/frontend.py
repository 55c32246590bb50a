"""CPU restatement of the reference's offline feature front-end (test infrastructure).

Follows, line by line:
  * ``WavFrontend.GetFbank``   /root/reference/AliParaformerAsr/WavFrontend.cs:31-37
      (x32768 scaling, then ``SpeechFeatures.OnlineFbank.GetFbank``)
  * ``WavFrontend.ApplyLfr``   WavFrontend.cs:73-111  (incl. quirks Q1, Q2 of SURVEY.md)
  * ``WavFrontend.ApplyCmvn``  WavFrontend.cs:53-71
  * ``WavFrontend.LoadCmvn``   WavFrontend.cs:112-153 (Kaldi-nnet ``am.mvn`` text parser)
  * ``PadHelper.PadSequence``  Utils/PadHelper.cs:23-65 (quirk Q4)

``OnlineFbank`` lives in the un-vendored NuGet package ``ManySpeech.SpeechFeatures 1.1.7``
(``AliParaformerAsr.csproj:49``).  It implements Kaldi ``compute-fbank-feats`` semantics
(kaldi-native-fbank); :func:`kaldi_fbank` restates that published algorithm in float32 and is
pinned against ``torchaudio.compliance.kaldi.fbank`` vectors in ``tests/golden``.

All arithmetic is float32 unless stated.  dither is not modelled (every parity config sets
``dither: 0.0`` because dither makes the reference non-deterministic).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# Constant used by PadHelper.PadSequence (Utils/PadHelper.cs:63): -23.025850929940457F * 32768
PAD_QUIRK_VALUE = np.float32(np.float32(-23.025850929940457) * np.float32(32768.0))

FLT_EPSILON = np.float32(1.1920928955078125e-07)


def mel_scale(freq):
    return 1127.0 * np.log(1.0 + np.asarray(freq, dtype=np.float64) / 700.0)


def mel_banks(num_bins: int = 80, sample_rate: int = 16000, n_fft: int = 512,
              low_freq: float = 20.0, high_freq: float = 0.0) -> np.ndarray:
    """Kaldi ``MelBanks`` triangular filters: dense [num_bins, n_fft/2] float32.

    Kaldi evaluates the triangles in the mel domain at FFT-bin centre frequencies and only
    over bins 0..n_fft/2-1 (the Nyquist bin is never used).
    """
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    num_fft_bins = n_fft // 2
    fft_bin_width = sample_rate / n_fft
    mel_low = float(mel_scale(low_freq))
    mel_high = float(mel_scale(high_freq))
    delta = (mel_high - mel_low) / (num_bins + 1)
    w = np.zeros((num_bins, num_fft_bins), dtype=np.float32)
    mel = mel_scale(fft_bin_width * np.arange(num_fft_bins))
    for b in range(num_bins):
        left = mel_low + b * delta
        center = mel_low + (b + 1) * delta
        right = mel_low + (b + 2) * delta
        up = (mel - left) / (center - left)
        down = (right - mel) / (right - center)
        tri = np.where(mel <= center, up, down)
        inside = (mel > left) & (mel < right)
        w[b] = np.where(inside, tri, 0.0).astype(np.float32)
    return w


def hamming_window(n: int = 400) -> np.ndarray:
    i = np.arange(n, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * math.pi * i / (n - 1))).astype(np.float32)


def num_frames(num_samples: int, snip_edges: bool, frame_len: int = 400, frame_shift: int = 160) -> int:
    if snip_edges:
        if num_samples < frame_len:
            return 0
        return 1 + (num_samples - frame_len) // frame_shift
    return (num_samples + frame_shift // 2) // frame_shift


def extract_frames(wave: np.ndarray, snip_edges: bool, frame_len: int = 400,
                   frame_shift: int = 160) -> np.ndarray:
    """Kaldi ``ExtractWindow`` framing (without the per-frame processing): [T, frame_len]."""
    n = int(wave.shape[0])
    t = num_frames(n, snip_edges, frame_len, frame_shift)
    if t == 0:
        return np.zeros((0, frame_len), dtype=np.float32)
    f = np.arange(t)[:, None]
    j = np.arange(frame_len)[None, :]
    if snip_edges:
        idx = f * frame_shift + j
    else:
        idx = f * frame_shift + (frame_shift // 2 - frame_len // 2) + j
        # Kaldi mirrors out-of-range indices (repeatedly if needed for very short inputs)
        for _ in range(8):
            idx = np.where(idx < 0, -idx - 1, idx)
            idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    return wave[idx].astype(np.float32)


def kaldi_fbank(samples: np.ndarray, snip_edges: bool = False, num_bins: int = 80,
                sample_rate: int = 16000, preemph: float = 0.97) -> np.ndarray:
    """``OnlineFbank.GetFbank`` restated: float PCM already scaled by 32768 -> [T, num_bins]."""
    wave = np.asarray(samples, dtype=np.float32)
    frames = extract_frames(wave, snip_edges)
    t = frames.shape[0]
    if t == 0:
        return np.zeros((0, num_bins), dtype=np.float32)
    # remove_dc_offset
    frames = frames - frames.mean(axis=1, keepdims=True, dtype=np.float32)
    # pre-emphasis: x[i] -= c*x[i-1] for i = N-1..1 ; x[0] -= c*x[0]
    pe = np.empty_like(frames)
    c = np.float32(preemph)
    pe[:, 1:] = frames[:, 1:] - c * frames[:, :-1]
    pe[:, 0] = frames[:, 0] - c * frames[:, 0]
    pe = pe * hamming_window(frames.shape[1])[None, :]
    n_fft = 512
    padded = np.zeros((t, n_fft), dtype=np.float32)
    padded[:, : frames.shape[1]] = pe
    spec = np.fft.rfft(padded, axis=1)
    power = (spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2).astype(np.float32)
    mel = power[:, : n_fft // 2] @ mel_banks(num_bins, sample_rate, n_fft).T
    mel = np.maximum(mel.astype(np.float32), FLT_EPSILON)
    return np.log(mel).astype(np.float32)


def get_fbank(samples: np.ndarray, snip_edges: bool = False, num_bins: int = 80) -> np.ndarray:
    """WavFrontend.GetFbank (WavFrontend.cs:31-37): scale by 32768f, then OnlineFbank."""
    scaled = np.asarray(samples, dtype=np.float32) * np.float32(32768.0)
    return kaldi_fbank(scaled, snip_edges=snip_edges, num_bins=num_bins)


def apply_lfr(fbank: np.ndarray, lfr_m: int = 7, lfr_n: int = 6) -> np.ndarray:
    """WavFrontend.ApplyLfr (WavFrontend.cs:73-111), flat semantics kept.

    Q1: the three left-pad frames are ZEROS (the copy loop writes frame 0 at offset tile_x*80
    every iteration and the bulk copy then overwrites it).  Q2: ``t_lfr = floor(T / lfr_n)``
    and the tail-replicate branch is unreachable for lfr_m=7, lfr_n=6, but is restated anyway.
    """
    inputs = np.asarray(fbank, dtype=np.float32).reshape(-1)
    dim = 80  # hard-coded in the reference (WavFrontend.cs:75)
    t = inputs.shape[0] // dim
    t_lfr = int(math.floor(t // lfr_n))
    tile_x = (lfr_m - 1) // 2
    t = t + tile_x
    temp = np.zeros(t * dim, dtype=np.float32)
    temp[tile_x * dim: tile_x * dim + inputs.shape[0]] = inputs
    out = np.zeros(t_lfr * lfr_m * dim, dtype=np.float32)
    for i in range(t_lfr):
        if lfr_m <= t - i * lfr_n:
            out[i * lfr_m * dim:(i + 1) * lfr_m * dim] = temp[i * lfr_n * dim: i * lfr_n * dim + lfr_m * dim]
        else:  # pragma: no cover - dead for (7, 6), kept for other settings
            num_padding = lfr_m - (t - i * lfr_n)
            frame = np.zeros(lfr_m * dim, dtype=np.float32)
            have = (t - i * lfr_n) * dim
            frame[:have] = temp[i * lfr_n * dim: i * lfr_n * dim + have]
            for j in range(num_padding):
                frame[(lfr_m - num_padding + j) * dim:(lfr_m - num_padding + j + 1) * dim] = temp[(t - 1) * dim: t * dim]
            out[i * lfr_m * dim:(i + 1) * lfr_m * dim] = frame
    return out.reshape(t_lfr, lfr_m * dim)


def apply_cmvn(feats: np.ndarray, add_shift: np.ndarray, rescale: np.ndarray) -> np.ndarray:
    """WavFrontend.ApplyCmvn (WavFrontend.cs:53-71): (x + neg_mean[k]) * inv_stddev[k] in float32."""
    x = np.asarray(feats, dtype=np.float32)
    return ((x + add_shift.astype(np.float32)[None, :]) * rescale.astype(np.float32)[None, :]).astype(np.float32)


def parse_am_mvn(text: str) -> Tuple[np.ndarray, np.ndarray]:
    """WavFrontend.LoadCmvn (WavFrontend.cs:112-153): take the bracketed vector on the
    ``<LearnRateCoef>`` line that follows ``<AddShift>`` / ``<Rescale>``."""
    state = 0
    shift: List[float] = []
    scale: List[float] = []
    for line in text.splitlines():
        if not line:
            continue
        if line.startswith("<AddShift>"):
            state = 1
            continue
        if line.startswith("<Rescale>"):
            state = 2
            continue
        if line.startswith("<LearnRateCoef>") and state in (1, 2):
            body = line[line.index("[") + 1: line.rindex("]")]
            vals = [float(tok) for tok in body.split(" ") if tok.strip()]
            if state == 1:
                shift = vals
            else:
                scale = vals
    return np.asarray(shift, dtype=np.float32), np.asarray(scale, dtype=np.float32)


def format_am_mvn(add_shift: Sequence[float], rescale: Sequence[float]) -> str:
    """Write an ``am.mvn`` file in the Kaldi-nnet text layout the reference parser expects."""
    d = len(add_shift)
    s1 = " ".join(repr(float(v)) for v in add_shift)
    s2 = " ".join(repr(float(v)) for v in rescale)
    return (f"<Nnet> \n<Splice> {d} {d}\n[ 0 ]\n<AddShift> {d} {d} \n<LearnRateCoef> 0 [ {s1} ]\n"
            f"<Rescale> {d} {d}\n<LearnRateCoef> 0 [ {s2} ]\n</Nnet> \n")


def extract_features(samples: np.ndarray, add_shift: np.ndarray, rescale: np.ndarray,
                     snip_edges: bool = False, lfr_m: int = 7, lfr_n: int = 6) -> np.ndarray:
    """One ``OfflineStream.AddSamples`` call (OfflineStream.cs:36-57): fbank -> LFR -> CMVN."""
    fb = get_fbank(samples, snip_edges=snip_edges)
    if fb.shape[0] == 0:
        return np.zeros((0, lfr_m * 80), dtype=np.float32)
    feats = fb
    if lfr_m != 1 or lfr_n != 1:
        feats = apply_lfr(fb, lfr_m, lfr_n)
    if feats.shape[0] == 0:
        return feats.reshape(0, lfr_m * 80)
    return apply_cmvn(feats, add_shift, rescale)


def pad_sequence(feats: Sequence[np.ndarray]) -> np.ndarray:
    """PadHelper.PadSequence (Utils/PadHelper.cs:23-65): right-pad with 0 to the longest item, then
    replace EVERY exact 0.0 by -23.025850929940457f*32768 (Q4).  Returns [B, T_max, 560]."""
    dim = feats[0].shape[-1]
    tmax = max(int(f.shape[0]) for f in feats)
    out = np.zeros((len(feats), tmax, dim), dtype=np.float32)
    for i, f in enumerate(feats):
        out[i, : f.shape[0]] = f
    out[out == 0] = PAD_QUIRK_VALUE
    return out
