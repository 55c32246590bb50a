"""CUDA front-end (fbank + LFR + CMVN + pad quirk) vs the CPU oracle and the committed Kaldi golden vectors."""
import os

import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from oracle import frontend as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

# log-mel tolerance: two float32 implementations of the same Kaldi recipe differ by up to ~5e-4 on weak bins
# (oracle vs torchaudio vs float64, see tests/test_oracle_frontend.py); the CUDA FFT sits in the same band.
FBANK_ATOL = 2e-3


@pytest.fixture(scope="module")
def engines():
    out = {}
    for snip in (False, True):
        cfg = synth.tiny()
        cfg.snip_edges = snip
        eng = Engine(cfg, synth.make_weights(cfg))
        shift, scale = synth.make_cmvn()
        eng.set_cmvn(shift, scale)
        out[snip] = eng
    yield out
    for e in out.values():
        e.close()


@pytest.mark.parametrize("snip", [False, True])
@pytest.mark.parametrize("n", [160000, 80000, 16000, 12345, 1000, 400, 399, 170])
def test_fbank_matches_oracle(engines, snip, n):
    x = synth.make_pcm(n % 97, n / 16000.0)[:n]
    ref = F.get_fbank(x, snip_edges=snip)
    got = engines[snip].fbank(x)
    assert got.shape == ref.shape
    if ref.size:
        assert np.abs(got - ref).max() < FBANK_ATOL
        assert np.abs(got - ref).mean() < 1e-4


@pytest.mark.parametrize("snip", [False, True])
def test_fbank_matches_kaldi_golden(engines, snip):
    g = np.load(os.path.join(GOLD, "kaldi_fbank.npz"))
    x = g["pcm"]
    ref = g["fbank_snip1" if snip else "fbank_snip0"]
    got = engines[snip].fbank(x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < FBANK_ATOL


@pytest.mark.parametrize("snip", [False, True])
@pytest.mark.parametrize("n", [160000, 80000, 12345, 1000, 960, 959])
def test_extract_matches_oracle(engines, snip, n):
    x = synth.make_pcm(3, n / 16000.0)[:n]
    shift = np.linspace(-9, -7, 560).astype(np.float32)
    scale = np.linspace(0.2, 0.3, 560).astype(np.float32)
    engines[snip].set_cmvn(shift, scale)
    ref = F.extract_features(x, shift, scale, snip_edges=snip)
    got = engines[snip].extract(x)
    engines[snip].set_cmvn(*synth.make_cmvn())
    assert got.shape == ref.shape
    if ref.size:
        assert np.abs(got - ref).max() < FBANK_ATOL


def test_silence_and_zero_quirk(engines):
    """1 s of zeros (reference test CreateStream_AddSamples, OfflineRecognizerTests .cs:213-226): log floor everywhere."""
    eng = engines[False]
    x = np.zeros(16000, np.float32)
    ref = F.extract_features(x, *synth.make_cmvn())
    got = eng.extract(x)
    assert got.shape == ref.shape == (16, 560)
    assert np.abs(got - ref).max() < 1e-4
    # Q1: the three left-pad frames are zeros before CMVN -> exactly shift*scale = -2.0
    assert np.all(got[0, :240] == -2.0)


def test_batch_features_ragged_pad_quirk(engines):
    """run_pcm's fused front-end == per-utterance oracle features + PadHelper.PadSequence (Q4)."""
    eng = engines[False]
    pcm = [synth.make_pcm(i, s) for i, s in enumerate([2.0, 1.0, 1.53])]
    shift, scale = synth.make_cmvn()
    ref = F.pad_sequence([F.extract_features(p, shift, scale) for p in pcm])
    eng.run_pcm(pcm)
    got = eng.tensor("feats")
    assert got.shape == ref.shape
    pad = ref == F.PAD_QUIRK_VALUE
    assert np.array_equal(got == F.PAD_QUIRK_VALUE, pad)
    assert np.abs(got[~pad] - ref[~pad]).max() < FBANK_ATOL
