"""End-to-end parity of the CUDA path (through the C-ABI) with the CPU oracle on seeded synthetic weights."""
import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from oracle import frontend as F, sanm
from _util import dims_of, margins

pytestmark = pytest.mark.gpu

# north_star tolerance: "logits within 1e-2" of the fp32 CPU path (fp16 operands, fp32 accumulation on the GPU)
LOGIT_ATOL = 1e-2


def _oracle_feats(pcm, cfg):
    shift, scale = synth.make_cmvn()
    return F.pad_sequence([F.extract_features(p, shift, scale, snip_edges=cfg.snip_edges) for p in pcm])


@pytest.fixture(scope="module")
def tiny_paraformer():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w)
    eng.set_cmvn(*synth.make_cmvn())
    yield cfg, w, eng
    eng.close()


def _compare(out, ref, cfg):
    assert np.array_equal(out.token_num, ref["token_num"])
    assert out.logits.shape == ref["logits"].shape
    err = np.abs(out.logits - ref["logits"]).max()
    assert err < LOGIT_ATOL, f"logits max abs err {err}"
    safe = margins(ref["logits"]) > 2 * LOGIT_ATOL
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])
    assert safe.mean() > 0.9


def test_run_feats_matches_oracle(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(i, 5.0) for i in range(3)]
    speech = _oracle_feats(pcm, cfg)
    ref = sanm.paraformer_forward(speech, w, dims_of(cfg))
    out = eng.run_feats(speech, want_logits=True)
    enc = eng.tensor("enc")
    assert np.abs(enc - ref["enc"]).max() < 1e-2
    assert np.abs(eng.tensor("alphas") - ref["alphas"]).max() < 2e-3
    _compare(out, ref, cfg)


def test_run_pcm_matches_oracle(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(10 + i, 5.0) for i in range(4)]
    ref = sanm.paraformer_forward(_oracle_feats(pcm, cfg), w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    _compare(out, ref, cfg)
    # tokens-only call returns the same ids
    out2 = eng.run_pcm(pcm)
    assert np.array_equal(out2.tokens, out.tokens) and out2.logits is None


def test_single_utterance_cfg1_shape(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(0, 5.0)]
    ref = sanm.paraformer_forward(_oracle_feats(pcm, cfg), w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert out.feat_frames == 83
    _compare(out, ref, cfg)


def test_too_short_audio_gives_empty_result(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    out = eng.run_pcm([np.full(500, 0.1, np.float32)])      # 3 fbank frames -> 0 LFR frames
    assert out.tokens.shape == (1, 0) and out.token_num.tolist() == [0]


def test_sensevoice_matches_oracle():
    cfg = synth.tiny("sensevoicesmall")
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w)
    eng.set_cmvn(*synth.make_cmvn())
    pcm = [synth.make_pcm(20 + i, 3.0) for i in range(2)]
    shift, scale = synth.make_cmvn()
    feats = [F.extract_features(p, shift, scale) for p in pcm]
    speech = np.stack([sanm.sensevoice_prepend(x, w["embed.weight"], cfg.use_itn) for x in feats])
    ref = sanm.sensevoice_forward(speech, w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert out.tokens.shape == ref["tokens"].shape == (2, 49 + 4)
    err = np.abs(out.logits - ref["logits"]).max()
    assert err < LOGIT_ATOL, err
    safe = margins(ref["logits"]) > 2 * LOGIT_ATOL
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])
    eng.close()
