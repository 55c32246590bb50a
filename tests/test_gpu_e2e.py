"""End-to-end parity of the CUDA path (through the C-ABI) with the CPU oracle on seeded synthetic weights."""
import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from oracle import frontend as F, sanm
from _util import dims_of, margins

pytestmark = pytest.mark.gpu

# north_star tolerance: "logits within 1e-2 fp16" of the fp32 CPU path.  What that bound can mean for an fp16-operand /
# fp32-accumulate implementation is MEASURED at full depth in tests/test_gpu_fulldepth.py and committed as
# profiles/parity_r02.md: the float32 oracle merely given fp16-rounded operands (oracle.sanm.OperandRounding) already
# sits rms 7.5e-3 / max 0.22 (CIF tail row) / 0.06 (other rows) from the float32 result on paraformer-large, and the CUDA
# path tracks that operand model to rms 3.2e-3 / max 1.9e-2.  The few-layer models of this file are held to two criteria
# that follow from it and involve no outlier allowance:
#   * against the operand model (the same three bounds as at full depth): every log-prob within LOGIT_MAX_MODEL, at
#     least 99 % within 1e-2, rms within LOGIT_RMS_MODEL;
#   * against float32: rms within LOGIT_RMS and max no further than 1.25 x the operand model's own distance + 1e-2.
# (The max sits on the CIF tail row: 1.9e-2 on the 3+2-layer model of the smoke test, the same as at 50+16 layers.)
LOGIT_MAX_MODEL = 3e-2
LOGIT_FRAC_OVER_1E2_MODEL = 1e-2
LOGIT_RMS_MODEL = 5e-3
LOGIT_RMS = 1e-2
# greedy ids must agree wherever the oracle's top-1/top-2 margin exceeds this (closer calls are decided by rounding)
TOKEN_MARGIN = 0.1


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


def _check_logits(got, ref, model, max_model=LOGIT_MAX_MODEL):
    """got: CUDA log-probs; ref: float32 oracle; model: the oracle with fp16 operand rounding."""
    d_model = np.abs(got - model)
    assert d_model.max() <= max_model, f"max abs err vs the fp16-operand oracle {d_model.max()}"
    assert float((d_model > 1e-2).mean()) <= LOGIT_FRAC_OVER_1E2_MODEL * (max_model / LOGIT_MAX_MODEL)
    assert float(np.sqrt(np.mean(d_model.astype(np.float64) ** 2))) <= LOGIT_RMS_MODEL
    diff = np.abs(got - ref)
    intrinsic = float(np.abs(model - ref).max())
    assert diff.max() <= 1.25 * intrinsic + 1e-2, f"max abs err vs float32 {diff.max()} (operand rounding alone: {intrinsic})"
    rms = float(np.sqrt(np.mean(diff.astype(np.float64) ** 2)))
    assert rms < LOGIT_RMS, f"logits rms err {rms}"


def _oracle_pair(fn, *args):
    ref = fn(*args)
    with sanm.OperandRounding():
        model = fn(*args)
    return ref, model


def _compare(out, refs, cfg, max_model=LOGIT_MAX_MODEL):
    ref, model = refs
    assert np.array_equal(out.token_num, ref["token_num"]) and np.array_equal(out.token_num, model["token_num"])
    assert out.logits.shape == ref["logits"].shape
    _check_logits(out.logits, ref["logits"], model["logits"], max_model)
    safe = margins(ref["logits"]) > TOKEN_MARGIN
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])
    assert safe.mean() > 0.8


def test_run_feats_matches_oracle(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(i, 5.0) for i in range(3)]
    speech = _oracle_feats(pcm, cfg)
    refs = _oracle_pair(sanm.paraformer_forward, speech, w, dims_of(cfg))
    ref = refs[0]
    out = eng.run_feats(speech, want_logits=True)
    enc = eng.tensor("enc")
    assert np.abs(enc - ref["enc"]).max() < 1e-2
    assert np.abs(eng.tensor("alphas") - ref["alphas"]).max() < 2e-3
    _compare(out, refs, cfg)


def test_run_pcm_matches_oracle(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(10 + i, 5.0) for i in range(4)]
    refs = _oracle_pair(sanm.paraformer_forward, _oracle_feats(pcm, cfg), w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    _compare(out, refs, cfg)
    # tokens-only call returns the same ids
    out2 = eng.run_pcm(pcm)
    assert np.array_equal(out2.tokens, out.tokens) and out2.logits is None


def test_single_utterance_cfg1_shape(tiny_paraformer):
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(0, 5.0)]
    refs = _oracle_pair(sanm.paraformer_forward, _oracle_feats(pcm, cfg), w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert out.feat_frames == 83
    _compare(out, refs, cfg)


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
    ref, model = _oracle_pair(sanm.sensevoice_forward, speech, w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert out.tokens.shape == ref["tokens"].shape == (2, 50 + 4)
    _check_logits(out.logits, ref["logits"], model["logits"])
    safe = margins(ref["logits"]) > TOKEN_MARGIN
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])
    eng.close()


def test_ragged_batch_with_pad_quirk(tiny_paraformer):
    """Different lengths in one batch: short items are right-padded, every exact zero becomes -23.0258509*32768 (Q4) and
    the padded frames are attended to because speech_lengths = T_max for all items (Q3)."""
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(20, 5.0), synth.make_pcm(21, 2.3), synth.make_pcm(22, 3.71), np.zeros(16000, np.float32)]
    refs = _oracle_pair(sanm.paraformer_forward, _oracle_feats(pcm, cfg), w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    # the padded frames hold -754511.06 in every dim (Q4): their first LayerNorm sees a variance of ~0.5 (the positional
    # encoding) on values of 1.7e7, which float32 statistics cannot resolve - the oracle (torch LN in float32) and the
    # kernel (float64 statistics on such rows) both produce SOME unit-variance row there, and every real frame attends to
    # them (Q3).  The comparison is therefore looser than for well-conditioned input: measured 4.8e-2.
    _compare(out, refs, cfg, max_model=0.1)


def test_long_utterance_takes_the_streaming_attention_path(tiny_paraformer):
    """T_lfr > 192 frames (11.5 s): the single-pass tcgen05 attention does not apply; FSMN kernel + streaming attention."""
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(30, 14.0), synth.make_pcm(31, 12.5)]
    speech = _oracle_feats(pcm, cfg)
    assert speech.shape[1] > 192
    refs = _oracle_pair(sanm.paraformer_forward, speech, w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert np.abs(eng.tensor("enc") - refs[0]["enc"]).max() < 1e-2
    _compare(out, refs, cfg)


def test_very_long_and_very_short_utterances_in_one_batch(tiny_paraformer):
    """Extreme raggedness: 45 s (T_lfr = 750: five 128-row GEMM tiles per utterance, streaming attention over 750 keys) next to
    0.4 s (T_lfr = 6) - the short one is 99 % PadSequence rows (Q4), which the reference does not mask (Q3)."""
    cfg, w, eng = tiny_paraformer
    pcm = [synth.make_pcm(40, 45.0), synth.make_pcm(41, 0.4)]
    speech = _oracle_feats(pcm, cfg)
    assert speech.shape[1] >= 740
    refs = _oracle_pair(sanm.paraformer_forward, speech, w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert np.array_equal(out.token_num, refs[0]["token_num"])
    _compare(out, refs, cfg, max_model=0.1)


def test_full_size_properties():
    """BASELINE configs[1] size (paraformer-large, 32 x 10 s), checked through size-independent properties: the step is
    deterministic, ids are valid, and an utterance decodes to the same ids alone or inside the batch (equal lengths =>
    no cross-utterance coupling) wherever the batch run's top-1/top-2 margin is not within rounding."""
    cfg = synth.paraformer_large()
    eng = Engine(cfg, synth.make_weights(cfg))
    eng.set_cmvn(*synth.make_cmvn())
    pcm = [synth.make_pcm(i, 10.0) for i in range(32)]
    a = eng.run_pcm(pcm, want_logits=True)
    b = eng.run_pcm(pcm)
    assert a.feat_frames == 166 and a.tokens.shape[0] == 32
    assert np.array_equal(a.tokens, b.tokens) and np.array_equal(a.token_num, b.token_num)
    assert a.tokens.min() >= 0 and a.tokens.max() < cfg.vocab
    assert (a.token_num <= a.tokens.shape[1]).all() and (a.token_num >= 1).all() and a.token_num.max() == a.tokens.shape[1]
    assert np.allclose(np.exp(a.logits.astype(np.float64)).sum(-1), 1.0, atol=1e-3)        # log-softmax rows
    sub = eng.run_pcm(pcm[:5], want_logits=True)
    assert np.array_equal(sub.token_num, a.token_num[:5])
    L = sub.tokens.shape[1]
    safe = margins(a.logits[:5, :L]) > 0.1
    assert safe.mean() > 0.8
    assert np.array_equal(sub.tokens[safe], a.tokens[:5, :L][safe])
    eng.close()


def test_engine_loads_the_reference_model_file_format(tmp_path, tiny_paraformer):
    """Engine(cfg, "model.onnx"): the ONNX initialisers are read, mapped and packed on the fly (row f3) and give the same
    ids and logits as the state dict they were exported from."""
    from _util import export_paraformer_onnx
    cfg, w, eng = tiny_paraformer
    path = tmp_path / "model.onnx"
    path.write_bytes(export_paraformer_onnx(w, cfg.enc_layers, cfg.dec_layers))
    eng2 = Engine(cfg, str(path))
    eng2.set_cmvn(*synth.make_cmvn())
    pcm = [synth.make_pcm(i, 2.0) for i in range(3)]
    a = eng.run_pcm(pcm, want_logits=True)
    b = eng2.run_pcm(pcm, want_logits=True)
    assert np.array_equal(a.tokens, b.tokens) and np.array_equal(a.logits, b.logits)
    eng2.close()
