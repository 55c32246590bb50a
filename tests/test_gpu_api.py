"""The reference's API contracts, on the device: the 10 xUnit facts of
/root/reference/AliParaformerAsr.Tests/OfflineRecognizerTests .cs:166-353 restated against
aliparaformerasr_b200.offline.OfflineRecognizer / OfflineStream (which drive libpfasr.so through the C-ABI), plus the
stream semantics the facts do not pin: Q10 (every AddSamples call is an independent fbank -> LFR -> CMVN, features are
concatenated, OfflineStream.cs:36-57), per-stream Hotwords on a SeACo handle (OfflineProjOfSeacoParaformer.cs:51-60)
and GetResults on a SenseVoiceSmall handle (OfflineProjOfSenseVoiceSmall.cs:53-175).

The reference fixture points at a downloaded sensevoice-small-int8 model directory; here the model directory is
synthetic (seeded weights as a PFW1 blob, asr.json, am.mvn, tokens.txt written to tmp_path)."""
import json
import threading

import numpy as np
import pytest

from aliparaformerasr_b200 import synth, weights as W
from aliparaformerasr_b200.offline import (ArgumentNullError, ObjectDisposedError, OfflineRecognizer,
                                            OfflineRecognizerResultEntity, get_hotwords, pad_sequence)
from oracle import frontend as F, sanm
from _util import dims_of, margins

pytestmark = pytest.mark.gpu


def _model_dir(tmp_path_factory, model):
    d = tmp_path_factory.mktemp(model)
    cfg = synth.tiny(model)
    w = synth.make_weights(cfg)
    W.save(str(d / "model.pfw"), w)
    conf = {"model": cfg.model, "use_itn": cfg.use_itn, "vocab_size": cfg.vocab, "ln_eps": cfg.ln_eps,
            "encoder_conf": {"output_size": cfg.d_model, "attention_heads": cfg.heads, "linear_units": cfg.ffn, "num_blocks": cfg.enc_layers,
                             "tp_blocks": cfg.tp_layers, "kernel_size": cfg.enc_kernel},
            "decoder_conf": {"num_blocks": cfg.dec_layers, "linear_units": cfg.dec_ffn, "kernel_size": cfg.dec_kernel},
            "predictor_conf": {"threshold": cfg.cif_threshold, "tail_threshold": cfg.cif_tail},
            "frontend_conf": {"fs": 16000, "n_mels": 80, "lfr_m": 7, "lfr_n": 6, "window": "hamming", "frame_length": 25, "frame_shift": 10,
                              "dither": 0.0, "snip_edges": False},
            "seaco_decoder_conf": {"num_blocks": cfg.seaco_layers, "linear_units": cfg.seaco_ffn, "kernel_size": cfg.seaco_kernel}}
    (d / "asr.json").write_text(json.dumps(conf), encoding="utf-8")
    (d / "am.mvn").write_text(F.format_am_mvn(*synth.make_cmvn()), encoding="utf-8")
    toks = ["<blank>", "<s>", "</s>"] + [chr(0x4E00 + i) for i in range(cfg.vocab - 3)]
    (d / "tokens.txt").write_text("\n".join(toks) + "\n", encoding="utf-8")
    (d / "hotwords.txt").write_text("\n".join(toks[10 + 3 * i] + toks[11 + 3 * i] for i in range(5)) + "\n", encoding="utf-8")
    return d, cfg, w, toks


def _rec(d, **kw):
    args = dict(model_file_path=str(d / "model.pfw"), config_file_path=str(d / "asr.json"), mvn_file_path=str(d / "am.mvn"),
                tokens_file_path=str(d / "tokens.txt"), modeleb_file_path="", hotword_file_path="", threads_num=2)
    args.update(kw)
    return OfflineRecognizer(**args)


@pytest.fixture(scope="module")
def sv(tmp_path_factory):
    d, cfg, w, toks = _model_dir(tmp_path_factory, "sensevoicesmall")
    rec = _rec(d)
    yield d, cfg, w, toks, rec
    rec.Dispose()


def _mock_audio(seconds=1, fs=16000):
    return np.zeros(seconds * fs, np.float32)            # GenerateMockAudioSamples: 1 s of silence


# ------------------------------------------------------------------ the ten facts (OfflineRecognizerTests .cs)
def test_fact1_init_with_valid_params_returns_non_null(sv):                      # :166-179
    assert sv[4] is not None


def test_fact2_init_with_missing_tokens_file_throws(sv):                         # :184-208
    with pytest.raises(Exception, match="tokens invalid"):
        _rec(sv[0], tokens_file_path="")


def test_fact3_create_stream_add_samples(sv):                                    # :213-226
    stream = sv[4].CreateOfflineStream()
    stream.AddSamples(_mock_audio())
    assert stream is not None


def test_fact4_get_result_with_valid_stream_returns_result_entity(sv):           # :231-249
    rec = sv[4]
    stream = rec.CreateOfflineStream()
    stream.AddSamples(_mock_audio())
    result = rec.GetResult(stream)
    assert isinstance(result, OfflineRecognizerResultEntity)
    assert result.Text is not None and isinstance(result.Text, str)
    assert result.Tokens is not None and result.Timestamps is not None
    assert len(stream.Tokens) == 16 + 4                   # 1 s -> 100 fbank frames -> 16 LFR frames + 4 prompt rows, one id per frame
    assert len(stream.Timestamps) == len(stream.Tokens)   # 3-output model: {0,0} per token (OfflineRecognizer.cs:151)


def test_fact5_add_samples_with_valid_samples_does_not_throw(sv):               # :266-280
    stream = sv[4].CreateOfflineStream()
    stream.AddSamples(np.full(1000, 0.1, np.float32))
    assert stream.features().shape == (1, 560)            # 6 fbank frames -> 1 LFR frame with snip_edges=false (Q2)


def test_fact6_add_samples_with_null_throws_argument_null_source(sv):           # :285-297
    stream = sv[4].CreateOfflineStream()
    with pytest.raises(ArgumentNullError) as ei:
        stream.AddSamples(None)
    assert ei.value.param_name == "source"


def test_fact7_set_hotwords_with_valid_text(sv):                                 # :302-318
    d, cfg, w, toks, rec = sv
    stream = rec.CreateOfflineStream()
    hot = get_hotwords(toks, str(d / "hotwords.txt"))
    stream.Hotwords = get_hotwords(toks, str(d / "hotwords.txt"))
    assert stream.Hotwords == hot and len(hot) == 6 and hot[-1] == [1] and all(len(h) == 2 for h in hot[:-1])


def test_fact8_set_hotwords_null_clears(sv):                                     # :323-335
    d, cfg, w, toks, rec = sv
    stream = rec.CreateOfflineStream()
    stream.Hotwords = get_hotwords(toks, str(d / "hotwords.txt"))
    stream.Hotwords = None
    assert stream.Hotwords is None
    stream.AddSamples(_mock_audio())
    assert rec.GetResult(stream) is not None              # a null list is skipped like an empty one (:52-60)


def test_fact9_dispose_then_create_stream_throws_object_disposed(sv):            # :340-353
    rec = _rec(sv[0])
    rec.Dispose()
    with pytest.raises(ObjectDisposedError) as ei:
        rec.CreateOfflineStream()
    assert ei.value.object_name == "OfflineRecognizer"
    rec.Dispose()                                          # idempotent, like the C# Dispose(bool) guard (:468-489)


def test_fact10_fixture_reuses_one_recognizer_for_many_streams(sv):             # InitRecognizer :355-370: created once, reused
    rec = sv[4]
    outs = []
    for k in range(3):
        s = rec.CreateOfflineStream()
        s.AddSamples(synth.make_pcm(k, 1.5))
        outs.append(rec.GetResult(s).Text)
    s = rec.CreateOfflineStream()
    s.AddSamples(synth.make_pcm(0, 1.5))
    assert rec.GetResult(s).Text == outs[0]               # deterministic across calls on the same handle


# ------------------------------------------------------------------ SenseVoice GetResults vs the oracle
def test_sensevoice_get_results_matches_oracle(sv):
    d, cfg, w, toks, rec = sv
    pcm = [synth.make_pcm(40 + i, 2.0) for i in range(3)]
    streams = []
    for p in pcm:
        s = rec.CreateOfflineStream()
        s.AddSamples(p)
        streams.append(s)
    results = rec.GetResults(streams)
    shift, scale = synth.make_cmvn()
    speech = np.stack([sanm.sensevoice_prepend(F.extract_features(p, shift, scale), w["embed.weight"], cfg.use_itn) for p in pcm])
    ref = sanm.sensevoice_forward(speech, w, dims_of(cfg))
    safe = margins(ref["logits"]) > 0.1
    got = np.asarray([s.Tokens for s in streams])
    assert got.shape == ref["tokens"].shape and np.array_equal(got[safe], ref["tokens"][safe]) and safe.mean() > 0.6
    assert len(results) == 3 and all(isinstance(r.Text, str) for r in results)


# ------------------------------------------------------------------ Q10: several AddSamples calls on one stream
@pytest.fixture(scope="module")
def pf(tmp_path_factory):
    d, cfg, w, toks = _model_dir(tmp_path_factory, "paraformer")
    rec = _rec(d)
    yield d, cfg, w, toks, rec
    rec.Dispose()


def test_q10_two_add_samples_concatenate_independent_features(pf):
    d, cfg, w, toks, rec = pf
    a, b = synth.make_pcm(50, 2.0), synth.make_pcm(51, 1.3)
    shift, scale = synth.make_cmvn()
    s1 = rec.CreateOfflineStream()
    s1.AddSamples(a)
    s1.AddSamples(b)                                       # no sample carry-over: a second, independent front-end pass
    s2 = rec.CreateOfflineStream()
    s2.AddSamples(synth.make_pcm(52, 2.5))
    want = np.concatenate([F.extract_features(a, shift, scale), F.extract_features(b, shift, scale)], axis=0)
    got = s1.features()
    assert got.shape == want.shape == (33 + 21, 560)
    assert np.abs(got - want).max() < 2e-3                 # same bound as tests/test_gpu_frontend.py
    # the mixed batch takes the host PadSequence + pf_offline_run_feats path (offline.py:_forward else-branch)
    results = rec.GetResults([s1, s2])
    feats = [want, F.extract_features(synth.make_pcm(52, 2.5), shift, scale)]
    ref = sanm.paraformer_forward(F.pad_sequence(feats), w, dims_of(cfg))
    assert np.array_equal(pad_sequence(feats), F.pad_sequence(feats))
    safe = margins(ref["logits"]) > 0.1
    got_ids = np.asarray([s1.Tokens, s2.Tokens])
    assert got_ids.shape == ref["tokens"].shape and np.array_equal(got_ids[safe], ref["tokens"][safe]) and safe.mean() > 0.7
    assert all(len(r.Timestamps) == len(r.Tokens) for r in results)
    assert s1._chunks == [] and s2._chunks == []           # RemoveChunk (OfflineStream.cs:69-79)


# ------------------------------------------------------------------ per-stream Hotwords on a SeACo handle
@pytest.fixture(scope="module")
def seaco(tmp_path_factory):
    d, cfg, w, toks = _model_dir(tmp_path_factory, "seacoparaformer")
    rec = _rec(d, hotword_file_path=str(d / "hotwords.txt"), lanes=2)
    yield d, cfg, w, toks, rec
    rec.Dispose()


def _seaco_ref(w, cfg, pcm, hot):
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(p, shift, scale) for p in pcm])
    rows = sanm.bias_embed_rows(sanm.hotword_embed(sanm.pad_hotwords(hot), w))
    return sanm.seaco_forward(speech, w, dims_of(cfg), rows)


def _decided(ref, cfg):
    ds = np.sort(ref["dha"], axis=-1)
    nob = ref["dha"][..., cfg.nobias_id]
    top_other = np.where(ref["dha_ids"] == cfg.nobias_id, ds[..., -2], ds[..., -1])
    return (np.abs(nob - top_other) > 0.25) & (margins(ref["logits"]) > 0.1)


def test_seaco_file_hotwords_then_per_stream_hotwords_then_file_again(seaco):
    d, cfg, w, toks, rec = seaco
    file_hot = get_hotwords(toks, str(d / "hotwords.txt"))
    pcm = [synth.make_pcm(60 + i, 3.0) for i in range(2)]

    def run(per_stream):
        streams = []
        for p, h in zip(pcm, per_stream):
            s = rec.CreateOfflineStream()
            s.AddSamples(p)
            s.Hotwords = h
            streams.append(s)
        rec.GetResults(streams)
        assert all(len(s.Timestamps) > 0 for s in streams)          # 4-output model: real timestamps (OfflineRecognizer.cs:172-183)
        return np.asarray([s.Tokens for s in streams])

    ref_file = _seaco_ref(w, cfg, pcm, file_hot)
    ok = _decided(ref_file, cfg)
    got = run([[], []])
    assert np.array_equal(got[ok], ref_file["tokens"][ok]) and ok.mean() > 0.5
    # hot words of all streams of the call are concatenated and replace the file ones for that call (:51-60)
    h0, h1 = [[7, 8, 9]], [[100, 101], [300]]
    ref_call = _seaco_ref(w, cfg, pcm, h0 + h1)
    ok2 = _decided(ref_call, cfg)
    got2 = run([h0, h1])
    assert np.array_equal(got2[ok2], ref_call["tokens"][ok2]) and ok2.mean() > 0.5
    got3 = run([None, []])                                            # ... and the file hot words are back afterwards
    assert np.array_equal(got3, got)


def test_seaco_per_call_hotwords_do_not_leak_between_threads_sharing_lanes(seaco):
    """set -> run -> restore is one leased section, and results live in the calling thread's storage: four
    threads on two lanes, two of them with per-call hot words, all get the ids of their own configuration."""
    d, cfg, w, toks, rec = seaco
    pcm = [synth.make_pcm(70, 3.0)]

    def once(hot):
        s = rec.CreateOfflineStream()
        s.AddSamples(pcm[0])
        s.Hotwords = hot
        rec.GetResults([s])
        return list(s.Tokens)

    base, alt = once([]), once([[7, 8, 9], [500, 600]])
    errors = []

    def worker(hot, want):
        try:
            for _ in range(6):
                if once(hot) != want:
                    errors.append((hot, "ids differ"))
        except Exception as ex:                                          # noqa: BLE001
            errors.append((hot, repr(ex)))

    threads = [threading.Thread(target=worker, args=(h, wnt)) for h, wnt in (([], base), ([[7, 8, 9], [500, 600]], alt)) * 2]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
