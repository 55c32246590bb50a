"""SeACo-paraformer parity (OfflineProjOfSeacoParaformer + EmbedSeacoModel) against oracle/sanm.py: hot-word LSTM
encoder, Q8 bias_embed layout, bias decoder (two passes), hot-word head and the NO_BIAS merge, through the C-ABI."""
import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from oracle import frontend as F, sanm
from _util import dims_of, margins

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_seaco():
    cfg = synth.tiny("seacoparaformer")
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w)
    eng.set_cmvn(*synth.make_cmvn())
    yield cfg, w, eng
    eng.close()


def _speech(n, seconds=3.0):
    shift, scale = synth.make_cmvn()
    pcm = [synth.make_pcm(i, seconds) for i in range(n)]
    return pcm, F.pad_sequence([F.extract_features(p, shift, scale) for p in pcm])


@pytest.mark.parametrize("nhot", [20, 200])
def test_seaco_parity(tiny_seaco, nhot):
    cfg, w, eng = tiny_seaco
    hot = synth.make_hotwords(nhot, cfg.vocab)
    hot[3] = list(range(3, 20))                       # longer than 10 ids: PadList truncates (EmbedSeacoModel.cs:119)
    eng.set_hotwords(hot)
    pcm, speech = _speech(4)
    rows = sanm.bias_embed_rows(sanm.hotword_embed(sanm.pad_hotwords(hot), w))
    assert rows.shape == ((nhot + 1) * 10, 512)      # Q8: all 10 LSTM steps of every hot word (+ the [sos] entry)
    ref = sanm.seaco_forward(speech, w, dims_of(cfg), rows)
    out = eng.run_pcm(pcm, want_logits=True)
    assert np.array_equal(out.token_num, ref["token_num"])
    # a row's branch (ASR vs hot-word posterior) is decided by the hot-word argmax: compare rows where that decision
    # and the final pick are not within rounding distance
    dha_sorted = np.sort(ref["dha"], axis=-1)
    nob = ref["dha"][..., cfg.nobias_id]
    top_other = np.where(ref["dha_ids"] == cfg.nobias_id, dha_sorted[..., -2], dha_sorted[..., -1])
    decided = np.abs(nob - top_other) > 0.25
    safe = decided & (margins(ref["logits"]) > 0.1)
    assert safe.mean() > 0.5
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])
    diff = np.abs(out.logits - ref["logits"])[decided]
    bound = 2e-2 + 1e-2 * np.abs(ref["logits"][decided])
    assert (diff <= bound).mean() > 0.999 and (diff <= 3 * bound).all(), float(diff.max())
    kept = (ref["dha_ids"] == cfg.nobias_id)
    assert 0.1 < kept.mean() < 0.9                    # both branches are exercised


def test_seaco_without_hotwords_is_plain_paraformer(tiny_seaco):
    cfg, w, eng = tiny_seaco
    eng.set_hotwords([])
    pcm, speech = _speech(2)
    ref = sanm.paraformer_forward(speech, w, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    assert np.array_equal(out.token_num, ref["token_num"])
    safe = margins(ref["logits"]) > 0.1
    assert np.array_equal(out.tokens[safe], ref["tokens"][safe])


def test_hotwords_rejected_on_plain_paraformer():
    from aliparaformerasr_b200 import _lib
    cfg = synth.tiny()
    eng = Engine(cfg, synth.make_weights(cfg))
    with pytest.raises(_lib.PfError):
        eng.set_hotwords([[5, 6]])
    eng.close()


def test_timestamp_branch(tiny_seaco):
    """Row f1: us_alphas / us_cif_peak of the CifPredictorV3 upsampler (persistent BiLSTM) and the host timestamps."""
    from aliparaformerasr_b200.offline import time_stamp_lfr6_onnx
    cfg, w, eng = tiny_seaco
    eng.set_hotwords([])
    pcm, speech = _speech(3, seconds=4.0)
    dims = dims_of(cfg)
    ref = sanm.paraformer_forward(speech, w, dims)
    ua, pk = sanm.upsample_timestamp(ref["enc"], ref["token_num"], w, dims)
    out = eng.run_pcm(pcm, want_timestamps=True)
    assert out.us_alphas.shape == ua.shape == (3, 3 * speech.shape[1])
    assert np.abs(out.us_alphas - ua).max() < 3e-3
    assert np.abs(out.us_cif_peak - pk).max() < 3e-2            # 3T-step running sum of the alphas
    assert np.allclose(out.us_alphas.sum(1), ref["token_num"], atol=1e-2)
    for i in range(3):
        fires_ref = np.nonzero(pk[i] > 1 - 1e-4)[0]
        fires = np.nonzero(out.us_cif_peak[i] > 1 - 1e-4)[0]
        assert len(fires) == len(fires_ref) and np.abs(fires - fires_ref).max() <= 1
        ts = time_stamp_lfr6_onnx(out.us_cif_peak[i], list(out.tokens[i]))
        ts_ref = sanm.time_stamp_lfr6_onnx(pk[i], list(ref["tokens"][i]))
        assert len(ts) == len(ts_ref)
        assert all(abs(a[0] - b[0]) <= 20 and abs(a[1] - b[1]) <= 20 for a, b in zip(ts, ts_ref))      # <= one upsampled frame


@pytest.mark.parametrize("n", [20, 36])
def test_timestamp_branch_wide_and_split_launches(tiny_seaco, n):
    """The persistent BiLSTM takes up to 32 utterances per launch (N of its tcgen05 product: a 16- and a 32-wide variant):
    20 utterances run the 32-wide kernel once, 36 run it once plus the 16-wide one for the remaining four."""
    cfg, w, eng = tiny_seaco
    eng.set_hotwords([])
    pcm, speech = _speech(n, seconds=1.5)
    dims = dims_of(cfg)
    ref = sanm.paraformer_forward(speech, w, dims)
    ua, pk = sanm.upsample_timestamp(ref["enc"], ref["token_num"], w, dims)
    out = eng.run_pcm(pcm, want_timestamps=True)
    assert out.us_alphas.shape == ua.shape
    assert np.abs(out.us_alphas - ua).max() < 3e-3
    assert np.allclose(out.us_alphas.sum(1), ref["token_num"], atol=1e-2)
    for i in range(n):
        fires_ref = np.nonzero(pk[i] > 1 - 1e-4)[0]
        fires = np.nonzero(out.us_cif_peak[i] > 1 - 1e-4)[0]
        assert len(fires) == len(fires_ref) and (len(fires) == 0 or np.abs(fires - fires_ref).max() <= 1)


def test_timestamps_rejected_without_branch():
    from aliparaformerasr_b200 import _lib
    cfg = synth.tiny()
    eng = Engine(cfg, synth.make_weights(cfg))
    eng.set_cmvn(*synth.make_cmvn())
    with pytest.raises(_lib.PfError):
        eng.run_pcm([synth.make_pcm(0, 2.0)], want_timestamps=True)
    eng.close()
