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
