"""Streaming (online) parity: pf_online_* through the C-ABI against oracle/online.py on seeded synthetic weights.
Covers OnlineStream chunking (Q13), window assembly (Q12, Q14, Q4-online), CIF with carry (Q15), cached decoder FSMN
with the reference's layer-0 cache quirk (Q11) and the greedy pick over padded rows (OnlineRecognizer.cs:390)."""
import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.online import OnlineEngine, OnlineRecognizer
from oracle import online as O
from _util import dims_of, margins

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 2e-2        # raw logits (not log-probs) of a 3+2 layer model, fp16 operands / fp32 accumulate
LOGIT_RTOL = 1e-2
TOKEN_MARGIN = 0.1


def _run_both(cfg, w, schedule, per_layer_cache=False, n_streams=3):
    """schedule: list of steps; each step = list of (stream index, samples or None) pushes before one Forward."""
    shift, scale = synth.make_cmvn()
    eng = OnlineEngine(cfg, w, per_layer_cache=per_layer_cache)
    eng.set_cmvn(shift, scale)
    rec = O.OnlineRecognizerOracle(w, dims_of(cfg), shift, scale, snip_edges=cfg.snip_edges, compat_layer0_cache=not per_layer_cache)
    sids = [eng.open_stream() for _ in range(n_streams)]
    ref_streams = [rec.create_stream() for _ in range(n_streams)]
    checked_rows = 0
    for step in schedule:
        for i, x in step:
            eng.push(sids[i], x)
            ref_streams[i].add_samples(x)
        col = {}
        ref_new = rec.forward(ref_streams, collect=col)
        out = eng.step(sids, want_logits=True)
        lens_ref = [len(t) for t in ref_new]
        assert list(out.appended) == lens_ref, (list(out.appended), lens_ref)
        if out.max_new > 0:
            widx = [i for i in range(n_streams) if lens_ref[i] > 0]
            ref_logits = col["logits"]
            assert len(widx) == ref_logits.shape[0]
            for b, i in enumerate(widx):
                got = out.logits[i]
                diff = np.abs(got - ref_logits[b])
                bound = LOGIT_ATOL + LOGIT_RTOL * np.abs(ref_logits[b])
                assert (diff <= bound).mean() > 0.999 and (diff <= 3 * bound).all(), float(diff.max())
                assert out.embeds_len[i] == col["lens"][b]
                safe = margins(ref_logits[b]) > TOKEN_MARGIN
                assert np.array_equal(out.new_tokens[i][safe], np.asarray(ref_new[i])[safe])
                checked_rows += int(safe.sum())
        # device-resident state follows the oracle's
        for i in range(n_streams):
            assert np.allclose(eng.state(sids[i], "cache_feats"), ref_streams[i].cache_feats, atol=2e-3, rtol=1e-4)
            assert abs(float(eng.state(sids[i], "cif_alpha")[0]) - float(ref_streams[i].cif_alpha[0])) < 2e-2
            fs = eng.state(sids[i], "fsmn")                                   # [layers, 10, 512]
            ref_fs = np.stack([c.T for c in ref_streams[i].states])            # reference keeps [512, 10]
            assert np.allclose(fs, ref_fs, atol=3e-2, rtol=1e-2), float(np.abs(fs - ref_fs).max())
    eng.close()
    return checked_rows


@pytest.fixture(scope="module")
def tiny():
    cfg = synth.tiny()
    return cfg, synth.make_weights(cfg)


def _schedule(n_steps=6):
    pcm = [synth.make_pcm(i, 4.0) for i in range(3)]
    steps = []
    for k in range(n_steps):
        step = [(0, pcm[0][k * 9600:(k + 1) * 9600])]                        # regular 600 ms producer
        if k % 2 == 0:
            step.append((1, pcm[1][k * 7000:(k + 1) * 7000 + 3000]))         # irregular sizes: skipped on some steps
        if k == 1:
            step.append((2, pcm[2][:3 * 9600]))                              # burst: one chunk per AddSamples (Q13)
        if k >= 2:
            step.append((2, pcm[2][3 * 9600 + k * 10:3 * 9600 + k * 10 + 5]))   # tiny pushes drain the backlog
        steps.append(step)
    return steps


def test_online_parity_compat(tiny):
    cfg, w = tiny
    assert _run_both(cfg, w, _schedule()) > 5


def test_online_parity_per_layer_cache(tiny):
    cfg, w = tiny
    assert _run_both(cfg, w, _schedule(), per_layer_cache=True) > 5


def test_online_snip_edges(tiny):
    cfg, w = tiny
    cfg2 = synth.tiny()
    cfg2.snip_edges = True                                                   # 58 frames per chunk: windows straddle chunks
    assert _run_both(cfg2, w, _schedule(8)) > 3


def test_online_recognizer_api(tiny, tmp_path):
    cfg, w = tiny
    tokens = tmp_path / "tokens.txt"
    tokens.write_text("\n".join(["<blank>", "<s>", "</s>"] + [f"t{i}" for i in range(3, cfg.vocab)]), encoding="utf-8")
    mvn = tmp_path / "am.mvn"
    from oracle import frontend as F
    mvn.write_text(F.format_am_mvn(*synth.make_cmvn()), encoding="utf-8")
    rec = OnlineRecognizer("", "", "", str(mvn), str(tokens), weights=w, config=cfg)
    s = rec.CreateOnlineStream()
    assert rec.GetResult(s).Text == ""                                       # nothing pushed: no window, no tokens
    pcm = synth.make_pcm(0, 3.0)
    for k in range(5):
        s.AddSamples(pcm[k * 9600:(k + 1) * 9600])
        r = rec.GetResults([s])[0]
    assert len(s.Tokens) > 2 and isinstance(r.Text, str)
    with pytest.raises(TypeError):
        s.AddSamples(None)
    s.Dispose()
    rec.Dispose()
    with pytest.raises(Exception):
        rec.CreateOnlineStream()


def test_online_errors(tiny):
    cfg, w = tiny
    eng = OnlineEngine(cfg, w)
    sid = eng.open_stream()
    from aliparaformerasr_b200 import _lib
    with pytest.raises(_lib.PfError):
        eng.step([sid, sid])                                                 # duplicate ids
    with pytest.raises(_lib.PfError):
        eng.push(sid + 7, np.zeros(10, np.float32))
    eng.close_stream(sid)
    with pytest.raises(_lib.PfError):
        eng.step([sid])
    assert eng.step([]).n_working == 0
    eng.close()
