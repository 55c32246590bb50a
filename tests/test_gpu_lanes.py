"""Execution lanes (pf_offline_create_mt): calls from different host threads run concurrently on one GPU and must return
exactly what the single-lane handle returns for the same batch."""
import threading

import numpy as np
import pytest

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_lanes_match_single_lane_results():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    batches = [[synth.make_pcm(10 * k + i, 1.0 + 0.3 * ((i + k) % 4)) for i in range(5)] for k in range(3)]
    one = Engine(cfg, w)
    one.set_cmvn(*synth.make_cmvn())
    ref = [one.run_pcm(b, want_logits=True) for b in batches]
    one.close()
    eng = Engine(cfg, w, lanes=3)
    eng.set_cmvn(*synth.make_cmvn())                     # configuration reaches every lane
    assert eng._lib.pf_offline_lanes(eng._handle()) == 3
    got = [None] * 3
    errs = []

    def worker(k):
        try:
            for _ in range(6):                            # repeated, interleaved with the other threads' batches
                o = eng.run_pcm(batches[k], want_logits=True)
                assert np.array_equal(o.tokens, ref[k].tokens)
            got[k] = o
        except Exception as ex:                           # noqa: BLE001
            errs.append(ex)

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for k in range(3):
        assert np.array_equal(got[k].token_num, ref[k].token_num)
        assert np.array_equal(got[k].logits, ref[k].logits)          # same kernels, same order: bit-identical
    # more threads than lanes share lanes and serialise on them
    extra = []
    ths = [threading.Thread(target=lambda k=k: extra.append((k, eng.run_pcm(batches[k % 3]).tokens))) for k in range(5)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert len(extra) == 5 and all(np.array_equal(tok, ref[k % 3].tokens) for k, tok in extra)
    eng.close()


def test_lane_local_hotwords_do_not_leak_between_threads():
    """Per-call hot words (OfflineProjOfSeacoParaformer.cs:51-60) set through pf_offline_set_hotwords_local only touch the
    calling thread's lane: a thread biased with hot words and a thread without them keep their own results."""
    cfg = synth.tiny("seacoparaformer")
    w = synth.make_weights(cfg)
    pcm = [synth.make_pcm(i, 1.5) for i in range(3)]
    hot = synth.make_hotwords(4, cfg.vocab)
    one = Engine(cfg, w)
    one.set_cmvn(*synth.make_cmvn())
    plain = one.run_pcm(pcm, want_logits=True)
    one.set_hotwords(hot)
    biased = one.run_pcm(pcm, want_logits=True)
    one.close()
    assert not np.array_equal(plain.logits, biased.logits)
    eng = Engine(cfg, w, lanes=2)
    eng.set_cmvn(*synth.make_cmvn())
    res = {}
    start = threading.Barrier(2)

    def worker(name, hotwords):
        eng.set_hotwords(hotwords, local=True)
        start.wait()
        for _ in range(4):
            res[name] = eng.run_pcm(pcm, want_logits=True)

    ths = [threading.Thread(target=worker, args=("plain", [])), threading.Thread(target=worker, args=("biased", hot))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert np.array_equal(res["plain"].logits, plain.logits)
    assert np.array_equal(res["biased"].logits, biased.logits)
    eng.close()


def test_gemm_replay_measures_the_profiled_step():
    cfg = synth.tiny()
    eng = Engine(cfg, synth.make_weights(cfg))
    eng.set_cmvn(*synth.make_cmvn())
    pcm = [synth.make_pcm(i, 2.0) for i in range(4)]
    ref = eng.run_pcm(pcm)
    assert eng.replay_gemms(2) == 0.0                      # nothing recorded without a profiled run
    eng.set_profile(1)
    eng.run_pcm(pcm)
    n = sum(p["launches"] for p in eng.profile() if p.get("name", "gemm") == "gemm")
    ms = eng.replay_gemms(3)
    assert n > 0 and 0.0 < ms < 50.0
    eng.set_profile(0)
    assert np.array_equal(eng.run_pcm(pcm).tokens, ref.tokens)   # the replay only scribbles over activations
    eng.close()
