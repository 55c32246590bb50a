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
