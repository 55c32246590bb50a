"""Shard equivalence: one handle spread over two GPUs returns exactly what one GPU returns - offline batch
split (Lmax exchanged between the device threads) and streaming with streams pinned to stream_id % ndev."""
import numpy as np
import pytest
import torch

from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from aliparaformerasr_b200.online import OnlineEngine

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def test_offline_two_devices_match_one():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    pcm = [synth.make_pcm(i, 2.0 + 0.5 * (i % 3)) for i in range(5)]        # ragged lengths, odd batch
    outs = []
    for devs in ([0], [0, 1]):
        eng = Engine(cfg, w, devices=devs)
        eng.set_cmvn(*synth.make_cmvn())
        outs.append(eng.run_pcm(pcm, want_logits=True))
        eng.close()
    a, b = outs
    assert np.array_equal(a.token_num, b.token_num)
    assert a.tokens.shape == b.tokens.shape and np.array_equal(a.tokens, b.tokens)
    assert np.allclose(a.logits, b.logits, atol=1e-5)


def test_online_two_devices_match_one():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    pcm = [synth.make_pcm(i, 3.0) for i in range(3)]
    hist = []
    for devs in ([0], [0, 1]):
        eng = OnlineEngine(cfg, w, devices=devs)
        eng.set_cmvn(*synth.make_cmvn())
        sids = [eng.open_stream() for _ in range(3)]
        steps = []
        for k in range(5):
            for i, s in enumerate(sids):
                if i != 1 or k % 2 == 0:
                    eng.push(s, pcm[i][k * 9600:(k + 1) * 9600])
            o = eng.step(sids)
            steps.append((o.max_new, o.appended.copy(), o.new_tokens.copy(), o.embeds_len.copy()))
        eng.close()
        hist.append(steps)
    for s1, s2 in zip(*hist):
        assert s1[0] == s2[0]
        for x, y in zip(s1[1:], s2[1:]):
            assert np.array_equal(x, y)
