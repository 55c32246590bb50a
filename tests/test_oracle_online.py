"""CPU checks of the streaming oracle (oracle/online.py) against the reference's host logic, restated independently
here from the C# line numbers quoted in each test."""
import math

import numpy as np

from aliparaformerasr_b200 import synth
from oracle import frontend as F, online as O, sanm
from _util import dims_of


def test_online_lfr_rule():
    # OnlineWavFrontend.cs:73-91: t=61 -> 61 % 6 = 1 >= 7-6 -> t_lfr = 10; t=60 -> 0 < 1 -> 60/6 - 1 = 9
    fb = np.arange(61 * 80, dtype=np.float32).reshape(61, 80)
    out = O.online_apply_lfr(fb)
    assert out.shape == (10, 560)
    assert np.array_equal(out[3], fb[18:25].reshape(-1))          # row i = frames [6i, 6i+7), no left padding
    assert O.online_apply_lfr(fb[:60]).shape == (9, 560)
    assert O.online_apply_lfr(fb[:5]).shape == (0, 560)


def test_online_position_encoding_q12():
    # OnlineWavFrontend.cs:163-177: inv_timescale_i = exp(-(i+1) * ln(1e4)/279); [sin(280) | cos(280)]; 1-based running pos
    pe = O.online_position_encoding(10, 560, start_idx=20)
    inc = math.log(10000.0) / 279.0
    for p in (0, 9):
        for i in (0, 100, 279):
            ang = (21 + p) * math.exp(-(i + 1) * inc)
            assert abs(pe[p, i] - math.sin(ang)) < 2e-5
            assert abs(pe[p, 280 + i] - math.cos(ang)) < 2e-5
    # differs from the offline (FunASR) table, whose first timescale is exp(0) = 1
    assert abs(pe[0, 0] - math.sin(21.0)) > 1e-3


def test_dynamic_mask():
    a = np.ones((2, 20), dtype=np.float32)
    m = O.dynamic_mask(a)
    assert m[:, :5].sum() == 0 and m[:, 15:].sum() == 0 and (m[:, 5:15] == 1).all()      # OnlineModel.cs:141-165


def test_host_cif_recurrence_and_carry():
    rng = np.random.default_rng(3)
    h = rng.standard_normal((21, 8)).astype(np.float32)
    a = rng.uniform(0.1, 0.6, size=21).astype(np.float32)
    fired, carry_a, carry_h = O.host_cif(h, a, 1.0)
    assert len(fired) == int(math.floor(float(a.astype(np.float64).sum()) + 1e-6))
    # weights of every frame sum to alpha: sum of fired frames + carry * carry_alpha = sum alpha_t h_t
    total = sum(fired) + carry_a * carry_h
    assert np.allclose(total, (a[:, None] * h).sum(0), atol=1e-4)
    assert 0.0 <= carry_a < 1.0
    # nothing fires below the threshold; carry keeps the weighted mean
    fired2, ca2, ch2 = O.host_cif(h[:3], np.asarray([0.2, 0.3, 0.1], np.float32), 1.0)
    assert not fired2 and abs(ca2 - 0.6) < 1e-6
    assert np.allclose(ch2, (0.2 * h[0] + 0.3 * h[1] + 0.1 * h[2]) / 0.6, atol=1e-5)


def test_stream_chunking_q13():
    shift, scale = synth.make_cmvn()
    s = O.OnlineStreamState(shift, scale)
    # the cache starts as 9600 zeros: the first AddSamples (any length > 0) consumes a chunk of silence
    s.add_samples(np.full(100, 0.1, np.float32))
    assert s.speech.shape[0] == 61 and s.cache_samples.shape[0] == 100      # 60 frames + the repeated first frame
    # exactly one chunk per call even when several are buffered (OnlineStream.cs:102-110)
    s.add_samples(np.zeros(3 * 9600, np.float32))
    assert s.speech.shape[0] == 121 and s.cache_samples.shape[0] == 100 + 2 * 9600
    c = s.get_decode_chunk()
    assert c.shape == (20, 560) and (c[:10] == 0).all() and s.start_idx == 10 and s.speech.shape[0] == 61
    c2 = s.get_decode_chunk()
    assert np.array_equal(c2[:10], c[10:]) and s.start_idx == 20            # feature cache = previous window's new rows
    assert s.get_decode_chunk() is None                                      # 1 frame left: no window


def test_online_forward_tiny_is_consistent():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    dims = dims_of(cfg)
    shift, scale = synth.make_cmvn()
    rec = O.OnlineRecognizerOracle(w, dims, shift, scale)
    streams = [rec.create_stream() for _ in range(2)]
    pcm = [synth.make_pcm(i, 3.0) for i in range(2)]
    total_new = [0, 0]
    for k in range(5):
        streams[0].add_samples(pcm[0][k * 9600:(k + 1) * 9600])
        if k % 2 == 0:
            streams[1].add_samples(pcm[1][k * 4800:(k + 1) * 4800])           # slower producer: skipped on some steps
        new = rec.forward(streams)
        for i in range(2):
            total_new[i] += len(new[i])
    assert len(streams[0].tokens) == 2 + total_new[0] and total_new[0] > 0
    assert all(0 <= t < cfg.vocab for t in streams[0].tokens)
    # Q11: with the reference's stack_states every layer saw the layer-0 cache; the per-layer variant differs
    rec2 = O.OnlineRecognizerOracle(w, dims, shift, scale, compat_layer0_cache=False)
    s2 = rec2.create_stream()
    for k in range(5):
        s2.add_samples(pcm[0][k * 9600:(k + 1) * 9600])
        rec2.forward([s2])
    assert len(s2.tokens) == len(streams[0].tokens)
