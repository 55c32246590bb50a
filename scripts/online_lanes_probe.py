"""Streaming path with the same GPU listed several times in `devices`: every entry is an independent device context
(own streams and per-stream state), so the step's streams are split over concurrent sub-steps.  Tuning aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.online import OnlineEngine

cfg = synth.paraformer_large()
w = synth.make_weights(cfg)
steps, nstreams = 12, 128
pcm = [synth.make_pcm(i, 0.6 * (steps + 6)) for i in range(nstreams)]
for devs in ([0], [0, 0], [0, 0, 0], [0, 0, 0, 0]):
    eng = OnlineEngine(cfg, w, devices=devs)
    eng.set_cmvn(*synth.make_cmvn())
    sids = [eng.open_stream() for _ in range(nstreams)]
    wall = []
    for k in range(steps + 4):
        for i, s in enumerate(sids):
            eng.push(s, pcm[i][k * 9600:(k + 1) * 9600])
        t0 = time.perf_counter()
        out = eng.step(sids)
        dt = time.perf_counter() - t0
        if k >= 4:
            wall.append(dt)
    print(f"contexts on GPU 0: {len(devs)}  step (128 streams) {np.median(wall) * 1e3:.3f} ms  -> {nstreams * 0.6 / np.median(wall):.0f} audio-s/s, Lmax {out.max_new}", flush=True)
    eng.close()
