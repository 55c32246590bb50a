"""Two (or more) full batches in flight on one GPU: K handles, one host thread each, every handle runs the whole
32 x 10 s batch.  Kernels of the two streams interleave at CTA granularity, so one batch's kernel tails / launch gaps
are filled by the other's work.  Reports aggregate audio-s/s, resident inputs and end to end (host PCM).  Tuning aid."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine

cfg = synth.paraformer_large()
w = synth.make_weights(cfg)
NB, SECONDS, STEPS = 32, 10.0, 20
for K in (1, 2, 3):
    engs, pcms = [], []
    for k in range(K):
        e = Engine(cfg, w, devices=[0])
        e.set_cmvn(*synth.make_cmvn())
        pcm = [synth.make_pcm(k * NB + i, SECONDS) for i in range(NB)]
        e.stage_pcm(pcm)
        for _ in range(3):
            e.run_staged()
        engs.append(e)
        pcms.append(pcm)
    for mode in ("resident", "e2e"):
        bar = threading.Barrier(K + 1)
        def worker(e, pcm):
            bar.wait()
            for _ in range(STEPS):
                if mode == "resident":
                    e.run_staged()
                else:
                    e.run_pcm(pcm)
            bar.wait()
        ths = [threading.Thread(target=worker, args=(e, p)) for e, p in zip(engs, pcms)]
        for t in ths: t.start()
        bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
        for t in ths: t.join()
        print(f"K={K} {mode:8s}: {dt / (STEPS * K) * 1e3:.3f} ms per batch of {NB}, {NB * SECONDS * STEPS * K / dt:.0f} audio-s/s", flush=True)
    for e in engs: e.close()
