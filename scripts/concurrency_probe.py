"""Is the step latency-bound enough that two independent half-batches on two streams beat one full batch?
Runs K handles (one host thread each) with batch 32/K on the same GPU and reports aggregate audio-s/s. Tuning aid."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine

cfg = synth.paraformer_large()
w = synth.make_weights(cfg)
for K in (1, 2, 4):
    nb = 32 // K
    engs = []
    for k in range(K):
        e = Engine(cfg, w, devices=[0])
        e.set_cmvn(*synth.make_cmvn())
        e.stage_pcm([synth.make_pcm(k * nb + i, 10.0) for i in range(nb)])
        for _ in range(3):
            e.run_staged()
        engs.append(e)
    steps = 20
    bar = threading.Barrier(K + 1)
    def worker(e):
        bar.wait()
        for _ in range(steps):
            e.run_staged()
        bar.wait()
    ths = [threading.Thread(target=worker, args=(e,)) for e in engs]
    for t in ths: t.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in ths: t.join()
    print(f"K={K} handles x batch {nb}: {dt / steps * 1e3:.3f} ms per 32 utterances, {32 * 10.0 * steps / dt:.0f} audio-s/s", flush=True)
    for e in engs: e.close()
