"""Host enqueue time vs device time of one step (pf_offline_get_timings [6..9]): is the launch path (CUDA driver lock
shared by the lanes' host threads) a bottleneck?  Tuning aid."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine

cfg = synth.paraformer_large()
w = synth.make_weights(cfg)
for L in (1, 3):
    eng = Engine(cfg, w, devices=[0], lanes=L)
    eng.set_cmvn(*synth.make_cmvn())
    rows = [[] for _ in range(L)]
    bar = threading.Barrier(L)

    def worker(l):
        eng.stage_pcm([synth.make_pcm(l * 32 + i, 10.0) for i in range(32)])
        for _ in range(3):
            eng.run_staged()
        bar.wait()
        for _ in range(12):
            eng.run_staged()
            rows[l].append(eng.timings())

    ths = [threading.Thread(target=worker, args=(l,)) for l in range(L)]
    for t in ths: t.start()
    for t in ths: t.join()
    med = {k: round(float(np.median([r[k] for rr in rows for r in rr])), 3) for k in rows[0][0]}
    print(f"lanes={L}", med, flush=True)
    eng.close()
