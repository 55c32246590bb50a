"""What does each kernel family cost INSIDE the pipelined step?  Runs the cfg2 step with a family skipped
(PFASR_DBG_SKIP bit mask, results are garbage) and prints the device time.  Tuning aid.
    PFASR_DBG_SKIP=<mask> python scripts/skip_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
cfg = synth.paraformer_large()
eng = Engine(cfg, synth.make_weights(cfg), devices=[0])
eng.set_cmvn(*synth.make_cmvn())
eng.stage_pcm([synth.make_pcm(i, 10.0) for i in range(32)])
ms = []
for i in range(13):
    eng.run_staged()
    if i >= 3:
        ms.append(eng.timings())
print("skip", os.environ.get("PFASR_DBG_SKIP", "0"), {k: round(float(np.median([m[k] for m in ms])), 3) for k in ms[0]}, "launches", eng.launch_count(), flush=True)
