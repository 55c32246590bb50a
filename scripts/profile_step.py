"""Minimal driver for ncu: paraformer-large, batch 32 x 10 s, N steps of the resident-input hot path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = synth.paraformer_large()
eng = Engine(cfg, synth.make_weights(cfg), devices=[0])
eng.set_cmvn(*synth.make_cmvn())
eng.stage_pcm([synth.make_pcm(i, 10.0) for i in range(batch)])
for _ in range(steps):
    out = eng.run_staged()
print("launches/step", eng.launch_count(), "Lmax", out.tokens.shape[1], eng.timings())
