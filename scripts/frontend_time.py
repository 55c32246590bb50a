import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
cfg = synth.tiny(); cfg.enc_layers = 1; cfg.dec_layers = 1
eng = Engine(cfg, synth.make_weights(cfg)); eng.set_cmvn(*synth.make_cmvn())
pcm = [synth.make_pcm(i, 10.0) for i in range(32)]
eng.stage_pcm(pcm)
ts = []
for _ in range(20):
    eng.run_staged(); ts.append(eng.timings()["h2d_frontend"])
print("frontend stage ms (median of 20, includes table copies + pad fill):", sorted(ts)[10])
