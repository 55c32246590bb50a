"""Per-family kernel times of cfg4 (SeACo 16 x 10 s + 200 hot words) with the timestamp outputs requested (profile mode: an event pair
around every launch group).      python scripts/ts_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
cfg = synth.seaco_paraformer()
eng = Engine(cfg, synth.make_weights(cfg))
eng.set_cmvn(*synth.make_cmvn())
eng.set_hotwords(synth.make_hotwords(200, cfg.vocab))
pcm = [synth.make_pcm(i, 10.0) for i in range(16)]
TS = "--no-ts" not in sys.argv
for _ in range(3):
    eng.run_pcm(pcm, want_timestamps=TS)
eng.set_profile(2)
eng.run_pcm(pcm, want_timestamps=TS)
k = {}
for p in eng.profile():
    k[p.get("name", "gemm")] = round(k.get(p.get("name", "gemm"), 0.0) + p["ms"], 3)
print(k, eng.timings())
