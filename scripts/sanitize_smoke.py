"""Tiny end-to-end passes of every model family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from aliparaformerasr_b200.online import OnlineEngine

which = sys.argv[1:] or ["paraformer", "sensevoicesmall", "seacoparaformer", "online"]
pcm = [synth.make_pcm(i, 2.0 + 0.7 * i) for i in range(3)]
for m in which:
    if m == "online":
        cfg = synth.tiny()
        eng = OnlineEngine(cfg, synth.make_weights(cfg))
        eng.set_cmvn(*synth.make_cmvn())
        sids = [eng.open_stream() for _ in range(2)]
        for k in range(3):
            for i, s in enumerate(sids):
                eng.push(s, pcm[i][k * 9600:(k + 1) * 9600])
            o = eng.step(sids)
        print("online ok", o.max_new, flush=True)
        eng.close()
        continue
    cfg = synth.tiny(m)
    eng = Engine(cfg, synth.make_weights(cfg))
    eng.set_cmvn(*synth.make_cmvn())
    if m == "seacoparaformer":
        eng.set_hotwords(synth.make_hotwords(5, cfg.vocab))
    out = eng.run_pcm(pcm, want_logits=True)
    print(m, "ok", out.tokens.shape, flush=True)
    eng.close()
