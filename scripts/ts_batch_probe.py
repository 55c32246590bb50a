import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
cfg = synth.seaco_paraformer()
eng = Engine(cfg, synth.make_weights(cfg))
eng.set_cmvn(*synth.make_cmvn())
eng.set_hotwords(synth.make_hotwords(200, cfg.vocab))
for B in (16, 32):
    pcm = [synth.make_pcm(i, 10.0) for i in range(B)]
    eng.stage_pcm(pcm)
    for ts in (False, True):
        for _ in range(3): eng.run_staged(want_timestamps=ts)
        ms = []
        for _ in range(7):
            eng.run_staged(want_timestamps=ts); ms.append(eng.timings()["total"])
        print("B", B, "timestamps", ts, "device ms", round(sorted(ms)[3], 3), flush=True)
