"""Tiny end-to-end passes of every model family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from aliparaformerasr_b200.online import OnlineEngine

which = sys.argv[1:] or ["paraformer", "sensevoicesmall", "seacoparaformer", "online", "audio", "lanes"]
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
    if m == "audio":                                     # device audio ingestion in front of the fbank kernel (csrc/audio.cu)
        from aliparaformerasr_b200 import _lib, audio
        cfg = synth.tiny()
        eng = Engine(cfg, synth.make_weights(cfg))
        eng.set_cmvn(*synth.make_cmvn())
        rng = np.random.default_rng(0)
        clips = [audio.Audio(rng.integers(-3000, 3000, 2 * 30000).astype(np.int16), _lib.PF_AUDIO_S16, 2, 44100),
                 audio.Audio((0.1 * rng.standard_normal(9000)).astype(np.float32), _lib.PF_AUDIO_F32, 1, 8000),
                 audio.Audio(rng.integers(0, 256, 3 * 20000, dtype=np.uint8), _lib.PF_AUDIO_S24, 1, 16000)]
        out = eng.run_audio(clips)
        print("audio ok", out.tokens.shape, flush=True)
        eng.close()
        continue
    if m == "lanes":                                     # two host threads on two execution lanes sharing the weights
        import threading
        cfg = synth.tiny()
        eng = Engine(cfg, synth.make_weights(cfg), lanes=2)
        eng.set_cmvn(*synth.make_cmvn())
        ths = [threading.Thread(target=lambda k=k: [eng.run_pcm(pcm[k:k + 2]) for _ in range(2)]) for k in range(2)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        print("lanes ok", flush=True)
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
