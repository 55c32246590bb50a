"""Device-timed throughput of the BASELINE.json configs other than the headline one (bench.py measures configs[1]):
cfg1 paraformer-large 1 x 5 s, cfg3 SenseVoiceSmall 64 x 8 s, cfg4 SeACo 16 x 10 s + 200 hot words, cfg5 streaming
(16 and 128 concurrent streams on one GPU, 600 ms chunks).  Synthetic weights / audio.  Prints one JSON
line per config; numbers go to profiles/ as context, they are not bench.py values.
    python scripts/bench_configs.py [cfg1 cfg3 cfg4 cfg5] [--steps K] [--lanes L]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aliparaformerasr_b200 import synth  # noqa: E402
from aliparaformerasr_b200.engine import Engine  # noqa: E402
from aliparaformerasr_b200.online import OnlineEngine  # noqa: E402


def offline_lanes(name, cfg, batch, seconds, steps, lanes, hotwords=None, timestamps=False):
    """The same config with `lanes` batches in flight (one host thread per execution lane): wall-clock throughput of
    the public call, host PCM in, token ids out."""
    import threading
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w, devices=[0], lanes=lanes)
    eng.set_cmvn(*synth.make_cmvn())
    if hotwords is not None:
        eng.set_hotwords(hotwords)
    pcms = [[synth.make_pcm(l * batch + i, seconds) for i in range(batch)] for l in range(lanes)]
    bar = threading.Barrier(lanes + 1)

    def worker(l):
        for _ in range(3):
            eng.run_pcm(pcms[l], want_timestamps=timestamps)
        bar.wait()
        for _ in range(steps):
            eng.run_pcm(pcms[l], want_timestamps=timestamps)
        bar.wait()

    ths = [threading.Thread(target=worker, args=(l,)) for l in range(lanes)]
    for t in ths:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for t in ths:
        t.join()
    print(json.dumps({"config": name, "lanes": lanes, "batch": batch, "seconds": seconds, "e2e_ms_per_step": dt / (steps * lanes) * 1e3,
                      "e2e_audio_s_per_s": batch * seconds * steps * lanes / dt}), flush=True)
    eng.close()


def offline(name, cfg, batch, seconds, steps, hotwords=None, timestamps=False):
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w, devices=[0])
    eng.set_cmvn(*synth.make_cmvn())
    if hotwords is not None:
        eng.set_hotwords(hotwords)
    pcm = [synth.make_pcm(i, seconds) for i in range(batch)]
    eng.stage_pcm(pcm)
    run = (lambda: eng.run_staged(want_timestamps=True)) if timestamps else eng.run_staged      # both with the PCM already resident
    for _ in range(3):
        out = run()
    ms = []
    for _ in range(steps):
        run()
        ms.append(eng.timings()["total"])
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.run_pcm(pcm, want_timestamps=timestamps)
    e2e = (time.perf_counter() - t0) / steps
    dev = float(np.median(ms))
    print(json.dumps({"config": name, "batch": batch, "seconds": seconds, "T": int(out.feat_frames), "Lmax": int(out.tokens.shape[1]),
                      "device_ms_per_step": dev, "audio_s_per_s": batch * seconds / (dev / 1e3), "e2e_ms_per_step": e2e * 1e3,
                      "e2e_audio_s_per_s": batch * seconds / e2e, "launches": eng.launch_count(), "gemm_tflop_per_step": eng.gemm_flops() / 1e12,
                      "stage_ms": eng.timings()}), flush=True)
    eng.close()


def online(name, nstreams, steps):
    cfg = synth.paraformer_large()
    w = synth.make_weights(cfg)
    eng = OnlineEngine(cfg, w, devices=[0])
    eng.set_cmvn(*synth.make_cmvn())
    sids = [eng.open_stream() for _ in range(nstreams)]
    pcm = [synth.make_pcm(i, 0.6 * (steps + 6)) for i in range(nstreams)]
    dev, wall, lmaxs = [], [], []
    for k in range(steps + 4):
        t0 = time.perf_counter()
        for i, s in enumerate(sids):
            eng.push(s, pcm[i][k * 9600:(k + 1) * 9600])
        out = eng.step(sids)
        dt = time.perf_counter() - t0
        if k >= 4:
            dev.append(eng.timings()["total"])
            wall.append(dt)
            lmaxs.append(out.max_new)
    d = float(np.median(dev))
    wl = float(np.median(wall))
    print(json.dumps({"config": name, "streams": nstreams, "chunk_seconds": 0.6, "device_ms_per_step": d, "e2e_ms_per_step": wl * 1e3,
                      "audio_s_per_s": nstreams * 0.6 / (d / 1e3), "e2e_audio_s_per_s": nstreams * 0.6 / wl, "median_Lmax": float(np.median(lmaxs)),
                      "launches": eng.launch_count(), "gemm_tflop_per_step": eng.gemm_flops() / 1e12, "stage_ms": eng.timings()}), flush=True)
    eng.close()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = 10
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
        args = [a for a in args if a != str(steps)]
    lanes = 0
    if "--lanes" in sys.argv:
        lanes = int(sys.argv[sys.argv.index("--lanes") + 1])
        args = [a for a in args if a != str(lanes)]
    which = args or ["cfg1", "cfg3", "cfg4", "cfg5"]
    if lanes > 1:                       # several batches in flight: end-to-end throughput only
        if "cfg1" in which:
            offline_lanes("cfg1 paraformer-large 1x5s", synth.paraformer_large(), 1, 5.0, steps, lanes)
        if "cfg3" in which:
            offline_lanes("cfg3 sensevoice-small 64x8s", synth.sensevoice_small(), 64, 8.0, steps, lanes)
        if "cfg4" in which:
            cfg = synth.seaco_paraformer()
            offline_lanes("cfg4 seaco-paraformer 16x10s + 200 hotwords", cfg, 16, 10.0, steps, lanes, hotwords=synth.make_hotwords(200, cfg.vocab))
        return
    if "cfg1" in which:
        offline("cfg1 paraformer-large 1x5s", synth.paraformer_large(), 1, 5.0, steps)
    if "cfg2" in which:
        offline("cfg2 paraformer-large 32x10s", synth.paraformer_large(), 32, 10.0, steps)
    if "cfg3" in which:
        offline("cfg3 sensevoice-small 64x8s", synth.sensevoice_small(), 64, 8.0, steps)
    if "cfg4" in which:
        cfg = synth.seaco_paraformer()
        offline("cfg4 seaco-paraformer 16x10s + 200 hotwords", cfg, 16, 10.0, steps, hotwords=synth.make_hotwords(200, cfg.vocab))
        offline("cfg4 + timestamps (us_alphas / us_cif_peak, 498-step BiLSTM)", cfg, 16, 10.0, steps, hotwords=synth.make_hotwords(200, cfg.vocab),
                timestamps=True)
    if "cfg5" in which:
        online("cfg5 streaming 16 streams/GPU (128 over 8 GPUs)", 16, steps)
        online("cfg5 streaming 128 streams on one GPU", 128, steps)


if __name__ == "__main__":
    main()
