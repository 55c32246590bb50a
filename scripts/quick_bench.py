"""Times the cfg2 step (32 x 10 s, paraformer-large) for one or more builds of libpfasr in ONE process tree on ONE box:

    python scripts/quick_bench.py [lib.so ...]        # no argument: the product library

Each library runs in its own subprocess (PFASR_LIB); per library: single-lane resident ms/step (median of 30, L2 not
flushed), 3-lane resident ms/step (>= 1 s loop) and the per-family kernel times of one profiled step.  Variants come
from scripts/build_variant.py.  Environment switches (PFASR_*) are passed through, so `env PFASR_GEMM_NO_REDADD=1` etc.
can be compared the same way: give the same library twice with NAME=VALUE prefixes:  "PFASR_GEMM_NO_REDADD=1:lib.so"."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import statistics
    import threading
    import time
    import torch
    from aliparaformerasr_b200 import synth
    from aliparaformerasr_b200.engine import Engine
    cfg = synth.paraformer_large()
    w = synth.make_weights(cfg)
    pcm = [[synth.make_pcm(32 * l + i, 10.0) for i in range(32)] for l in range(3)]
    out = {}
    eng = Engine(cfg, w, lanes=1)
    eng.set_cmvn(*synth.make_cmvn())
    eng.stage_pcm(pcm[0])
    for _ in range(5):
        eng.run_staged()
    st = torch.cuda.ExternalStream(eng.stream_ptr(0))
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        eng.run_staged()
        b.record(st)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    out["lane1_ms"] = statistics.median(ts)
    eng.set_profile(2)
    eng.run_staged()
    k = {}
    for p in eng.profile():
        k[p.get("name", "gemm")] = round(k.get(p.get("name", "gemm"), 0.0) + p["ms"], 3)
    eng.set_profile(1)
    eng.run_staged()
    out["gemm_replay_ms"] = eng.replay_gemms(5)
    out["gemm_tflops"] = eng.gemm_flops() / out["gemm_replay_ms"] / 1e9
    eng.set_profile(0)
    out["kernels"] = k
    eng.close()
    eng = Engine(cfg, w, lanes=3)
    eng.set_cmvn(*synth.make_cmvn())
    n = 90

    def worker(l):
        torch.cuda.set_device(0)
        eng.stage_pcm(pcm[l])
        for _ in range(4):
            eng.run_staged()
        gate.wait()
        for _ in range(n):
            eng.run_staged()

    gate = threading.Barrier(4)
    th = [threading.Thread(target=worker, args=(l,)) for l in range(3)]
    for t in th:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    out["lane3_ms"] = (time.perf_counter() - t0) * 1e3 / (3 * n)
    eng.set_profile(1)
    res = [None] * 3

    def rep(l):
        eng.run_staged()
        fl = eng.gemm_flops()
        gate2.wait()
        res[l] = (fl, eng.replay_gemms(5))

    gate2 = threading.Barrier(3)
    th = [threading.Thread(target=rep, args=(l,)) for l in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    out["gemm_tflops_3lanes_concurrent"] = sum(r[0] for r in res) / max(r[1] for r in res) / 1e9
    eng.close()
    print("RESULT " + json.dumps(out))


def main():
    if os.environ.get("PFASR_QB_CHILD"):
        return child()
    libs = sys.argv[1:] or [""]
    for spec in libs:
        env = dict(os.environ, PFASR_QB_CHILD="1")
        *sets, lib = spec.split(":")
        for kv in sets:
            k, v = kv.split("=", 1)
            env[k] = v
        if lib:
            env["PFASR_LIB"] = os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(spec or "product", line[0][7:] if line else ("FAILED " + r.stderr[-800:]), flush=True)


if __name__ == "__main__":
    main()
