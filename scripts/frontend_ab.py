"""A/B of two libpfasr builds on the front-end kernel: features of a ragged batch must be bit-identical, and the front-end stage time of
32 x 10 s is printed for each.     python scripts/frontend_ab.py libA.so libB.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out):
    import numpy as np
    from aliparaformerasr_b200 import synth
    from aliparaformerasr_b200.engine import Engine
    cfg = synth.tiny(); cfg.enc_layers = 1; cfg.dec_layers = 1
    eng = Engine(cfg, synth.make_weights(cfg)); eng.set_cmvn(*synth.make_cmvn())
    feats = [eng.extract(synth.make_pcm(100 + i, s)) for i, s in enumerate((0.4, 1.0, 2.35, 5.0, 10.0))]
    pcm = [synth.make_pcm(i, 10.0) for i in range(32)]
    eng.stage_pcm(pcm)
    ts = []
    for _ in range(30):
        eng.run_staged(); ts.append(eng.timings()["h2d_frontend"])
    np.savez(out, *feats)
    print("frontend stage ms (median of 30):", sorted(ts)[15])


if __name__ == "__main__":
    if os.environ.get("PFASR_AB_CHILD"):
        child(os.environ["PFASR_AB_CHILD"])
        sys.exit(0)
    import numpy as np
    outs = []
    for i, lib in enumerate(sys.argv[1:3]):
        out = f"/tmp/frontend_ab_{i}.npz"
        env = dict(os.environ, PFASR_AB_CHILD=out)
        if lib:
            env["PFASR_LIB"] = os.path.join(ROOT, lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        print(lib or "product", r.stdout.strip(), r.stderr[-300:] if r.returncode else "")
        outs.append(np.load(out))
    same = all(np.array_equal(outs[0][k], outs[1][k]) for k in outs[0].files)
    worst = max(float(np.abs(outs[0][k] - outs[1][k]).max()) for k in outs[0].files)
    print("bit-identical:", same, "max |diff|:", worst)
