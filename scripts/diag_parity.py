"""Per-stage error of the CUDA path vs the fp32 oracle (and vs an oracle with fp16-rounded weights)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aliparaformerasr_b200 import synth
from aliparaformerasr_b200.engine import Engine
from oracle import frontend as F, sanm

def dims_of(cfg):
    return sanm.ModelDims(**{k: v for k, v in cfg.as_dict().items() if k in sanm.ModelDims.__dataclass_fields__})

def run(cfg, B, secs, tag):
    w = synth.make_weights(cfg)
    eng = Engine(cfg, w); eng.set_cmvn(*synth.make_cmvn())
    pcm = [synth.make_pcm(i, secs) for i in range(B)]
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(p, shift, scale) for p in pcm])
    t = time.time(); ref = sanm.paraformer_forward(speech, w, dims_of(cfg)); t_cpu = time.time() - t
    w16 = {k: (v.astype(np.float16).astype(np.float32) if v.ndim >= 2 and 'fsmn' not in k and 'cif_output' not in k else v) for k, v in w.items()}
    ref16 = sanm.paraformer_forward(speech, w16, dims_of(cfg))
    out = eng.run_pcm(pcm, want_logits=True)
    enc = eng.tensor("enc"); al = eng.tensor("alphas")
    def e(a, b): return float(np.abs(a - b).max()), float(np.sqrt(np.mean((a - b) ** 2)))
    print(f"== {tag}: B={B} T={speech.shape[1]} L={out.tokens.shape[1]} cpu {t_cpu:.2f}s timings {eng.timings()} launches {eng.launch_count()}")
    print("  token_num", out.token_num.tolist(), ref["token_num"].tolist())
    print("  enc    max/rms err vs fp32 oracle", e(enc, ref["enc"]), " vs w16 oracle", e(enc, ref16["enc"]), " w16-vs-fp32", e(ref16["enc"], ref["enc"]))
    print("  alphas", e(al, ref["alphas"]), e(al, ref16["alphas"]))
    if out.logits.shape == ref["logits"].shape:
        print("  logits", e(out.logits, ref["logits"]), " vs w16", e(out.logits, ref16["logits"]), " w16-vs-fp32", e(ref16["logits"], ref["logits"]))
        s = np.sort(ref["logits"], -1); m = s[..., -1] - s[..., -2]
        mism = out.tokens != ref["tokens"]
        print("  tokens mismatching", int(mism.sum()), "of", mism.size, " margins at mismatches", np.round(m[mism], 4).tolist()[:10], " median margin", float(np.median(m)))
        print("  distinct tokens", len(np.unique(ref["tokens"])))
    eng.close()

run(synth.tiny(), 3, 5.0, "tiny")
run(synth.paraformer_large(), 2, 5.0, "paraformer-large")
