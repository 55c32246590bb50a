"""Where does the distance between an fp16-operand implementation and the float32 graph come from?  (CPU only.)

Runs the float32 oracle at full depth against the same oracle with tensor-core operand rounding switched on site by
site (oracle.sanm.OperandRounding) and prints max / rms of the log-prob difference, the alpha difference and the
encoder difference.  Output feeds profiles/parity_r02.md.

    python scripts/precision_study.py [--utts 2] [--seconds 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aliparaformerasr_b200 import synth                      # noqa: E402
from oracle import frontend as F, sanm                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--first", type=int, default=0)
    a = ap.parse_args()
    cfg = synth.paraformer_large()
    w = synth.make_weights(cfg)
    dims = sanm.ModelDims(**{k: v for k, v in cfg.as_dict().items() if k in sanm.ModelDims.__dataclass_fields__})
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(synth.make_pcm(a.first + i, a.seconds), shift, scale) for i in range(a.utts)])
    t0 = time.time()
    ref = sanm.paraformer_forward(speech, w, dims)
    print(f"fp32 oracle: {time.time() - t0:.1f} s, token_num {ref['token_num'].tolist()}", file=sys.stderr)
    variants = {
        "weights only": dict(weights=True, activations=False),
        "activations only": dict(weights=False, activations=True),
        "all operands": dict(),
        "all but pred.conv": dict(skip=("pred.conv",)),
        "all but pred.conv+head": dict(skip=("pred.conv", "head")),
        "all but encoder": dict(skip=("enc.qkv", "enc.att", "enc.out", "enc.ffn1", "enc.ffn2")),
        "encoder only": dict(skip=("pred.conv", "dec.kv", "dec.ffn1", "dec.ffn2", "dec.q", "dec.att", "dec.out", "head")),
    }
    rows = []
    for name, kw in variants.items():
        with sanm.OperandRounding(**kw):
            got = sanm.paraformer_forward(speech, w, dims)
        same_n = bool(np.array_equal(got["token_num"], ref["token_num"]) and got["logits"].shape == ref["logits"].shape)
        row = {"variant": name, "token_num_equal": same_n}
        for key in ("enc", "alphas"):
            d = np.abs(got[key] - ref[key])
            row[key + "_max"] = float(d.max())
            row[key + "_rms"] = float(np.sqrt((d.astype(np.float64) ** 2).mean()))
        if same_n:
            d = np.abs(got["logits"] - ref["logits"])
            row["logp_max"] = float(d.max())
            row["logp_rms"] = float(np.sqrt((d.astype(np.float64) ** 2).mean()))
            # tail token (last emitted row of each utterance) vs the rest
            tail = np.zeros(d.shape[:2], bool)
            for b, n in enumerate(ref["fires"]):
                tail[b, n - 1] = True
            row["logp_max_tail_rows"] = float(d[tail].max())
            row["logp_max_other_rows"] = float(d[~tail].max())
            s = np.sort(ref["logits"], axis=-1)
            margin = s[..., -1] - s[..., -2]
            row["token_mismatch"] = int((got["tokens"] != ref["tokens"]).sum())
            row["token_mismatch_margin_max"] = float(margin[got["tokens"] != ref["tokens"]].max()) if row["token_mismatch"] else 0.0
        rows.append(row)
        print(json.dumps(row))
    return rows


if __name__ == "__main__":
    main()
