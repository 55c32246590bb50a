"""Full-depth parity measurements: the CUDA path (through the C-ABI) against the CPU oracle at the sizes of the five
BASELINE configs (paraformer-large 50+16, SenseVoiceSmall 50+20, SeACo 50+16+4 with 200 hot words, streaming 50+16).

Every comparison is made twice:

  * against the FLOAT32 oracle (the reference arithmetic: OfflineProjOfParaformer.cs:39-87,
    OfflineProjOfSenseVoiceSmall.cs:53-175, OfflineProjOfSeacoParaformer.cs:48-135, OnlineRecognizer.cs:341-401), and
  * against the same oracle with tensor-core OPERAND ROUNDING switched on (oracle.sanm.OperandRounding: weights and GEMM /
    attention inputs rounded to fp16, everything else float32).  The second comparison isolates implementation error
    from the quantisation the north star allows ("logits within 1e-2 fp16"); the distance between the two oracles is the
    error any fp16-operand implementation carries.

``python tests/_parity.py`` (on a GPU box) writes gpurun_out/parity_r02.json, which scripts/parity_table.py turns into
profiles/parity_r02.md.  tests/test_gpu_fulldepth.py asserts on the same numbers.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from aliparaformerasr_b200 import synth                      # noqa: E402
from aliparaformerasr_b200.engine import Engine              # noqa: E402
from oracle import frontend as F, sanm                       # noqa: E402
from _util import dims_of, margins                           # noqa: E402

TOKEN_MARGIN = 0.1
TAP_LAYERS = (1, 10, 25, 50)


def err(got, ref):
    d = np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64))
    return {"max": float(d.max()) if d.size else 0.0, "rms": float(np.sqrt((d ** 2).mean())) if d.size else 0.0}


def tail_mask(fires, shape):
    m = np.zeros(shape, bool)
    for b, n in enumerate(fires):
        if n > 0:
            m[b, n - 1] = True
    return m


def logit_stats(got, ref_logits, ref_tokens, got_tokens, fires=None):
    out = {"logp": err(got, ref_logits)}
    d = np.abs(got.astype(np.float64) - ref_logits)
    if fires is not None:
        tm = tail_mask(fires, d.shape[:2])
        out["logp_tail_rows_max"] = float(d[tm].max())
        out["logp_other_rows_max"] = float(d[~tm].max())
    out["frac_over_1e-2"] = float((d > 1e-2).mean())
    m = margins(ref_logits)
    safe = m > TOKEN_MARGIN
    out["rows"] = int(safe.size)
    out["rows_safe"] = int(safe.sum())
    out["token_mismatch_safe"] = int((got_tokens[safe] != ref_tokens[safe]).sum())
    out["token_mismatch_all"] = int((got_tokens != ref_tokens).sum())
    mm = got_tokens != ref_tokens
    out["largest_margin_of_a_mismatch"] = float(m[mm].max()) if mm.any() else 0.0
    return out


def oracle_feats(pcm, cfg):
    shift, scale = synth.make_cmvn()
    return F.pad_sequence([F.extract_features(p, shift, scale, snip_edges=cfg.snip_edges) for p in pcm])


# ---------------------------------------------------------------------------------------------- paraformer-large
def paraformer_fulldepth(n_utts=4, seconds=10.0, first=0):
    cfg = synth.paraformer_large()
    w = synth.make_weights(cfg)
    dims = dims_of(cfg)
    pcm = [synth.make_pcm(first + i, seconds) for i in range(n_utts)]
    speech = oracle_feats(pcm, cfg)
    refs = {}
    for name, ctx in (("fp32", None), ("fp16_operands", sanm.OperandRounding()), ("fp16_weights_only", sanm.OperandRounding(activations=False))):
        col = {"tap_layers": TAP_LAYERS}
        t0 = time.time()
        if ctx is None:
            refs[name] = sanm.paraformer_forward(speech, w, dims, col)
        else:
            with ctx:
                refs[name] = sanm.paraformer_forward(speech, w, dims, col)
        refs[name]["col"] = col
        refs[name]["seconds"] = time.time() - t0
    eng = Engine(cfg, w)
    eng.set_cmvn(*synth.make_cmvn())
    # the network is compared on IDENTICAL input features (pf_offline_run_feats, the IOfflineProj.ModelProj contract:
    # speech [B,T,560] in, log-probs out); the PCM entry (own fbank kernel, <= 2e-3 on the log-mel) is reported beside it
    out_pcm = eng.run_pcm(pcm, want_logits=True)
    out = eng.run_feats(speech, want_logits=True)
    got = {"enc": eng.tensor("enc"), "alphas": eng.tensor("alphas"), "acoustic_embeds": eng.tensor("acoustic_embeds"),
           f"enc_after_{cfg.enc_layers}": eng.tensor("x")}
    eng.close()
    # residual stream after 1 / 10 / 25 layers: the same weights driven through a shallower encoder
    for n in TAP_LAYERS:
        if n >= cfg.enc_layers:
            continue
        c2 = synth.paraformer_large()
        c2.enc_layers = n
        e2 = Engine(c2, w)
        e2.set_cmvn(*synth.make_cmvn())
        e2.run_feats(speech)
        got[f"enc_after_{n}"] = e2.tensor("x")
        e2.close()
    res = {"config": f"paraformer-large 50+16, {n_utts} x {seconds:g} s (BASELINE configs[1] shape, utterances {first}..{first + n_utts - 1})",
           "token_num_cuda": out.token_num.tolist(), "oracle_seconds": refs["fp32"]["seconds"]}
    for name, ref in refs.items():
        r = {"token_num_equal": bool(np.array_equal(out.token_num, ref["token_num"])), "stages": {}}
        for n in TAP_LAYERS:
            r["stages"][f"enc_after_{n}"] = err(got[f"enc_after_{n}"], ref["col"][f"enc_after_{n}"])
        r["stages"]["enc (after_norm)"] = err(got["enc"], ref["enc"])
        r["stages"]["alphas"] = err(got["alphas"], ref["alphas"])
        if r["token_num_equal"] and ref["logits"].shape == out.logits.shape:
            r["stages"]["acoustic_embeds"] = err(got["acoustic_embeds"], ref["acoustic_embeds"])
            r.update(logit_stats(out.logits, ref["logits"], ref["tokens"], out.tokens, ref["fires"]))
        if np.array_equal(out_pcm.token_num, ref["token_num"]) and ref["logits"].shape == out_pcm.logits.shape:
            r["from_pcm"] = logit_stats(out_pcm.logits, ref["logits"], ref["tokens"], out_pcm.tokens, ref["fires"])
        res["vs_" + name] = r
    res["fp16_operands_vs_fp32"] = logit_stats(refs["fp16_operands"]["logits"], refs["fp32"]["logits"], refs["fp32"]["tokens"],
                                               refs["fp16_operands"]["tokens"], refs["fp32"]["fires"]) \
        if refs["fp16_operands"]["logits"].shape == refs["fp32"]["logits"].shape else None
    res["fp16_weights_only_vs_fp32"] = logit_stats(refs["fp16_weights_only"]["logits"], refs["fp32"]["logits"], refs["fp32"]["tokens"],
                                                   refs["fp16_weights_only"]["tokens"], refs["fp32"]["fires"]) \
        if refs["fp16_weights_only"]["logits"].shape == refs["fp32"]["logits"].shape else None
    return res


# ---------------------------------------------------------------------------------------------- SenseVoiceSmall
def sensevoice_fulldepth(n_utts=4, seconds=8.0, first=100):
    cfg = synth.sensevoice_small()
    w = synth.make_weights(cfg)
    dims = dims_of(cfg)
    pcm = [synth.make_pcm(first + i, seconds) for i in range(n_utts)]
    shift, scale = synth.make_cmvn()
    feats = [F.extract_features(p, shift, scale) for p in pcm]
    speech = np.stack([sanm.sensevoice_prepend(x, w["embed.weight"], cfg.use_itn) for x in feats])
    refs = {"fp32": sanm.sensevoice_forward(speech, w, dims)}
    with sanm.OperandRounding():
        refs["fp16_operands"] = sanm.sensevoice_forward(speech, w, dims)
    eng = Engine(cfg, w)
    eng.set_cmvn(shift, scale)
    out_pcm = eng.run_pcm(pcm, want_logits=True)
    out = eng.run_feats(speech, want_logits=True)          # prompted rows prepended by the caller, as the C# does (Q7)
    enc = eng.tensor("enc")
    eng.close()
    res = {"config": f"SenseVoiceSmall 50+20, {n_utts} x {seconds:g} s, use_itn (BASELINE configs[2] shape)", "frames": int(out.tokens.shape[1])}
    for name, ref in refs.items():
        r = {"shape_equal": out.logits.shape == ref["logits"].shape, "stages": {"enc (tp_norm)": err(enc, ref["enc"])}}
        r.update(logit_stats(out.logits, ref["logits"], ref["tokens"], out.tokens))
        r["from_pcm"] = logit_stats(out_pcm.logits, ref["logits"], ref["tokens"], out_pcm.tokens)
        res["vs_" + name] = r
    res["fp16_operands_vs_fp32"] = logit_stats(refs["fp16_operands"]["logits"], refs["fp32"]["logits"], refs["fp32"]["tokens"], refs["fp16_operands"]["tokens"])
    return res


# ---------------------------------------------------------------------------------------------- SeACo
def seaco_fulldepth(n_utts=2, seconds=10.0, nhot=200, first=200, timestamps=True):
    cfg = synth.seaco_paraformer()
    w = synth.make_weights(cfg)
    dims = dims_of(cfg)
    pcm = [synth.make_pcm(first + i, seconds) for i in range(n_utts)]
    speech = oracle_feats(pcm, cfg)
    hot = synth.make_hotwords(nhot, cfg.vocab)
    refs = {}
    for name, ctx in (("fp32", None), ("fp16_operands", sanm.OperandRounding())):
        if ctx is None:
            rows = sanm.bias_embed_rows(sanm.hotword_embed(sanm.pad_hotwords(hot), w))
            refs[name] = sanm.seaco_forward(speech, w, dims, rows)
        else:
            with ctx:
                rows = sanm.bias_embed_rows(sanm.hotword_embed(sanm.pad_hotwords(hot), w))
                refs[name] = sanm.seaco_forward(speech, w, dims, rows)
    eng = Engine(cfg, w)
    eng.set_cmvn(*synth.make_cmvn())
    eng.set_hotwords(hot)
    out = eng.run_feats(speech, want_logits=True, want_timestamps=timestamps)
    enc_cuda = eng.tensor("enc")
    eng.close()
    res = {"config": f"SeACo-paraformer 50+16+4, {n_utts} x {seconds:g} s, {nhot} hot words = {rows.shape[0]} bias rows (BASELINE configs[3] shape)",
           "token_num_cuda": out.token_num.tolist()}
    for name, ref in refs.items():
        r = {"token_num_equal": bool(np.array_equal(out.token_num, ref["token_num"]))}
        if r["token_num_equal"] and ref["logits"].shape == out.logits.shape:
            # a row's branch (ASR vs hot-word posterior) is decided by the hot-word argmax: rows whose decision is within
            # rounding distance may legitimately take the other branch
            ds = np.sort(ref["dha"], axis=-1)
            nob = ref["dha"][..., cfg.nobias_id]
            top_other = np.where(ref["dha_ids"] == cfg.nobias_id, ds[..., -2], ds[..., -1])
            decided = np.abs(nob - top_other) > 0.25
            r["rows_branch_decided"] = int(decided.sum())
            r["rows_kept_asr"] = int((ref["dha_ids"] == cfg.nobias_id).sum())
            st = logit_stats(out.logits[decided][None], ref["logits"][decided][None], ref["tokens"][decided][None], out.tokens[decided][None])
            r.update(st)
        res["vs_" + name] = r
    if timestamps:
        from aliparaformerasr_b200.offline import time_stamp_lfr6_onnx

        def ts_stats(ua, pk, tokens):
            ts = {"us_alphas": err(out.us_alphas, ua), "us_cif_peak": err(out.us_cif_peak, pk), "fire_index_max_abs_diff": 0, "fire_count_equal": True,
                  "timestamp_ms_max_abs_diff": 0, "timestamps_identical": True, "fires": 0}
            for i in range(n_utts):
                fr = np.nonzero(pk[i] > 1 - 1e-4)[0]
                fg = np.nonzero(out.us_cif_peak[i] > 1 - 1e-4)[0]
                ts["fires"] += len(fr)
                if len(fr) != len(fg):
                    ts["fire_count_equal"] = False
                    ts["timestamps_identical"] = False
                    continue
                ts["fire_index_max_abs_diff"] = max(ts["fire_index_max_abs_diff"], int(np.abs(fr - fg).max()) if len(fr) else 0)
                a = time_stamp_lfr6_onnx(out.us_cif_peak[i], list(tokens[i]))
                b = sanm.time_stamp_lfr6_onnx(pk[i], list(tokens[i]))
                if a != b:
                    ts["timestamps_identical"] = False
                if len(a) == len(b) and a:
                    ts["timestamp_ms_max_abs_diff"] = max(ts["timestamp_ms_max_abs_diff"], int(np.abs(np.asarray(a) - np.asarray(b)).max()))
            return ts

        ref = refs["fp32"]
        ua, pk = sanm.upsample_timestamp(ref["enc"], ref["token_num"], w, dims)
        res["timestamps_vs_fp32"] = ts_stats(ua, pk, ref["tokens"])
        # row f1 in isolation: the oracle's CifPredictorV3 upsampler driven by the CUDA path's OWN encoder output, so
        # the encoder's fp16 noise is common to both sides and any difference is the timestamp kernels'
        ua2, pk2 = sanm.upsample_timestamp(enc_cuda, out.token_num, w, dims)
        res["timestamps_vs_oracle_on_cuda_enc"] = ts_stats(ua2, pk2, out.tokens)
    return res


# ---------------------------------------------------------------------------------------------- streaming
def online_fulldepth(n_streams=4, n_steps=5):
    from aliparaformerasr_b200.online import OnlineEngine
    from oracle import online as O
    cfg = synth.paraformer_large()
    w = synth.make_weights(cfg)
    shift, scale = synth.make_cmvn()
    res = {"config": f"paraformer-large streaming 50+16, {n_streams} streams x {n_steps} steps of 600 ms (BASELINE configs[4] shape)"}
    pcm = [synth.make_pcm(300 + i, 0.6 * (n_steps + 1)) for i in range(n_streams)]
    for name, ctx in (("fp32", None), ("fp16_operands", sanm.OperandRounding())):
        eng = OnlineEngine(cfg, w)
        eng.set_cmvn(shift, scale)
        rec = O.OnlineRecognizerOracle(w, dims_of(cfg), shift, scale, snip_edges=cfg.snip_edges)
        sids = [eng.open_stream() for _ in range(n_streams)]
        rs = [rec.create_stream() for _ in range(n_streams)]
        r = {"appended_equal": True, "logits": {"max": 0.0, "rms_sq_sum": 0.0, "n": 0}, "rows_safe": 0, "token_mismatch_safe": 0,
             "cif_alpha_max": 0.0, "fsmn_max": 0.0, "cache_feats_max": 0.0, "steps_with_tokens": 0}
        for k in range(n_steps):
            for i in range(n_streams):
                x = pcm[i][k * 9600:(k + 1) * 9600]
                eng.push(sids[i], x)
                rs[i].add_samples(x)
            col = {}
            if ctx is None:
                ref_new = rec.forward(rs, collect=col)
            else:
                with ctx:
                    ref_new = rec.forward(rs, collect=col)
            out = eng.step(sids, want_logits=True)
            lens_ref = [len(t) for t in ref_new]
            if list(out.appended) != lens_ref:
                r["appended_equal"] = False
                break
            if out.max_new > 0:
                r["steps_with_tokens"] += 1
                widx = [i for i in range(n_streams) if lens_ref[i] > 0]
                for b, i in enumerate(widx):
                    d = np.abs(out.logits[i].astype(np.float64) - col["logits"][b])
                    r["logits"]["max"] = max(r["logits"]["max"], float(d.max()))
                    r["logits"]["rms_sq_sum"] += float((d ** 2).sum())
                    r["logits"]["n"] += d.size
                    safe = margins(col["logits"][b]) > TOKEN_MARGIN
                    r["rows_safe"] += int(safe.sum())
                    r["token_mismatch_safe"] += int((out.new_tokens[i][safe] != np.asarray(ref_new[i])[safe]).sum())
            for i in range(n_streams):
                r["cache_feats_max"] = max(r["cache_feats_max"], float(np.abs(eng.state(sids[i], "cache_feats") - rs[i].cache_feats).max()))
                r["cif_alpha_max"] = max(r["cif_alpha_max"], abs(float(eng.state(sids[i], "cif_alpha")[0]) - float(rs[i].cif_alpha[0])))
                fs = eng.state(sids[i], "fsmn")
                ref_fs = np.stack([c.T for c in rs[i].states])
                r["fsmn_max"] = max(r["fsmn_max"], float(np.abs(fs - ref_fs).max()))
        n = max(r["logits"]["n"], 1)
        r["logits"] = {"max": r["logits"]["max"], "rms": float(np.sqrt(r["logits"]["rms_sq_sum"] / n))}
        eng.close()
        res["vs_" + name] = r
    return res


def main():
    out = {}
    for name, fn in (("paraformer", paraformer_fulldepth), ("sensevoice", sensevoice_fulldepth), ("seaco", seaco_fulldepth), ("online", online_fulldepth)):
        t0 = time.time()
        out[name] = fn()
        out[name]["wall_s"] = time.time() - t0
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_r02.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
