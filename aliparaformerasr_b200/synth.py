"""Synthetic model artefacts (random-init weights of the reference architectures, ``am.mvn``,
16 kHz PCM) for tests and benchmarks.

No model files exist offline (the reference git-clones them from modelscope at run time,
AliParaformerAsr.Examples/Utils/GitHelper.cs:78-165), so parity and throughput are measured on
seeded random weights with the exact shapes of paraformer-large / SenseVoiceSmall.  Parameter names
are the FunASR state-dict names so a real checkpoint converter can emit the same blob.

Recipe = SURVEY.md section 8(d): Linear/conv ~ N(0, 1/fan_in), FSMN taps ~ N(0, 0.1^2), LN gamma ~ 1,
beta ~ 0 (slightly perturbed so both are exercised), CIF head bias chosen so mean alpha ~ 0.25,
output layer scaled x4 to widen argmax margins.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Tuple

import numpy as np
import torch

WEIGHT_SEED = 20260917


@dataclass
class ModelConfig:
    """Flat mirror of the ``asr.yaml`` fields the hot path consumes (Model/ConfEntity.cs:5-43,
    Model/EncoderConfEntity.cs:13-25, Model/DecoderConfEntity.cs:7-16, Model/PredictorConfEntity.cs:13-17,
    Model/FrontendConfEntity.cs:7-15).  Field order matches ``pf_config`` in include/pf_abi.h."""
    model: str = "paraformer"
    input_size: int = 560
    d_model: int = 512
    heads: int = 4
    ffn: int = 2048
    enc_layers: int = 50
    tp_layers: int = 0
    enc_kernel: int = 11
    dec_layers: int = 16
    dec_ffn: int = 2048
    dec_kernel: int = 11
    vocab: int = 8404
    ln_eps: float = 1e-12
    cif_threshold: float = 1.0
    cif_tail: float = 0.45
    smooth_factor: float = 1.0
    noise_threshold: float = 0.0
    # frontend_conf
    fs: int = 16000
    n_mels: int = 80
    lfr_m: int = 7
    lfr_n: int = 6
    snip_edges: bool = False
    dither: float = 0.0          # parsed for completeness; the device front-end never dithers (reference default 1.0 is random)
    use_itn: bool = True
    # SeACo bias decoder (seaco_decoder_conf [EXT]) and NO_BIAS class id; unused by the other models
    seaco_layers: int = 4
    seaco_ffn: int = 1024
    seaco_kernel: int = 21
    nobias_id: int = 8377
    # CifPredictorV3 timestamp branch (present in the -timestamp- and SeACo models)
    timestamps: bool = False
    upsample_times: int = 3
    smooth_factor2: float = 0.25
    noise_threshold2: float = 0.01

    def as_dict(self):
        return asdict(self)


def paraformer_large() -> ModelConfig:
    return ModelConfig()


def sensevoice_small() -> ModelConfig:
    return ModelConfig(model="sensevoicesmall", enc_layers=50, tp_layers=20, dec_layers=0, vocab=25055,
                       ln_eps=1e-5)


def seaco_paraformer() -> ModelConfig:
    return ModelConfig(model="seacoparaformer", timestamps=True)


def tiny(model: str = "paraformer") -> ModelConfig:
    """Few-layer variant with the real per-layer shapes: the oracle finishes in seconds."""
    if model == "seacoparaformer":
        return ModelConfig(model=model, enc_layers=3, dec_layers=2, seaco_layers=2, timestamps=True)
    if model == "sensevoicesmall":
        return ModelConfig(model=model, enc_layers=3, tp_layers=2, dec_layers=0, vocab=25055, ln_eps=1e-5)
    return ModelConfig(model=model, enc_layers=3, dec_layers=2)


def _normal(g: torch.Generator, shape, std: float) -> np.ndarray:
    return (torch.randn(shape, generator=g, dtype=torch.float32) * std).numpy()


def _ln_params(g, n) -> Tuple[np.ndarray, np.ndarray]:
    return (1.0 + _normal(g, (n,), 0.05)).astype(np.float32), _normal(g, (n,), 0.05)


def _enc_layer(w: Dict[str, np.ndarray], g, p: str, in_size: int, cfg: ModelConfig):
    d, f, k = cfg.d_model, cfg.ffn, cfg.enc_kernel
    w[p + ".norm1.weight"], w[p + ".norm1.bias"] = _ln_params(g, in_size)
    w[p + ".self_attn.linear_q_k_v.weight"] = _normal(g, (3 * d, in_size), 1.0 / math.sqrt(in_size))
    w[p + ".self_attn.linear_q_k_v.bias"] = _normal(g, (3 * d,), 0.02)
    w[p + ".self_attn.fsmn_block.weight"] = _normal(g, (d, 1, k), 0.1)
    w[p + ".self_attn.linear_out.weight"] = _normal(g, (d, d), 1.0 / math.sqrt(d))
    w[p + ".self_attn.linear_out.bias"] = _normal(g, (d,), 0.02)
    w[p + ".norm2.weight"], w[p + ".norm2.bias"] = _ln_params(g, d)
    w[p + ".feed_forward.w_1.weight"] = _normal(g, (f, d), 1.0 / math.sqrt(d))
    w[p + ".feed_forward.w_1.bias"] = _normal(g, (f,), 0.02)
    w[p + ".feed_forward.w_2.weight"] = _normal(g, (d, f), 1.0 / math.sqrt(f))
    w[p + ".feed_forward.w_2.bias"] = _normal(g, (d,), 0.02)


def make_weights(cfg: ModelConfig, seed: int = WEIGHT_SEED) -> Dict[str, np.ndarray]:
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, np.ndarray] = {}
    d = cfg.d_model
    _enc_layer(w, g, "encoder.encoders0.0", cfg.input_size, cfg)
    for i in range(cfg.enc_layers - 1):
        _enc_layer(w, g, f"encoder.encoders.{i}", d, cfg)
    w["encoder.after_norm.weight"], w["encoder.after_norm.bias"] = _ln_params(g, d)
    for i in range(cfg.tp_layers):
        _enc_layer(w, g, f"encoder.tp_encoders.{i}", d, cfg)
    if cfg.tp_layers:
        w["encoder.tp_norm.weight"], w["encoder.tp_norm.bias"] = _ln_params(g, d)
    if cfg.model == "sensevoicesmall":
        w["ctc.ctc_lo.weight"] = _normal(g, (cfg.vocab, d), 4.0 / math.sqrt(d))
        w["ctc.ctc_lo.bias"] = _normal(g, (cfg.vocab,), 0.02)
        w["embed.weight"] = _normal(g, (16, cfg.input_size), 1.0)
        return w
    # CifPredictorV2
    w["predictor.cif_conv1d.weight"] = _normal(g, (d, d, 3), 1.0 / math.sqrt(3 * d))
    w["predictor.cif_conv1d.bias"] = _normal(g, (d,), 0.02)
    w["predictor.cif_output.weight"] = _normal(g, (1, d), 1.0 / math.sqrt(d))
    w["predictor.cif_output.bias"] = np.asarray([-1.1], dtype=np.float32)
    # ParaformerSANMDecoder
    _sanm_decoder(w, g, "decoder", cfg.dec_layers, cfg.dec_ffn, cfg.dec_kernel, d)
    pending_v3 = cfg.timestamps
    w["decoder.output_layer.weight"] = _normal(g, (cfg.vocab, d), 4.0 / math.sqrt(d))
    w["decoder.output_layer.bias"] = _normal(g, (cfg.vocab,), 0.02)
    if cfg.model == "seacoparaformer":
        # bias decoder (no output layer), hot-word head, hot-word encoder (Embedding + 2-layer LSTM)
        _sanm_decoder(w, g, "seaco_decoder", cfg.seaco_layers, cfg.seaco_ffn, cfg.seaco_kernel, d)
        w["hotword_output_layer.weight"] = _normal(g, (cfg.vocab, d), 4.0 / math.sqrt(d))
        hb = _normal(g, (cfg.vocab,), 0.02)
        hb[cfg.nobias_id] = SEACO_NOBIAS_BIAS        # so that a good share of the rows keeps the ASR posterior
        w["hotword_output_layer.bias"] = hb
        w["bias_embed.weight"] = _normal(g, (cfg.vocab, d), 1.0)
        for layer in range(2):
            w[f"bias_encoder.weight_ih_l{layer}"] = _normal(g, (4 * d, d), 1.0 / math.sqrt(d))
            w[f"bias_encoder.weight_hh_l{layer}"] = _normal(g, (4 * d, d), 1.0 / math.sqrt(d))
            w[f"bias_encoder.bias_ih_l{layer}"] = _normal(g, (4 * d,), 0.05)
            w[f"bias_encoder.bias_hh_l{layer}"] = _normal(g, (4 * d,), 0.05)
    if pending_v3:
        # CifPredictorV3 upsampler (drawn last so the other tensors keep their values with or without it)
        w["predictor.upsample_cnn.weight"] = _normal(g, (d, d, cfg.upsample_times), 1.0 / math.sqrt(d))
        w["predictor.upsample_cnn.bias"] = _normal(g, (d,), 0.02)
        for sfx in ("", "_reverse"):
            w["predictor.blstm.weight_ih_l0" + sfx] = _normal(g, (4 * d, d), 1.0 / math.sqrt(d))
            w["predictor.blstm.weight_hh_l0" + sfx] = _normal(g, (4 * d, d), 1.0 / math.sqrt(d))
            w["predictor.blstm.bias_ih_l0" + sfx] = _normal(g, (4 * d,), 0.05)
            w["predictor.blstm.bias_hh_l0" + sfx] = _normal(g, (4 * d,), 0.05)
        w["predictor.cif_output2.weight"] = _normal(g, (1, 2 * d), 2.0 / math.sqrt(2 * d))
        w["predictor.cif_output2.bias"] = np.asarray([0.3], dtype=np.float32)
    return w


SEACO_NOBIAS_BIAS = 21.0


def _sanm_decoder(w: Dict[str, np.ndarray], g, prefix: str, layers: int, f: int, k: int, d: int):
    for i in range(layers):
        p = f"{prefix}.decoders.{i}"
        w[p + ".norm1.weight"], w[p + ".norm1.bias"] = _ln_params(g, d)
        w[p + ".feed_forward.w_1.weight"] = _normal(g, (f, d), 1.0 / math.sqrt(d))
        w[p + ".feed_forward.w_1.bias"] = _normal(g, (f,), 0.02)
        w[p + ".feed_forward.norm.weight"], w[p + ".feed_forward.norm.bias"] = _ln_params(g, f)
        w[p + ".feed_forward.w_2.weight"] = _normal(g, (d, f), 1.0 / math.sqrt(f))
        w[p + ".norm2.weight"], w[p + ".norm2.bias"] = _ln_params(g, d)
        w[p + ".self_attn.fsmn_block.weight"] = _normal(g, (d, 1, k), 0.1)
        w[p + ".norm3.weight"], w[p + ".norm3.bias"] = _ln_params(g, d)
        w[p + ".src_attn.linear_q.weight"] = _normal(g, (d, d), 1.0 / math.sqrt(d))
        w[p + ".src_attn.linear_q.bias"] = _normal(g, (d,), 0.02)
        w[p + ".src_attn.linear_k_v.weight"] = _normal(g, (2 * d, d), 1.0 / math.sqrt(d))
        w[p + ".src_attn.linear_k_v.bias"] = _normal(g, (2 * d,), 0.02)
        w[p + ".src_attn.linear_out.weight"] = _normal(g, (d, d), 1.0 / math.sqrt(d))
        w[p + ".src_attn.linear_out.bias"] = _normal(g, (d,), 0.02)
    p = f"{prefix}.decoders3.0"
    w[p + ".norm1.weight"], w[p + ".norm1.bias"] = _ln_params(g, d)
    w[p + ".feed_forward.w_1.weight"] = _normal(g, (f, d), 1.0 / math.sqrt(d))
    w[p + ".feed_forward.w_1.bias"] = _normal(g, (f,), 0.02)
    w[p + ".feed_forward.norm.weight"], w[p + ".feed_forward.norm.bias"] = _ln_params(g, f)
    w[p + ".feed_forward.w_2.weight"] = _normal(g, (d, f), 1.0 / math.sqrt(f))
    w[f"{prefix}.after_norm.weight"], w[f"{prefix}.after_norm.bias"] = _ln_params(g, d)


def make_hotwords(n: int = 200, vocab: int = 8404, seed: int = 7):
    """``n`` id lists of length U{2..6}, ids U{3..vocab-1}, plus the trailing ``[sos]`` entry the
    reference appends (OfflineRecognizer.cs:87)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ln = int(torch.randint(2, 7, (1,), generator=g).item())
        out.append([int(t) for t in torch.randint(3, vocab, (ln,), generator=g)])
    out.append([1])
    return out


def make_cmvn(dim: int = 560) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic ``am.mvn`` content: AddShift = -8.0, Rescale = 0.25 for every dim."""
    return np.full(dim, -8.0, dtype=np.float32), np.full(dim, 0.25, dtype=np.float32)


def make_pcm(index: int, seconds: float, fs: int = 16000) -> np.ndarray:
    """Utterance ``index``: six amplitude-modulated sinusoids + Gaussian noise, clipped to [-1, 1]
    (SURVEY 8d).  Deterministic per global utterance index, so shards see the same audio as 1 GPU."""
    g = torch.Generator().manual_seed(1000 + index)
    n = int(round(seconds * fs))
    t = torch.arange(n, dtype=torch.float64) / fs
    x = torch.zeros(n, dtype=torch.float64)
    for _ in range(6):
        f = 80.0 + (3800.0 - 80.0) * torch.rand(1, generator=g).item()
        a = 0.02 + 0.18 * torch.rand(1, generator=g).item()
        fm = 2.0 + 4.0 * torch.rand(1, generator=g).item()
        ph = 2 * math.pi * torch.rand(1, generator=g).item()
        x += a * torch.sin(2 * math.pi * f * t + ph) * (0.6 + 0.4 * torch.sin(2 * math.pi * fm * t))
    x += 0.01 * torch.randn(n, generator=g, dtype=torch.float64)
    return x.clamp_(-1.0, 1.0).to(torch.float32).numpy()
