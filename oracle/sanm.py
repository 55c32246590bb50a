"""CPU restatement (torch, float32) of the network graph the reference executes through
OnnxRuntime for the offline path (test infrastructure; parity UNPINNED, see oracle/__init__.py).

Call sites in the reference that define the tensor contract:
  * ``OfflineProjOfParaformer.ModelProj``       OfflineProjOfParaformer.cs:39-87
      inputs  ``speech [B,T,560] f32``, ``speech_lengths [B] i32`` (= T for every item, Q3)
      outputs ``logits [B,L,V] f32`` (log-softmax), ``token_num [B] i32``
  * ``OfflineProjOfSenseVoiceSmall.ModelProj``  OfflineProjOfSenseVoiceSmall.cs:53-175
      prompt rows from ``data/embed.onnx`` prepended (Q6, Q7), CTC log-softmax over T+4 frames
  * greedy pick ``OfflineRecognizer.Forward``   OfflineRecognizer.cs:139-152 (Q5: last max wins)

Graph semantics = the FunASR ONNX export of paraformer-large / SenseVoiceSmall
(SURVEY.md section 2.5): SANMEncoder (encoders0 + encoders + after_norm [+ tp_encoders + tp_norm]),
CifPredictorV2 (+ tail 0.45, threshold 1.0), ParaformerSANMDecoder (16 + 1 layers).

Weights are a ``dict[str, np.ndarray]`` keyed by FunASR state-dict names, float32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as Fn


@dataclass
class ModelDims:
    """Subset of ``asr.yaml`` the graph depends on (Model/EncoderConfEntity.cs:13-25,
    Model/DecoderConfEntity.cs:7-16, Model/PredictorConfEntity.cs:13-17)."""
    model: str = "paraformer"          # "paraformer" | "sensevoicesmall"
    input_size: int = 560
    d_model: int = 512
    heads: int = 4
    ffn: int = 2048
    enc_layers: int = 50               # encoders0 (1) + encoders (49)
    tp_layers: int = 0                 # SenseVoice tp_encoders
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
    # SeACo bias decoder (seaco_decoder_conf of the seaco-paraformer asr.yaml [EXT]) and its NO_BIAS class id
    seaco_layers: int = 4
    seaco_ffn: int = 1024
    seaco_kernel: int = 21
    nobias_id: int = 8377
    # CifPredictorV3 timestamp branch (predictor_conf of the -timestamp- / seaco models [EXT])
    upsample_times: int = 3
    smooth_factor2: float = 0.25
    noise_threshold2: float = 0.01


def _t(w: Dict[str, np.ndarray], name: str) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w[name]))


def _ln(x: torch.Tensor, w, name: str, eps: float) -> torch.Tensor:
    return Fn.layer_norm(x, (x.shape[-1],), _t(w, name + ".weight"), _t(w, name + ".bias"), eps)


# --------------------------------------------------------------------------- operand-rounding model (test aid)
class OperandRounding:
    """Where a half-precision tensor-core implementation rounds: GEMM / attention OPERANDS go to IEEE fp16 (weights and
    activations), everything else (accumulation, bias, residual stream, LayerNorm, softmax, CIF) stays float32.  Used as
    a context manager it turns this float32 oracle into the "same arithmetic, fp16 operands" model that
    tests/test_gpu_fulldepth.py uses to PROVE where the distance between the CUDA path and the float32 graph comes from:

        with OperandRounding(weights=True, activations=False): ...   # the fp32 graph merely GIVEN fp16-rounded weights
        with OperandRounding(): ...                                  # every tensor-core operand rounded

    ``skip`` names sites that stay float32 (e.g. {"pred.conv", "head"}); sites: enc.qkv enc.att enc.out enc.ffn1
    enc.ffn2 pred.conv dec.kv dec.ffn1 dec.ffn2 dec.q dec.att dec.out head hw.lstm (SeACo stacks use the dec.* names)."""
    current: "Optional[OperandRounding]" = None

    def __init__(self, weights: bool = True, activations: bool = True, skip=()):
        self.weights, self.activations, self.skip = weights, activations, frozenset(skip)

    def __enter__(self):
        self._prev = OperandRounding.current
        OperandRounding.current = self
        return self

    def __exit__(self, *exc):
        OperandRounding.current = self._prev
        return False


def _h(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(torch.float32)


def _ra(x: torch.Tensor, site: str) -> torch.Tensor:
    """activation operand of ``site``"""
    r = OperandRounding.current
    return _h(x) if r is not None and r.activations and site not in r.skip else x


def _rw(x: torch.Tensor, site: str) -> torch.Tensor:
    """weight operand of ``site``"""
    r = OperandRounding.current
    return _h(x) if r is not None and r.weights and site not in r.skip else x


def _lin(x: torch.Tensor, w, wname: str, bname: Optional[str], site: str) -> torch.Tensor:
    return Fn.linear(_ra(x, site), _rw(_t(w, wname), site), _t(w, bname) if bname else None)


def sinusoidal_pe(t: int, depth: int, start: int = 1) -> torch.Tensor:
    """FunASR ``SinusoidalPositionEncoder``: positions start..start+t-1, [sin | cos] halves,
    ``inv_timescale_i = exp(-i * ln(1e4) / (depth/2 - 1))``."""
    pos = torch.arange(start, start + t, dtype=torch.float32)
    half = depth // 2
    inc = math.log(10000.0) / (half - 1)
    inv = torch.exp(torch.arange(half, dtype=torch.float32) * (-inc))
    ang = pos[:, None] * inv[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def _fsmn(v: torch.Tensor, weight: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Depthwise memory block: ``conv1d(pad(v*mask)) + v*mask`` then ``*mask``; kernel k, pad (k-1)/2
    both sides (sanm_shfit = 0), no bias.  v: [B,T,D]; weight: [D,1,k]; mask: [B,T,1] or None."""
    if mask is not None:
        v = v * mask
    k = weight.shape[-1]
    left = (k - 1) // 2
    x = Fn.pad(v.transpose(1, 2), (left, k - 1 - left))
    x = Fn.conv1d(x, weight, None, groups=weight.shape[0]).transpose(1, 2)
    x = x + v
    if mask is not None:
        x = x * mask
    return x


def _mha(q, k, v, heads: int, site: str = "att") -> torch.Tensor:
    """softmax(q k^T / sqrt(d_k)) v over [B,Tq,D] x [B,Tk,D]; masks are all-ones (Q3)."""
    b, tq, d = q.shape
    tk = k.shape[1]
    dk = d // heads
    r = OperandRounding.current
    if r is not None and r.activations and site not in r.skip:
        # tensor-core form: q, k, v are fp16 operands; the scale is applied to the fp32 scores; the un-normalised
        # probabilities exp(s - max) are the fp16 operand of the second product and the row sum divides its fp32 result
        qh = _h(q).view(b, tq, heads, dk).transpose(1, 2)
        kh = _h(k).view(b, tk, heads, dk).transpose(1, 2)
        vh = _h(v).view(b, tk, heads, dk).transpose(1, 2)
        sc = (qh @ kh.transpose(-2, -1)) * (dk ** -0.5)
        e = torch.exp(sc - sc.max(dim=-1, keepdim=True).values)
        return ((_h(e) @ vh) / e.sum(dim=-1, keepdim=True)).transpose(1, 2).reshape(b, tq, d)
    qh = q.view(b, tq, heads, dk).transpose(1, 2) * (dk ** -0.5)
    kh = k.view(b, tk, heads, dk).transpose(1, 2)
    vh = v.view(b, tk, heads, dk).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-2, -1), dim=-1)
    return (att @ vh).transpose(1, 2).reshape(b, tq, d)


def encoder_layer(x: torch.Tensor, w, p: str, dims: ModelDims) -> torch.Tensor:
    """EncoderLayerSANM (export form): pre-LN SAN-M attention (+FSMN memory) then FFN."""
    in_size = x.shape[-1]
    d = dims.d_model
    h = _ln(x, w, p + ".norm1", dims.ln_eps)
    qkv = _lin(h, w, p + ".self_attn.linear_q_k_v.weight", p + ".self_attn.linear_q_k_v.bias", "enc.qkv")
    q, k, v = torch.split(qkv, d, dim=-1)
    mem = _fsmn(_ra(v, "enc.att"), _t(w, p + ".self_attn.fsmn_block.weight"), None)      # the memory reads the stored V
    ctx = _mha(q, k, v, dims.heads, "enc.att")
    att = _lin(ctx, w, p + ".self_attn.linear_out.weight", p + ".self_attn.linear_out.bias", "enc.out") + mem
    x = att + x if in_size == d else att
    h = _ln(x, w, p + ".norm2", dims.ln_eps)
    h = torch.relu(_lin(h, w, p + ".feed_forward.w_1.weight", p + ".feed_forward.w_1.bias", "enc.ffn1"))
    h = _lin(h, w, p + ".feed_forward.w_2.weight", p + ".feed_forward.w_2.bias", "enc.ffn2")
    return x + h


def encoder(speech: torch.Tensor, w, dims: ModelDims, collect: Optional[dict] = None) -> torch.Tensor:
    """SANMEncoder: ``speech [B,T,560]`` -> ``[B,T,512]``."""
    b, t, din = speech.shape
    x = speech * (dims.d_model ** 0.5) + sinusoidal_pe(t, din)[None]
    x = encoder_layer(x, w, "encoder.encoders0.0", dims)
    if collect is not None:
        collect["enc_layer0"] = x.clone()
        collect["enc_after_1"] = x.numpy().copy()
    for i in range(dims.enc_layers - 1):
        x = encoder_layer(x, w, f"encoder.encoders.{i}", dims)
        if collect is not None and (i + 2) in collect.get("tap_layers", ()):
            collect[f"enc_after_{i + 2}"] = x.numpy().copy()          # residual stream after i+2 layers
    x = _ln(x, w, "encoder.after_norm", dims.ln_eps)
    if dims.tp_layers:
        for i in range(dims.tp_layers):
            x = encoder_layer(x, w, f"encoder.tp_encoders.{i}", dims)
        x = _ln(x, w, "encoder.tp_norm", dims.ln_eps)
    return x


def predictor_alphas(enc: torch.Tensor, w, dims: ModelDims) -> torch.Tensor:
    """CifPredictorV2 alpha head + tail: returns alphas [B, T+1] (last column = tail_threshold)."""
    ctx = Fn.pad(_ra(enc, "pred.conv").transpose(1, 2), (1, 1))
    out = torch.relu(Fn.conv1d(ctx, _rw(_t(w, "predictor.cif_conv1d.weight"), "pred.conv"), _t(w, "predictor.cif_conv1d.bias")))
    out = Fn.linear(out.transpose(1, 2), _t(w, "predictor.cif_output.weight"), _t(w, "predictor.cif_output.bias"))
    alphas = torch.sigmoid(out).squeeze(-1)
    alphas = torch.relu(alphas * dims.smooth_factor - dims.noise_threshold)
    # mask is all ones (speech_lengths == T for every item, Q3); tail: one extra step of 0.45 at t = T
    tail = torch.full((enc.shape[0], 1), dims.cif_tail, dtype=alphas.dtype)
    return torch.cat([alphas, tail], dim=1)


def cif(hidden: np.ndarray, alphas: np.ndarray, threshold: float = 1.0):
    """Continuous integrate-and-fire, the export recurrence (same rule as the reference's host
    version ``OnlineRecognizer.cs:149-200``, Q15), float32 sequential arithmetic.

    hidden [B,T1,D] (T1 = T+1, last row zeros), alphas [B,T1].
    Returns (acoustic_embeds [B,Lmax,D], token_num [B] int32 = floor(sum alpha), fires [B] counts,
             cif_peak [B,T1] = integrate value before reset at each step).
    """
    b, t1, d = hidden.shape
    thr = np.float32(threshold)
    one = np.float32(1.0)
    frames: List[List[np.ndarray]] = []
    peaks = np.zeros((b, t1), dtype=np.float32)
    token_num = np.zeros(b, dtype=np.int32)
    for bi in range(b):
        integrate = np.float32(0.0)
        frame = np.zeros(d, dtype=np.float32)
        total = np.float32(0.0)
        out: List[np.ndarray] = []
        for t in range(t1):
            a = np.float32(alphas[bi, t])
            total = np.float32(total + a)
            completion = np.float32(one - integrate)
            integrate = np.float32(integrate + a)
            peaks[bi, t] = integrate
            fire = integrate >= thr
            cur = completion if fire else a
            frame = frame + cur * hidden[bi, t]
            if fire:
                integrate = np.float32(integrate - one)
                out.append(frame)
                frame = np.float32(a - cur) * hidden[bi, t]
        frames.append(out)
        token_num[bi] = int(math.floor(float(total)))
    fires = np.array([len(f) for f in frames], dtype=np.int32)
    lmax = int(fires.max()) if b else 0
    emb = np.zeros((b, lmax, d), dtype=np.float32)
    for bi, out in enumerate(frames):
        if out:
            emb[bi, : len(out)] = np.stack(out)
    return emb, token_num, fires, peaks


def _dec_ffn(x: torch.Tensor, w, p: str, dims: ModelDims) -> torch.Tensor:
    """PositionwiseFeedForwardDecoderSANM: w_2(LN(relu(w_1 x))), w_2 has no bias."""
    h = torch.relu(_lin(x, w, p + ".w_1.weight", p + ".w_1.bias", "dec.ffn1"))
    h = _ln(h, w, p + ".norm", dims.ln_eps)
    return _lin(h, w, p + ".w_2.weight", None, "dec.ffn2")


def decoder(enc: torch.Tensor, embeds: torch.Tensor, token_num: torch.Tensor, w, dims: ModelDims,
            collect: Optional[dict] = None, prefix: str = "decoder", layers: Optional[int] = None,
            return_hidden: bool = False):
    """ParaformerSANMDecoder (export form) -> logits [B,L,V] before log-softmax (or, with ``return_hidden``, the
    after_norm output [B,L,512] that the SeACo branch consumes; ``prefix="seaco_decoder"`` runs the bias decoder, which
    has no output layer)."""
    b, l, d = embeds.shape
    tgt_mask = (torch.arange(l)[None, :] < token_num[:, None]).to(embeds.dtype)[:, :, None]
    x = embeds
    for i in range(dims.dec_layers if layers is None else layers):
        p = f"{prefix}.decoders.{i}"
        t = _dec_ffn(_ln(x, w, p + ".norm1", dims.ln_eps), w, p + ".feed_forward", dims)
        tn = _ln(t, w, p + ".norm2", dims.ln_eps)
        x = x + _fsmn(tn, _t(w, p + ".self_attn.fsmn_block.weight"), tgt_mask)
        h = _ln(x, w, p + ".norm3", dims.ln_eps)
        q = _lin(h, w, p + ".src_attn.linear_q.weight", p + ".src_attn.linear_q.bias", "dec.q")
        kv = _lin(enc, w, p + ".src_attn.linear_k_v.weight", p + ".src_attn.linear_k_v.bias", "dec.kv")
        k, v = torch.split(kv, d, dim=-1)
        ctx = _mha(q, k, v, dims.heads, "dec.att")
        x = x + _lin(ctx, w, p + ".src_attn.linear_out.weight", p + ".src_attn.linear_out.bias", "dec.out")
        if collect is not None and i == 0:
            collect["dec_layer0"] = x.clone()
    p = f"{prefix}.decoders3.0"
    x = _dec_ffn(_ln(x, w, p + ".norm1", dims.ln_eps), w, p + ".feed_forward", dims)
    x = _ln(x, w, f"{prefix}.after_norm", dims.ln_eps)
    if prefix != "decoder":
        return x
    logits = _lin(x, w, "decoder.output_layer.weight", "decoder.output_layer.bias", "head")
    return (logits, x) if return_hidden else logits


def greedy_pick(logp: np.ndarray) -> np.ndarray:
    """OfflineRecognizer.Forward argmax (OfflineRecognizer.cs:139-152): ``best = x[best] > x[k] ? best : k``
    for k = 1..V-1, i.e. ties (and NaN) resolve to the LARGEST index (Q5).  [.., V] -> [..] int32."""
    v = logp.shape[-1]
    flat = logp.reshape(-1, v)
    mx = flat.max(axis=1, keepdims=True)
    # last index attaining the maximum
    rev = np.argmax((flat == mx)[:, ::-1], axis=1)
    best = (v - 1 - rev).astype(np.int32)
    # NaN rows: comparison is always false -> walks to the last NaN-or-later index; emulate exactly
    nan_rows = np.isnan(flat).any(axis=1)
    for r in np.nonzero(nan_rows)[0]:
        cur = 0
        row = flat[r]
        for k in range(1, v):
            cur = cur if row[cur] > row[k] else k
        best[r] = cur
    return best.reshape(logp.shape[:-1])


def paraformer_forward(speech: np.ndarray, w, dims: ModelDims, collect: Optional[dict] = None):
    """The whole ``InferenceSession.Run`` of OfflineProjOfParaformer.cs:68 on ``speech [B,T,560]``.

    Returns dict(logits [B,L,V] log-probs, token_num [B], tokens [B,L] greedy ids, enc, alphas, ...)."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(speech, dtype=np.float32))
        enc = encoder(x, w, dims, collect)
        alphas = predictor_alphas(enc, w, dims)
        hidden = torch.cat([enc, torch.zeros(enc.shape[0], 1, enc.shape[2])], dim=1)
        emb, token_num, fires, peaks = cif(hidden.numpy(), alphas.numpy(), dims.cif_threshold)
        logits = decoder(enc, torch.from_numpy(emb), torch.from_numpy(token_num.astype(np.int64)), w, dims, collect)
        logp = torch.log_softmax(logits, dim=-1).numpy()
    return {
        "logits": logp,
        "token_num": token_num,
        "tokens": greedy_pick(logp),
        "enc": enc.numpy(),
        "alphas": alphas.numpy(),
        "acoustic_embeds": emb,
        "fires": fires,
        "cif_peak": peaks,
    }


# --------------------------------------------------------------------------- SenseVoice

SENSEVOICE_LID = {"auto": 0, "zh": 3, "en": 4, "yue": 7, "ja": 11, "ko": 12, "nospeech": 13}
SENSEVOICE_TEXTNORM = {"withitn": 14, "woitn": 15}


def sensevoice_prompt_ids(use_itn: bool) -> Tuple[int, int]:
    """OfflineProjOfSenseVoiceSmall.cs:57-74 as written (Q6): ``languageId`` is first set from the
    hard-coded ``"ja"`` (11) and then OVERWRITTEN by the textnorm lookup (``out languageId``), while
    ``textnormId`` keeps its initial value 15.  -> language = 14 (use_itn) or 15; textnorm = 15."""
    language_id = SENSEVOICE_LID["ja"]
    textnorm_id = 15
    language_id = SENSEVOICE_TEXTNORM["withitn" if use_itn else "woitn"]
    return language_id, textnorm_id


def sensevoice_prepend(feats: np.ndarray, table: np.ndarray, use_itn: bool) -> np.ndarray:
    """Split-embed layout per utterance (OfflineProjOfSenseVoiceSmall.cs:84-100, Q7):
    ``[embed(language), embed(1), embed(2), embed(textnorm), speech...]`` -> [T+4, 560]."""
    lang, tn = sensevoice_prompt_ids(use_itn)
    rows = table[[lang, 1, 2, tn]].astype(np.float32)
    return np.concatenate([rows, feats.astype(np.float32)], axis=0)


def sensevoice_forward(speech: np.ndarray, w, dims: ModelDims):
    """SenseVoiceSmall graph on already-prompted ``speech [B,T+4,560]``: encoder (+tp) -> CTC head ->
    log-softmax; the reference does NO CTC collapse (OfflineRecognizer.cs:153-168 is commented out)."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(speech, dtype=np.float32))
        enc = encoder(x, w, dims)
        logits = _lin(enc, w, "ctc.ctc_lo.weight", "ctc.ctc_lo.bias", "head")
        logp = torch.log_softmax(logits, dim=-1).numpy()
    return {"logits": logp, "tokens": greedy_pick(logp), "enc": enc.numpy(),
            "token_num": np.full(speech.shape[0], speech.shape[1], dtype=np.int32)}


# --------------------------------------------------------------------------- SeACo-paraformer (hot-word bias)

def pad_hotwords(hotwords, max_length: int = 10) -> np.ndarray:
    """EmbedSeacoModel.PadList (EmbedSeacoModel.cs:110-123): truncate to 10 ids, right-pad with 0 -> [N,10] int32."""
    out = np.zeros((len(hotwords), max_length), dtype=np.int32)
    for i, h in enumerate(hotwords):
        h = list(h)[:max_length]
        out[i, : len(h)] = h
    return out


def hotword_ids_from_text(tokens, lines, sos_eos_id: int = 1):
    """OfflineRecognizer.GetHotwords (OfflineRecognizer.cs:72-90, Q9): one hot word per line, tokenised per UTF-16 char
    with ``Array.IndexOf(tokens, ch)``; unknown chars are dropped; a trailing ``[sos]`` entry is appended."""
    index = {}
    for i, t in enumerate(tokens):
        index.setdefault(t, i)                      # IndexOf = first match
    out = []
    for line in lines:
        units = line.encode("utf-16-le")
        chars = [units[i:i + 2].decode("utf-16-le", errors="surrogatepass") for i in range(0, len(units), 2)]
        out.append([index[c] for c in chars if c in index])
    out.append([sos_eos_id])
    return out


def hotword_embed(hotword: np.ndarray, w) -> np.ndarray:
    """``model_eb.onnx`` (EmbedSeacoModel.Forward, EmbedSeacoModel.cs:70-108): ``hotword [N,10] i32`` ->
    Embedding(V,512) -> 2-layer LSTM(512) run time-major -> ``hw_embed [10,N,512]`` (all time steps)."""
    with torch.no_grad():
        ids = torch.from_numpy(np.asarray(hotword, dtype=np.int64))
        x = Fn.embedding(ids, _t(w, "bias_embed.weight")).transpose(0, 1)          # [10, N, 512]
        for layer in range(2):
            wih, whh = _t(w, f"bias_encoder.weight_ih_l{layer}"), _t(w, f"bias_encoder.weight_hh_l{layer}")
            bih, bhh = _t(w, f"bias_encoder.bias_ih_l{layer}"), _t(w, f"bias_encoder.bias_hh_l{layer}")
            n, hdim = x.shape[1], whh.shape[1]
            h = torch.zeros(n, hdim)
            c = torch.zeros(n, hdim)
            outs = []
            for t in range(x.shape[0]):
                gates = Fn.linear(_ra(x[t], "hw.lstm"), _rw(wih, "hw.lstm"), bih) + Fn.linear(_ra(h, "hw.lstm"), _rw(whh, "hw.lstm"), bhh)
                i, f, g, o = gates.chunk(4, dim=1)                                    # PyTorch gate order
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            x = torch.stack(outs, dim=0)
    return x.numpy()


def bias_embed_rows(hw_embed: np.ndarray) -> np.ndarray:
    """OfflineProjOfSeacoParaformer.cs:85-108 (Q8): ALL 10 LSTM steps of every hot word, hot-word major:
    ``[10,N,512]`` -> ``[N*10, 512]`` (the C# then tiles this over the batch)."""
    return np.ascontiguousarray(np.transpose(hw_embed, (1, 0, 2)).reshape(-1, hw_embed.shape[2]))


def seaco_forward(speech: np.ndarray, w, dims: ModelDims, bias_rows: np.ndarray):
    """The ``InferenceSession.Run`` of OfflineProjOfSeacoParaformer.cs:116 on ``speech [B,T,560]`` and
    ``bias_embed [B, Nb, 512]`` (identical rows for every batch item).  FunASR SeacoParaformer export [EXT]:
    ASR decoder (returning its hidden), bias decoder over the hot-word rows queried by the CIF embeds and by the
    decoder hidden, ``hotword_output_layer`` -> log-softmax ``dha``; rows whose dha argmax is NO_BIAS keep the ASR
    posterior, the others take ``dha``."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(speech, dtype=np.float32))
        enc = encoder(x, w, dims)
        alphas = predictor_alphas(enc, w, dims)
        hidden = torch.cat([enc, torch.zeros(enc.shape[0], 1, enc.shape[2])], dim=1)
        emb, token_num, fires, peaks = cif(hidden.numpy(), alphas.numpy(), dims.cif_threshold)
        embeds = torch.from_numpy(emb)
        tn = torch.from_numpy(token_num.astype(np.int64))
        logits, dec_hidden = decoder(enc, embeds, tn, w, dims, return_hidden=True)
        asr = torch.log_softmax(logits, dim=-1)
        b = enc.shape[0]
        mem = torch.from_numpy(np.ascontiguousarray(bias_rows, dtype=np.float32))[None].expand(b, -1, -1)
        sdims = ModelDims(**{**dims.__dict__, "dec_ffn": dims.seaco_ffn, "dec_kernel": dims.seaco_kernel})
        cif_att = decoder(mem, embeds, tn, w, sdims, prefix="seaco_decoder", layers=dims.seaco_layers)
        dec_att = decoder(mem, dec_hidden, tn, w, sdims, prefix="seaco_decoder", layers=dims.seaco_layers)
        merged = cif_att + dec_att
        dha = torch.log_softmax(_lin(merged, w, "hotword_output_layer.weight", "hotword_output_layer.bias", "head"), dim=-1)
        dha_ids = dha.argmax(dim=-1)
        keep_asr = (dha_ids == dims.nobias_id)[..., None]
        out = torch.where(keep_asr, asr, dha).numpy()
    return {"logits": out, "token_num": token_num, "tokens": greedy_pick(out), "asr_logits": asr.numpy(), "dha": dha.numpy(),
            "dha_ids": dha_ids.numpy(), "enc": enc.numpy(), "acoustic_embeds": emb, "dec_hidden": dec_hidden.numpy(),
            "cif_peak": peaks}


# --------------------------------------------------------------------------- timestamps (CifPredictorV3, row f1)

def upsample_timestamp(enc: np.ndarray, token_num: np.ndarray, w, dims: ModelDims):
    """``CifPredictorV3.get_upsample_timestmap`` of the FunASR export [EXT] (outputs #3/#4 of the ``-timestamp-`` and
    SeACo graphs, read at OfflineProjOfParaformer.cs:75-79): ConvTranspose1d(512,512,k=3,stride=3) -> BiLSTM(512) ->
    Linear(1024,1) -> sigmoid -> relu(a * smooth_factor2 - noise_threshold2) -> rescale so each utterance sums to its
    token count -> ``cif_wo_hidden`` with threshold 1 - 1e-4.  Returns (us_alphas, us_cif_peak), both [B, 3T]."""
    with torch.no_grad():
        h = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
        up = Fn.conv_transpose1d(h.transpose(1, 2), _t(w, "predictor.upsample_cnn.weight"), _t(w, "predictor.upsample_cnn.bias"),
                                 stride=dims.upsample_times).transpose(1, 2)                      # [B, 3T, 512]
        lstm = torch.nn.LSTM(h.shape[2], h.shape[2], 1, bias=True, batch_first=True, bidirectional=True)
        for name in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            getattr(lstm, name).copy_(_t(w, "predictor.blstm." + name))
            getattr(lstm, name + "_reverse").copy_(_t(w, "predictor.blstm." + name + "_reverse"))
        y, _ = lstm(up)
        a2 = torch.sigmoid(Fn.linear(y, _t(w, "predictor.cif_output2.weight"), _t(w, "predictor.cif_output2.bias"))).squeeze(-1)
        a2 = torch.relu(a2 * dims.smooth_factor2 - dims.noise_threshold2)
        tn = torch.from_numpy(np.asarray(token_num, dtype=np.float32))
        a2 = a2 * (tn / a2.sum(-1))[:, None]
        us_alphas = a2.numpy().astype(np.float32)
    thr = np.float32(dims.cif_threshold - 1e-4)
    b, t3 = us_alphas.shape
    peaks = np.zeros((b, t3), dtype=np.float32)
    for bi in range(b):
        integrate = np.float32(0.0)
        for t in range(t3):
            integrate = np.float32(integrate + us_alphas[bi, t])
            peaks[bi, t] = integrate
            if integrate >= thr:
                integrate = np.float32(integrate - thr)
    return us_alphas, peaks


def time_stamp_lfr6_onnx(us_cif_peak, tokens, begin_time: float = 0.0, total_offset: float = -1.5):
    """``OfflineRecognizer.time_stamp_lfr6_onnx`` (OfflineRecognizer.cs:200-302), float32 arithmetic like the C#:
    fire places (peak > 1 - 1e-4) shifted by ``total_offset`` -> [start, end] ms per emitted token."""
    f32 = np.float32
    START_END_THRESHOLD, MAX_TOKEN_DURATION = 5, 30
    TIME_RATE = f32(f32(10.0) * 6 / 1000 / 3)
    us = np.asarray(us_cif_peak, dtype=np.float32)
    num_frames = us.shape[0]
    tokens = list(tokens)
    if tokens and tokens[-1] == 2:
        tokens = tokens[:-1]
    fire_place = [f32(i + total_offset) for i in range(num_frames) if float(us[i]) > 1.0 - 1e-4]
    ts_list, new_char = [], []
    if fire_place[0] > START_END_THRESHOLD:
        ts_list.append([f32(0.0), f32(fire_place[0] * TIME_RATE)])
        new_char.append(False)
    for i in range(len(fire_place) - 1):
        new_char.append(tokens[i] != 1)
        if i == len(fire_place) - 2 or MAX_TOKEN_DURATION < 0 or fire_place[i + 1] - fire_place[i] < MAX_TOKEN_DURATION:
            ts_list.append([f32(fire_place[i] * TIME_RATE), f32(fire_place[i + 1] * TIME_RATE)])
        else:
            split = f32(fire_place[i] + MAX_TOKEN_DURATION)
            ts_list.append([f32(fire_place[i] * TIME_RATE), f32(split * TIME_RATE)])
            ts_list.append([f32(split * TIME_RATE), f32(fire_place[i + 1] * TIME_RATE)])
            new_char.append(False)
    if num_frames - fire_place[-1] > START_END_THRESHOLD:
        end = f32(f32(num_frames + fire_place[-1]) / 2)
        ts_list[-1][1] = f32(end * TIME_RATE)
        ts_list.append([f32(end * TIME_RATE), f32(f32(num_frames) * TIME_RATE)])
        new_char.append(False)
    else:
        ts_list[-1][1] = f32(f32(num_frames) * TIME_RATE)
    if begin_time > 0.0:
        for t in ts_list:
            t[0] = f32(t[0] + f32(begin_time) / f32(1000.0))
            t[1] = f32(t[1] + f32(begin_time) / f32(1000.0))
    new_char.append(True)
    return [[int(f32(t[0] * f32(1000))), int(f32(t[1] * f32(1000)))] for c, t in zip(new_char, ts_list) if c]
