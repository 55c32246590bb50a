"""CPU restatement of the reference's streaming (online) path (test infrastructure; network graph parity UNPINNED,
see oracle/__init__.py).  Host logic follows the C# line by line:

  * ``OnlineStream``        /root/reference/AliParaformerAsr/OnlineStream.cs:38-65 (state), :84-112 (AddSamples, Q13),
                            :114-167 (InputSpeech: per-chunk state-less fbank, first-frame repeat), :170-216
                            (GetDecodeChunk: splice frame, LFR, CMVN, x sqrt(512), PE, 10-frame feature cache, Q14)
  * ``OnlineWavFrontend``   OnlineWavFrontend.cs:73-91 (streaming LFR rule), :53-71 (CMVN), :152-188 (PE, Q12)
  * ``OnlineModel``         OnlineModel.cs:15-16,30-31 (chunk_size 5, lfr 10, chunk length 60), :141-165 (DynamicMask),
                            :207-229 (stack_states, Q11: every layer receives the stream's layer-0 cache), :230-254
  * ``OnlineRecognizer``    OnlineRecognizer.cs:50-123 (EncoderProj), :125-234 (PredictorProj = host CIF with carry,
                            Q15), :236-339 (DecoderProj + greedy pick, Q5), :341-401 (Forward), :459-471 (Q4 online)

The two ONNX graphs (``encoder.onnx`` / ``decoder.onnx``) are not vendored; their semantics are the FunASR
paraformer-online export: the encoder is the offline SAN-M encoder WITHOUT the sqrt(d) scale and
positional encoding (the host applies both) followed by the CifPredictorV2 alpha head without tail; the decoder is the
offline SANM decoder whose FSMN memory is causal over ``[in_cache (kernel-1 frames) | new tokens]`` and returns the
last kernel-1 frames as ``out_cache``; logits are returned raw (no log-softmax; the greedy pick is unaffected).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as Fn

from . import frontend as F
from . import sanm

ONLINE_PAD_VALUE = np.float32(-23.025850929940457)      # OnlineRecognizer.cs:469
CHUNK_SIZE = 5                                           # OnlineModel.cs:15
LFR = 10                                                 # OnlineModel.cs:16
CHUNK_LENGTH = LFR * CHUNK_SIZE + 10                     # OnlineModel.cs:30 -> 60 fbank frames per decode chunk


# --------------------------------------------------------------------------- OnlineWavFrontend
def online_apply_lfr(fbank: np.ndarray, lfr_m: int = 7, lfr_n: int = 6) -> np.ndarray:
    """OnlineWavFrontend.ApplyLfr (OnlineWavFrontend.cs:73-91): no left padding; ``t_lfr = t/n - 1`` when
    ``t % n < m - n`` else ``t/n``; output i = frames [i*n, i*n + m) flattened."""
    t = fbank.shape[0]
    t_lfr = 0
    if t % lfr_n < lfr_m - lfr_n:
        t_lfr = t // lfr_n - 1
    if t % lfr_n >= lfr_m - lfr_n:
        t_lfr = t // lfr_n
    t_lfr = max(t_lfr, 0)
    flat = np.ascontiguousarray(fbank, dtype=np.float32).reshape(-1)
    d = fbank.shape[1]
    out = np.zeros((t_lfr, lfr_m * d), dtype=np.float32)
    for i in range(t_lfr):
        out[i] = flat[i * lfr_n * d: i * lfr_n * d + lfr_m * d]
    return out


def online_position_encoding(timesteps: int, dim: int, start_idx: int) -> np.ndarray:
    """OnlineWavFrontend.SinusoidalPositionEncoder (OnlineWavFrontend.cs:152-188, Q12): positions are 1-based and
    continue across chunks (``start_idx``); ``inv_timescale_i = exp(-(i+1) * ln(1e4)/(dim/2 - 1))`` (FunASR uses i);
    the product position * inv_timescale is a float32 multiply, sin/cos are evaluated in double and rounded."""
    half = dim // 2
    inc = np.float32(np.float32(math.log(np.float32(10000.0))) / np.float32(half - 1))
    inv = (np.arange(1, half + 1, dtype=np.float32) * np.float32(-inc)).astype(np.float32)
    inv = np.exp(inv.astype(np.float64)).astype(np.float32)
    pos = np.arange(start_idx + 1, start_idx + timesteps + 1, dtype=np.float32)
    ang = (pos[:, None] * inv[None, :]).astype(np.float32)                # float * float in C#
    return np.concatenate([np.sin(ang.astype(np.float64)), np.cos(ang.astype(np.float64))], axis=1).astype(np.float32)


class OnlineStreamState:
    """``OnlineStream`` (OnlineStream.cs): sample cache, fbank FIFO, splice frame, feature cache, CIF carry,
    16 decoder FSMN caches, token list."""

    def __init__(self, add_shift: np.ndarray, rescale: np.ndarray, snip_edges: bool = False, d_model: int = 512,
                 dec_layers: int = 16, dec_kernel: int = 11, n_mels: int = 80, lfr_m: int = 7, lfr_n: int = 6):
        self.add_shift = np.asarray(add_shift, dtype=np.float32)
        self.rescale = np.asarray(rescale, dtype=np.float32)
        self.snip_edges = snip_edges
        self.n_mels, self.lfr_m, self.lfr_n = n_mels, lfr_m, lfr_n
        self.cache_samples = np.zeros(160 * CHUNK_LENGTH, dtype=np.float32)        # OnlineStream.cs:61 (Q13)
        self.speech = np.zeros((0, n_mels), dtype=np.float32)                      # OnlineInputEntity.Speech
        self.cache_input_len = 0                                                   # only its emptiness is used
        self.splice: Optional[np.ndarray] = None                                   # _cachelfrSplice
        self.cache_feats = np.zeros((10, lfr_m * n_mels), dtype=np.float32)        # InitCacheFeats
        self.start_idx = 0
        self.cif_hidden = np.zeros((1, d_model), dtype=np.float32)                 # InitHidden
        self.cif_alpha = np.zeros((1,), dtype=np.float32)                          # InitAlpha
        self.states = [np.zeros((d_model, dec_kernel - 1), dtype=np.float32) for _ in range(dec_layers)]
        self.tokens: List[int] = [0, 0]                                            # {_blank_id, _blank_id}

    # OnlineStream.AddSamples (:84-112): at most ONE chunk is consumed per call, and only when buffer > chunk
    def add_samples(self, samples: np.ndarray) -> None:
        if samples is None:
            raise ValueError("samples is null")
        self.cache_samples = np.concatenate([self.cache_samples, np.asarray(samples, dtype=np.float32)])
        chunk = 160 * CHUNK_LENGTH
        if self.cache_samples.shape[0] > chunk:
            self._input_speech(self.cache_samples[:chunk].copy())
            self.cache_samples = self.cache_samples[chunk:].copy()

    # OnlineStream.InputSpeech (:114-167)
    def _input_speech(self, samples: np.ndarray) -> None:
        feats = F.get_fbank(samples, snip_edges=self.snip_edges, num_bins=self.n_mels)
        if self.cache_input_len == 0 and feats.shape[0] > 0:
            feats = np.concatenate([feats[:1], feats], axis=0)                     # _repeatNum = 1 (:64,141-153)
        frame_num = int(math.ceil((samples.shape[0] - 400) / 160.0))
        if frame_num < 1 or samples.shape[0] < 400:
            frame_num = 0
        self.cache_input_len = samples.shape[0] - frame_num * 160
        self.speech = np.concatenate([self.speech, feats.astype(np.float32)], axis=0)

    # OnlineStream.GetDecodeChunk (:170-216) -> [20, 560] or None
    def get_decode_chunk(self) -> Optional[np.ndarray]:
        if CHUNK_LENGTH > self.speech.shape[0]:
            return None
        pad = self.speech[:CHUNK_LENGTH]
        head = self.splice if self.splice is not None else pad[:1]
        pad = np.concatenate([head, pad], axis=0)                                  # 61 frames
        self.splice = pad[-1:].copy()
        x = online_apply_lfr(pad, self.lfr_m, self.lfr_n)
        x = F.apply_cmvn(x, self.add_shift, self.rescale)
        x = (x.astype(np.float64) * math.pow(512, 0.5)).astype(np.float32)         # (float)(x * Math.Pow(512, 0.5))
        t = x.shape[0]
        x = (x + online_position_encoding(t, x.shape[1], self.start_idx)).astype(np.float32)
        chunk = np.concatenate([self.cache_feats, x], axis=0)
        self.start_idx += t
        self.cache_feats = chunk[-self.cache_feats.shape[0]:].copy()
        self.speech = self.speech[CHUNK_LENGTH:].copy()                            # RemoveChunk
        return chunk


# --------------------------------------------------------------------------- graphs (FunASR online export, [EXT])
def online_encoder(speech: np.ndarray, w: Dict[str, np.ndarray], dims: sanm.ModelDims):
    """``encoder.onnx``: pre-scaled, pre-PE'd ``speech [B,20,560]`` -> ``enc [B,20,512]``, ``alphas [B,20]``
    (OnlineRecognizer.cs:84-115).  No tail alpha in streaming."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(speech, dtype=np.float32))
        x = sanm.encoder_layer(x, w, "encoder.encoders0.0", dims)
        for i in range(dims.enc_layers - 1):
            x = sanm.encoder_layer(x, w, f"encoder.encoders.{i}", dims)
        enc = sanm._ln(x, w, "encoder.after_norm", dims.ln_eps)
        alphas = sanm.predictor_alphas(enc, w, dims)[:, :-1]
    return enc.numpy(), alphas.numpy()


def _fsmn_cached(tn: torch.Tensor, weight: torch.Tensor, mask: torch.Tensor, cache: torch.Tensor):
    """Streaming decoder FSMN: x = cat(cache [B,D,k-1], (tn*mask)^T); conv1d without padding; + tn*mask; * mask.
    Returns (out [B,L,D], out_cache = last k-1 columns of x)."""
    v = tn * mask
    x = torch.cat([cache, v.transpose(1, 2)], dim=2)
    k = weight.shape[-1]
    out_cache = x[:, :, -(k - 1):].clone()
    y = Fn.conv1d(x, weight, None, groups=weight.shape[0]).transpose(1, 2)
    y = (y + v) * mask
    return y, out_cache


def online_decoder(enc: np.ndarray, embeds: np.ndarray, embeds_len: np.ndarray, in_caches: Sequence[np.ndarray],
                   w: Dict[str, np.ndarray], dims: sanm.ModelDims):
    """``decoder.onnx``: ``enc [B,20,512]``, ``acoustic_embeds [B,L,512]``, lengths, ``in_cache_i [B,512,10]`` ->
    raw ``logits [B,L,V]`` and ``out_cache_i`` (OnlineRecognizer.cs:255-330)."""
    with torch.no_grad():
        memory = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
        x = torch.from_numpy(np.ascontiguousarray(embeds, dtype=np.float32))
        b, l, d = x.shape
        lens = torch.from_numpy(np.asarray(embeds_len, dtype=np.int64))
        mask = (torch.arange(l)[None, :] < lens[:, None]).to(x.dtype)[:, :, None]
        out_caches = []
        for i in range(dims.dec_layers):
            p = f"decoder.decoders.{i}"
            t = sanm._dec_ffn(sanm._ln(x, w, p + ".norm1", dims.ln_eps), w, p + ".feed_forward", dims)
            tn = sanm._ln(t, w, p + ".norm2", dims.ln_eps)
            y, oc = _fsmn_cached(tn, sanm._t(w, p + ".self_attn.fsmn_block.weight"), mask,
                                 torch.from_numpy(np.ascontiguousarray(in_caches[i], dtype=np.float32)))
            out_caches.append(oc.numpy())
            x = x + y
            h = sanm._ln(x, w, p + ".norm3", dims.ln_eps)
            q = sanm._lin(h, w, p + ".src_attn.linear_q.weight", p + ".src_attn.linear_q.bias", "dec.q")
            kv = sanm._lin(memory, w, p + ".src_attn.linear_k_v.weight", p + ".src_attn.linear_k_v.bias", "dec.kv")
            k, v = torch.split(kv, d, dim=-1)
            ctx = sanm._mha(q, k, v, dims.heads, "dec.att")
            x = x + sanm._lin(ctx, w, p + ".src_attn.linear_out.weight", p + ".src_attn.linear_out.bias", "dec.out")
        p = "decoder.decoders3.0"
        x = sanm._dec_ffn(sanm._ln(x, w, p + ".norm1", dims.ln_eps), w, p + ".feed_forward", dims)
        x = sanm._ln(x, w, "decoder.after_norm", dims.ln_eps)
        logits = sanm._lin(x, w, "decoder.output_layer.weight", "decoder.output_layer.bias", "head")
    return logits.numpy(), out_caches


# --------------------------------------------------------------------------- OnlineModel helpers
def dynamic_mask(alphas: np.ndarray) -> np.ndarray:
    """OnlineModel.DynamicMask (OnlineModel.cs:141-165): zero alphas [0, chunk_size) and [chunk_size + lfr, end)."""
    a = np.array(alphas, dtype=np.float32, copy=True)
    a[:, :CHUNK_SIZE] = 0.0
    if a.shape[1] > CHUNK_SIZE + LFR:
        a[:, CHUNK_SIZE + LFR:] = 0.0
    return a


def host_cif(hiddens: np.ndarray, alphas: np.ndarray, threshold: float):
    """PredictorProj's recurrence for ONE stream (OnlineRecognizer.cs:149-200, Q15), float32 with separate multiply
    and add.  hiddens [T,D] / alphas [T] already include the carried entry at index 0.
    Returns (frames list of [D], carry_alpha, carry_hidden [D])."""
    thr = np.float32(threshold)
    integrate = np.float32(0.0)
    frames = np.zeros(hiddens.shape[1], dtype=np.float32)
    fired: List[np.ndarray] = []
    for j in range(alphas.shape[0]):
        alpha = np.float32(alphas[j])
        h = hiddens[j].astype(np.float32)
        if np.float32(alpha + integrate) < thr:
            integrate = np.float32(integrate + alpha)
            frames = (frames + (alpha * h).astype(np.float32)).astype(np.float32)
        else:
            frames = (frames + (np.float32(thr - integrate) * h).astype(np.float32)).astype(np.float32)
            fired.append(frames)
            integrate = np.float32(integrate + alpha)
            integrate = np.float32(integrate - thr)
            frames = (integrate * h).astype(np.float32)
    carry_h = (frames / integrate).astype(np.float32) if integrate > 0.0 else frames
    return fired, integrate, carry_h


class OnlineRecognizerOracle:
    """``OnlineRecognizer.Forward`` (OnlineRecognizer.cs:341-401) over :class:`OnlineStreamState` objects."""

    def __init__(self, w: Dict[str, np.ndarray], dims: sanm.ModelDims, add_shift, rescale, snip_edges: bool = False,
                 compat_layer0_cache: bool = True):
        self.w, self.dims = w, dims
        self.add_shift, self.rescale, self.snip_edges = add_shift, rescale, snip_edges
        self.compat = compat_layer0_cache

    def create_stream(self) -> OnlineStreamState:
        return OnlineStreamState(self.add_shift, self.rescale, self.snip_edges, self.dims.d_model, self.dims.dec_layers,
                                 self.dims.dec_kernel)

    def forward(self, streams: Sequence[OnlineStreamState], collect: Optional[dict] = None) -> List[List[int]]:
        """Returns the ids appended to each stream by this call (empty for streams without a full chunk)."""
        new_tokens: List[List[int]] = [[] for _ in streams]
        working, chunks = [], []
        for i, s in enumerate(streams):
            c = s.get_decode_chunk()
            if c is None:
                continue
            working.append(i)
            chunks.append(c)
        if not working:
            return new_tokens
        d = self.dims
        # stack_states (OnlineModel.cs:207-229, Q11): layer i of the batch tensor = each stream's layer-0 cache
        caches = []
        for layer in range(d.dec_layers):
            src = 0 if self.compat else layer
            caches.append(np.stack([streams[i].states[src] for i in working]))
        speech = np.stack(chunks).astype(np.float32)
        speech = np.where(speech == 0, ONLINE_PAD_VALUE, speech).astype(np.float32)    # PadSequence_unittest (:469)
        enc, alphas = online_encoder(speech, self.w, d)
        alphas = dynamic_mask(alphas)
        frames_all, lens = [], []
        for bi, i in enumerate(working):
            s = streams[i]
            hid = np.concatenate([s.cif_hidden, enc[bi]], axis=0)
            alp = np.concatenate([s.cif_alpha, alphas[bi]], axis=0)
            fired, carry_a, carry_h = host_cif(hid, alp, d.cif_threshold)
            s.cif_alpha = np.asarray([carry_a], dtype=np.float32)
            s.cif_hidden = carry_h[None, :].astype(np.float32)
            frames_all.append(fired)
            lens.append(len(fired))
        lmax = max(lens)
        if collect is not None:
            collect.update(speech=speech, enc=enc, alphas=alphas, lens=np.asarray(lens, dtype=np.int32))
        if lmax == 0:
            return new_tokens
        embeds = np.zeros((len(working), lmax, d.d_model), dtype=np.float32)
        for bi, fired in enumerate(frames_all):
            if fired:
                embeds[bi, : len(fired)] = np.stack(fired)
        logits, out_caches = online_decoder(enc, embeds, np.asarray(lens), caches, self.w, d)
        ids = sanm.greedy_pick(logits)                                      # all Lmax rows, padded ones included
        if collect is not None:
            collect.update(embeds=embeds, logits=logits, out_caches=out_caches)
        for bi, i in enumerate(working):
            streams[i].tokens.extend(int(t) for t in ids[bi])
            new_tokens[i] = [int(t) for t in ids[bi]]
            streams[i].states = [out_caches[layer][bi].copy() for layer in range(d.dec_layers)]
        return new_tokens
