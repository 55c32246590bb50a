"""Host-side mirror of the reference's public offline API, driving libpfasr.so through the C-ABI.

The reference's host language is C#; this image has no dotnet, so the host side above the C-ABI is written in
Python with the same type names, call sequence, argument meaning and error behaviour:

  * ``OfflineRecognizer``  /root/reference/AliParaformerAsr/OfflineRecognizer.cs:23 (ctor), :92 CreateOfflineStream,
                           :102 GetResult, :110 GetResults, :441 DisposeOfflineStream, :468 Dispose
  * ``OfflineStream``      OfflineStream.cs:30-36 (Tokens / Timestamps / Hotwords / AddSamples)
  * result entity          Model/OfflineRecognizerResultEntity.cs:9-29

What moved to the GPU: everything between ``AddSamples`` and the argmax loop of ``Forward``
(WavFrontend, PadHelper, IOfflineProj.ModelProj, greedy pick).  What stays on the host, as in the reference:
config / token-table parsing and ``DecodeMulti`` (ids -> text).
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .engine import Engine
from .synth import ModelConfig

PAD_QUIRK_VALUE = np.float32(np.float32(-23.025850929940457) * np.float32(32768.0))   # Utils/PadHelper.cs:63


class ArgumentNullError(ValueError):
    """``ArgumentNullException``; ``param_name`` mirrors ``ParamName`` (tests expect "source", WavFrontend.cs:34)."""

    def __init__(self, param_name: str):
        super().__init__(f"Value cannot be null. (Parameter '{param_name}')")
        self.param_name = param_name


class ObjectDisposedError(RuntimeError):
    """``ObjectDisposedException``; ``object_name`` mirrors ``ObjectName`` (OfflineRecognizer.cs:94-97)."""

    def __init__(self, object_name: str):
        super().__init__(f"Cannot access a disposed object. Object name: '{object_name}'.")
        self.object_name = object_name


# --------------------------------------------------------------------------- config / asset loading (host, as in C#)
def read_tokens(path: str) -> Optional[List[str]]:
    """Utils/PreloadHelper.cs:120-141 ``ReadTokens``: one token per line; empty path -> null."""
    if not path:
        return None
    with open(path, "rb") as f:
        lines = re.split(r"\r\n|\r|\n", f.read().decode("utf-8-sig"))   # File.ReadAllLines: CR / LF / CRLF, BOM dropped
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def load_cmvn(path: str):
    """``WavFrontend.LoadCmvn`` (WavFrontend.cs:112-153): the bracketed vectors on the ``<LearnRateCoef>`` lines
    following ``<AddShift>`` and ``<Rescale>`` of a Kaldi-nnet ``am.mvn``."""
    state = 0
    shift: List[float] = []
    scale: List[float] = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f.read().splitlines():
            if not line:
                continue
            if line.startswith("<AddShift>"):
                state = 1
            elif line.startswith("<Rescale>"):
                state = 2
            elif line.startswith("<LearnRateCoef>") and state in (1, 2):
                body = line[line.index("[") + 1: line.rindex("]")]
                vals = [float(t) for t in body.split(" ") if t.strip()]
                if state == 1:
                    shift = vals
                else:
                    scale = vals
    return np.asarray(shift, dtype=np.float32), np.asarray(scale, dtype=np.float32)


def load_conf(path: str) -> ModelConfig:
    """``OfflineRecognizer.LoadConf`` (OfflineRecognizer.cs:55-71): ``.json`` or ``.yaml`` -> the ConfEntity fields the
    hot path consumes.  Missing keys keep the paraformer-large defaults baked into the C# entities."""
    cfg = ModelConfig()
    if not path:
        return cfg
    raw = {}
    low = path.lower()
    if low.endswith(".json"):
        with open(path, "r", encoding="utf-8") as f:
            raw = json.load(f) or {}
    elif low.endswith(".yaml"):
        import yaml
        with open(path, "r", encoding="utf-8") as f:
            raw = yaml.safe_load(f) or {}
    cfg.model = str(raw.get("model", cfg.model) or cfg.model)
    cfg.use_itn = bool(raw.get("use_itn", cfg.use_itn))
    enc = raw.get("encoder_conf") or {}
    cfg.d_model = int(enc.get("output_size", cfg.d_model))
    cfg.heads = int(enc.get("attention_heads", cfg.heads))
    cfg.ffn = int(enc.get("linear_units", cfg.ffn))
    cfg.enc_layers = int(enc.get("num_blocks", cfg.enc_layers))
    cfg.tp_layers = int(enc.get("tp_blocks", cfg.tp_layers))
    cfg.enc_kernel = int(enc.get("kernel_size", cfg.enc_kernel))
    dec = raw.get("decoder_conf") or {}
    cfg.dec_layers = int(dec.get("num_blocks", cfg.dec_layers))
    cfg.dec_ffn = int(dec.get("linear_units", cfg.dec_ffn))
    cfg.dec_kernel = int(dec.get("kernel_size", cfg.dec_kernel))
    pred = raw.get("predictor_conf") or {}
    cfg.cif_threshold = float(pred.get("threshold", cfg.cif_threshold))
    cfg.cif_tail = float(pred.get("tail_threshold", cfg.cif_tail))
    cfg.smooth_factor = float(pred.get("smooth_factor", cfg.smooth_factor))
    cfg.noise_threshold = float(pred.get("noise_threshold", cfg.noise_threshold))
    fe = raw.get("frontend_conf") or {}
    cfg.fs = int(fe.get("fs", cfg.fs))
    cfg.n_mels = int(fe.get("n_mels", cfg.n_mels))
    cfg.lfr_m = int(fe.get("lfr_m", cfg.lfr_m))
    cfg.lfr_n = int(fe.get("lfr_n", cfg.lfr_n))
    cfg.snip_edges = bool(fe.get("snip_edges", cfg.snip_edges))
    cfg.input_size = cfg.lfr_m * cfg.n_mels
    # the reference forwards window / frame sizes / dither to the fbank (WavFrontend.cs:21-26, FrontendConfEntity.cs:7-15);
    # the device front-end is built for the values every published model uses, so anything else is an error here
    # instead of silently different features
    window = str(fe.get("window", "hamming")).lower()
    frame_length, frame_shift = int(fe.get("frame_length", 25)), int(fe.get("frame_shift", 10))
    if window != "hamming" or frame_length != 25 or frame_shift != 10:
        raise NotImplementedError(f"frontend_conf window={window!r} frame_length={frame_length} frame_shift={frame_shift}: the device "
                                  "front-end computes Hamming windows of 25 ms every 10 ms (PF_ERR_UNSUPPORTED)")
    # dither: the C# entity defaults to 1.0, which makes the reference non-deterministic; the device path never dithers
    # (pf_abi.h), so a non-zero value is accepted and recorded, not applied
    cfg.dither = float(fe.get("dither", 0.0))
    sd = raw.get("seaco_decoder_conf") or {}
    cfg.seaco_layers = int(sd.get("num_blocks", cfg.seaco_layers))
    cfg.seaco_ffn = int(sd.get("linear_units", cfg.seaco_ffn))
    cfg.seaco_kernel = int(sd.get("kernel_size", cfg.seaco_kernel))
    if "vocab_size" in raw:
        cfg.vocab = int(raw["vocab_size"])
    if "ln_eps" in raw:
        cfg.ln_eps = float(raw["ln_eps"])
    if cfg.model.lower() == "sensevoicesmall":
        cfg.dec_layers = 0
    return cfg


def get_hotwords(tokens: Sequence[str], hotword_file_path: str, sos_eos_id: int = 1) -> List[List[int]]:
    """``OfflineRecognizer.GetHotwords`` (OfflineRecognizer.cs:72-90, Q9): one hot word per line, tokenised per UTF-16
    char with ``Array.IndexOf(tokens, ch)`` (first match; unknown chars dropped), plus a trailing ``[sos]`` entry.
    A missing file yields an empty list."""
    if not hotword_file_path or not os.path.exists(hotword_file_path):
        return []
    index = {}
    for i, t in enumerate(tokens):
        index.setdefault(t, i)
    out: List[List[int]] = []
    with open(hotword_file_path, "r", encoding="utf-8-sig") as f:
        for line in f.read().splitlines():
            units = line.encode("utf-16-le")
            chars = [units[i:i + 2].decode("utf-16-le", errors="surrogatepass") for i in range(0, len(units), 2)]
            out.append([index[c] for c in chars if c in index])
    out.append([sos_eos_id])
    return out


# --------------------------------------------------------------------------- entities
@dataclass
class OfflineRecognizerResultEntity:
    """Model/OfflineRecognizerResultEntity.cs:9-29"""
    text: str = ""
    text_len: int = 0
    tokens: List[str] = field(default_factory=list)
    timestamps: List[List[int]] = field(default_factory=list)

    # C#-style aliases
    @property
    def Text(self):
        return self.text

    @property
    def Tokens(self):
        return self.tokens

    @property
    def Timestamps(self):
        return self.timestamps


class OfflineStream:
    """OfflineStream.cs: holds one utterance's input until ``GetResults`` consumes it."""

    def __init__(self, recognizer: "OfflineRecognizer"):
        self._rec = recognizer
        self._chunks: List[np.ndarray] = []      # one entry per AddSamples call (Q10: features are concatenated)
        self.hotwords: Optional[List[List[int]]] = []
        self.tokens: List[int] = [0, 0]          # OfflineStream.cs:26 {blank, blank}
        self.timestamps: List[List[int]] = []
        self._disposed = False

    def add_samples(self, samples) -> None:
        """OfflineStream.AddSamples (OfflineStream.cs:36-57).  ``None`` raises like the LINQ ``Select`` in
        WavFrontend.GetFbank does (ArgumentNullException, ParamName "source")."""
        if samples is None:
            raise ArgumentNullError("source")
        self._chunks.append(np.ascontiguousarray(samples, dtype=np.float32).reshape(-1))

    AddSamples = add_samples

    # C#-style member names (OfflineStream.cs:30-34)
    @property
    def Hotwords(self):
        return self.hotwords

    @Hotwords.setter
    def Hotwords(self, value):
        self.hotwords = value

    @property
    def Tokens(self):
        return self.tokens

    @property
    def Timestamps(self):
        return self.timestamps

    def features(self) -> np.ndarray:
        """What ``OfflineInputEntity.Speech`` holds: per-call fbank->LFR->CMVN, concatenated (Q10)."""
        eng = self._rec._engine
        parts = [eng.extract(c) for c in self._chunks]
        dim = eng.cfg.lfr_m * eng.cfg.n_mels
        return np.concatenate(parts, axis=0) if parts else np.zeros((0, dim), np.float32)

    def remove_chunk(self) -> None:
        """OfflineStream.RemoveChunk (OfflineStream.cs:69-79): input is dropped once tokens were produced."""
        if len(self.tokens) > 2:
            self._chunks = []

    def dispose(self) -> None:
        self._disposed = True
        self._chunks = []

    Dispose = dispose


from .text import TokenTable, time_stamp_lfr6_onnx  # noqa: E402  (native post-processing, csrc/text.cu)


def pad_sequence(feats: Sequence[np.ndarray]) -> np.ndarray:
    """PadHelper.PadSequence (Utils/PadHelper.cs:23-65) on the host, used only when a stream received several
    AddSamples calls: right-pad with 0 to the longest item, then every exact 0.0 -> -23.0258509f*32768 (Q4)."""
    dim = feats[0].shape[-1]
    tmax = max(int(f.shape[0]) for f in feats)
    out = np.zeros((len(feats), tmax, dim), dtype=np.float32)
    for i, f in enumerate(feats):
        out[i, : f.shape[0]] = f
    out[out == 0] = PAD_QUIRK_VALUE
    return out


class OfflineRecognizer:
    """OfflineRecognizer.cs:23 — same constructor arguments; ``model_file_path`` is the PFW1 weight blob that replaces
    ``model.onnx``.  ``threads_num`` is accepted and ignored (it only set ORT inter-op threads, Q17).  ``lanes`` > 1 lets
    ``get_results`` calls from different host threads run concurrently on the GPU (pf_offline_create_mt)."""

    def __init__(self, model_file_path: str, config_file_path: str, mvn_file_path: str, tokens_file_path: str,
                 modeleb_file_path: str = "", hotword_file_path: str = "", batch_size: int = 1, threads_num: int = 1,
                 devices: Optional[Sequence[int]] = None, weights=None, config: Optional[ModelConfig] = None, lanes: int = 1):
        self._disposed = False
        self._conf = config if config is not None else load_conf(config_file_path)
        self._tokens = read_tokens(tokens_file_path)
        if not self._tokens:
            raise Exception("tokens invalid")                      # OfflineRecognizer.cs:30-33
        self._token_table = TokenTable(path=tokens_file_path)      # the same file, held by libpfasr for DecodeMulti
        self._mvn_file_path = mvn_file_path
        self._engine = Engine(self._conf, weights if weights is not None else model_file_path, devices=devices, lanes=lanes)
        if mvn_file_path:
            shift, scale = load_cmvn(mvn_file_path)
            self._engine.set_cmvn(shift, scale)
        # SeACo: the hot-word encoder (model_eb.onnx in the reference; here part of the same PFW1 blob, so
        # modeleb_file_path is accepted and unused) runs once on the file hot words (OfflineProjOfSeacoParaformer.cs:29-34)
        self._seaco = self._conf.model.lower() == "seacoparaformer"
        self._hotwords = get_hotwords(self._tokens, hotword_file_path) if self._seaco else []
        if self._seaco and self._hotwords:
            self._engine.set_hotwords(self._hotwords)

    # -- API surface of the reference
    def create_offline_stream(self) -> OfflineStream:
        if self._disposed:
            raise ObjectDisposedError("OfflineRecognizer")          # OfflineRecognizer.cs:94-97
        return OfflineStream(self)

    def get_result(self, stream: OfflineStream) -> OfflineRecognizerResultEntity:
        return self.get_results([stream])[0]

    def get_results(self, streams: List[OfflineStream]) -> List[OfflineRecognizerResultEntity]:
        self._forward(streams)
        return self._decode_multi(streams)

    def dispose_offline_stream(self, stream: Optional[OfflineStream]) -> None:
        if stream is not None:
            stream.dispose()

    def dispose(self) -> None:
        if not self._disposed:
            self._engine.close()
            self._token_table.close()
            self._tokens = None
            self._disposed = True

    CreateOfflineStream = create_offline_stream
    GetResult = get_result
    GetResults = get_results
    DisposeOfflineStream = dispose_offline_stream
    Dispose = dispose

    # -- OfflineRecognizer.Forward (OfflineRecognizer.cs:118-198)
    def _forward(self, streams: List[OfflineStream]) -> None:
        if not streams:
            return
        if self._disposed:
            raise ObjectDisposedError("OfflineRecognizer")
        # per-stream hot words replace the file hot words for this call (OfflineProjOfSeacoParaformer.cs:51-60)
        call_hotwords = [list(h) for s in streams for h in (s.hotwords or [])] if self._seaco else []
        # set -> run -> restore is one leased section: another thread sharing the lane cannot run with these hot words
        with self._engine.lease():
            try:
                if call_hotwords:
                    self._engine.set_hotwords(call_hotwords, local=True)
                want_ts = bool(getattr(self._conf, "timestamps", False))      # 4-output models (OfflineProjOfParaformer.cs:75-79)
                if all(len(s._chunks) == 1 for s in streams):
                    # one AddSamples per stream: fused fbank+LFR+CMVN+PadSequence on the device
                    out = self._engine.run_pcm([s._chunks[0] for s in streams], want_timestamps=want_ts)
                else:
                    feats = [s.features() for s in streams]
                    if max(f.shape[0] for f in feats) == 0:
                        raise ValueError("no input samples")
                    out = self._engine.run_feats(pad_sequence(feats), want_timestamps=want_ts)
            except _lib.PfError as ex:
                raise Exception("Offline recognition failed") from ex   # OfflineRecognizer.cs:194-197
            finally:
                if call_hotwords:
                    self._engine.set_hotwords(self._hotwords, local=True)
        for i, s in enumerate(streams):
            s.tokens = [int(t) for t in out.tokens[i]]
            if out.us_cif_peak is not None:                         # cif_peak_tensor != null (OfflineRecognizer.cs:172-183)
                s.timestamps.extend(time_stamp_lfr6_onnx(out.us_cif_peak[i], s.tokens))
            else:
                s.timestamps.extend([[0, 0] for _ in s.tokens])     # 3-output models: {0,0} per token (:151)
            s.remove_chunk()

    # -- OfflineRecognizer.DecodeMulti (OfflineRecognizer.cs:304-418), native: pf_decode_offline (csrc/text.cu)
    def _decode_multi(self, streams: List[OfflineStream]) -> List[OfflineRecognizerResultEntity]:
        results = []
        for s in streams:
            ent = OfflineRecognizerResultEntity()
            ent.text, ent.text_len, ent.tokens, ent.timestamps = self._token_table.decode_offline(s.tokens, s.timestamps)
            results.append(ent)
        return results
