"""Host-side mirror of the reference's streaming API, driving ``pf_online_*`` of libpfasr.so.

  * ``OnlineRecognizer``  /root/reference/AliParaformerAsr/OnlineRecognizer.cs:22 (ctor), :27 CreateOnlineStream,
                          :32 GetResult, :41 GetResults, :403-436 DecodeMulti (ids -> lower-cased text)
  * ``OnlineStream``      OnlineStream.cs:67-84 (Tokens / AddSamples), :279 IsFinished
  * result entity         Model/OnlineRecognizerResultEntity.cs:9-28

All per-stream model state (fbank FIFO, splice frame, feature cache, CIF carry, FSMN caches) lives on the GPU inside
the handle; this module keeps only what the reference's host keeps for the user: the token list and the text decode.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from .engine import to_pf_config
from .offline import ObjectDisposedError, load_cmvn, load_conf, read_tokens
from .text import TokenTable
from .synth import ModelConfig
from .weights import pack



@dataclass
class OnlineRecognizerResultEntity:
    """Model/OnlineRecognizerResultEntity.cs:9-28"""
    text: str = ""

    @property
    def Text(self):
        return self.text


@dataclass
class StepOutput:
    """One ``Forward`` (OnlineRecognizer.cs:341-401) over the listed streams."""
    max_new: int
    n_working: int
    appended: np.ndarray            # [n] ids appended to each stream
    new_tokens: np.ndarray          # [n, max_new]
    embeds_len: np.ndarray          # [n] acoustic_embeds_len
    logits: Optional[np.ndarray] = None


class OnlineEngine:
    """One ``pf_online`` handle (what ``OnlineModel`` + the three ``*Proj`` methods are in the reference)."""

    def __init__(self, cfg: ModelConfig, weights: Union[str, Dict[str, np.ndarray], np.ndarray],
                 devices: Optional[Sequence[int]] = None, per_layer_cache: bool = False):
        self._lib = _lib.load()
        self.cfg = cfg
        self._h = C.c_void_p()
        pcfg = to_pf_config(cfg)
        pcfg.online_flags = 1 if per_layer_cache else 0        # 0 = bug-compatible stack_states (Q11)
        dev_arr, ndev = None, 0
        if devices is not None:
            dev_np = np.asarray(list(devices), dtype=np.int32)
            dev_arr, ndev = _lib.iptr(dev_np), len(dev_np)
        if isinstance(weights, str):
            st = self._lib.pf_online_create(C.byref(pcfg), weights.encode(), dev_arr, ndev, C.byref(self._h))
        else:
            blob = pack(weights) if isinstance(weights, dict) else np.ascontiguousarray(weights, dtype=np.uint8)
            st = self._lib.pf_online_create_from_memory(C.byref(pcfg), blob.ctypes.data_as(C.c_void_p), blob.nbytes,
                                                        dev_arr, ndev, C.byref(self._h))
        _lib.check(st)

    def close(self) -> None:
        if self._h:
            self._lib.pf_online_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise _lib.PfError(_lib.PF_ERR_DISPOSED, "engine disposed")
        return self._h

    def set_cmvn(self, add_shift, rescale) -> None:
        a = np.ascontiguousarray(add_shift, dtype=np.float32)
        b = np.ascontiguousarray(rescale, dtype=np.float32)
        _lib.check(self._lib.pf_online_set_cmvn(self._handle(), _lib.fptr(a), _lib.fptr(b), a.shape[0]))

    def open_stream(self) -> int:
        sid = C.c_int32(-1)
        _lib.check(self._lib.pf_online_stream_open(self._handle(), C.byref(sid)))
        return int(sid.value)

    def close_stream(self, sid: int) -> None:
        _lib.check(self._lib.pf_online_stream_close(self._handle(), int(sid)))

    def push(self, sid: int, samples: np.ndarray) -> None:
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        _lib.check(self._lib.pf_online_stream_push(self._handle(), int(sid), _lib.fptr(x), x.shape[0]))

    def ready(self, sid: int) -> bool:
        return self._lib.pf_online_stream_ready(self._handle(), int(sid)) == 1

    def step(self, sids: Sequence[int], want_logits: bool = False) -> StepOutput:
        ids = np.asarray(list(sids), dtype=np.int32)
        res = _lib.PfOnlineResult()
        flags = _lib.PF_RUN_WANT_LOGITS if want_logits else 0
        _lib.check(self._lib.pf_online_step(self._handle(), _lib.iptr(ids), len(ids), flags, C.byref(res)))
        n, l, v = res.n_streams, res.max_new, res.vocab
        out = StepOutput(max_new=l, n_working=res.n_working,
                         appended=np.ctypeslib.as_array(res.appended, shape=(n,)).copy() if n else np.zeros(0, np.int32),
                         new_tokens=(np.ctypeslib.as_array(res.new_tokens, shape=(n, l)).copy() if n and l else np.zeros((n, 0), np.int32)),
                         embeds_len=np.ctypeslib.as_array(res.embeds_len, shape=(n,)).copy() if n else np.zeros(0, np.int32))
        if want_logits and res.logits and l:
            out.logits = np.ctypeslib.as_array(res.logits, shape=(n, l, v)).copy()
        return out

    def state(self, sid: int, name: str) -> np.ndarray:
        sizes = {"cache_feats": (10, self.cfg.input_size), "cif_alpha": (1,), "cif_hidden": (self.cfg.d_model,),
                 "splice": (self.cfg.n_mels,), "fsmn": (self.cfg.dec_layers, self.cfg.dec_kernel - 1, self.cfg.d_model)}
        out = np.zeros(sizes[name], dtype=np.float32)
        _lib.check(self._lib.pf_online_get_state(self._handle(), int(sid), name.encode(), _lib.fptr(out), out.size))
        return out

    def timings(self) -> Dict[str, float]:
        ms = (C.c_float * 6)()
        self._lib.pf_online_get_timings(self._handle(), ms, 6)
        keys = ["fbank", "assemble_encoder", "predictor_cif", "decoder", "head_pick", "total"]
        return {k: float(ms[i]) for i, k in enumerate(keys)}

    def launch_count(self) -> int:
        return int(self._lib.pf_online_get_launch_count(self._handle()))

    def gemm_flops(self) -> float:
        return float(self._lib.pf_online_get_gemm_flops(self._handle()))


class OnlineStream:
    """OnlineStream.cs: the user-visible half (token list, AddSamples); the model state is device resident."""

    def __init__(self, recognizer: "OnlineRecognizer"):
        self._rec = recognizer
        self._sid = recognizer._engine.open_stream()
        self.tokens: List[int] = [0, 0]                 # OnlineStream.cs:54 {_blank_id, _blank_id}
        self._disposed = False

    @property
    def Tokens(self):
        return self.tokens

    def add_samples(self, samples) -> None:
        """OnlineStream.AddSamples (OnlineStream.cs:84-112); null samples fail like ``samples.Length`` does."""
        if samples is None:
            raise TypeError("Object reference not set to an instance of an object (samples)")
        self._rec._engine.push(self._sid, samples)

    AddSamples = add_samples

    def dispose(self) -> None:
        if not self._disposed:
            self._rec._engine.close_stream(self._sid)
            self._disposed = True

    Dispose = dispose


class OnlineRecognizer:
    """OnlineRecognizer.cs:22 — ``encoder_file_path`` is the PFW1 blob replacing encoder.onnx + decoder.onnx
    (``decoder_file_path`` is accepted for signature compatibility and may be empty)."""

    def __init__(self, encoder_file_path: str, decoder_file_path: str, config_file_path: str, mvn_file_path: str,
                 tokens_file_path: str, threads_num: int = 1, devices: Optional[Sequence[int]] = None, weights=None,
                 config: Optional[ModelConfig] = None, per_layer_cache: bool = False):
        self._disposed = False
        self._conf = config if config is not None else load_conf(config_file_path)
        self._tokens = read_tokens(tokens_file_path)
        # null when the path is empty: the reference then fails inside DecodeMulti (NullReferenceException)
        self._token_table = TokenTable(path=tokens_file_path) if self._tokens else None
        self._engine = OnlineEngine(self._conf, weights if weights is not None else encoder_file_path, devices=devices,
                                    per_layer_cache=per_layer_cache)
        if mvn_file_path:
            shift, scale = load_cmvn(mvn_file_path)
            self._engine.set_cmvn(shift, scale)

    def create_online_stream(self) -> OnlineStream:
        if self._disposed:
            raise ObjectDisposedError("OnlineRecognizer")
        return OnlineStream(self)

    def get_result(self, stream: OnlineStream) -> OnlineRecognizerResultEntity:
        return self.get_results([stream])[0]

    def get_results(self, streams: List[OnlineStream]) -> List[OnlineRecognizerResultEntity]:
        self._forward(streams)
        return self._decode_multi(streams)

    def dispose(self) -> None:
        if not self._disposed:
            self._engine.close()
            if self._token_table is not None:
                self._token_table.close()
            self._disposed = True

    CreateOnlineStream = create_online_stream
    GetResult = get_result
    GetResults = get_results
    Dispose = dispose

    # OnlineRecognizer.Forward (OnlineRecognizer.cs:341-401)
    def _forward(self, streams: List[OnlineStream]) -> None:
        if not streams:
            return
        try:
            out = self._engine.step([s._sid for s in streams])
        except _lib.PfError as ex:
            raise Exception("Online recognition failed") from ex       # OnlineRecognizer.cs:396-399
        for i, s in enumerate(streams):
            s.tokens.extend(int(t) for t in out.new_tokens[i, : int(out.appended[i])])

    # OnlineRecognizer.DecodeMulti (OnlineRecognizer.cs:403-436), native: pf_decode_online (csrc/text.cu)
    def _decode_multi(self, streams: List[OnlineStream]) -> List[OnlineRecognizerResultEntity]:
        if self._token_table is None:
            raise TypeError("tokens table is null")
        return [OnlineRecognizerResultEntity(text=self._token_table.decode_online(s.tokens)) for s in streams]
