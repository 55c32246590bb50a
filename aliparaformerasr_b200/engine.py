"""Thin host wrapper over the C-ABI handle (``pf_offline``): what ``OfflineProjOfCuda : IOfflineProj`` is on the
C# side (csharp/OfflineProjOfCuda.cs).  Device work happens in libpfasr.so; this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from .synth import ModelConfig
from .weights import pack


def to_pf_config(cfg: ModelConfig) -> _lib.PfConfig:
    kinds = {"paraformer": _lib.PF_MODEL_PARAFORMER, "sensevoicesmall": _lib.PF_MODEL_SENSEVOICE_SMALL,
             "seacoparaformer": _lib.PF_MODEL_SEACO_PARAFORMER}
    # OfflineRecognizer.cs:39-53: unknown model names fall back to the paraformer projection
    kind = kinds.get(cfg.model.lower(), _lib.PF_MODEL_PARAFORMER)
    c = _lib.PfConfig()
    c.struct_bytes = C.sizeof(_lib.PfConfig)
    c.model_kind = kind
    for f in ("input_size", "d_model", "heads", "ffn", "enc_layers", "tp_layers", "enc_kernel", "dec_layers",
              "dec_ffn", "dec_kernel", "vocab", "fs", "n_mels", "lfr_m", "lfr_n"):
        setattr(c, f, int(getattr(cfg, f)))
    for f in ("ln_eps", "cif_threshold", "cif_tail", "smooth_factor", "noise_threshold"):
        setattr(c, f, float(getattr(cfg, f)))
    c.seaco_layers, c.seaco_ffn, c.seaco_kernel = int(cfg.seaco_layers), int(cfg.seaco_ffn), int(cfg.seaco_kernel)
    c.seaco_nobias_id = int(cfg.nobias_id)
    c.smooth_factor2, c.noise_threshold2 = float(cfg.smooth_factor2), float(cfg.noise_threshold2)
    c.snip_edges = int(bool(cfg.snip_edges))
    c.use_itn = int(bool(cfg.use_itn))
    return c


@dataclass
class ModelOutput:
    """``ModelOutputEntity`` (Model/ModelOutputEntity.cs:10-19) + the greedy ids of OfflineRecognizer.cs:139-152."""
    tokens: np.ndarray                 # [B, L] int32
    token_num: np.ndarray              # [B] int32  (model_out_lens)
    feat_frames: int
    logits: Optional[np.ndarray] = None      # [B, L, V] log-probs (model_out) when requested
    cif_peak: Optional[np.ndarray] = None
    us_alphas: Optional[np.ndarray] = None   # [B, 3T] (models with the CifPredictorV3 timestamp branch, on request)
    us_cif_peak: Optional[np.ndarray] = None


class Engine:
    """One ``pf_offline`` handle."""

    def __init__(self, cfg: ModelConfig, weights: Union[str, Dict[str, np.ndarray], np.ndarray],
                 devices: Optional[Sequence[int]] = None, lanes: int = 1):
        self._lib = _lib.load()
        self.cfg = cfg
        self._h = C.c_void_p()
        pcfg = to_pf_config(cfg)
        dev_arr = None
        ndev = 0
        if devices is not None:
            dev_np = np.asarray(list(devices), dtype=np.int32)
            dev_arr, ndev = _lib.iptr(dev_np), len(dev_np)
        if isinstance(weights, str) and weights.lower().endswith(".onnx"):
            # the reference's own model file (OfflineModel.cs:35-70 loads model.onnx / model.int8.onnx into ORT): read the
            # initialisers, de-quantise, map them to FunASR names (onnx_weights.py) and hand the blob to the library
            from . import onnx_weights
            graph = onnx_weights.read_onnx(weights)
            if pcfg.model_kind == _lib.PF_MODEL_PARAFORMER:
                weights = onnx_weights.paraformer_state_dict(graph, cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn, cfg.input_size,
                                                             cfg.dec_ffn, cfg.vocab)
            elif pcfg.model_kind == _lib.PF_MODEL_SENSEVOICE_SMALL:
                weights = onnx_weights.sensevoice_state_dict(graph, cfg.enc_layers, cfg.tp_layers, cfg.d_model, cfg.ffn, cfg.input_size, cfg.vocab)
            else:
                raise NotImplementedError("ONNX ingestion is mapped for the paraformer and SenseVoiceSmall exports; convert SeACo models to a PFW1 blob")
        if pcfg.model_kind == _lib.PF_MODEL_SENSEVOICE_SMALL and isinstance(weights, dict) and "embed.weight" not in weights:
            # split-embed export: the prompt rows come from the reference's own data/embed.onnx (EmbedSVModel.cs:45-77,
            # OfflineProjOfSenseVoiceSmall.cs:84-100), shipped as sha256-pinned package data
            from . import onnx_weights
            weights = dict(weights)
            weights["embed.weight"] = onnx_weights.packaged_sensevoice_embed()
        # lanes > 1: calls from different host threads run concurrently on the GPU (pf_offline_create_mt)
        if isinstance(weights, str):
            st = self._lib.pf_offline_create_mt(C.byref(pcfg), weights.encode(), dev_arr, ndev, int(lanes), C.byref(self._h))
        else:
            blob = pack(weights) if isinstance(weights, dict) else np.ascontiguousarray(weights, dtype=np.uint8)
            st = self._lib.pf_offline_create_from_memory_mt(C.byref(pcfg), blob.ctypes.data_as(C.c_void_p), blob.nbytes,
                                                            dev_arr, ndev, int(lanes), C.byref(self._h))
        _lib.check(st)

    # -- lifecycle
    def close(self) -> None:
        if self._h:
            self._lib.pf_offline_destroy(self._h)
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

    def lease(self):
        """Context manager: the calling thread's execution lane stays locked for the enclosed calls (per-call hot words:
        set -> run -> restore must not interleave with another thread sharing the lane; pf_offline_lane_acquire)."""
        import contextlib

        @contextlib.contextmanager
        def _lease():
            if self._lib.pf_offline_lane_acquire(self._handle()) < 0:
                raise _lib.PfError(_lib.PF_ERR_DISPOSED, "engine disposed")
            try:
                yield self
            finally:
                self._lib.pf_offline_lane_release(self._handle())
        return _lease()

    def set_cmvn(self, add_shift: np.ndarray, rescale: np.ndarray) -> None:
        a = np.ascontiguousarray(add_shift, dtype=np.float32)
        b = np.ascontiguousarray(rescale, dtype=np.float32)
        _lib.check(self._lib.pf_offline_set_cmvn(self._handle(), _lib.fptr(a), _lib.fptr(b), a.shape[0]))

    def set_hotwords(self, hotwords: Sequence[Sequence[int]], local: bool = False) -> None:
        """``EmbedSeacoModel.Forward`` + the Q8 bias_embed assembly; ``hotwords`` = id lists (PadList applied here:
        truncate to 10, pad with 0, EmbedSeacoModel.cs:110-123).  Empty list clears them."""
        ids = np.zeros((len(hotwords), 10), dtype=np.int32)
        for i, h in enumerate(hotwords):
            h = list(h)[:10]
            ids[i, : len(h)] = h
        # local: only the calling thread's execution lane (per-call hot words while other threads use the handle)
        fn = self._lib.pf_offline_set_hotwords_local if local else self._lib.pf_offline_set_hotwords
        _lib.check(fn(self._handle(), _lib.iptr(ids) if len(hotwords) else None, len(hotwords)))

    # -- front-end only (OfflineStream.AddSamples)
    def num_frames(self, nsamp: int) -> int:
        return int(self._lib.pf_frontend_num_frames(self._handle(), int(nsamp)))

    def extract(self, samples: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        t = max(self.num_frames(x.shape[0]), 0)
        out = np.zeros((t, self.cfg.lfr_m * self.cfg.n_mels), dtype=np.float32)
        n = C.c_int32(0)
        _lib.check(self._lib.pf_frontend_extract(self._handle(), _lib.fptr(x), x.shape[0], _lib.fptr(out), t, C.byref(n)))
        return out[: n.value]

    def fbank(self, samples: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        cap = x.shape[0] // 160 + 2
        out = np.zeros((cap, self.cfg.n_mels), dtype=np.float32)
        n = C.c_int32(0)
        _lib.check(self._lib.pf_frontend_fbank(self._handle(), _lib.fptr(x), x.shape[0], _lib.fptr(out), cap, C.byref(n)))
        return out[: n.value]

    # -- hot path
    def _collect(self, res: _lib.PfResult, want_logits: bool, want_peak: bool) -> ModelOutput:
        b, l, v = res.batch, res.max_len, res.vocab
        tokens = np.ctypeslib.as_array(res.tokens, shape=(b, l)).copy() if l > 0 else np.zeros((b, 0), np.int32)
        token_num = np.ctypeslib.as_array(res.token_num, shape=(b,)).copy()
        out = ModelOutput(tokens=tokens, token_num=token_num, feat_frames=res.feat_frames)
        if want_logits and res.logits:
            out.logits = np.ctypeslib.as_array(res.logits, shape=(b, l, v)).copy()
        if want_peak and res.cif_peak:
            out.cif_peak = np.ctypeslib.as_array(res.cif_peak, shape=(b, res.feat_frames + 1)).copy()
        if res.us_frames > 0 and res.us_cif_peak:
            out.us_alphas = np.ctypeslib.as_array(res.us_alphas, shape=(b, res.us_frames)).copy()
            out.us_cif_peak = np.ctypeslib.as_array(res.us_cif_peak, shape=(b, res.us_frames)).copy()
        return out

    @staticmethod
    def _pcm_args(pcm: Sequence[np.ndarray]):
        arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in pcm]
        ptrs = (C.POINTER(C.c_float) * len(arrs))(*[_lib.fptr(a) for a in arrs])
        ns = np.asarray([a.shape[0] for a in arrs], dtype=np.int32)
        return arrs, ptrs, ns

    def run_pcm(self, pcm: Sequence[np.ndarray], want_logits: bool = False, want_cif_peak: bool = False,
                want_timestamps: bool = False) -> ModelOutput:
        arrs, ptrs, ns = self._pcm_args(pcm)
        flags = ((_lib.PF_RUN_WANT_LOGITS if want_logits else 0) | (_lib.PF_RUN_WANT_CIF_PEAK if want_cif_peak else 0) |
                 (_lib.PF_RUN_WANT_TIMESTAMPS if want_timestamps else 0))
        res = _lib.PfResult()
        _lib.check(self._lib.pf_offline_run_pcm(self._handle(), ptrs, _lib.iptr(ns), len(arrs), flags, C.byref(res)))
        return self._collect(res, want_logits, want_cif_peak)

    def run_audio(self, clips: Sequence["Audio"], want_logits: bool = False, want_cif_peak: bool = False,
                  want_timestamps: bool = False) -> ModelOutput:
        """GetFileSample + AddSamples + Forward on raw file samples (``audio.Audio``): conversion to float, the
        stereo down-mix and the resample to 16 kHz run on the device (csrc/audio.cu)."""
        table = (_lib.PfAudio * len(clips))(*[c.as_pf_audio() for c in clips])
        flags = ((_lib.PF_RUN_WANT_LOGITS if want_logits else 0) | (_lib.PF_RUN_WANT_CIF_PEAK if want_cif_peak else 0) |
                 (_lib.PF_RUN_WANT_TIMESTAMPS if want_timestamps else 0))
        res = _lib.PfResult()
        _lib.check(self._lib.pf_offline_run_audio(self._handle(), table, len(clips), flags, C.byref(res)))
        return self._collect(res, want_logits, want_cif_peak)

    def run_feats(self, speech: np.ndarray, want_logits: bool = False, want_cif_peak: bool = False,
                  want_timestamps: bool = False) -> ModelOutput:
        x = np.ascontiguousarray(speech, dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != self.cfg.input_size:
            raise ValueError("speech must be [B, T, input_size]")
        flags = ((_lib.PF_RUN_WANT_LOGITS if want_logits else 0) | (_lib.PF_RUN_WANT_CIF_PEAK if want_cif_peak else 0) |
                 (_lib.PF_RUN_WANT_TIMESTAMPS if want_timestamps else 0))
        res = _lib.PfResult()
        _lib.check(self._lib.pf_offline_run_feats(self._handle(), _lib.fptr(x), x.shape[0], x.shape[1], flags, C.byref(res)))
        return self._collect(res, want_logits, want_cif_peak)

    def stage_pcm(self, pcm: Sequence[np.ndarray]) -> None:
        arrs, ptrs, ns = self._pcm_args(pcm)
        self._staged_keepalive = arrs
        _lib.check(self._lib.pf_offline_stage_pcm(self._handle(), ptrs, _lib.iptr(ns), len(arrs)))

    def run_staged(self, want_logits: bool = False, want_timestamps: bool = False) -> ModelOutput:
        res = _lib.PfResult()
        flags = (_lib.PF_RUN_WANT_LOGITS if want_logits else 0) | (_lib.PF_RUN_WANT_TIMESTAMPS if want_timestamps else 0)
        _lib.check(self._lib.pf_offline_run_staged(self._handle(), flags, C.byref(res)))
        return self._collect(res, want_logits, False)

    # -- introspection
    def tensor(self, name: str, dev_index: int = 0) -> np.ndarray:
        dims = (C.c_int32 * 4)()
        nd = C.c_int32(0)
        _lib.check(self._lib.pf_offline_get_tensor(self._handle(), dev_index, name.encode(), None, 0, dims, C.byref(nd)))
        shape = tuple(int(dims[i]) for i in range(nd.value))
        out = np.zeros(shape, dtype=np.float32)
        _lib.check(self._lib.pf_offline_get_tensor(self._handle(), dev_index, name.encode(), _lib.fptr(out), out.size, dims, C.byref(nd)))
        return out

    def timings(self) -> Dict[str, float]:
        ms = (C.c_float * 10)()
        self._lib.pf_offline_get_timings(self._handle(), ms, 10)
        keys = ["h2d_frontend", "encoder", "predictor_cif", "decoder", "head_pick", "total",
                "host_enqueued_1", "host_counts_ready", "host_enqueued_2", "host_done"]
        return {k: float(ms[i]) for i, k in enumerate(keys)}

    def set_profile(self, on: int) -> None:
        _lib.check(self._lib.pf_offline_set_profile(self._handle(), int(on)))

    def gemm_ms(self) -> float:
        return float(self._lib.pf_offline_get_gemm_ms(self._handle()))

    def replay_gemms(self, iters: int = 5) -> float:
        """ms per pass over the GEMMs of the last ``set_profile(1)`` run, launched back to back (see pf_abi.h)."""
        return float(self._lib.pf_offline_replay_gemms(self._handle(), int(iters)))

    def profile(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        n = self._lib.pf_offline_get_profile_json(self._handle(), buf, len(buf))
        return json.loads(buf.value[:n].decode()) if n else []

    def stream_ptr(self, dev_index: int = 0) -> int:
        return int(self._lib.pf_offline_get_stream(self._handle(), dev_index) or 0)

    def launch_count(self) -> int:
        return int(self._lib.pf_offline_get_launch_count(self._handle()))

    def gemm_flops(self) -> float:
        return float(self._lib.pf_offline_get_gemm_flops(self._handle()))
