"""ctypes binding of ``libpfasr.so`` (include/pf_abi.h).

This is the same stub a C# maintainer writes with ``[DllImport("pfasr")]``; Python is used
here because the build image has no dotnet.  Loading fails loudly when the CUDA library has not been built: there
is no CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libpfasr.so"

PF_OK = 0
PF_ERR_BAD_ARG = -1
PF_ERR_CUDA = -2
PF_ERR_OOM = -3
PF_ERR_SHAPE = -4
PF_ERR_DISPOSED = -5
PF_ERR_WEIGHTS = -6
PF_ERR_UNSUPPORTED = -7

PF_MODEL_PARAFORMER = 0
PF_MODEL_SENSEVOICE_SMALL = 1
PF_MODEL_SEACO_PARAFORMER = 2
PF_RUN_WANT_LOGITS = 1
PF_RUN_WANT_CIF_PEAK = 2
PF_RUN_WANT_TIMESTAMPS = 4


class PfConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("model_kind", C.c_int32), ("input_size", C.c_int32), ("d_model", C.c_int32),
        ("heads", C.c_int32), ("ffn", C.c_int32), ("enc_layers", C.c_int32), ("tp_layers", C.c_int32),
        ("enc_kernel", C.c_int32), ("dec_layers", C.c_int32), ("dec_ffn", C.c_int32), ("dec_kernel", C.c_int32),
        ("vocab", C.c_int32), ("ln_eps", C.c_float), ("cif_threshold", C.c_float), ("cif_tail", C.c_float),
        ("smooth_factor", C.c_float), ("noise_threshold", C.c_float), ("fs", C.c_int32), ("n_mels", C.c_int32),
        ("lfr_m", C.c_int32), ("lfr_n", C.c_int32), ("snip_edges", C.c_int32), ("use_itn", C.c_int32),
        ("online_flags", C.c_int32), ("seaco_layers", C.c_int32), ("seaco_ffn", C.c_int32),
        ("seaco_kernel", C.c_int32), ("seaco_nobias_id", C.c_int32), ("smooth_factor2", C.c_float), ("noise_threshold2", C.c_float),
    ]


class PfResult(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("max_len", C.c_int32), ("vocab", C.c_int32), ("feat_frames", C.c_int32),
        ("tokens", C.POINTER(C.c_int32)), ("token_num", C.POINTER(C.c_int32)),
        ("logits", C.POINTER(C.c_float)), ("cif_peak", C.POINTER(C.c_float)),
        ("us_frames", C.c_int32), ("us_alphas", C.POINTER(C.c_float)), ("us_cif_peak", C.POINTER(C.c_float)),
    ]


class PfOnlineResult(C.Structure):
    _fields_ = [
        ("n_streams", C.c_int32), ("max_new", C.c_int32), ("vocab", C.c_int32), ("n_working", C.c_int32),
        ("appended", C.POINTER(C.c_int32)), ("new_tokens", C.POINTER(C.c_int32)), ("embeds_len", C.POINTER(C.c_int32)),
        ("logits", C.POINTER(C.c_float)),
    ]


PF_AUDIO_U8, PF_AUDIO_S16, PF_AUDIO_S24, PF_AUDIO_S32, PF_AUDIO_F32 = range(5)


class PfAudio(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_values", C.c_int64), ("format", C.c_int32), ("channels", C.c_int32),
                ("sample_rate", C.c_int32), ("reserved", C.c_int32)]


class PfTextResult(C.Structure):
    _fields_ = [
        ("text", C.c_void_p), ("text_capacity", C.c_size_t), ("text_bytes", C.c_size_t),
        ("text_len", C.c_int32), ("n_tokens", C.c_int32), ("n_timestamps", C.c_int32), ("reserved", C.c_int32),
        ("tokens", C.c_void_p), ("tokens_capacity", C.c_size_t), ("tokens_bytes", C.c_size_t),
        ("ts", C.POINTER(C.c_int32)), ("ts_capacity", C.c_size_t), ("ts_count", C.c_size_t),
        ("ts_offsets", C.POINTER(C.c_int32)), ("ts_offsets_capacity", C.c_size_t),
    ]


class PfError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libpfasr error {code}: {message}")
        self.code = code
        self.message = message


_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/pf_abi.h declares
SIGNATURES = {
    "pf_offline_create": (C.c_int32, [C.POINTER(PfConfig), C.c_char_p, _I, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_offline_create_from_memory": (C.c_int32, [C.POINTER(PfConfig), C.c_void_p, C.c_size_t, _I, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_offline_create_mt": (C.c_int, [C.POINTER(PfConfig), C.c_char_p, _I, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_offline_create_from_memory_mt": (C.c_int, [C.POINTER(PfConfig), C.c_void_p, C.c_size_t, _I, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_build_experiments": (C.c_int32, []),
    "pf_offline_lanes": (C.c_int32, [C.c_void_p]),
    "pf_offline_lane_acquire": (C.c_int32, [C.c_void_p]),
    "pf_offline_lane_release": (C.c_int32, [C.c_void_p]),
    "pf_offline_destroy": (C.c_int32, [C.c_void_p]),
    "pf_offline_set_cmvn": (C.c_int32, [C.c_void_p, _F, _F, C.c_int32]),
    "pf_offline_set_hotwords": (C.c_int32, [C.c_void_p, _I, C.c_int32]),
    "pf_offline_set_hotwords_local": (C.c_int32, [C.c_void_p, _I, C.c_int32]),
    "pf_frontend_extract": (C.c_int32, [C.c_void_p, _F, C.c_int32, _F, C.c_int32, _I]),
    "pf_frontend_fbank": (C.c_int32, [C.c_void_p, _F, C.c_int32, _F, C.c_int32, _I]),
    "pf_frontend_num_frames": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pf_offline_run_pcm": (C.c_int32, [C.c_void_p, C.POINTER(_F), _I, C.c_int32, C.c_uint32, C.POINTER(PfResult)]),
    "pf_offline_run_feats": (C.c_int32, [C.c_void_p, _F, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(PfResult)]),
    "pf_offline_stage_pcm": (C.c_int32, [C.c_void_p, C.POINTER(_F), _I, C.c_int32]),
    "pf_offline_run_staged": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(PfResult)]),
    "pf_offline_get_tensor": (C.c_int32, [C.c_void_p, C.c_int32, C.c_char_p, _F, C.c_size_t, _I, _I]),
    "pf_offline_get_timings": (C.c_int32, [C.c_void_p, _F, C.c_int32]),
    "pf_offline_get_launch_count": (C.c_int64, [C.c_void_p]),
    "pf_offline_get_gemm_flops": (C.c_double, [C.c_void_p]),
    "pf_offline_set_profile": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pf_offline_get_gemm_ms": (C.c_double, [C.c_void_p]),
    "pf_offline_replay_gemms": (C.c_double, [C.c_void_p, C.c_int32]),
    "pf_offline_get_profile_json": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int32]),
    "pf_offline_get_stream": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "pf_online_create": (C.c_int32, [C.POINTER(PfConfig), C.c_char_p, _I, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_online_create_from_memory": (C.c_int32, [C.POINTER(PfConfig), C.c_void_p, C.c_size_t, _I, C.c_int32, C.POINTER(C.c_void_p)]),
    "pf_online_destroy": (C.c_int32, [C.c_void_p]),
    "pf_online_set_cmvn": (C.c_int32, [C.c_void_p, _F, _F, C.c_int32]),
    "pf_online_stream_open": (C.c_int32, [C.c_void_p, _I]),
    "pf_online_stream_close": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pf_online_stream_push": (C.c_int32, [C.c_void_p, C.c_int32, _F, C.c_int32]),
    "pf_online_stream_ready": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pf_online_step": (C.c_int32, [C.c_void_p, _I, C.c_int32, C.c_uint32, C.POINTER(PfOnlineResult)]),
    "pf_online_get_state": (C.c_int32, [C.c_void_p, C.c_int32, C.c_char_p, _F, C.c_size_t]),
    "pf_online_get_timings": (C.c_int32, [C.c_void_p, _F, C.c_int32]),
    "pf_online_get_launch_count": (C.c_int64, [C.c_void_p]),
    "pf_online_get_gemm_flops": (C.c_double, [C.c_void_p]),
    "pf_wav_parse": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(PfAudio)]),
    "pf_audio_num_samples": (C.c_int64, [C.POINTER(PfAudio)]),
    "pf_offline_run_audio": (C.c_int, [C.c_void_p, C.POINTER(PfAudio), C.c_int32, C.c_uint32, C.POINTER(PfResult)]),
    "pf_dbg_audio_convert": (C.c_int, [C.POINTER(PfAudio), _F, C.c_int64, C.POINTER(C.c_int64)]),
    "pf_tokens_create": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "pf_tokens_create_from_memory": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "pf_tokens_destroy": (C.c_int, [C.c_void_p]),
    "pf_tokens_count": (C.c_int32, [C.c_void_p]),
    "pf_tokens_get": (C.c_int32, [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]),
    "pf_timestamps_lfr6": (C.c_int, [_F, C.c_int32, _I, C.c_int32, C.c_float, C.c_float, _I, C.c_int32, _I]),
    "pf_decode_offline": (C.c_int, [C.c_void_p, _I, C.c_int32, _I, C.c_int32, C.POINTER(PfTextResult)]),
    "pf_decode_offline_result": (C.c_int, [C.c_void_p, C.POINTER(PfResult), C.c_int32, C.POINTER(PfTextResult)]),
    "pf_decode_online": (C.c_int, [C.c_void_p, _I, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "pf_last_error": (C.c_char_p, []),
    "pf_abi_version": (C.c_int32, []),
    "pf_dbg_gemm": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _F, _F, C.c_int32, C.c_int32, C.c_int32, _F, _F, C.c_int32]),
    "pf_dbg_ln_gemm": (C.c_int32, [C.c_int32, C.c_int32, _F, _F, _F, C.c_float, _F, _F, C.c_int32, _F, _F, C.c_int32]),
    "pf_dbg_gemm_pick": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, C.c_int32, C.POINTER(C.c_int32)]),
    "pf_dbg_ffn_chain": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _F, _F, _F, _F, _F, C.c_int32]),
    "pf_dbg_gemm_ln": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _F, _F, _F, C.c_float, _F, _F]),
    "pf_dbg_layernorm": (C.c_int32, [C.c_int32, C.c_int32, _F, _F, _F, C.c_float, _F]),
    "pf_dbg_embed_pe_ln": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, C.c_float, _F, _F, C.c_float, _F]),
    "pf_dbg_attention": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _F]),
    "pf_dbg_attention_fsmn": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _F]),
    "pf_dbg_fsmn": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F, _F, _F, _I, C.c_int32, _F]),
    "pf_dbg_cif": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _F, _F, C.c_float, C.c_int32, _F, _I, _I, _F]),
    "pf_dbg_logsoftmax_argmax": (C.c_int32, [C.c_int32, C.c_int32, _F, _I]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpfasr.so and attach the signatures.  Raises if the CUDA library is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PFASR_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m aliparaformerasr_b200.build` "
            "(nvcc, sm_100a). aliparaformerasr_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    variant = "PFASR_LIB" in os.environ            # an A/B build of another revision (scripts/build_variant.py) may lack newer hooks
    for name, (res, args) in SIGNATURES.items():
        if variant and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != PF_OK:
        raise PfError(status, load().pf_last_error().decode(errors="replace"))


def fptr(a):
    return a.ctypes.data_as(_F)


def iptr(a):
    return a.ctypes.data_as(_I)
