"""Shared helpers for the parity tests (oracle side = torch CPU fp32)."""
import ctypes as C

import numpy as np

from aliparaformerasr_b200 import _lib, synth
from oracle import sanm


def dims_of(cfg: synth.ModelConfig) -> sanm.ModelDims:
    return sanm.ModelDims(**{k: v for k, v in cfg.as_dict().items() if k in sanm.ModelDims.__dataclass_fields__})


def f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def half_round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def dbg_gemm(lib, A, W, bias=None, resid=None, addend=None, relu=0, out_half=0, tile_n=0, iters=0):
    M, K = A.shape
    N = W.shape[0]
    out = np.zeros((M, N), dtype=np.float32)
    ms = C.c_float(0)
    A, W = f(A), f(W)
    keep = [f(x) if x is not None else None for x in (bias, resid, addend)]
    ptr = [(_lib.fptr(x) if x is not None else None) for x in keep]
    _lib.check(lib.pf_dbg_gemm(M, N, K, _lib.fptr(A), _lib.fptr(W), ptr[0], ptr[1], ptr[2], relu, out_half, tile_n,
                               _lib.fptr(out), C.byref(ms), iters))
    return out, ms.value


def margins(logp):
    s = np.sort(logp, axis=-1)
    return s[..., -1] - s[..., -2]


def dbg_ffn_chain(lib, a, w1, b1, w2, b2, x, iters=0):
    M, D = a.shape
    F = w1.shape[0]
    out = np.zeros((M, D), dtype=np.float32)
    ms = C.c_float(0)
    a, w1, b1, w2, b2, x = (f(v) for v in (a, w1, b1, w2, b2, x))
    _lib.check(lib.pf_dbg_ffn_chain(M, D, F, _lib.fptr(a), _lib.fptr(w1), _lib.fptr(b1), _lib.fptr(w2), _lib.fptr(b2), _lib.fptr(x),
                                    _lib.fptr(out), C.byref(ms), iters))
    return out, ms.value
