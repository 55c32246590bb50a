"""Host post-processing through libpfasr (include/pf_abi.h, "host post-processing" block; csrc/text.cu): the tokens
table, ``DecodeMulti`` of both recognisers and ``time_stamp_lfr6_onnx``.

Replaces the managed loops of OfflineRecognizer.cs:200-302 / :304-418 and OnlineRecognizer.cs:403-436.  Error behaviour
follows the C#: an id outside the tokens table, an empty ``us_cif_peak`` fire list or more fires than tokens raise
``IndexError`` (IndexOutOfRangeException / ArgumentOutOfRangeException there)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

_I = C.POINTER(C.c_int32)
_F = C.POINTER(C.c_float)


def _check(status: int) -> None:
    if status == _lib.PF_ERR_SHAPE:
        raise IndexError(_lib.load().pf_last_error().decode(errors="replace"))
    _lib.check(status)


def _ids(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).astype(np.int32).reshape(-1))


class TokenTable:
    """``string[] _tokens`` of the recognisers, held by the native library."""

    def __init__(self, path: Optional[str] = None, lines: Optional[Sequence[str]] = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if path is not None:
            _check(self._lib.pf_tokens_create(path.encode(), C.byref(self._h)))
        else:
            data = "\n".join(lines or []).encode("utf-8")
            _check(self._lib.pf_tokens_create_from_memory(data, len(data), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self._lib.pf_tokens_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.pf_tokens_count(self._h))

    def __getitem__(self, i: int) -> str:
        n = self._lib.pf_tokens_get(self._h, int(i), None, 0)
        if n < 0:
            raise IndexError(i)
        buf = C.create_string_buffer(n + 1)
        self._lib.pf_tokens_get(self._h, int(i), buf, n + 1)
        return buf.value.decode("utf-8")

    def lines(self) -> List[str]:
        return [self[i] for i in range(len(self))]

    # -- OfflineRecognizer.DecodeMulti, one stream
    def decode_offline(self, ids, timestamps=None) -> Tuple[str, int, List[str], List[List[int]]]:
        """-> (Text, TextLen, Tokens, Timestamps).  ``timestamps`` is [n, 2] or None ({0, 0} per id)."""
        ids = _ids(ids)
        ts = None
        n_ts = 0
        if timestamps is not None:
            ts = np.ascontiguousarray(np.asarray(timestamps, dtype=np.int32).reshape(-1, 2))
            n_ts = ts.shape[0]
        res = _lib.PfTextResult()
        call = lambda: self._lib.pf_decode_offline(self._h, ids.ctypes.data_as(_I), ids.size,
                                                   ts.ctypes.data_as(_I) if ts is not None else None, n_ts, C.byref(res))
        _check(call())                                               # size query: all buffers NULL
        return self._fill(res, call)

    def decode_offline_result(self, pf_result, utt: int) -> Tuple[str, int, List[str], List[List[int]]]:
        res = _lib.PfTextResult()
        call = lambda: self._lib.pf_decode_offline_result(self._h, C.byref(pf_result), int(utt), C.byref(res))
        _check(call())
        return self._fill(res, call)

    @staticmethod
    def _fill(res, call):
        text = C.create_string_buffer(res.text_bytes + 1)
        toks = C.create_string_buffer(max(1, res.tokens_bytes))
        ts = np.zeros(max(1, res.ts_count), np.int32)
        off = np.zeros(res.n_timestamps + 1, np.int32)
        res.text, res.text_capacity = C.cast(text, C.c_void_p), len(text)
        res.tokens, res.tokens_capacity = C.cast(toks, C.c_void_p), len(toks)
        res.ts, res.ts_capacity = ts.ctypes.data_as(_I), ts.size
        res.ts_offsets, res.ts_offsets_capacity = off.ctypes.data_as(_I), off.size
        _check(call())
        tokens = [t.decode("utf-8") for t in toks.raw[: res.tokens_bytes].split(b"\0")[: res.n_tokens]]
        stamps = [[int(v) for v in ts[off[i]: off[i + 1]]] for i in range(res.n_timestamps)]
        return text.value.decode("utf-8"), int(res.text_len), tokens, stamps

    # -- OnlineRecognizer.DecodeMulti, one stream
    def decode_online(self, ids) -> str:
        ids = _ids(ids)
        need = C.c_size_t(0)
        _check(self._lib.pf_decode_online(self._h, ids.ctypes.data_as(_I), ids.size, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 1)
        _check(self._lib.pf_decode_online(self._h, ids.ctypes.data_as(_I), ids.size, buf, len(buf), C.byref(need)))
        return buf.value.decode("utf-8")


def time_stamp_lfr6_onnx(us_cif_peak, tokens, begin_time: float = 0.0, total_offset: float = -1.5) -> List[List[int]]:
    """``OfflineRecognizer.time_stamp_lfr6_onnx`` (OfflineRecognizer.cs:200-302) via ``pf_timestamps_lfr6``."""
    lib = _lib.load()
    us = np.ascontiguousarray(np.asarray(us_cif_peak, dtype=np.float32).reshape(-1))
    ids = _ids(tokens)
    n = C.c_int32(0)
    args = (us.ctypes.data_as(_F), us.size, ids.ctypes.data_as(_I), ids.size, float(begin_time), float(total_offset))
    _check(lib.pf_timestamps_lfr6(*args, None, 0, C.byref(n)))
    out = np.zeros((max(1, n.value), 2), np.int32)
    _check(lib.pf_timestamps_lfr6(*args, out.ctypes.data_as(_I), out.shape[0], C.byref(n)))
    return [[int(a), int(b)] for a, b in out[: n.value]]
