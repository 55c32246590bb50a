"""PFW1 weight blob: the flat file ``pf_offline_create`` loads (csrc/engine.cu ``Blob::parse``).

It stands where ``model.onnx`` stands for the reference (``OfflineModel.initModel``,
/root/reference/AliParaformerAsr/OfflineModel.cs:35-70): a one-time conversion of the model's initialisers
into named, 256-byte aligned float32 tensors keyed by FunASR state-dict names.

Layout (little endian)::

    "PFW1" | u32 version=1 | u32 count | u32 reserved
    count x { char name[96]; u32 dtype (0 = f32); u32 ndim; u64 dims[4]; u64 offset; u64 nbytes }
    payload (each tensor 256-byte aligned)
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_ENTRY = struct.Struct("<96sII4QQQ")
_ALIGN = 256


def pack(weights: Dict[str, np.ndarray]) -> np.ndarray:
    """Serialise ``name -> float32 array`` into one contiguous uint8 buffer."""
    names = list(weights.keys())
    table_bytes = 16 + _ENTRY.size * len(names)
    offset = (table_bytes + _ALIGN - 1) // _ALIGN * _ALIGN
    entries = []
    for n in names:
        a = np.ascontiguousarray(weights[n], dtype=np.float32)
        if a.ndim > 4:
            raise ValueError(f"{n}: more than 4 dims")
        if len(n.encode()) > 95:
            raise ValueError(f"{n}: name too long")
        dims = list(a.shape) + [1] * (4 - a.ndim)
        entries.append((n, a, dims, offset))
        offset = (offset + a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    buf = np.zeros(offset, dtype=np.uint8)
    buf[:16] = np.frombuffer(b"PFW1" + struct.pack("<III", 1, len(names), 0), dtype=np.uint8)
    pos = 16
    for n, a, dims, off in entries:
        raw = _ENTRY.pack(n.encode(), 0, a.ndim, *dims, off, a.nbytes)
        buf[pos:pos + _ENTRY.size] = np.frombuffer(raw, dtype=np.uint8)
        pos += _ENTRY.size
        buf[off:off + a.nbytes] = a.view(np.uint8).reshape(-1)
    return buf


def unpack(buf: np.ndarray) -> Dict[str, np.ndarray]:
    raw = np.ascontiguousarray(buf, dtype=np.uint8)
    if bytes(raw[:4]) != b"PFW1":
        raise ValueError("bad magic")
    version, count, _ = struct.unpack("<III", bytes(raw[4:16]))
    if version != 1:
        raise ValueError("unsupported PFW version")
    out: Dict[str, np.ndarray] = {}
    for i in range(count):
        name, dtype, ndim, d0, d1, d2, d3, off, nbytes = _ENTRY.unpack(bytes(raw[16 + i * _ENTRY.size: 16 + (i + 1) * _ENTRY.size]))
        shape = (d0, d1, d2, d3)[:ndim]
        out[name.rstrip(b"\0").decode()] = raw[off:off + nbytes].view(np.float32).reshape(shape).copy()
    return out


def save(path: str, weights: Dict[str, np.ndarray]) -> None:
    pack(weights).tofile(path)
