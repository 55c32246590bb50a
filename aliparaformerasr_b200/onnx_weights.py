"""Model-file ingestion: ONNX initialisers -> the PFW1 weight blob, without the ``onnx`` package.

Stands where ``OfflineModel.initModel`` / ``EmbedSVModel`` hand a ``model.onnx`` / ``embed.onnx`` to OnnxRuntime
(/root/reference/AliParaformerAsr/OfflineModel.cs:35-70, EmbedSVModel.cs:20-43).  Only what a weight converter needs is
parsed: the protobuf wire format of ``ModelProto.graph`` -> ``initializer`` (TensorProto) and ``node`` (NodeProto: op
type + input / output names).  Quantised exports (``model.int8.onnx`` / ``model_quant.onnx``: ORT dynamic quantisation,
``<w>_quantized`` int8/uint8 + ``<w>_scale`` + ``<w>_zero_point``) are de-quantised to float32.

Pinned by the only ONNX file the reference ships, ``AliParaformerAsr/data/embed.onnx`` (the 16 x 560 SenseVoice prompt
table): tests/test_onnx_weights.py reads a committed copy of its initialiser bytes and compares with
tests/golden/sensevoice_embed.npy.  The MatMul-order -> FunASR-name mapping below cannot be pinned without a model file
([EXT]: MatMul weights are anonymous ``onnx::MatMul_<n>`` tensors stored [in, out]).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Tuple

import numpy as np

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field number, wire type, value) of one protobuf message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _packed_varints(v) -> List[int]:
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """TensorProto: dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9, double_data=10."""
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    for num, wt, val in _fields(buf):
        if num == 1:
            dims.extend(_packed_varints(val) if wt == 2 else [val])
        elif num == 2:
            dtype = val
        elif num == 4:
            floats.extend(struct.unpack(f"<{len(val) // 4}f", bytes(val)) if wt == 2 else struct.unpack("<f", val))
        elif num in (5, 7):
            ints.extend(_packed_varints(val) if wt == 2 else [val])
        elif num == 8:
            name = bytes(val).decode("utf-8")
        elif num == 9:
            raw = bytes(val)
    if dtype not in _DTYPES:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype}")
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    else:
        vals = [(x - (1 << 64)) if x >= (1 << 63) else x for x in ints]     # zig-zag-free negative int64 varints
        arr = np.asarray(vals, dtype=np_dtype)
    return name, arr.reshape(dims) if dims else arr.reshape(())


@dataclass
class OnnxGraph:
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    nodes: List[Tuple[str, List[str], List[str]]] = field(default_factory=list)      # (op_type, inputs, outputs) in graph order


def read_onnx(path_or_bytes) -> OnnxGraph:
    """Initialisers and node list of an ONNX file.  A truncated or corrupted file raises ``ValueError``."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    try:
        return _read_onnx(data)
    except ValueError:
        raise
    except (IndexError, TypeError, UnicodeDecodeError, OverflowError, struct.error) as ex:
        raise ValueError(f"corrupt ONNX file: {type(ex).__name__}: {ex}") from ex


def _read_onnx(data) -> OnnxGraph:
    g = OnnxGraph()
    for num, wt, val in _fields(memoryview(data)):
        if num != 7 or wt != 2:                        # ModelProto.graph
            continue
        for gnum, gwt, gval in _fields(val):
            if gnum == 5 and gwt == 2:                 # GraphProto.initializer
                name, arr = _tensor(gval)
                g.initializers[name] = arr
            elif gnum == 1 and gwt == 2:               # GraphProto.node
                op, ins, outs = "", [], []
                for nnum, nwt, nval in _fields(gval):
                    if nnum == 1:
                        ins.append(bytes(nval).decode("utf-8"))
                    elif nnum == 2:
                        outs.append(bytes(nval).decode("utf-8"))
                    elif nnum == 4:
                        op = bytes(nval).decode("utf-8")
                g.nodes.append((op, ins, outs))
    return g


def dequantize(initializers: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """ORT dynamic-quantisation triplets ``X_quantized`` / ``X_scale`` / ``X_zero_point`` -> float32 ``X``
    ((q - zero_point) * scale, per tensor or per output channel); everything else is passed through as float32."""
    out: Dict[str, np.ndarray] = {}
    for name, arr in initializers.items():
        if name.endswith("_quantized"):
            stem = name[: -len("_quantized")]
            scale = np.asarray(initializers[stem + "_scale"], dtype=np.float32)
            zp = np.asarray(initializers.get(stem + "_zero_point", 0)).astype(np.float32)
            q = arr.astype(np.float32)
            if scale.ndim == 1 and scale.size > 1:   # per-channel along the last axis (MatMul weight [in, out])
                out[stem] = (q - zp.reshape(1, -1)) * scale.reshape(1, -1)
            else:
                out[stem] = (q - zp) * scale
        elif name.endswith("_scale") or name.endswith("_zero_point"):
            continue
        elif arr.dtype.kind == "f":
            out[name] = arr.astype(np.float32)
        else:
            out[name] = arr
    return out


def matmul_weights_in_order(g: OnnxGraph) -> List[str]:
    """Initialiser names consumed as the weight operand of MatMul / MatMulInteger / Gemm nodes, in graph order."""
    names = []
    for op, ins, _ in g.nodes:
        if op in ("MatMul", "Gemm", "MatMulInteger", "DynamicQuantizeMatMul") and len(ins) >= 2:
            w = ins[1]
            stem = w[: -len("_quantized")] if w.endswith("_quantized") else w
            if w in g.initializers or stem + "_quantized" in g.initializers:
                names.append(stem)
    return names


def sensevoice_embed_table(path_or_bytes) -> np.ndarray:
    """The ``weight [16, 560]`` table of ``data/embed.onnx`` (EmbedSVModel.cs:45-77: a single Gather)."""
    g = read_onnx(path_or_bytes)
    tabs = [a for a in g.initializers.values() if a.ndim == 2 and a.dtype == np.float32]
    if len(tabs) != 1:
        raise ValueError("embed.onnx: expected exactly one 2-D float initialiser")
    return tabs[0]


def paraformer_state_dict(g: OnnxGraph, enc_layers: int = 50, dec_layers: int = 16) -> Dict[str, np.ndarray]:
    """[EXT, unpinned] FunASR paraformer export -> FunASR state-dict names.  Named initialisers (biases, LayerNorm, FSMN
    and conv kernels, output layer) keep their module names in the export; the anonymous MatMul weights are assigned by
    graph order - per encoder layer linear_q_k_v, linear_out, w_1, w_2; predictor cif_output; per decoder layer w_1,
    w_2, linear_q, linear_k_v, linear_out; decoders3 w_1, w_2 - and transposed from [in, out] to torch's [out, in]."""
    init = dequantize(g.initializers)
    sd = {k: v for k, v in init.items() if not k.startswith("onnx::") and v.dtype == np.float32}
    order = [n for n in matmul_weights_in_order(g) if n.startswith("onnx::")]
    want: List[str] = []
    for i in range(enc_layers):
        p = "encoder.encoders0.0" if i == 0 else f"encoder.encoders.{i - 1}"
        want += [p + ".self_attn.linear_q_k_v.weight", p + ".self_attn.linear_out.weight", p + ".feed_forward.w_1.weight",
                 p + ".feed_forward.w_2.weight"]
    want.append("predictor.cif_output.weight")
    for i in range(dec_layers):
        p = f"decoder.decoders.{i}"
        want += [p + ".feed_forward.w_1.weight", p + ".feed_forward.w_2.weight", p + ".src_attn.linear_q.weight",
                 p + ".src_attn.linear_k_v.weight", p + ".src_attn.linear_out.weight"]
    want += ["decoder.decoders3.0.feed_forward.w_1.weight", "decoder.decoders3.0.feed_forward.w_2.weight"]
    if len(order) < len(want):
        raise ValueError(f"graph has {len(order)} anonymous MatMul weights, the paraformer layout needs {len(want)}")
    for name, src in zip(want, order):
        sd[name] = np.ascontiguousarray(init[src].T)
    return sd
