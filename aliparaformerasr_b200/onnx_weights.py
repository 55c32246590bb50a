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


# sha256 of the float32 payload of the reference's AliParaformerAsr/data/embed.onnx (file sha256 5c69dceb...52f9b1)
SENSEVOICE_EMBED_SHA256 = "9d27e2547e90d9a0348ad97cfc1d86ce2ae7453c03f8f9d96e86b7aecca801a8"


def packaged_sensevoice_embed() -> np.ndarray:
    """The reference's own prompt table, shipped as package data (``data/sensevoice_embed.npy``): EmbedSVModel loads
    ``data/embed.onnx`` as an embedded resource of the assembly (EmbedSVModel.cs:20-43), so a split-embed SenseVoice
    ``model.onnx`` carries no ``embed.weight``.  The bytes are checked against the pinned digest on every load."""
    import hashlib
    import os
    tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sensevoice_embed.npy"))
    tab = np.ascontiguousarray(tab, dtype=np.float32)
    if tab.shape != (16, 560) or hashlib.sha256(tab.tobytes()).hexdigest() != SENSEVOICE_EMBED_SHA256:
        raise ValueError("packaged SenseVoice prompt table does not match the reference's data/embed.onnx")
    return tab


def _enc_slots(prefix_layers, d_in0: int, d: int, f: int):
    """(name, (out, in)) of the four Linear weights of every SAN-M encoder layer, in execution order."""
    out = []
    for k, p in enumerate(prefix_layers):
        din = d_in0 if k == 0 and d_in0 else d
        out += [(p + ".self_attn.linear_q_k_v.weight", (3 * d, din)), (p + ".self_attn.linear_out.weight", (d, d)),
                (p + ".feed_forward.w_1.weight", (f, d)), (p + ".feed_forward.w_2.weight", (d, f))]
    return out


def _assign_matmul_weights(g: OnnxGraph, init: Dict[str, np.ndarray], slots, sd: Dict[str, np.ndarray]) -> Dict[str, str]:
    """Give every anonymous MatMul weight its FunASR name.  torch.onnx exports ``nn.Linear`` on a 3-D input as
    ``MatMul(x, onnx::MatMul_N)`` followed by ``Add(bias, y)`` where the bias KEEPS its module name, so the weight's
    name is read off the bias of the Add that consumes the MatMul output ("<module>.bias" -> "<module>.weight").  Only
    bias-less layers (the decoder's feed_forward.w_2) fall back to graph order: they take the first still-unassigned
    slot that follows the slot of the previous MatMul.  Every assignment is checked against the slot's [out, in] shape,
    so two same-sized tensors cannot be swapped silently."""
    shape_of = dict(slots)
    consumers: Dict[str, List[Tuple[str, List[str]]]] = {}
    for op, ins, outs in g.nodes:
        for x in ins:
            consumers.setdefault(x, []).append((op, ins))
    order = [s for s, _ in slots]
    pos = {s: i for i, s in enumerate(order)}
    assigned: Dict[str, str] = {}
    last = -1
    for op, ins, outs in g.nodes:
        if op not in ("MatMul", "Gemm", "MatMulInteger", "DynamicQuantizeMatMul") or len(ins) < 2:
            continue
        w = ins[1]
        stem = w[: -len("_quantized")] if w.endswith("_quantized") else w
        if stem not in init or init[stem].ndim != 2:
            continue
        if not stem.startswith("onnx::") and stem in shape_of:          # the export kept the module name
            last = pos[stem]
            continue
        name = None
        if op == "Gemm" and len(ins) >= 3 and ins[2].endswith(".bias"):
            name = ins[2][: -len("bias")] + "weight"
        for cop, cins in consumers.get(outs[0], []) if outs else []:
            if cop == "Add":
                for b in cins:
                    if b.endswith(".bias") and b in init:
                        name = b[: -len("bias")] + "weight"
        if name is None or name not in shape_of or name in assigned.values():
            nxt = [s for s in order[last + 1:] if s not in assigned.values() and s not in sd]
            cand = [s for s in nxt if tuple(init[stem].shape) == shape_of[s][::-1]]
            if not cand:
                continue                                                  # not one of ours (e.g. a helper MatMul)
            name = cand[0]
        want = shape_of[name]
        if tuple(init[stem].shape) != want[::-1]:
            raise ValueError(f"{name}: MatMul weight {stem} has shape {tuple(init[stem].shape)}, expected [in, out] = {want[::-1]}")
        assigned[stem] = name
        last = pos[name]
    return assigned


def _finish_state_dict(g: OnnxGraph, slots, what: str) -> Dict[str, np.ndarray]:
    init = dequantize(g.initializers)
    sd = {k: v for k, v in init.items() if not k.startswith("onnx::") and v.dtype == np.float32}
    for name, arr in list(sd.items()):
        want = dict(slots).get(name)
        if want is not None and tuple(arr.shape) == want[::-1] and want[0] != want[1]:
            sd[name] = np.ascontiguousarray(arr.T)                        # a named weight stored [in, out]
    assigned = _assign_matmul_weights(g, init, slots, sd)
    for src, name in assigned.items():
        sd[name] = np.ascontiguousarray(init[src].T)
    missing = [s for s, _ in slots if s not in sd]
    if missing:
        raise ValueError(f"{what}: {len(missing)} Linear weights could not be located in the graph, first: {missing[0]}")
    for name, want in slots:
        if tuple(sd[name].shape) != want:
            raise ValueError(f"{what}: {name} has shape {tuple(sd[name].shape)}, expected {want}")
    return sd


def paraformer_state_dict(g: OnnxGraph, enc_layers: int = 50, dec_layers: int = 16, d: int = 512, f: int = 2048, din: int = 560,
                          dec_f: int = 2048, vocab: int = 0) -> Dict[str, np.ndarray]:
    """[EXT, EXPERIMENTAL until pinned against a real export] FunASR paraformer ``model.onnx`` -> FunASR state-dict names.
    Named initialisers (biases, LayerNorm, FSMN and conv kernels) keep their module names in the export; the anonymous
    ``onnx::MatMul_N`` weights ([in, out]) get theirs from the named bias of the Add behind them, bias-less ones by
    order, all shape-checked (see ``_assign_matmul_weights``)."""
    enc = ["encoder.encoders0.0"] + [f"encoder.encoders.{i}" for i in range(enc_layers - 1)]
    slots = _enc_slots(enc, din, d, f)
    slots.append(("predictor.cif_output.weight", (1, d)))
    for i in range(dec_layers):
        p = f"decoder.decoders.{i}"
        slots += [(p + ".feed_forward.w_1.weight", (dec_f, d)), (p + ".feed_forward.w_2.weight", (d, dec_f)),
                  (p + ".src_attn.linear_q.weight", (d, d)), (p + ".src_attn.linear_k_v.weight", (2 * d, d)),
                  (p + ".src_attn.linear_out.weight", (d, d))]
    slots += [("decoder.decoders3.0.feed_forward.w_1.weight", (dec_f, d)), ("decoder.decoders3.0.feed_forward.w_2.weight", (d, dec_f))]
    if vocab:
        slots.append(("decoder.output_layer.weight", (vocab, d)))
    return _finish_state_dict(g, slots, "paraformer model.onnx")


def sensevoice_state_dict(g: OnnxGraph, enc_layers: int = 50, tp_layers: int = 20, d: int = 512, f: int = 2048, din: int = 560,
                          vocab: int = 25055) -> Dict[str, np.ndarray]:
    """[EXT, EXPERIMENTAL] SenseVoiceSmall ``model.onnx`` -> state-dict names: encoders0 + encoders + tp_encoders (four
    Linear weights each) and the CTC head.  A split-embed export has no ``embed.weight``; the caller adds the reference's
    own table (``packaged_sensevoice_embed``), which is what EmbedSVModel does with data/embed.onnx."""
    enc = ["encoder.encoders0.0"] + [f"encoder.encoders.{i}" for i in range(enc_layers - 1)]
    slots = _enc_slots(enc, din, d, f) + _enc_slots([f"encoder.tp_encoders.{i}" for i in range(tp_layers)], 0, d, f)
    slots.append(("ctc.ctc_lo.weight", (vocab, d)))
    return _finish_state_dict(g, slots, "sensevoice model.onnx")
