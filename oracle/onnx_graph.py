"""A numpy evaluator for ONNX graphs: the REAL-GRAPH oracle (test infrastructure).

The reference never restates its networks - it hands ``model.onnx`` / ``encoder.onnx`` / ``decoder.onnx`` /
``model_eb.onnx`` / ``data/embed.onnx`` to ``Microsoft.ML.OnnxRuntime`` (OfflineModel.cs:35-70, OnlineModel.cs:60-120,
EmbedSeacoModel.cs:20-43, EmbedSVModel.cs:20-43).  Neither onnxruntime nor the ``onnx`` package exists in the build image
and the model files are not vendored, so ``oracle/sanm.py`` restates the FunASR export from prior knowledge ([EXT]).  This
module closes the loop from the other side: it executes an ONNX file node by node with float32 numpy, so that the moment
a model directory is mounted (``baseline/_ref/<model>/model.onnx``) the restatement, the weight-name mapping of
``aliparaformerasr_b200/onnx_weights.py`` and the CUDA path can all be checked against the graph the reference really
runs - including the int8 files, whose ``DynamicQuantizeLinear`` / ``MatMulInteger`` pairs are evaluated the way
OnnxRuntime's CPU provider defines them.

Pinned today on the one ONNX file the reference ships (``data/embed.onnx``, a single Gather) and on synthetic graphs
written with the node vocabulary of the torch.onnx exports (tests/test_onnx_graph.py).  Operator semantics follow the
published ONNX operator specification (opset 11-17 forms: attributes or trailing inputs for axes / pads / split).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from aliparaformerasr_b200.onnx_weights import _fields, _packed_varints, _tensor     # protobuf wire helpers (no onnx package)


@dataclass
class Node:
    op: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)
    name: str = ""


@dataclass
class Graph:
    nodes: List[Node] = field(default_factory=list)
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    inputs: List[str] = field(default_factory=list)
    outputs: List[str] = field(default_factory=list)
    opset: int = 13


def _sint(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


def _attribute(buf):
    """AttributeProto: name=1, f=2, i=3, s=4, t=5, g=6, floats=7, ints=8, strings=9."""
    name, val = "", None
    floats, ints, strings = [], [], []
    for num, wt, v in _fields(buf):
        if num == 1:
            name = bytes(v).decode()
        elif num == 2:
            val = struct.unpack("<f", v)[0]
        elif num == 3:
            val = _sint(v)
        elif num == 4:
            val = bytes(v)
        elif num == 5:
            val = _tensor(v)[1]
        elif num == 6:
            val = _graph(v)
        elif num == 7:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)) if wt == 2 else struct.unpack("<f", v))
        elif num == 8:
            ints.extend([_sint(x) for x in _packed_varints(v)] if wt == 2 else [_sint(v)])
        elif num == 9:
            strings.append(bytes(v))
    if val is None:
        val = floats or ints or strings or 0
    return name, val


def _graph(buf) -> Graph:
    g = Graph()
    for num, wt, v in _fields(buf):
        if num == 1 and wt == 2:                        # node
            n = Node("", [], [])
            for nnum, _, nv in _fields(v):
                if nnum == 1:
                    n.inputs.append(bytes(nv).decode())
                elif nnum == 2:
                    n.outputs.append(bytes(nv).decode())
                elif nnum == 3:
                    n.name = bytes(nv).decode()
                elif nnum == 4:
                    n.op = bytes(nv).decode()
                elif nnum == 5:
                    k, a = _attribute(nv)
                    n.attrs[k] = a
            g.nodes.append(n)
        elif num == 5 and wt == 2:                      # initializer
            name, arr = _tensor(v)
            g.initializers[name] = arr
        elif num in (11, 12) and wt == 2:               # input / output ValueInfoProto
            for vnum, _, vv in _fields(v):
                if vnum == 1:
                    (g.inputs if num == 11 else g.outputs).append(bytes(vv).decode())
    g.inputs = [i for i in g.inputs if i not in g.initializers]
    return g


def load(path_or_bytes) -> Graph:
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    graph, opset = None, 13
    for num, wt, v in _fields(memoryview(data)):
        if num == 7 and wt == 2:
            graph = _graph(v)
        elif num == 8 and wt == 2:                      # opset_import
            dom, ver = "", 0
            for onum, _, ov in _fields(v):
                if onum == 1:
                    dom = bytes(ov).decode()
                elif onum == 2:
                    ver = ov
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph is None:
        raise ValueError("no graph in the ONNX file")
    graph.opset = opset
    return graph


# ---------------------------------------------------------------------------------------------- operators
_NP_OF_ONNX = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}
OPS: Dict[str, Callable] = {}


def op(*names):
    def deco(fn):
        for n in names:
            OPS[n] = fn
        return fn
    return deco


def _axes(n: Node, ins, pos: int, default=None):
    """axes / pads / split / ... come as an attribute up to some opset and as a trailing input afterwards."""
    if len(ins) > pos and ins[pos] is not None:
        return [int(x) for x in np.asarray(ins[pos]).reshape(-1)]
    a = n.attrs.get("axes", default)
    return None if a is None else [int(x) for x in (a if isinstance(a, (list, tuple)) else [a])]


for _name, _fn in (("Add", np.add), ("Sub", np.subtract), ("Mul", np.multiply), ("Equal", np.equal), ("Less", np.less), ("Greater", np.greater),
                   ("LessOrEqual", np.less_equal), ("GreaterOrEqual", np.greater_equal), ("And", np.logical_and), ("Or", np.logical_or)):
    OPS[_name] = (lambda f: lambda n, ins: [f(ins[0], ins[1])])(_fn)
for _name, _fn in (("Sqrt", np.sqrt), ("Exp", np.exp), ("Log", np.log), ("Neg", np.negative), ("Abs", np.abs), ("Floor", np.floor), ("Ceil", np.ceil),
                   ("Tanh", np.tanh), ("Not", np.logical_not), ("Sin", np.sin), ("Cos", np.cos), ("Reciprocal", np.reciprocal)):
    OPS[_name] = (lambda f: lambda n, ins: [f(ins[0])])(_fn)


@op("Div")
def _div(n, ins):
    a, b = ins
    if np.issubdtype(np.asarray(a).dtype, np.integer) and np.issubdtype(np.asarray(b).dtype, np.integer):
        return [np.trunc(np.asarray(a, np.float64) / np.asarray(b, np.float64)).astype(np.asarray(a).dtype)]     # C-style integer division
    return [np.divide(a, b)]


@op("Pow")
def _pow(n, ins):
    return [np.power(ins[0], ins[1]).astype(np.asarray(ins[0]).dtype)]


@op("Relu")
def _relu(n, ins):
    return [np.maximum(ins[0], 0)]


@op("Sigmoid")
def _sigmoid(n, ins):
    x = np.asarray(ins[0])
    return [(1.0 / (1.0 + np.exp(-x))).astype(x.dtype)]


@op("Erf")
def _erf(n, ins):
    x = np.asarray(ins[0])
    return [np.vectorize(math.erf, otypes=[np.float64])(x).astype(x.dtype)]


@op("Identity", "Dropout")
def _identity(n, ins):
    return [ins[0]]


@op("Min")
def _min(n, ins):
    out = ins[0]
    for x in ins[1:]:
        out = np.minimum(out, x)
    return [out]


@op("Max")
def _max(n, ins):
    out = ins[0]
    for x in ins[1:]:
        out = np.maximum(out, x)
    return [out]


@op("Clip")
def _clip(n, ins):
    lo = ins[1] if len(ins) > 1 and ins[1] is not None else n.attrs.get("min")
    hi = ins[2] if len(ins) > 2 and ins[2] is not None else n.attrs.get("max")
    return [np.clip(ins[0], lo, hi)]


@op("Where")
def _where(n, ins):
    return [np.where(ins[0], ins[1], ins[2])]


@op("Cast")
def _cast(n, ins):
    return [np.asarray(ins[0]).astype(_NP_OF_ONNX[int(n.attrs["to"])])]


@op("Constant")
def _constant(n, ins):
    if "value" in n.attrs:
        return [np.asarray(n.attrs["value"])]
    if "value_float" in n.attrs:
        return [np.asarray(n.attrs["value_float"], np.float32)]
    if "value_int" in n.attrs:
        return [np.asarray(n.attrs["value_int"], np.int64)]
    if "value_ints" in n.attrs:
        return [np.asarray(n.attrs["value_ints"], np.int64)]
    if "value_floats" in n.attrs:
        return [np.asarray(n.attrs["value_floats"], np.float32)]
    raise NotImplementedError("Constant without a value attribute")


@op("ConstantOfShape")
def _constant_of_shape(n, ins):
    v = np.asarray(n.attrs.get("value", np.zeros(1, np.float32))).reshape(-1)
    return [np.full([int(x) for x in np.asarray(ins[0]).reshape(-1)], v[0], dtype=v.dtype)]


@op("Shape")
def _shape(n, ins):
    return [np.asarray(np.asarray(ins[0]).shape, dtype=np.int64)]


@op("Size")
def _size(n, ins):
    return [np.asarray(np.asarray(ins[0]).size, dtype=np.int64)]


@op("Reshape")
def _reshape(n, ins):
    x = np.asarray(ins[0])
    shape = [int(s) for s in np.asarray(ins[1]).reshape(-1)]
    shape = [x.shape[i] if s == 0 and not n.attrs.get("allowzero", 0) else s for i, s in enumerate(shape)]
    return [x.reshape(shape)]


@op("Flatten")
def _flatten(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", 1))
    return [x.reshape(int(np.prod(x.shape[:ax], dtype=np.int64)), -1)]


@op("Transpose")
def _transpose(n, ins):
    perm = n.attrs.get("perm")
    return [np.transpose(ins[0], perm)]


@op("Squeeze")
def _squeeze(n, ins):
    ax = _axes(n, ins, 1)
    return [np.squeeze(ins[0], axis=None if ax is None else tuple(ax))]


@op("Unsqueeze")
def _unsqueeze(n, ins):
    x = np.asarray(ins[0])
    ax = _axes(n, ins, 1)
    rank = x.ndim + len(ax)
    for a in sorted(a % rank for a in ax):
        x = np.expand_dims(x, a)
    return [x]


@op("Concat")
def _concat(n, ins):
    return [np.concatenate([np.asarray(x) for x in ins], axis=int(n.attrs["axis"]))]


@op("Split")
def _split(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", 0))
    sizes = [int(s) for s in np.asarray(ins[1]).reshape(-1)] if len(ins) > 1 and ins[1] is not None else n.attrs.get("split")
    if not sizes:
        k = len(n.outputs)
        sizes = [x.shape[ax] // k] * k
    return list(np.split(x, np.cumsum(sizes)[:-1], axis=ax))


@op("Slice")
def _slice(n, ins):
    x = np.asarray(ins[0])
    if len(ins) > 1:
        starts, ends = np.asarray(ins[1]).reshape(-1), np.asarray(ins[2]).reshape(-1)
        axes = np.asarray(ins[3]).reshape(-1) if len(ins) > 3 and ins[3] is not None else np.arange(len(starts))
        steps = np.asarray(ins[4]).reshape(-1) if len(ins) > 4 and ins[4] is not None else np.ones(len(starts), np.int64)
    else:
        starts, ends = np.asarray(n.attrs["starts"]), np.asarray(n.attrs["ends"])
        axes = np.asarray(n.attrs.get("axes", list(range(len(starts)))))
        steps = np.ones(len(starts), np.int64)
    sl = [slice(None)] * x.ndim
    for s, e, a, st in zip(starts, ends, axes, steps):
        d = x.shape[int(a)]
        s, e, st = int(s), int(e), int(st)
        e = max(min(e, d), -d - 1) if st > 0 else max(min(e, d - 1), -d - 1)       # INT_MAX / INT_MIN sentinels
        if st < 0 and e == -d - 1:
            e = None
        sl[int(a)] = slice(s, e, st)
    return [x[tuple(sl)]]


@op("Gather")
def _gather(n, ins):
    return [np.take(ins[0], np.asarray(ins[1]).astype(np.int64), axis=int(n.attrs.get("axis", 0)))]


@op("GatherElements")
def _gather_elements(n, ins):
    return [np.take_along_axis(np.asarray(ins[0]), np.asarray(ins[1]).astype(np.int64), axis=int(n.attrs.get("axis", 0)))]


@op("ScatterND")
def _scatter_nd(n, ins):
    out = np.array(ins[0], copy=True)
    idx, upd = np.asarray(ins[1]).astype(np.int64), np.asarray(ins[2])
    for i in np.ndindex(idx.shape[:-1]):
        out[tuple(idx[i])] = upd[i]
    return [out]


@op("Expand")
def _expand(n, ins):
    x = np.asarray(ins[0])
    shape = [int(s) for s in np.asarray(ins[1]).reshape(-1)]
    return [x * np.ones(shape, dtype=x.dtype)] if x.dtype != np.bool_ else [np.logical_and(x, np.ones(shape, np.bool_))]


@op("Tile")
def _tile(n, ins):
    return [np.tile(ins[0], [int(r) for r in np.asarray(ins[1]).reshape(-1)])]


@op("Range")
def _range(n, ins):
    s, l, d = (np.asarray(v).reshape(()) for v in ins)
    return [np.arange(s, l, d).astype(s.dtype)]


@op("Pad")
def _pad(n, ins):
    x = np.asarray(ins[0])
    pads = [int(p) for p in (np.asarray(ins[1]).reshape(-1) if len(ins) > 1 and ins[1] is not None else n.attrs["pads"])]
    value = np.asarray(ins[2]).reshape(-1)[0] if len(ins) > 2 and ins[2] is not None and np.asarray(ins[2]).size else n.attrs.get("value", 0.0)
    mode = n.attrs.get("mode", b"constant")
    mode = mode.decode() if isinstance(mode, bytes) else mode
    width = [(pads[i], pads[i + x.ndim]) for i in range(x.ndim)]
    if mode == "constant":
        return [np.pad(x, width, mode="constant", constant_values=value)]
    return [np.pad(x, width, mode={"reflect": "reflect", "edge": "edge"}[mode])]


def _reduce(fn):
    def run(n, ins):
        x = np.asarray(ins[0])
        ax = _axes(n, ins, 1)
        keep = bool(n.attrs.get("keepdims", 1))
        if ax is None and n.attrs.get("noop_with_empty_axes", 0):
            return [x]
        return [fn(x, axis=None if ax is None else tuple(ax), keepdims=keep).astype(x.dtype)]
    return run


OPS["ReduceMean"] = _reduce(np.mean)
OPS["ReduceSum"] = _reduce(np.sum)
OPS["ReduceMax"] = _reduce(np.max)
OPS["ReduceMin"] = _reduce(np.min)


@op("ArgMax")
def _argmax(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", 0))
    if n.attrs.get("select_last_index", 0):
        r = x.shape[ax] - 1 - np.argmax(np.flip(x, ax), axis=ax)
    else:
        r = np.argmax(x, axis=ax)
    return [np.expand_dims(r, ax).astype(np.int64) if n.attrs.get("keepdims", 1) else r.astype(np.int64)]


@op("CumSum")
def _cumsum(n, ins):
    x = np.asarray(ins[0])
    ax = int(np.asarray(ins[1]).reshape(()))
    if n.attrs.get("reverse", 0):
        x = np.flip(x, ax)
    # sequential float32 accumulation, the way a CPU kernel walks the axis
    y = np.cumsum(x.astype(np.float64) if False else x, axis=ax, dtype=x.dtype)
    if n.attrs.get("exclusive", 0):
        y = y - x
    if n.attrs.get("reverse", 0):
        y = np.flip(y, ax)
    return [y]


@op("Softmax")
def _softmax(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", -1))
    e = np.exp(x - x.max(axis=ax, keepdims=True))
    return [(e / e.sum(axis=ax, keepdims=True)).astype(x.dtype)]


@op("LogSoftmax")
def _log_softmax(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", -1))
    s = x - x.max(axis=ax, keepdims=True)
    return [(s - np.log(np.exp(s).sum(axis=ax, keepdims=True))).astype(x.dtype)]


@op("LayerNormalization")
def _layer_norm(n, ins):
    x = np.asarray(ins[0])
    ax = int(n.attrs.get("axis", -1)) % x.ndim
    axes = tuple(range(ax, x.ndim))
    mean = x.mean(axis=axes, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=axes, keepdims=True)
    y = (x - mean) / np.sqrt(var + np.float32(n.attrs.get("epsilon", 1e-5)))
    y = y * ins[1]
    if len(ins) > 2 and ins[2] is not None:
        y = y + ins[2]
    return [y.astype(x.dtype)]


@op("MatMul")
def _matmul(n, ins):
    return [np.matmul(ins[0], ins[1])]


@op("Gemm")
def _gemm(n, ins):
    a = np.asarray(ins[0]).T if n.attrs.get("transA", 0) else np.asarray(ins[0])
    b = np.asarray(ins[1]).T if n.attrs.get("transB", 0) else np.asarray(ins[1])
    y = np.float32(n.attrs.get("alpha", 1.0)) * (a @ b)
    if len(ins) > 2 and ins[2] is not None:
        y = y + np.float32(n.attrs.get("beta", 1.0)) * ins[2]
    return [y.astype(a.dtype)]


@op("Conv")
def _conv(n, ins):
    """1-D and 2-D convolution (NCW / NCHW), groups, strides, dilations, explicit pads."""
    x, w = np.asarray(ins[0]), np.asarray(ins[1])
    nd = x.ndim - 2
    strides = list(n.attrs.get("strides", [1] * nd))
    dil = list(n.attrs.get("dilations", [1] * nd))
    pads = list(n.attrs.get("pads", [0] * (2 * nd)))
    groups = int(n.attrs.get("group", 1))
    if nd == 1:                                           # run as 2-D with a unit height
        x, w = x[:, :, None, :], w[:, :, None, :]
        strides, dil, pads = [1] + strides, [1] + dil, [0, pads[0], 0, pads[1]]
    x = np.pad(x, ((0, 0), (0, 0), (pads[0], pads[2]), (pads[1], pads[3])))
    b, cin, h, wd = x.shape
    cout, cpg, kh, kw = w.shape
    oh = (h - (dil[0] * (kh - 1) + 1)) // strides[0] + 1
    ow = (wd - (dil[1] * (kw - 1) + 1)) // strides[1] + 1
    y = np.zeros((b, cout, oh, ow), dtype=np.float32)
    opg = cout // groups
    for g in range(groups):
        xg = x[:, g * cpg:(g + 1) * cpg]
        wg = w[g * opg:(g + 1) * opg]
        for i in range(kh):
            for j in range(kw):
                patch = xg[:, :, i * dil[0]: i * dil[0] + (oh - 1) * strides[0] + 1: strides[0], j * dil[1]: j * dil[1] + (ow - 1) * strides[1] + 1: strides[1]]
                y[:, g * opg:(g + 1) * opg] += np.einsum("bchw,oc->bohw", patch, wg[:, :, i, j], optimize=True)
    if len(ins) > 2 and ins[2] is not None:
        y += np.asarray(ins[2]).reshape(1, -1, 1, 1)
    return [y[:, :, 0, :] if nd == 1 else y]


@op("ConvTranspose")
def _conv_transpose(n, ins):
    """1-D transposed convolution (NCW), stride / pads / output_padding, group 1."""
    x, w = np.asarray(ins[0]), np.asarray(ins[1])                       # w: [cin, cout, k]
    stride = int(n.attrs.get("strides", [1])[0])
    pads = list(n.attrs.get("pads", [0, 0]))
    opad = int(n.attrs.get("output_padding", [0])[0])
    b, cin, t = x.shape
    _, cout, k = w.shape
    full = (t - 1) * stride + k + opad
    y = np.zeros((b, cout, full), dtype=np.float32)
    for j in range(k):
        y[:, :, j: j + (t - 1) * stride + 1: stride] += np.einsum("bct,co->bot", x, w[:, :, j], optimize=True)
    y = y[:, :, pads[0]: full - pads[1]]
    if len(ins) > 2 and ins[2] is not None:
        y = y + np.asarray(ins[2]).reshape(1, -1, 1)
    return [y]


@op("LSTM")
def _lstm(n, ins):
    """ONNX LSTM, layout 0 ([T, B, I]), forward / reverse / bidirectional, default activations, gate order i o f c."""
    x, w, r = np.asarray(ins[0]), np.asarray(ins[1]), np.asarray(ins[2])
    bias = np.asarray(ins[3]) if len(ins) > 3 and ins[3] is not None else None
    h0 = np.asarray(ins[5]) if len(ins) > 5 and ins[5] is not None else None
    c0 = np.asarray(ins[6]) if len(ins) > 6 and ins[6] is not None else None
    hs = int(n.attrs["hidden_size"])
    direction = n.attrs.get("direction", b"forward")
    direction = direction.decode() if isinstance(direction, bytes) else direction
    t, b, _ = x.shape
    ndir = w.shape[0]
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))                            # noqa: E731
    y = np.zeros((t, ndir, b, hs), dtype=np.float32)
    yh = np.zeros((ndir, b, hs), dtype=np.float32)
    yc = np.zeros((ndir, b, hs), dtype=np.float32)
    for d in range(ndir):
        rev = direction == "reverse" or (direction == "bidirectional" and d == 1)
        h = h0[d].astype(np.float32) if h0 is not None else np.zeros((b, hs), np.float32)
        c = c0[d].astype(np.float32) if c0 is not None else np.zeros((b, hs), np.float32)
        wb = bias[d, : 4 * hs] + bias[d, 4 * hs:] if bias is not None else 0.0
        for step in (range(t - 1, -1, -1) if rev else range(t)):
            g = x[step] @ w[d].T + h @ r[d].T + wb
            i, o, f, cc = g[:, :hs], g[:, hs:2 * hs], g[:, 2 * hs:3 * hs], g[:, 3 * hs:]
            c = sig(f) * c + sig(i) * np.tanh(cc)
            h = sig(o) * np.tanh(c)
            y[step, d] = h
        yh[d], yc[d] = h, c
    return [y, yh, yc]


@op("DynamicQuantizeLinear")
def _dynamic_quantize_linear(n, ins):
    """uint8 dynamic quantisation as OnnxRuntime's CPU provider computes it (ONNX spec): the range always includes 0."""
    x = np.asarray(ins[0], np.float32)
    lo, hi = min(float(x.min()), 0.0), max(float(x.max()), 0.0)
    scale = np.float32((hi - lo) / 255.0) if hi > lo else np.float32(1.0)
    zp = np.uint8(np.clip(np.round(-lo / scale), 0, 255))
    q = np.clip(np.round(x / scale) + zp, 0, 255).astype(np.uint8)           # np.round = round half to even, like the spec
    return [q, np.asarray(scale, np.float32), np.asarray(zp, np.uint8)]


@op("MatMulInteger")
def _matmul_integer(n, ins):
    a, b = np.asarray(ins[0]).astype(np.int32), np.asarray(ins[1]).astype(np.int32)
    if len(ins) > 2 and ins[2] is not None:
        a = a - np.asarray(ins[2]).astype(np.int32)
    if len(ins) > 3 and ins[3] is not None:
        b = b - np.asarray(ins[3]).astype(np.int32)
    return [np.matmul(a, b).astype(np.int32)]


@op("DequantizeLinear")
def _dequantize_linear(n, ins):
    x = np.asarray(ins[0]).astype(np.float32)
    zp = np.asarray(ins[2]).astype(np.float32) if len(ins) > 2 and ins[2] is not None else 0.0
    scale = np.asarray(ins[1], np.float32)
    ax = int(n.attrs.get("axis", 1))
    if scale.ndim == 1 and scale.size > 1:
        shape = [1] * x.ndim
        shape[ax] = -1
        scale, zp = scale.reshape(shape), (zp.reshape(shape) if isinstance(zp, np.ndarray) and zp.ndim == 1 else zp)
    return [((x - zp) * scale).astype(np.float32)]


# ---------------------------------------------------------------------------------------------- execution
def run(graph: Graph, feeds: Dict[str, np.ndarray], outputs: Optional[Sequence[str]] = None, keep: Optional[Sequence[str]] = None,
        trace: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """Execute ``graph`` in node order (ONNX graphs are stored topologically sorted).  ``feeds`` names the graph inputs;
    ``keep`` asks for intermediate values by tensor name (per-stage taps); ``trace`` (a dict) receives the op histogram."""
    env: Dict[str, object] = dict(graph.initializers)
    for k, v in feeds.items():
        env[k] = np.asarray(v)
    missing = [i for i in graph.inputs if i not in env]
    if missing:
        raise ValueError(f"missing graph inputs: {missing}")
    wanted = list(outputs) if outputs is not None else list(graph.outputs)
    kept = {}
    for n in graph.nodes:
        fn = OPS.get(n.op)
        if fn is None:
            raise NotImplementedError(f"ONNX operator {n.op!r} (node {n.name or n.outputs}) is not in the evaluator's vocabulary")
        ins = [env[i] if i != "" else None for i in n.inputs]
        outs = fn(n, ins)
        for name, val in zip(n.outputs, outs):
            if name:
                env[name] = val
        if trace is not None:
            trace[n.op] = trace.get(n.op, 0) + 1
        if keep:
            for name in n.outputs:
                if name in keep:
                    kept[name] = np.asarray(env[name])
    res = {k: np.asarray(env[k]) for k in wanted}
    res.update(kept)
    return res


def supported_ops() -> List[str]:
    return sorted(OPS)
