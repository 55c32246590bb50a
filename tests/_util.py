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


# ---------------------------------------------------------------- minimal ONNX (protobuf) writer for the ingestion tests
def _pb_varint(x):
    out = b""
    while True:
        b = x & 0x7F
        x >>= 7
        out += bytes([b | (0x80 if x else 0)])
        if not x:
            return out


def _pb_ld(num, payload):
    return _pb_varint((num << 3) | 2) + _pb_varint(len(payload)) + payload


def onnx_tensor(name, arr):
    dt = {np.dtype(np.float32): 1, np.dtype(np.uint8): 2, np.dtype(np.int8): 3, np.dtype(np.int64): 7}[arr.dtype]
    return (b"".join(_pb_varint((1 << 3) | 0) + _pb_varint(int(d)) for d in arr.shape) + _pb_varint((2 << 3) | 0) + _pb_varint(dt) +
            _pb_ld(8, name.encode()) + _pb_ld(9, np.ascontiguousarray(arr).tobytes()))


def _pb_sint(x):
    return _pb_varint(x & ((1 << 64) - 1))


def onnx_attr(name, val):
    """AttributeProto: name=1, f=2, i=3, s=4, t=5, floats=7, ints=8, type=20 (FLOAT 1, INT 2, STRING 3, TENSOR 4, FLOATS 6, INTS 7)."""
    import struct
    out = _pb_ld(1, name.encode())
    if isinstance(val, float):
        out += _pb_varint((2 << 3) | 5) + struct.pack("<f", val) + _pb_varint((20 << 3) | 0) + _pb_varint(1)
    elif isinstance(val, (int, np.integer)):
        out += _pb_varint((3 << 3) | 0) + _pb_sint(int(val)) + _pb_varint((20 << 3) | 0) + _pb_varint(2)
    elif isinstance(val, (bytes, str)):
        out += _pb_ld(4, val.encode() if isinstance(val, str) else val) + _pb_varint((20 << 3) | 0) + _pb_varint(3)
    elif isinstance(val, np.ndarray):
        out += _pb_ld(5, onnx_tensor("", val)) + _pb_varint((20 << 3) | 0) + _pb_varint(4)
    elif isinstance(val, (list, tuple)) and val and isinstance(val[0], float):
        out += _pb_ld(7, struct.pack(f"<{len(val)}f", *val)) + _pb_varint((20 << 3) | 0) + _pb_varint(6)
    else:
        out += _pb_ld(8, b"".join(_pb_sint(int(v)) for v in val)) + _pb_varint((20 << 3) | 0) + _pb_varint(7)
    return out


def onnx_model(tensors, nodes, inputs=(), outputs=(), opset=13):
    """nodes: (op, inputs, outputs) or (op, inputs, outputs, attrs dict)."""
    def node(op, ins, outs, attrs=None):
        return (b"".join(_pb_ld(1, i.encode()) for i in ins) + b"".join(_pb_ld(2, o.encode()) for o in outs) + _pb_ld(4, op.encode()) +
                b"".join(_pb_ld(5, onnx_attr(k, v)) for k, v in (attrs or {}).items()))
    graph = (b"".join(_pb_ld(1, node(*n)) for n in nodes) + _pb_ld(2, b"g") + b"".join(_pb_ld(5, t) for t in tensors) +
             b"".join(_pb_ld(11, _pb_ld(1, i.encode())) for i in inputs) + b"".join(_pb_ld(12, _pb_ld(1, o.encode())) for o in outputs))
    return (_pb_varint((1 << 3) | 0) + _pb_varint(8) + _pb_ld(7, graph) +
            _pb_ld(8, _pb_ld(1, b"") + _pb_varint((2 << 3) | 0) + _pb_varint(opset)))


class GraphBuilder:
    """Emits a RUNNABLE ONNX graph with the node vocabulary of a torch.onnx (opset 13) export: Linear = MatMul with an anonymous
    [in, out] weight + Add with the module's named bias, LayerNorm decomposed into ReduceMean / Sub / Pow / Sqrt / Div / Mul / Add,
    ConstantPad1d + Conv for the FSMN, Split / Reshape / Transpose around the attention products, CumSum-based CIF."""

    def __init__(self, weights):
        self.w = weights
        self.nodes, self.inits, self.k = [], {}, 0

    def name(self, tag):
        self.k += 1
        return f"/{tag}_{self.k}"

    def const(self, arr, name=None):
        name = name or self.name("Constant")
        self.inits[name] = np.asarray(arr)
        return name

    def weight(self, name):
        if name not in self.inits:
            self.inits[name] = np.asarray(self.w[name], np.float32)
        return name

    def op(self, op, ins, nout=1, **attrs):
        outs = [self.name(op) for _ in range(nout)]
        self.nodes.append((op, list(ins), outs, attrs))
        return outs[0] if nout == 1 else outs

    def linear(self, x, mod):
        a = f"onnx::MatMul_{3000 + 11 * len(self.inits)}"
        self.inits[a] = np.ascontiguousarray(np.asarray(self.w[mod + ".weight"], np.float32).T)
        y = self.op("MatMul", [x, a])
        return self.op("Add", [self.weight(mod + ".bias"), y]) if mod + ".bias" in self.w else y

    def layer_norm(self, x, mod, eps):
        mean = self.op("ReduceMean", [x], axes=[-1], keepdims=1)
        d = self.op("Sub", [x, mean])
        var = self.op("ReduceMean", [self.op("Pow", [d, self.const(np.float32(2.0))])], axes=[-1], keepdims=1)
        std = self.op("Sqrt", [self.op("Add", [var, self.const(np.float32(eps))])])
        y = self.op("Mul", [self.op("Div", [d, std]), self.weight(mod + ".weight")])
        return self.op("Add", [y, self.weight(mod + ".bias")])

    def fsmn(self, v, mod, kernel, mask=None):
        if mask is not None:
            v = self.op("Mul", [v, mask])
        left = (kernel - 1) // 2
        x = self.op("Transpose", [v], perm=[0, 2, 1])
        x = self.op("Pad", [x, self.const(np.asarray([0, 0, left, 0, 0, kernel - 1 - left], np.int64))], mode="constant")
        x = self.op("Conv", [x, self.weight(mod + ".weight")], group=int(self.w[mod + ".weight"].shape[0]), kernel_shape=[kernel],
                    dilations=[1], strides=[1], pads=[0, 0])
        x = self.op("Add", [self.op("Transpose", [x], perm=[0, 2, 1]), v])
        return self.op("Mul", [x, mask]) if mask is not None else x

    def heads(self, x, h, dk):
        return self.op("Transpose", [self.op("Reshape", [x, self.const(np.asarray([0, 0, h, dk], np.int64))])], perm=[0, 2, 1, 3])

    def mha(self, q, k, v, h, d):
        dk = d // h
        qh = self.op("Mul", [self.heads(q, h, dk), self.const(np.float32(dk ** -0.5))])
        kt = self.op("Transpose", [self.op("Reshape", [k, self.const(np.asarray([0, 0, h, dk], np.int64))])], perm=[0, 2, 3, 1])
        att = self.op("Softmax", [self.op("MatMul", [qh, kt])], axis=-1)
        ctx = self.op("Transpose", [self.op("MatMul", [att, self.heads(v, h, dk)])], perm=[0, 2, 1, 3])
        return self.op("Reshape", [ctx, self.const(np.asarray([0, 0, d], np.int64))])

    def model(self, inputs, outputs, opset=13):
        tensors = [onnx_tensor(k, v) for k, v in self.inits.items()]
        return onnx_model(tensors, self.nodes, inputs, outputs, opset)


def export_paraformer_graph(weights, cfg):
    """Runnable paraformer ``model.onnx`` for a synthetic state dict: inputs ``speech [B,T,560]`` (+ the unused
    ``speech_lengths``), outputs ``logits`` (log-softmax) and ``token_num``, graph semantics of SURVEY.md 2.5."""
    from oracle import sanm
    g = GraphBuilder(weights)
    d, h, eps = cfg.d_model, cfg.heads, cfg.ln_eps
    t_max = 2048
    pe = sanm.sinusoidal_pe(t_max, cfg.input_size).numpy()
    x = g.op("Mul", ["speech", g.const(np.float32(d ** 0.5))])
    tlen = g.op("Gather", [g.op("Shape", ["speech"]), g.const(np.asarray(1, np.int64))], axis=0)
    pe_t = g.op("Slice", [g.const(pe[None]), g.const(np.asarray([0], np.int64)), g.op("Unsqueeze", [tlen, g.const(np.asarray([0], np.int64))]),
                          g.const(np.asarray([1], np.int64))])
    x = g.op("Add", [x, pe_t])

    def enc_layer(x, p, residual):
        hx = g.layer_norm(x, p + ".norm1", eps)
        q, k, v = g.op("Split", [g.linear(hx, p + ".self_attn.linear_q_k_v"), g.const(np.asarray([d, d, d], np.int64))], nout=3, axis=-1)
        mem = g.fsmn(v, p + ".self_attn.fsmn_block", cfg.enc_kernel)
        att = g.op("Add", [g.linear(g.mha(q, k, v, h, d), p + ".self_attn.linear_out"), mem])
        x = g.op("Add", [x, att]) if residual else att
        hx = g.layer_norm(x, p + ".norm2", eps)
        hx = g.linear(g.op("Relu", [g.linear(hx, p + ".feed_forward.w_1")]), p + ".feed_forward.w_2")
        return g.op("Add", [x, hx])

    x = enc_layer(x, "encoder.encoders0.0", False)
    for i in range(cfg.enc_layers - 1):
        x = enc_layer(x, f"encoder.encoders.{i}", True)
    enc = g.layer_norm(x, "encoder.after_norm", eps)
    # CifPredictorV2: ConstantPad1d(1, 1) + Conv1d(k=3) + ReLU + Linear(512, 1) + sigmoid, tail 0.45, then the CumSum form of CIF
    c = g.op("Pad", [g.op("Transpose", [enc], perm=[0, 2, 1]), g.const(np.asarray([0, 0, 1, 0, 0, 1], np.int64))], mode="constant")
    c = g.op("Relu", [g.op("Conv", [c, g.weight("predictor.cif_conv1d.weight"), g.weight("predictor.cif_conv1d.bias")], group=1, kernel_shape=[3],
                           dilations=[1], strides=[1], pads=[0, 0])])
    al = g.op("Sigmoid", [g.linear(g.op("Transpose", [c], perm=[0, 2, 1]), "predictor.cif_output")])
    al = g.op("Relu", [g.op("Sub", [g.op("Mul", [al, g.const(np.float32(cfg.smooth_factor))]), g.const(np.float32(cfg.noise_threshold))])])
    al = g.op("Squeeze", [al, g.const(np.asarray([2], np.int64))])
    first = g.op("Slice", [al, g.const(np.asarray([0], np.int64)), g.const(np.asarray([1], np.int64)), g.const(np.asarray([1], np.int64))])
    tail = g.op("Add", [g.op("Mul", [first, g.const(np.float32(0.0))]), g.const(np.float32(cfg.cif_tail))])
    alphas = g.op("Concat", [al, tail], axis=1)                                                   # [B, T+1]
    hid0 = g.op("Slice", [enc, g.const(np.asarray([0], np.int64)), g.const(np.asarray([1], np.int64)), g.const(np.asarray([1], np.int64))])
    hidden = g.op("Concat", [enc, g.op("Mul", [hid0, g.const(np.float32(0.0))])], axis=1)       # [B, T+1, d], last row zeros
    csum = g.op("CumSum", [alphas, g.const(np.asarray(1, np.int64))])
    prev = g.op("Sub", [csum, alphas])
    tn = g.op("Floor", [g.op("ReduceSum", [alphas, g.const(np.asarray([1], np.int64))], keepdims=0)])    # [B]
    lmax = g.op("Cast", [g.op("ReduceMax", [tn], keepdims=0)], to=7)
    lidx = g.op("Cast", [g.op("Range", [g.const(np.asarray(0, np.int64)), lmax, g.const(np.asarray(1, np.int64))])], to=1)
    lo = g.op("Unsqueeze", [lidx, g.const(np.asarray([0, 2], np.int64))])                         # [1, L, 1]
    hi = g.op("Add", [lo, g.const(np.float32(1.0))])
    cs3, pv3 = (g.op("Unsqueeze", [v, g.const(np.asarray([1], np.int64))]) for v in (csum, prev))
    wgt = g.op("Relu", [g.op("Sub", [g.op("Min", [cs3, hi]), g.op("Max", [pv3, lo])])])        # overlap of token l with frame t
    mask = g.op("Cast", [g.op("Less", [g.op("Unsqueeze", [lidx, g.const(np.asarray([0], np.int64))]),
                                       g.op("Unsqueeze", [tn, g.const(np.asarray([1], np.int64))])])], to=1)
    mask = g.op("Unsqueeze", [mask, g.const(np.asarray([2], np.int64))])                          # [B, L, 1]
    x = g.op("Mul", [g.op("MatMul", [wgt, hidden]), mask])                                       # acoustic_embeds

    def dec_ffn(x, p):
        hx = g.op("Relu", [g.linear(x, p + ".w_1")])
        return g.linear(g.layer_norm(hx, p + ".norm", eps), p + ".w_2")

    for i in range(cfg.dec_layers):
        p = f"decoder.decoders.{i}"
        tt = dec_ffn(g.layer_norm(x, p + ".norm1", eps), p + ".feed_forward")
        x = g.op("Add", [x, g.fsmn(g.layer_norm(tt, p + ".norm2", eps), p + ".self_attn.fsmn_block", cfg.dec_kernel, mask)])
        hx = g.layer_norm(x, p + ".norm3", eps)
        k, v = g.op("Split", [g.linear(enc, p + ".src_attn.linear_k_v"), g.const(np.asarray([d, d], np.int64))], nout=2, axis=-1)
        x = g.op("Add", [x, g.linear(g.mha(g.linear(hx, p + ".src_attn.linear_q"), k, v, h, d), p + ".src_attn.linear_out")])
    x = dec_ffn(g.layer_norm(x, "decoder.decoders3.0.norm1", eps), "decoder.decoders3.0.feed_forward")
    x = g.linear(g.layer_norm(x, "decoder.after_norm", eps), "decoder.output_layer")
    g.nodes.append(("LogSoftmax", [x], ["logits"], {"axis": -1}))
    g.nodes.append(("Cast", [tn], ["token_num"], {"to": 6}))
    g.nodes.append(("Identity", [enc], ["enc"], {}))
    return g.model(["speech", "speech_lengths"], ["logits", "token_num"])


def _export_linears(weights, linears, extra_nodes=()):
    """FunASR / torch.onnx convention: ``nn.Linear`` on a 3-D input becomes ``MatMul(x, onnx::MatMul_N)`` with the weight
    stored [in, out] under an anonymous name, followed by ``Add(<module>.bias, y)`` whose bias keeps its module name;
    a bias-less Linear is a bare MatMul.  ``linears``: module names in the order the nodes are written."""
    tensors, nodes = [], []
    anon = set()
    for k, mod in enumerate(linears):
        a = f"onnx::MatMul_{1000 + 7 * k}"
        anon.add(mod + ".weight")
        tensors.append(onnx_tensor(a, np.ascontiguousarray(weights[mod + ".weight"].T)))
        nodes.append(("MatMul", [f"x{k}", a], [f"y{k}"]))
        if mod + ".bias" in weights:
            nodes.append(("Add", [mod + ".bias", f"y{k}"], [f"x{k + 1}"]))
        else:
            nodes.append(("Relu", [f"y{k}"], [f"x{k + 1}"]))
    nodes += list(extra_nodes)
    for name, arr in weights.items():
        if name not in anon:
            tensors.append(onnx_tensor(name, np.asarray(arr, np.float32)))
    return onnx_model(tensors, nodes)


def export_paraformer_onnx(weights, enc_layers, dec_layers, anonymous_head=True):
    """A model.onnx image in the FunASR export convention for a synthetic paraformer state dict.  The node order is
    deliberately NOT the order onnx_weights lists its slots in (the cross-attention K/V projection is written before the
    query projection, the alpha head after the decoder), so the round trip only succeeds if names are derived from the
    graph (the bias of the Add behind each MatMul), not from position."""
    order = []
    for i in range(enc_layers):
        p = "encoder.encoders0.0" if i == 0 else f"encoder.encoders.{i - 1}"
        order += [p + ".self_attn.linear_q_k_v", p + ".self_attn.linear_out", p + ".feed_forward.w_1", p + ".feed_forward.w_2"]
    for i in range(dec_layers):
        p = f"decoder.decoders.{i}"
        order += [p + ".src_attn.linear_k_v", p + ".feed_forward.w_1", p + ".feed_forward.w_2", p + ".src_attn.linear_q", p + ".src_attn.linear_out"]
    order += ["decoder.decoders3.0.feed_forward.w_1", "decoder.decoders3.0.feed_forward.w_2", "predictor.cif_output"]
    if anonymous_head:
        order.append("decoder.output_layer")
    return _export_linears(weights, order)


def export_sensevoice_onnx(weights, enc_layers, tp_layers):
    """Split-embed SenseVoiceSmall export: encoder + tp_encoder Linears and the CTC head, no ``embed.weight``."""
    order = []
    for p in ["encoder.encoders0.0"] + [f"encoder.encoders.{i}" for i in range(enc_layers - 1)] + [f"encoder.tp_encoders.{i}" for i in range(tp_layers)]:
        order += [p + ".self_attn.linear_q_k_v", p + ".self_attn.linear_out", p + ".feed_forward.w_1", p + ".feed_forward.w_2"]
    order.append("ctc.ctc_lo")
    w = {k: v for k, v in weights.items() if k != "embed.weight"}
    return _export_linears(w, order)
