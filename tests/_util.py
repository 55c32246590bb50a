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


def onnx_model(tensors, nodes):
    def node(op, ins, outs):
        return b"".join(_pb_ld(1, i.encode()) for i in ins) + b"".join(_pb_ld(2, o.encode()) for o in outs) + _pb_ld(4, op.encode())
    graph = b"".join(_pb_ld(1, node(*n)) for n in nodes) + _pb_ld(2, b"g") + b"".join(_pb_ld(5, t) for t in tensors)
    return _pb_varint((1 << 3) | 0) + _pb_varint(8) + _pb_ld(7, graph)


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
