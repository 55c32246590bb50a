"""ONNX initialiser reader (row f3): round trips through a minimal protobuf writer, de-quantisation, MatMul order, and
- when the reference tree is mounted (build container) - the real ``data/embed.onnx`` against the committed golden."""
import os
import struct

import numpy as np
import pytest

from aliparaformerasr_b200 import onnx_weights as ow

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sensevoice_embed.npy")
REAL = "/root/reference/AliParaformerAsr/data/embed.onnx"


def _vi(x):
    out = b""
    while True:
        b = x & 0x7F
        x >>= 7
        out += bytes([b | (0x80 if x else 0)])
        if not x:
            return out


def _ld(num, payload):
    return _vi((num << 3) | 2) + _vi(len(payload)) + payload


def _tensor(name, arr, raw=True):
    dt = {np.dtype(np.float32): 1, np.dtype(np.uint8): 2, np.dtype(np.int8): 3, np.dtype(np.int64): 7}[arr.dtype]
    body = b"".join(_vi((1 << 3) | 0) + _vi(int(d)) for d in arr.shape) + _vi((2 << 3) | 0) + _vi(dt) + _ld(8, name.encode())
    if raw:
        body += _ld(9, arr.tobytes())
    else:
        body += _ld(4, struct.pack(f"<{arr.size}f", *arr.reshape(-1)))
    return body


def _node(op, ins, outs):
    return b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())


def _model(tensors, nodes):
    graph = b"".join(_ld(1, _node(*n)) for n in nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in tensors)
    return _vi((1 << 3) | 0) + _vi(8) + _ld(7, graph)


def test_roundtrip_and_matmul_order():
    rng = np.random.default_rng(0)
    w1 = rng.standard_normal((4, 6)).astype(np.float32)
    w2 = rng.standard_normal((6, 3)).astype(np.float32)
    b = rng.standard_normal(3).astype(np.float32)
    data = _model([_tensor("onnx::MatMul_7", w1), _tensor("onnx::MatMul_9", w2, raw=False), _tensor("enc.bias", b),
                   _tensor("shape", np.asarray([1, -1, 3], np.int64))],
                  [("MatMul", ["x", "onnx::MatMul_7"], ["h"]), ("Relu", ["h"], ["r"]), ("MatMul", ["r", "onnx::MatMul_9"], ["y"]),
                   ("Add", ["y", "enc.bias"], ["z"])])
    g = ow.read_onnx(data)
    assert np.array_equal(g.initializers["onnx::MatMul_7"], w1)
    assert np.array_equal(g.initializers["onnx::MatMul_9"], w2)
    assert np.array_equal(g.initializers["shape"], [1, -1, 3])
    assert ow.matmul_weights_in_order(g) == ["onnx::MatMul_7", "onnx::MatMul_9"]
    assert [n[0] for n in g.nodes] == ["MatMul", "Relu", "MatMul", "Add"]


def test_dequantize_dynamic_quantisation_triplets():
    rng = np.random.default_rng(1)
    q = rng.integers(-128, 128, size=(5, 4), dtype=np.int8)
    scale = np.asarray([0.02, 0.03, 0.01, 0.05], np.float32)
    zp = np.zeros(4, np.int8)
    qu = rng.integers(0, 256, size=(3, 2), dtype=np.uint8)
    data = _model([_tensor("w_quantized", q), _tensor("w_scale", scale), _tensor("w_zero_point", zp),
                   _tensor("u_quantized", qu), _tensor("u_scale", np.asarray(0.1, np.float32).reshape(())),
                   _tensor("u_zero_point", np.asarray(128, np.uint8).reshape(())), _tensor("ln.weight", np.ones(4, np.float32))],
                  [("MatMulInteger", ["a", "w_quantized"], ["y"])])
    g = ow.read_onnx(data)
    d = ow.dequantize(g.initializers)
    assert set(d) == {"w", "u", "ln.weight"}
    assert np.allclose(d["w"], q.astype(np.float32) * scale[None, :])
    assert np.allclose(d["u"], (qu.astype(np.float32) - 128.0) * 0.1)
    assert ow.matmul_weights_in_order(g) == ["w"]


def test_embed_table_from_generated_file_matches_golden():
    gold = np.load(GOLD)
    data = _model([_tensor("weight", gold)], [("Gather", ["weight", "x"], ["y"])])
    assert np.array_equal(ow.sensevoice_embed_table(data), gold)


@pytest.mark.skipif(not os.path.exists(REAL), reason="reference tree not mounted (GPU box)")
def test_real_embed_onnx_matches_golden():
    # the one numeric artefact the reference ships (EmbedSVModel.cs:20-43 loads it as an embedded resource)
    assert np.array_equal(ow.sensevoice_embed_table(REAL), np.load(GOLD))
    g = ow.read_onnx(REAL)
    assert g.nodes == [("Gather", ["weight", "x"], ["y"])]


def test_paraformer_export_maps_back_to_the_state_dict():
    """model.onnx ingestion end to end on the host: a synthetic paraformer state dict written in the FunASR export
    convention (anonymous MatMul weights, named biases, node order different from the mapper's slot order) maps back to
    the same names and values - names come from the bias behind each MatMul, bias-less w_2 from its position."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from _util import export_paraformer_onnx
    from aliparaformerasr_b200 import synth
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    for anon_head in (True, False):
        g = ow.read_onnx(export_paraformer_onnx(w, cfg.enc_layers, cfg.dec_layers, anonymous_head=anon_head))
        sd = ow.paraformer_state_dict(g, cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn, cfg.input_size, cfg.dec_ffn, cfg.vocab)
        assert set(sd) == set(w)
        for k in w:
            assert sd[k].shape == w[k].shape and np.array_equal(sd[k], w[k]), k


def test_same_shaped_weights_cannot_be_swapped_silently():
    """linear_out and linear_q are both [512, 512]: writing them in swapped node order still maps each to its own name
    (bias-derived), and a weight whose shape does not fit its slot is an error, not a silent mis-assignment."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from _util import _export_linears
    from aliparaformerasr_b200 import synth
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    mods = sorted({k[: -len(".weight")] for k in w if k.endswith(".weight") and w[k].ndim == 2 and
                   ("linear" in k or ".w_" in k or k.startswith("predictor.cif_output"))})
    # bias-less w_2 must still follow its w_1; everything else is shuffled
    rng = np.random.default_rng(3)
    rest = [m for m in mods if not (m.startswith("decoder.") and m.endswith("feed_forward.w_2"))]
    rng.shuffle(rest)
    order = []
    for m in rest:
        order.append(m)
        if m.startswith("decoder.") and m.endswith("feed_forward.w_1"):
            order.append(m[:-1] + "2")
    sd = ow.paraformer_state_dict(ow.read_onnx(_export_linears(w, order)), cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn,
                                  cfg.input_size, cfg.dec_ffn)
    for k in w:
        assert np.array_equal(sd[k], w[k]), k
    bad = dict(w)
    bad["encoder.encoders.0.self_attn.linear_out.weight"] = np.zeros((512, 256), np.float32)
    with pytest.raises(ValueError):
        ow.paraformer_state_dict(ow.read_onnx(_export_linears(bad, order)), cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn,
                                 cfg.input_size, cfg.dec_ffn)


def test_sensevoice_split_embed_export_and_packaged_prompt_table():
    """A split-embed SenseVoice model.onnx has no embed.weight: the mapper finds the 70 x 4 + 1 Linear weights and the
    packaged table is the reference's data/embed.onnx payload (sha256-pinned, EmbedSVModel.cs:20-43)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from _util import export_sensevoice_onnx
    from aliparaformerasr_b200 import synth
    cfg = synth.tiny("sensevoicesmall")
    w = synth.make_weights(cfg)
    sd = ow.sensevoice_state_dict(ow.read_onnx(export_sensevoice_onnx(w, cfg.enc_layers, cfg.tp_layers)), cfg.enc_layers, cfg.tp_layers,
                                  cfg.d_model, cfg.ffn, cfg.input_size, cfg.vocab)
    assert "embed.weight" not in sd
    for k in w:
        if k != "embed.weight":
            assert np.array_equal(sd[k], w[k]), k
    tab = ow.packaged_sensevoice_embed()
    assert tab.shape == (16, 560) and np.array_equal(tab, np.load(GOLD))
    if os.path.exists(REAL):
        assert np.array_equal(tab, ow.sensevoice_embed_table(REAL))


def test_corrupted_files_raise_value_error():
    rng = np.random.default_rng(0)
    base = _model([_tensor("onnx::MatMul_1", rng.standard_normal((4, 6)).astype(np.float32)), _tensor("b", rng.standard_normal(6).astype(np.float32))],
                  [("MatMul", ["x", "onnx::MatMul_1"], ["y"]), ("Add", ["y", "b"], ["z"])])
    ok = 0
    for _ in range(3000):
        blob = bytearray(base)
        if rng.random() < 0.5:
            blob = blob[: int(rng.integers(0, len(blob) + 1))]
        for _ in range(int(rng.integers(0, 5))):
            if blob:
                blob[int(rng.integers(0, len(blob)))] = int(rng.integers(0, 256))
        try:
            ow.read_onnx(bytes(blob))
            ok += 1
        except ValueError:
            pass
    assert ok > 0
