"""oracle/onnx_graph.py - the numpy ONNX evaluator that stands in for OnnxRuntime as the real-graph oracle.

Pinned on the one ONNX file the reference ships (data/embed.onnx, golden copy of its table under tests/golden/), checked
operator by operator against torch on random inputs, and run end to end on a synthetic paraformer ``model.onnx`` written with
the node vocabulary of a torch.onnx export (tests/_util.GraphBuilder), where it must agree with the hand-written restatement
oracle/sanm.py AND the initialisers must map back through aliparaformerasr_b200/onnx_weights.py.  When a real model directory
is mounted under baseline/_ref/ the last test runs the evaluator on it and compares with oracle/sanm.py on mapped weights."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from aliparaformerasr_b200 import onnx_weights as ow, synth
from oracle import frontend as F, onnx_graph as G, sanm
from _util import GraphBuilder, dims_of, export_paraformer_graph, onnx_model, onnx_tensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sensevoice_embed.npy")
REAL_EMBED = "/root/reference/AliParaformerAsr/data/embed.onnx"


def _run1(op, ins, attrs=None, nout=1, inits=None):
    names = [f"i{k}" for k in range(len(ins))]
    outs = [f"o{k}" for k in range(nout)]
    data = onnx_model([onnx_tensor(k, v) for k, v in (inits or {}).items()], [(op, names, outs, attrs or {})], names, outs)
    g = G.load(data)
    res = G.run(g, dict(zip(names, ins)))
    return [res[o] for o in outs]


def test_embed_onnx_the_reference_ships():
    gold = np.load(GOLD)
    ids = np.asarray([[14, 1, 2, 15, 0, 7]], np.int64)
    data = onnx_model([onnx_tensor("weight", gold)], [("Gather", ["weight", "x"], ["y"], {})], ["x"], ["y"])
    assert np.array_equal(G.run(G.load(data), {"x": ids})["y"], gold[ids])
    if os.path.exists(REAL_EMBED):                 # EmbedSVModel.Forward (EmbedSVModel.cs:45-77) on the real file
        g = G.load(REAL_EMBED)
        assert [n.op for n in g.nodes] == ["Gather"] and g.inputs == ["x"] and g.outputs == ["y"]
        assert np.array_equal(G.run(g, {"x": ids})["y"], gold[ids])


def test_operators_against_torch():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 6, 16)).astype(np.float32)
    # Conv1d with groups / pads / dilation; ConvTranspose1d stride 3
    w = rng.standard_normal((16, 1, 5)).astype(np.float32)
    xt = torch.from_numpy(x.transpose(0, 2, 1).copy())
    got = _run1("Conv", [x.transpose(0, 2, 1).copy(), w], {"group": 16, "kernel_shape": [5], "pads": [2, 2], "strides": [1], "dilations": [1]})[0]
    assert np.allclose(got, Fn.conv1d(xt, torch.from_numpy(w), padding=2, groups=16).numpy(), atol=1e-5)
    w2 = rng.standard_normal((8, 16, 3)).astype(np.float32)
    b2 = rng.standard_normal(8).astype(np.float32)
    got = _run1("Conv", [x.transpose(0, 2, 1).copy(), w2, b2], {"group": 1, "kernel_shape": [3], "pads": [0, 0], "strides": [2], "dilations": [1]})[0]
    assert np.allclose(got, Fn.conv1d(xt, torch.from_numpy(w2), torch.from_numpy(b2), stride=2).numpy(), atol=1e-5)
    wt = rng.standard_normal((16, 16, 3)).astype(np.float32)
    got = _run1("ConvTranspose", [x.transpose(0, 2, 1).copy(), wt], {"strides": [3], "kernel_shape": [3]})[0]
    assert np.allclose(got, Fn.conv_transpose1d(xt, torch.from_numpy(wt), stride=3).numpy(), atol=1e-5)
    # LayerNormalization (opset 17 form), Softmax, LogSoftmax, CumSum, Pad, Slice with negative steps, Split
    gm, bt = rng.standard_normal(16).astype(np.float32), rng.standard_normal(16).astype(np.float32)
    got = _run1("LayerNormalization", [x, gm, bt], {"axis": -1, "epsilon": 1e-12})[0]
    assert np.allclose(got, Fn.layer_norm(torch.from_numpy(x), (16,), torch.from_numpy(gm), torch.from_numpy(bt), 1e-12).numpy(), atol=1e-5)
    assert np.allclose(_run1("LogSoftmax", [x], {"axis": -1})[0], torch.log_softmax(torch.from_numpy(x), -1).numpy(), atol=1e-6)
    assert np.allclose(_run1("CumSum", [x, np.asarray(1, np.int64)])[0], np.cumsum(x, 1), atol=1e-6)
    assert np.array_equal(_run1("Pad", [x, np.asarray([0, 1, 0, 0, 2, 0], np.int64)], {"mode": "constant"})[0], np.pad(x, ((0, 0), (1, 2), (0, 0))))
    got = _run1("Slice", [x, np.asarray([-1], np.int64), np.asarray([-(2 ** 62)], np.int64), np.asarray([1], np.int64), np.asarray([-1], np.int64)])[0]
    assert np.array_equal(got, x[:, ::-1])
    a, b = _run1("Split", [x, np.asarray([4, 12], np.int64)], {"axis": -1}, nout=2)
    assert np.array_equal(a, x[..., :4]) and np.array_equal(b, x[..., 4:])
    # bidirectional LSTM vs torch (ONNX gate order i o f c, torch i f g o)
    lstm = torch.nn.LSTM(16, 8, 1, bidirectional=True)
    xs = torch.from_numpy(x.transpose(1, 0, 2).copy())            # [T, B, I]
    want, _ = lstm(xs)

    def onnx_w(t):                                                   # torch [i f g o] -> onnx [i o f c]
        i, f, g_, o = t.detach().numpy().reshape(4, 8, -1)
        return np.concatenate([i, o, f, g_]).reshape(32, -1) if t.ndim == 2 else np.concatenate([i, o, f, g_]).reshape(-1)
    W = np.stack([onnx_w(lstm.weight_ih_l0), onnx_w(lstm.weight_ih_l0_reverse)])
    R = np.stack([onnx_w(lstm.weight_hh_l0), onnx_w(lstm.weight_hh_l0_reverse)])
    Bv = np.stack([np.concatenate([onnx_w(lstm.bias_ih_l0), onnx_w(lstm.bias_hh_l0)]), np.concatenate([onnx_w(lstm.bias_ih_l0_reverse), onnx_w(lstm.bias_hh_l0_reverse)])])
    y = _run1("LSTM", [xs.numpy(), W.astype(np.float32), R.astype(np.float32), Bv.astype(np.float32)], {"hidden_size": 8, "direction": "bidirectional"}, nout=3)[0]
    assert np.allclose(np.concatenate([y[:, 0], y[:, 1]], -1), want.detach().numpy(), atol=1e-5)


def test_dynamic_quantisation_ops_follow_the_onnx_definition():
    """model.int8.onnx / model_quant.onnx (AliParaformerAsr.Examples/Program.cs:100): DynamicQuantizeLinear + MatMulInteger."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 32)).astype(np.float32) * 3
    q, scale, zp = _run1("DynamicQuantizeLinear", [x], nout=3)
    assert q.dtype == np.uint8 and 0 <= int(zp) <= 255
    assert np.abs((q.astype(np.float32) - zp) * scale - x).max() <= scale * 0.5 + 1e-6
    w = rng.integers(-128, 128, size=(32, 7), dtype=np.int8)
    acc = _run1("MatMulInteger", [q, w, zp, np.asarray(0, np.int8)])[0]
    assert acc.dtype == np.int32 and np.array_equal(acc, (q.astype(np.int32) - int(zp)) @ w.astype(np.int32))


@pytest.fixture(scope="module")
def tiny_graph():
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    return cfg, w, export_paraformer_graph(w, cfg)


def test_paraformer_graph_agrees_with_the_restatement(tiny_graph):
    """The evaluator on a model.onnx written with the export's node vocabulary == oracle/sanm.py on the same weights: two
    independent formulations (decomposed LayerNorm, Split/Reshape/Transpose attention, ConstantPad + grouped Conv FSMN,
    CumSum-overlap CIF vs the sequential integrate-and-fire recurrence) of SURVEY 2.5."""
    cfg, w, data = tiny_graph
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(synth.make_pcm(i, 3.0), shift, scale) for i in range(2)])
    g = G.load(data)
    trace = {}
    res = G.run(g, {"speech": speech, "speech_lengths": np.full(2, speech.shape[1], np.int32)}, keep=["enc"], trace=trace)
    ref = sanm.paraformer_forward(speech, w, dims_of(cfg))
    assert np.array_equal(res["token_num"], ref["token_num"])
    assert np.abs(res["enc"] - ref["enc"]).max() < 1e-4
    assert res["logits"].shape == ref["logits"].shape and np.abs(res["logits"] - ref["logits"]).max() < 2e-3
    assert np.array_equal(res["logits"].argmax(-1)[ref["logits"].max(-1) - np.sort(ref["logits"], -1)[..., -2] > 0.05],
                          ref["tokens"][ref["logits"].max(-1) - np.sort(ref["logits"], -1)[..., -2] > 0.05])
    for needed in ("MatMul", "Add", "ReduceMean", "Pow", "Sqrt", "Div", "Split", "Reshape", "Transpose", "Softmax", "Pad", "Conv", "CumSum", "Range",
                   "LogSoftmax"):
        assert trace.get(needed, 0) > 0, needed


def test_runnable_graph_maps_back_through_onnx_weights(tiny_graph):
    cfg, w, data = tiny_graph
    sd = ow.paraformer_state_dict(ow.read_onnx(data), cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn, cfg.input_size, cfg.dec_ffn, cfg.vocab)
    for k, v in w.items():
        assert k in sd and np.array_equal(sd[k], v), k


def _mounted_models():
    return sorted(glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "model.onnx"), recursive=True))


@pytest.mark.skipif(not _mounted_models(), reason="no model.onnx mounted under baseline/_ref/ (none exists offline)")
def test_real_model_graph_vs_restatement():
    """With a real paraformer model directory mounted: the reference's own graph (this evaluator) vs oracle/sanm.py on the
    weights mapped by onnx_weights.paraformer_state_dict - pins both the restatement and the name mapping."""
    path = _mounted_models()[0]
    g = G.load(path)
    cfg = synth.paraformer_large()
    sd = ow.paraformer_state_dict(ow.read_onnx(path), cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.ffn, cfg.input_size, cfg.dec_ffn)
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(synth.make_pcm(0, 3.0), shift, scale)])
    feeds = {g.inputs[0]: speech, g.inputs[1]: np.full(1, speech.shape[1], np.int32)}
    res = G.run(g, feeds)
    cfg.vocab = int(sd["decoder.output_layer.bias"].shape[0])
    ref = sanm.paraformer_forward(speech, sd, dims_of(cfg))
    out = res[g.outputs[0]]
    assert out.shape == ref["logits"].shape and np.abs(out - ref["logits"]).max() < 1e-2
