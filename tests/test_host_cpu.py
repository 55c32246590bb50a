"""CPU tests of the host-side logic: the reference-API mirror's contracts, config/asset parsing, sharding (gloo, 2 ranks)."""
import json
import os

import numpy as np
import pytest

from aliparaformerasr_b200 import offline, shard, synth
from oracle import frontend as F


def test_load_cmvn_matches_reference_parser(tmp_path):
    shift = np.linspace(-9, -7, 560).astype(np.float32)
    scale = np.linspace(0.2, 0.3, 560).astype(np.float32)
    p = tmp_path / "am.mvn"
    p.write_text(F.format_am_mvn(shift, scale))
    s, c = offline.load_cmvn(str(p))
    assert np.array_equal(s, shift) and np.array_equal(c, scale)


def test_load_conf_yaml_and_json(tmp_path):
    y = tmp_path / "asr.yaml"
    y.write_text("model: sensevoicesmall\nuse_itn: true\nencoder_conf:\n  output_size: 512\n  num_blocks: 50\n  tp_blocks: 20\n"
                 "frontend_conf:\n  snip_edges: true\n  lfr_m: 7\n  lfr_n: 6\nvocab_size: 25055\n")
    c = offline.load_conf(str(y))
    assert (c.model, c.tp_layers, c.snip_edges, c.vocab, c.dec_layers) == ("sensevoicesmall", 20, True, 25055, 0)
    j = tmp_path / "asr.json"
    j.write_text(json.dumps({"model": "paraformer", "decoder_conf": {"num_blocks": 16, "kernel_size": 11}, "predictor_conf": {"threshold": 1.0}}))
    c = offline.load_conf(str(j))
    assert (c.model, c.dec_layers, c.enc_layers, c.input_size) == ("paraformer", 16, 50, 560)
    assert offline.load_conf("").model == "paraformer"          # defaults baked into the C# entities


def test_missing_tokens_file_raises_tokens_invalid():
    """Init_WithMissingTokensFile (OfflineRecognizerTests .cs:184-208)"""
    with pytest.raises(Exception) as ei:
        offline.OfflineRecognizer("model.pfw", "", "", "")
    assert "tokens invalid" in str(ei.value)


def test_add_samples_null_raises_argument_null_source():
    """AddSamples_WithNull (OfflineRecognizerTests .cs:285-297)"""
    s = offline.OfflineStream.__new__(offline.OfflineStream)
    s._chunks = []
    with pytest.raises(offline.ArgumentNullError) as ei:
        s.add_samples(None)
    assert ei.value.param_name == "source"


def test_host_pad_sequence_equals_reference_padhelper():
    a = np.ones((3, 560), np.float32)
    b = np.ones((5, 560), np.float32)
    b[1, 3] = 0
    assert np.array_equal(offline.pad_sequence([a, b]), F.pad_sequence([a, b]))


def test_decode_multi_text_rules():
    rec = offline.OfflineRecognizer.__new__(offline.OfflineRecognizer)
    from aliparaformerasr_b200.text import TokenTable
    rec._token_table = TokenTable(lines=["<blank>", "<s>", "</s>", "你", "好", "hel@@", "lo", "world", "<unk>"])
    s = offline.OfflineStream.__new__(offline.OfflineStream)
    s.tokens = [3, 4, 5, 6, 7, 2, 3]
    s.timestamps = [[0, 0]] * 7
    r = rec._decode_multi([s])[0]
    assert r.tokens == ["你", "好", "hello", "world"]
    assert r.text.replace(" ", "") == "你好helloworld" and "hello" in r.text
    assert r.text_len == len(r.text)


def test_split_batch_contiguous():
    assert shard.split_batch(32, 8) == [(4 * i, 4) for i in range(8)]
    assert shard.split_batch(5, 2) == [(0, 3), (3, 2)]
    assert shard.split_batch(1, 4) == [(0, 1), (1, 0), (1, 0), (1, 0)]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = shard.split_batch(6, world)[rank]
    l = 3 + rank                                     # ranks see different Lmax
    toks = (np.arange(count * l, dtype=np.int32).reshape(count, l) + 100 * rank)
    tn = np.full(count, l, np.int32)
    full, num = shard.gather_tokens(toks, tn, width=8)
    if rank == 0:
        q.put((full.tolist(), num.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_tokens_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, num = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.asarray(full)
    assert full.shape == (6, 8) and num == [3, 3, 3, 4, 4, 4]
    assert full[0, :3].tolist() == [0, 1, 2] and full[0, 3] == -1
    assert full[3, :4].tolist() == [100, 101, 102, 103] and full[3, 4] == -1
