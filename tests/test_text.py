"""Host post-processing (SURVEY.md §8 f2 + the consumer half of f1): the native DecodeMulti / time_stamp_lfr6_onnx /
tokens table of libpfasr (csrc/text.cu) against the oracle restatement (oracle/text.py, oracle/sanm.py) and against
hand traces of the C# (OfflineRecognizer.cs:304-418, OnlineRecognizer.cs:403-436).  CPU only: no device work."""
import ctypes as C

import numpy as np
import pytest

from aliparaformerasr_b200 import _lib
from aliparaformerasr_b200.text import TokenTable, time_stamp_lfr6_onnx
from oracle import sanm
from oracle import text as otext

VOCAB = ["<blank>", "<s>", "</s>", "你", "好", "hel@@", "lo", "world", "<unk>", "▁the", "re", "▁cat", "s", "x@@", "y",
         "ab@@", "d", "A@@", "B\t77", "龥", "一", "é", "Ünï", "ПРИВЕТ", "ΑΒΓ", "𝄞", "▁", "@@", "好\t9"]
IDX = {t: i for i, t in enumerate(VOCAB)}


@pytest.fixture(scope="module")
def table():
    t = TokenTable(lines=VOCAB)
    yield t
    t.close()


def ids(*names):
    return [IDX[n] for n in names]


def stamps(n):
    return [[10 * i, 10 * i + 7] for i in range(n)]


def both(table, seq, ts=None):
    ts = stamps(len(seq)) if ts is None else ts
    a = table.decode_offline(seq, ts)
    b = otext.decode_multi_offline(VOCAB, seq, ts)
    assert a == b
    return a


def test_tokens_table_read_all_lines_semantics(tmp_path):
    raw = "﻿<blank>\r\n<s>\n</s>\r你\n\n好\n".encode("utf-8")
    p = tmp_path / "tokens.txt"
    p.write_bytes(raw)
    t = TokenTable(path=str(p))
    expect = ["<blank>", "<s>", "</s>", "你", "", "好"]       # BOM dropped, CR / LF / CRLF, no empty last line
    assert t.lines() == expect == otext.read_all_lines(raw)
    from aliparaformerasr_b200.offline import read_tokens
    assert read_tokens(str(p)) == expect
    assert len(t) == 6 and t[3] == "你"
    with pytest.raises(IndexError):
        t[6]
    t.close()
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.pf_tokens_create_from_memory(b"", 0, C.byref(h)) == _lib.PF_ERR_BAD_ARG       # "tokens invalid"
    assert lib.pf_tokens_create(b"/nonexistent/tokens.txt", C.byref(h)) == _lib.PF_ERR_BAD_ARG
    assert lib.pf_tokens_destroy(None) == _lib.PF_ERR_DISPOSED
    assert lib.pf_tokens_count(None) == -1


def test_offline_bpe_join_hand_trace(table):
    """你 好 hel@@ lo world: 'hel@@' + 'lo' meet at "@@▁▁" (branch 1), 'world' opens a new word; the text keeps the
    trailing blank the final Replace chain leaves (OfflineRecognizer.cs:402-405)."""
    text, text_len, toks, ts = both(table, ids("你", "好", "hel@@", "lo", "world", "</s>", "你"))
    assert text == "你好 hello world "
    assert text_len == len(text)
    assert toks == ["你", "好", "hello", "world"]
    assert ts == [[0, 7], [10, 17], [20, 27, 30, 37], [40, 47]]


def test_offline_plain_words_and_specials(table):
    text, _, toks, ts = both(table, ids("<s>", "hel@@", "lo", "<unk>", "<blank>", "world"))
    assert text == " hello world " and toks == ["hello", "world"]
    assert ts == [[10, 17, 20, 27], [50, 57]]
    # 'B\t77': Split('\t')[0]
    assert both(table, ids("B\t77"))[2] == ["B"]
    # stops at id 2, empty input, shorter timestamp list ends the Zip
    assert both(table, ids("</s>", "lo"))[0] == ""
    assert both(table, [])[0] == ""
    assert both(table, ids("lo", "world", "s"), stamps(2))[2] == ["lo", "world"]


def test_offline_sentencepiece_style_hand_trace(table):
    """'▁the' 're' '▁cat' 's': bar counts 3 and 5 take the second merge branch (:366-392), the final text goes through
    the "▁▁▁" -> blank chain."""
    text, _, toks, ts = both(table, ids("▁the", "re", "▁cat", "s"))
    assert text == "there cats"
    assert toks == ["there", "cats"]
    assert ts == [[0, 7, 10, 17], [20, 27, 30, 37]]


def test_offline_remove_by_value_quirk(table):
    """List<string>.Remove(Tokens.Last()) deletes the FIRST equal entry (:356): with 好 x@@ 好 y the merge of x@@+y drops
    the first 好 from Tokens while Timestamps (reference equality) loses its last entry."""
    text, _, toks, ts = both(table, ids("好", "x@@", "好", "y"))
    assert toks == ["x@@", "好", "xy"]
    assert ts == [[0, 7], [10, 17], [10, 17, 30, 37]]
    # the text is "好▁x@@▁好▁y▁": no "@@▁▁" forms because 好 sits between the pieces, so "@@" and each bar turn into blanks
    assert text == "好 x  好 y "


def test_offline_chinese_range_edges(table):
    # U+4E00 and U+9FA5 are inside the class; a supplementary-plane symbol counts two UTF-16 units in TextLen
    text, text_len, toks, _ = both(table, ids("一", "龥", "𝄞"))
    assert toks == ["一", "龥", "𝄞"] and text.startswith("一龥")
    assert text_len == len(text.encode("utf-16-le")) // 2 == len(text) + 1
    # '好\t9' is split at the tab before the Chinese test
    assert both(table, ids("好\t9"))[0] == "好"


def test_offline_out_of_table_id_raises(table):
    with pytest.raises(IndexError):
        table.decode_offline([len(VOCAB)], [[0, 0]])
    with pytest.raises(IndexError):
        otext.decode_multi_offline(VOCAB, [len(VOCAB)], [[0, 0]])
    with pytest.raises(IndexError):
        table.decode_online([-1])


def test_offline_buffer_too_small_reports_sizes(table):
    lib = _lib.load()
    seq = np.asarray(ids("你", "好", "hel@@", "lo"), np.int32)
    res = _lib.PfTextResult()
    buf = C.create_string_buffer(4)
    res.text, res.text_capacity = C.cast(buf, C.c_void_p), 4
    st = lib.pf_decode_offline(table._h, seq.ctypes.data_as(C.POINTER(C.c_int32)), seq.size, None, 0, C.byref(res))
    assert st == _lib.PF_ERR_BAD_ARG and b"too small" in lib.pf_last_error()
    assert res.text_bytes == len("你好 hello ".encode()) and res.n_tokens == 3 and res.ts_count == 8


def test_offline_random_sequences_match_oracle(table):
    rng = np.random.default_rng(7)
    for _ in range(400):
        n = int(rng.integers(0, 24))
        seq = [int(v) for v in rng.integers(0, len(VOCAB), n)]
        ts = [[int(a), int(a + b)] for a, b in zip(rng.integers(0, 9000, n), rng.integers(0, 500, n))]
        both(table, seq, ts)
        assert table.decode_online(seq) == otext.decode_multi_online(VOCAB, seq)


def test_online_hand_trace(table):
    """OnlineRecognizer.cs:430: "@@▁▁" and "@@▁" vanish, "▁▁" becomes one blank, the rest of the bars vanish, ToLower."""
    assert table.decode_online(ids("你", "好", "hel@@", "lo", "world")) == "你好hello world"
    assert table.decode_online(ids("A@@", "B\t77", "</s>", "lo")) == "ab\t77"      # whole line, no tab split (:417-425)
    assert table.decode_online(ids("Ünï", "ПРИВЕТ", "ΑΒΓ", "é")) == "ünï привет αβγ é"
    assert table.decode_online([]) == ""


def test_timestamps_native_matches_oracle_and_hand_trace():
    pk = np.zeros(150, np.float32)
    for i in (20, 35, 80, 100):
        pk[i] = 1.0
    toks = [5, 6, 7, 8, 2]
    a = time_stamp_lfr6_onnx(pk, toks)
    assert a == sanm.time_stamp_lfr6_onnx(pk, toks) == [[370, 669], [669, 1270], [1569, 2485]]
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(8, 400))
        k = int(rng.integers(1, 20))
        pk = (rng.random(n) * 0.9).astype(np.float32)
        pos = np.sort(rng.choice(n, size=min(k, n), replace=False))
        pk[pos] = 1.0
        toks = [int(v) for v in rng.integers(1, 9, len(pos))]
        if rng.random() < 0.5:
            toks[-1] = 2
        bt = float(rng.choice([0.0, 0.0, 250.0]))
        assert time_stamp_lfr6_onnx(pk, toks, begin_time=bt) == sanm.time_stamp_lfr6_onnx(pk, toks, begin_time=bt)
    with pytest.raises(IndexError):                       # fire_place[0] on an empty list
        time_stamp_lfr6_onnx(np.zeros(10, np.float32), [3, 4])
    with pytest.raises(IndexError):                       # tokens[i] past the row
        pk = np.zeros(50, np.float32)
        pk[[5, 10, 15, 20]] = 1.0
        time_stamp_lfr6_onnx(pk, [3])


def test_decode_survives_arbitrary_bytes_in_the_tokens_table():
    """tokens.txt is user data: lines with invalid UTF-8, stray bars / '@@' fragments and truncated sequences must decode
    without faults (the text functions work on bytes)."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    frag = [b"\xe2\x96\x81", b"@@", b"\xe4\xbd", b"\xe4\xbd\xa0", b"\xf0\x9d\x84", b"\xc3", b"A", b"\t", b"</s>", b"\xff", b""]
    for _ in range(300):
        lines = []
        for _ in range(int(rng.integers(3, 40))):
            lines.append(b"".join(frag[int(k)] for k in rng.integers(0, len(frag), int(rng.integers(0, 5)))))
        blob = b"\n".join(lines)
        h = C.c_void_p()
        st = lib.pf_tokens_create_from_memory(blob, len(blob), C.byref(h))
        if st != _lib.PF_OK:
            continue
        n = lib.pf_tokens_count(h)
        ids = rng.integers(0, max(n, 1), int(rng.integers(0, 30))).astype(np.int32)
        res = _lib.PfTextResult()
        p = ids.ctypes.data_as(C.POINTER(C.c_int32))
        st = lib.pf_decode_offline(h, p, ids.size, None, 0, C.byref(res))
        # PF_ERR_SHAPE: a table line such as "@@▁" makes the C# call Last() on an empty list (InvalidOperationException)
        assert st in (_lib.PF_OK, _lib.PF_ERR_SHAPE)
        if st != _lib.PF_OK:
            lib.pf_tokens_destroy(h)
            continue
        text = C.create_string_buffer(res.text_bytes + 1)
        toks = C.create_string_buffer(max(1, res.tokens_bytes))
        ts = np.zeros(max(1, res.ts_count), np.int32)
        off = np.zeros(res.n_timestamps + 1, np.int32)
        res.text, res.text_capacity = C.cast(text, C.c_void_p), len(text)
        res.tokens, res.tokens_capacity = C.cast(toks, C.c_void_p), len(toks)
        res.ts, res.ts_capacity = ts.ctypes.data_as(C.POINTER(C.c_int32)), ts.size
        res.ts_offsets, res.ts_offsets_capacity = off.ctypes.data_as(C.POINTER(C.c_int32)), off.size
        assert lib.pf_decode_offline(h, p, ids.size, None, 0, C.byref(res)) == _lib.PF_OK
        assert off[res.n_timestamps] == res.ts_count and len(text.value) <= res.text_bytes
        need = C.c_size_t(0)
        assert lib.pf_decode_online(h, p, ids.size, None, 0, C.byref(need)) == _lib.PF_OK
        buf = C.create_string_buffer(need.value + 1)
        assert lib.pf_decode_online(h, p, ids.size, buf, len(buf), C.byref(need)) == _lib.PF_OK
        lib.pf_tokens_destroy(h)
