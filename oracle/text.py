"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's host
post-processing after the graph: tokens.txt loading and the DecodeMulti string joining of both recognisers.
(``time_stamp_lfr6_onnx`` lives in oracle/sanm.py.)

Pinned by hand-traced cases in tests/test_text.py that follow the C# statement by statement; the reference ships no
golden strings for this path.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

BAR = "▁"
_CHINESE = re.compile(r"^[一-龥]+$")          # OfflineRecognizer.cs:427-439 IsChinese(allMatch: true)
_SPECIAL = ("</s>", "<s>", "<blank>", "<unk>")


def read_all_lines(data: bytes) -> List[str]:
    """File.ReadAllLines (Utils/PreloadHelper.cs:138): CR, LF, CRLF end a line; no empty line after a final
    terminator; the StreamReader drops a UTF-8 BOM."""
    text = data.decode("utf-8-sig")
    lines = re.split(r"\r\n|\r|\n", text)
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def _remove_first_equal(lst: list, value) -> None:
    # List<T>.Remove(item): removes the first element that Equals(item)
    for i, v in enumerate(lst):
        if v == value:
            del lst[i]
            return


def decode_multi_offline(tokens_table: Sequence[str], ids: Sequence[int], timestamps: Sequence[Sequence[int]]
                         ) -> Tuple[str, int, List[str], List[List[int]]]:
    """OfflineRecognizer.DecodeMulti, per-stream body (OfflineRecognizer.cs:310-412) -> (Text, TextLen, Tokens,
    Timestamps).  Timestamps are compared by identity when removed (int[] reference equality in the C#)."""
    out_tokens: List[str] = []
    out_ts: List[list] = []
    text = ""
    last_token = ""
    last_ts: Optional[list] = None
    for token, ts in zip(ids, timestamps):
        if token == 2:
            break
        if not 0 <= token < len(tokens_table):
            raise IndexError(token)
        ts = list(ts)
        cur = tokens_table[token].split("\t")[0]
        if cur in _SPECIAL:
            continue
        if _CHINESE.match(cur):
            text += cur
            out_tokens.append(cur)
            out_ts.append(ts)
            continue
        text += BAR + cur + BAR
        joined = last_token + BAR + cur + BAR
        if joined.find("@@" + BAR + BAR) > 0:
            cur_token = joined.replace("@@" + BAR + BAR, "")
            cur_ts = ts if last_ts is None else list(last_ts) + list(ts)
            _remove_first_equal(out_tokens, out_tokens[-1])
            out_tokens.append(cur_token.replace(BAR, ""))
            out_ts.pop()                                   # reference equality: always the last entry
            out_ts.append(cur_ts)
            last_token, last_ts = cur_token, cur_ts
        elif joined.count(BAR) in (3, 5) and joined.find(BAR * 3) < 0:
            cur_token = joined.replace(BAR + BAR, "")
            cur_ts = ts if last_ts is None else list(last_ts) + list(ts)
            if out_tokens:
                _remove_first_equal(out_tokens, out_tokens[-1])
            out_tokens.append(cur_token.replace(BAR, ""))
            if out_ts:
                out_ts.pop()
            out_ts.append(cur_ts)
            last_token, last_ts = cur_token, cur_ts
        else:
            out_tokens.append(cur.replace(BAR, ""))
            out_ts.append(ts)
            last_token, last_ts = BAR + cur + BAR, ts
    if text.find("@@" + BAR + BAR) > 0 or text.find(BAR * 3) < 0:
        text = text.replace("@@" + BAR + BAR, "").replace(BAR + BAR, " ").replace("@@", " ").replace(BAR, " ")
    else:
        text = text.replace(BAR * 3, " ").replace(BAR + BAR, "").replace(BAR, "")
    text_len = len(text.encode("utf-16-le")) // 2          # string.Length counts UTF-16 code units
    return text, text_len, out_tokens, [list(t) for t in out_ts]


def decode_multi_online(tokens_table: Sequence[str], ids: Sequence[int]) -> str:
    """OnlineRecognizer.DecodeMulti, per-stream body (OnlineRecognizer.cs:408-432)."""
    text = ""
    for token in ids:
        if token == 2:
            break
        if not 0 <= token < len(tokens_table):
            raise IndexError(token)
        cur = tokens_table[token]
        if cur in _SPECIAL:
            continue
        text += cur if _CHINESE.match(cur) else BAR + cur + BAR
    return text.replace("@@" + BAR + BAR, "").replace("@@" + BAR, "").replace(BAR + BAR, " ").replace(BAR, "").lower()
