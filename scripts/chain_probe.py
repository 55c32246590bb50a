"""Fused feed-forward kernel (csrc/ffn_chain.cu) vs the two separate GEMM launches it replaces.  Tuning aid only.
    python scripts/chain_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_ffn_chain, dbg_gemm  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
D, F = 512, 2048
for M in (5344, 1328, 8768, 2656, 320, 10688):
    a = rng.standard_normal((M, D)).astype(np.float32)
    w1 = (rng.standard_normal((F, D)) / np.sqrt(D)).astype(np.float32)
    b1 = rng.standard_normal(F).astype(np.float32)
    w2 = (rng.standard_normal((D, F)) / np.sqrt(F)).astype(np.float32)
    b2 = rng.standard_normal(D).astype(np.float32)
    x = rng.standard_normal((M, D)).astype(np.float32)
    _, t1 = dbg_gemm(lib, a, w1, b1, None, None, relu=1, out_half=1, tile_n=0, iters=50)
    hh = np.maximum(a @ w1.T + b1, 0).astype(np.float32)
    _, t2 = dbg_gemm(lib, hh, w2, b2, x, None, relu=0, out_half=0, tile_n=0, iters=50)
    _, tc = dbg_ffn_chain(lib, a, w1, b1, w2, b2, x, iters=50)
    print(f"M={M:6d}: ffn1 {t1 * 1e3:6.2f} us + ffn2 {t2 * 1e3:6.2f} us = {(t1 + t2) * 1e3:6.2f} us | chain {tc * 1e3:6.2f} us "
          f"({4.0 * M * D * F / tc / 1e9:6.1f} TFLOP/s)", flush=True)
