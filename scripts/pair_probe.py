"""CTA-pair (cta_group::2) vs single-CTA tiles of the tcgen05 GEMM on the encoder's shapes.  Tuning aid only.
    PFASR_GEMM_PAIR=1 python scripts/pair_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_gemm  # noqa: E402

SHAPES = [(5344, 1536, 512, "f16", "enc qkv"), (5344, 512, 512, "f32+res", "enc out"), (5344, 2048, 512, "f16+relu", "enc ffn1"),
          (5344, 512, 2048, "f32+res", "enc ffn2"), (10688, 2048, 512, "f16+relu", "2x ffn1"), (10688, 512, 2048, "f32+res", "2x ffn2")]
lib = _lib.load()
rng = np.random.default_rng(0)
for M, N, K, epi, name in SHAPES:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    resid = rng.standard_normal((M, N)).astype(np.float32) if "res" in epi else None
    rows = []
    for tile in (0, 128, 192, 256):
        for cm in ((0,) if tile == 0 else (1, 2)):
            _, ms = dbg_gemm(lib, A, W, bias, resid, None, relu=int("relu" in epi), out_half=int(epi.startswith("f16")),
                             tile_n=tile | (cm << 12), iters=50)
            rows.append(f"{'auto' if tile == 0 else str(tile) + ('pair' if cm == 2 else '')} {ms * 1e3:6.2f}us")
    print(f"{name:9s} {M:5d}x{N:5d}x{K:4d} | " + " | ".join(rows), flush=True)
