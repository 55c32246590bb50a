"""Steady-state k-block rate of the tcgen05 GEMM main loop: one tile per CTA (or CTA pair) with a long K, so the launch,
prologue and epilogue are amortised; per-SM operand ingest (bytes / k-block) is what differs between the variants.
    PFASR_LIB=aliparaformerasr_b200/_ab/libpfasr_exp.so python scripts/kblock_probe.py      (experiments build for the pairs)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_gemm  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
K = 8192
for ctas in (148, 74, 32):
    for N in (256,):
        M = 128 * ctas
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        for tile, cm in ((256, 1), (256, 2), (192, 1), (128, 1), (128, 2)):
            try:
                _, ms = dbg_gemm(lib, A, W, None, None, None, relu=0, out_half=1, tile_n=tile | (cm << 12), iters=20)
            except Exception as e:      # product build: no pairs
                print(f"tile {tile} cm {cm}: {e}")
                continue
            kb = K // 64 * ((N + tile - 1) // tile)
            print(f"ctas {ctas:3d} N {N} tile {tile:3d} cm {cm}: {ms * 1e3:8.2f} us  -> {ms * 1e3 / kb:6.3f} us per k-block "
                  f"({(16384 + tile * 128 // cm) / (ms * 1e3 / kb) / 1e3:6.1f} GB/s per SM, {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s)", flush=True)
