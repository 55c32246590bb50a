"""Bottleneck probe for the tcgen05 GEMM: times a few path shapes with parts of the kernel disabled
(PFASR_GEMM_DBG bit 1 = no epilogue, 2 = no TMA, 4 = no MMA; results are garbage by design).  Tuning aid only.
    PFASR_GEMM_DBG=<mask> python scripts/gemm_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_gemm  # noqa: E402

SHAPES = [(5312, 2048, 512, "f16+relu"), (5312, 512, 2048, "f32+res"), (5312, 512, 512, "f32+res"), (5312, 16384, 512, "f16"),
          (8192, 8192, 2048, "f16")]
lib = _lib.load()
rng = np.random.default_rng(0)
mask = os.environ.get("PFASR_GEMM_DBG", "0")
for M, N, K, epi in SHAPES:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    resid = rng.standard_normal((M, N)).astype(np.float32) if "res" in epi else None
    for tile, cm in ((256, 1), (224, 1), (192, 1), (160, 1), (128, 1), (0, 0)):
        _, ms = dbg_gemm(lib, A, W, bias, resid, None, relu=int("relu" in epi), out_half=int(epi.startswith("f16")),
                         tile_n=tile | (cm << 12), iters=30)
        print(f"dbg={mask} {M:6d} {N:6d} {K:5d} {epi:10s} tile {tile} pair {cm}: {ms * 1e3:8.2f} us {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
