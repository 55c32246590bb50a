"""One GEMM shape through the C-ABI test hook (for ncu captures):  python scripts/gemm_one.py M N K tile_code [epi] [iters]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_gemm  # noqa: E402

M, N, K, code = (int(x, 0) for x in sys.argv[1:5])
epi = sys.argv[5] if len(sys.argv) > 5 else "f16"
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32)
W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
bias = rng.standard_normal(N).astype(np.float32)
resid = rng.standard_normal((M, N)).astype(np.float32) if "res" in epi else None
addend = rng.standard_normal((M, N)).astype(np.float32) if "add" in epi else None
_, ms = dbg_gemm(_lib.load(), A, W, bias, resid, addend, relu=int("relu" in epi), out_half=int(epi.startswith("f16")), tile_n=code, iters=iters)
print(f"{M}x{N}x{K} code {code:#x} {epi}: {ms * 1e3:.2f} us, {2.0 * M * N * K / (ms * 1e-3) / 1e12:.1f} TFLOP/s")
