"""Times the tcgen05 GEMM through the C-ABI test hook for the shapes of the offline path x every tile width.
Back-to-back launches (warm L2): use for relative tile/shape tuning, not as a bench value.
    python scripts/gemm_sweep.py [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aliparaformerasr_b200 import _lib  # noqa: E402
from _util import dbg_gemm  # noqa: E402

SHAPES = [  # (M, N, K, epilogue) of cfg2: encoder QKV / out / FFN1 / FFN2, decoder, head
    (5312, 1536, 512, "f16"), (5312, 512, 512, "f32+res+add"), (5312, 2048, 512, "f16+relu"), (5312, 512, 2048, "f32+res"),
    (5312, 16384, 512, "f16"), (1600, 2048, 512, "f32+relu"), (1600, 512, 2048, "f32"), (1600, 512, 512, "f32+res"),
    (1600, 8404, 512, "f32"),
]


def main():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    rows = []
    for M, N, K, epi in SHAPES:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        resid = rng.standard_normal((M, N)).astype(np.float32) if "res" in epi else None
        addend = rng.standard_normal((M, N)).astype(np.float32) if "add" in epi else None
        ref = None
        for tile, cm, cn in [(64, 1, 1)] + [(t, cm, cn) for t in (128, 256) for cm, cn in ((1, 1), (2, 1))] + [(0, 0, 0)]:
            code = tile | (cm << 12) | (cn << 16)
            out, ms = dbg_gemm(lib, A, W, bias, resid, addend, relu=int("relu" in epi), out_half=int(epi.startswith("f16")), tile_n=code, iters=50)
            if ref is None:
                ref = out
            err = float(np.abs(out - ref).max())      # every configuration must give the same result as 64 / 1x1
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            rows.append({"M": M, "N": N, "K": K, "epi": epi, "tile_n": tile, "cluster": f"{cm}x{cn}", "us": ms * 1e3, "tflops": tf, "max_diff_vs_first": err})
            print(f"{M:6d} {N:6d} {K:5d} {epi:12s} tile {tile:3d} cluster {cm}x{cn}: {ms * 1e3:8.2f} us  {tf:7.1f} TFLOP/s  diff {err:.2e}", flush=True)
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
