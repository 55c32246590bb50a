"""cuBLAS (torch.matmul, fp16) timings for the GEMM shapes of the offline path: a calibration of what a tuned library
reaches on these mid-size shapes (NOT part of the product path or of any reported bench value).
"""
import sys
import torch

SHAPES = [(5312, 1536, 512), (5312, 512, 512), (5312, 2048, 512), (5312, 512, 2048), (5312, 16384, 512),
          (1600, 2048, 512), (1600, 512, 2048), (1600, 512, 512), (1600, 8404, 512), (8192, 8192, 8192)]
for M, N, K in SHAPES:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16)
    for _ in range(5):
        torch.matmul(a, w.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 50 if M * N * K < 1e11 else 5
    e0.record()
    for _ in range(it):
        torch.matmul(a, w.t())
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / it
    print(f"{M:6d} {N:6d} {K:5d}: {us:8.2f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
