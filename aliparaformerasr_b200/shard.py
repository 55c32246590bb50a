"""Batch sharding helpers for the multi-GPU path.

Utterances never interact, so a batch shards data-parallel with no data-path collective: every GPU (one process per
GPU under torchrun, or one device thread inside a multi-device handle) runs the whole model on a contiguous slice.
The only exchange is the trivial result gather of token ids ("C1" in SURVEY 2.2), done here with torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def split_batch(batch: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous split of ``batch`` items over ``parts`` shards -> [(begin, count)], same rule as csrc/abi.cu."""
    base, rem = divmod(batch, parts)
    out, pos = [], 0
    for i in range(parts):
        cnt = base + (1 if i < rem else 0)
        out.append((pos, cnt))
        pos += cnt
    return out


def gather_tokens(tokens: np.ndarray, token_num: np.ndarray, width: int, device=None, dst: int = 0):
    """Gather per-rank ``tokens [B_r, L_r]`` / ``token_num [B_r]`` on ``dst``: rows are padded to ``width`` with -1 so
    ranks with different Lmax (it is data dependent) exchange fixed-size buffers.  Equal per-rank batch sizes.
    Returns (tokens [world*B_r, width], token_num [world*B_r]) on dst, (None, None) elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    b, l = tokens.shape
    if l > width:
        raise ValueError(f"token rows ({l}) wider than the exchange buffer ({width})")
    buf = torch.full((b, width + 1), -1, dtype=torch.int32)
    buf[:, :l] = torch.from_numpy(np.ascontiguousarray(tokens, dtype=np.int32))
    buf[:, width] = torch.from_numpy(np.ascontiguousarray(token_num, dtype=np.int32))
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None, None
    full = torch.cat(out, dim=0).cpu().numpy()
    return full[:, :width], full[:, width]
