"""Builds ``libpfasr.so`` in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m aliparaformerasr_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libpfasr.so"
# PFASR_BUILD_EXPERIMENTS=1 additionally compiles the variants that were measured SLOWER than the selected path and are kept
# for A/B work only: the fused FFN1->FFN2 kernel (ffn_chain.cu), the CTA-pair (cta_group::2) GEMM, the
# sixteen-epilogue-warp GEMM, the fused-LayerNorm GEMM epilogue, the half-SM GEMM (gemm_half.cu: two CTAs per SM) and the
# row-tile-stationary LayerNorm + GEMM kernel (gemm_ln.cu).
# The product library carries only the selected path.
EXPERIMENTS = os.environ.get("PFASR_BUILD_EXPERIMENTS", "0") == "1"
SOURCES = ["gemm.cu", "frontend.cu", "audio.cu", "ops.cu", "attention.cu", "attention_tc.cu", "online.cu", "timestamp.cu", "text.cu", "engine.cu", "abi.cu", "dbg.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
if EXPERIMENTS:
    SOURCES.insert(1, "ffn_chain.cu")
    SOURCES.insert(1, "gemm_ln.cu")
    SOURCES.insert(1, "gemm_half.cu")
    NVCC_FLAGS.append("-DPFASR_EXPERIMENTS=1")


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libpfasr can only be built with the CUDA toolkit")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    h.update(b"experiments" if EXPERIMENTS else b"product")
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "pf_abi.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / "stamp"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = OBJ / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
