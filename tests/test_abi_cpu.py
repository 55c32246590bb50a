"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/pf_abi.h declares, validates arguments
before touching a device, and (on a box without a GPU) refuses to run instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from aliparaformerasr_b200 import _lib, synth, weights
from aliparaformerasr_b200.engine import Engine, to_pf_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pf_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pf_abi.h but not exported by libpfasr.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.pf_abi_version() == 5


def test_config_struct_layout():
    assert C.sizeof(_lib.PfConfig) == 4 * 31
    c = to_pf_config(synth.paraformer_large())
    assert (c.input_size, c.d_model, c.enc_layers, c.dec_layers, c.vocab, c.lfr_m, c.lfr_n) == (560, 512, 50, 16, 8404, 7, 6)
    assert to_pf_config(synth.sensevoice_small()).model_kind == _lib.PF_MODEL_SENSEVOICE_SMALL


def test_bad_arguments_are_rejected_before_any_device_work(lib):
    h = C.c_void_p()
    cfg = to_pf_config(synth.tiny())
    blob = weights.pack({"x": np.zeros(4, np.float32)})
    bad = to_pf_config(synth.tiny())
    bad.struct_bytes = 12
    assert lib.pf_offline_create_from_memory(C.byref(bad), blob.ctypes.data_as(C.c_void_p), blob.nbytes, None, 0, C.byref(h)) == _lib.PF_ERR_BAD_ARG
    assert b"struct_bytes" in lib.pf_last_error()
    bad = to_pf_config(synth.tiny())
    bad.d_model = 256
    assert lib.pf_offline_create_from_memory(C.byref(bad), blob.ctypes.data_as(C.c_void_p), blob.nbytes, None, 0, C.byref(h)) == _lib.PF_ERR_UNSUPPORTED
    assert lib.pf_offline_create(C.byref(cfg), b"", None, 0, C.byref(h)) == _lib.PF_ERR_WEIGHTS
    assert lib.pf_offline_create(C.byref(cfg), b"/nonexistent/model.pfw", None, 0, C.byref(h)) == _lib.PF_ERR_WEIGHTS
    # disposed / null handle (ObjectDisposedException in the reference)
    res = _lib.PfResult()
    assert lib.pf_offline_run_staged(None, 0, C.byref(res)) == _lib.PF_ERR_DISPOSED
    assert lib.pf_offline_destroy(None) == _lib.PF_ERR_DISPOSED
    assert lib.pf_frontend_num_frames(None, 100) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = synth.tiny()
    with pytest.raises(_lib.PfError) as ei:
        Engine(cfg, synth.make_weights(cfg))
    assert ei.value.code == _lib.PF_ERR_CUDA and "no CPU fallback" in ei.value.message


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aliparaformerasr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_blob_roundtrip():
    w = synth.make_weights(synth.tiny())
    back = weights.unpack(weights.pack(w))
    assert set(back) == set(w)
    for k in w:
        assert back[k].shape == w[k].shape and np.array_equal(back[k], w[k])


def test_header_is_plain_c(tmp_path):
    """include/pf_abi.h is what a foreign-function binding reads: it must compile as C99 on its own (no C++-isms, no
    missing includes) and its structs must have the sizes the ctypes mirror uses."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_check.c"
    src.write_text(
        '#include "pf_abi.h"\n#include <stdio.h>\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %d\\n", sizeof(pf_config), sizeof(pf_result), sizeof(pf_online_result),\n'
        '                        sizeof(pf_text_result), sizeof(pf_audio), PF_ABI_VERSION); return 0; }\n')
    exe = tmp_path / "abi_check"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes[:5] == [C.sizeof(_lib.PfConfig), C.sizeof(_lib.PfResult), C.sizeof(_lib.PfOnlineResult), C.sizeof(_lib.PfTextResult),
                         C.sizeof(_lib.PfAudio)]
    assert sizes[5] == 5


def _build_example(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    exe = tmp_path / "offline_wav"
    libdir = os.path.join(ROOT, "aliparaformerasr_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "offline_wav.c"), "-L", libdir, "-lpfasr", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_consumer_links_and_reports_errors(tmp_path):
    """examples/offline_wav.c: a C99 program linked against libpfasr.so through include/pf_abi.h only."""
    import subprocess
    exe = _build_example(tmp_path)
    tok = tmp_path / "tokens.txt"
    tok.write_text("<blank>\n<s>\n</s>\n")
    r = subprocess.run([str(exe), "/nonexistent/model.pfw", str(tok), "/nonexistent/a.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open weights file" in r.stderr
    r = subprocess.run([str(exe), "/nonexistent/model.pfw", "/nonexistent/tokens.txt", "a.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "tokens" in r.stderr


def test_library_sass_uses_tcgen05_and_tma():
    """What the hot kernels are made of, read from the built library itself (cuobjdump -sass, sm_100a): 5th-generation
    tensor-core MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM), TMA loads, stores and reduce-adds (UTMALDG / UTMASTG /
    UTMAREDG) - not mma.sync re-compiles."""
    import shutil
    import subprocess
    from collections import Counter
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    r = subprocess.run([cuobjdump, "-sass", "-arch", "sm_100a", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0 or not r.stdout:
        r = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    per_kernel, cur = {}, None
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), Counter())
            continue
        if cur is not None:
            for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA"):
                if re.search(r"\b" + op + r"(\b|\.)", line):
                    cur[op] += 1
    gemm = [c for k, c in per_kernel.items() if "pf_gemm_f16_tn_tcgen05" in k]
    attn = [c for k, c in per_kernel.items() if "pf_sanm_attention_tc" in k]
    chain = [c for k, c in per_kernel.items() if "pf_ffn_chain_tcgen05" in k]
    assert gemm and attn
    assert bool(chain) == bool(lib_experiments()), "the fused feed-forward kernel ships only in PFASR_BUILD_EXPERIMENTS=1 builds"
    for c in gemm + chain:
        assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["UTMALDG"] > 0 and c["HMMA"] == 0
    assert any(c["UTMASTG"] > 0 for c in gemm)                       # asynchronous TMA-store epilogue
    assert any(c["UTMAREDG"] > 0 for c in gemm)                      # in-place residual: fp32 reduce-add by the TMA engine
    for c in attn:
        assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["UTMALDG"] > 0 and c["UTMAREDG"] > 0   # FSMN memory leaves by TMA reduce-add


def lib_experiments():
    return _lib.load().pf_build_experiments()


def test_corrupted_weight_blobs_are_refused():
    """PFW1 files are untrusted input: truncations and byte flips end in PF_ERR_WEIGHTS (or, when the table still parses,
    in whatever comes next - no device here, a missing tensor on a GPU box), never in a fault."""
    import torch
    lib = _lib.load()
    rng = np.random.default_rng(3)
    base = weights.pack({"a.weight": rng.standard_normal((3, 5)).astype(np.float32), "b": rng.standard_normal(7).astype(np.float32)})
    cfg = to_pf_config(synth.tiny())
    seen = set()
    for _ in range(2000):
        blob = bytearray(base.tobytes())
        if rng.random() < 0.4:
            blob = blob[: int(rng.integers(0, len(blob) + 1))]
        for _ in range(int(rng.integers(1, 8))):
            if blob:
                blob[int(rng.integers(0, min(len(blob), 400)))] = int(rng.integers(0, 256))     # header + entry table
        buf = np.frombuffer(bytes(blob) + b"\0", dtype=np.uint8)
        h = C.c_void_p()
        st = lib.pf_offline_create_from_memory(C.byref(cfg), buf.ctypes.data_as(C.c_void_p), len(blob), None, 0, C.byref(h))
        seen.add(st)
        assert st != _lib.PF_OK and not h
    allowed = {_lib.PF_ERR_WEIGHTS, _lib.PF_ERR_CUDA} if not torch.cuda.is_available() else {_lib.PF_ERR_WEIGHTS}
    assert _lib.PF_ERR_WEIGHTS in seen and seen <= allowed | {_lib.PF_ERR_BAD_ARG}
