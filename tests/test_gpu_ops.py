"""Op-level parity of every CUDA kernel against the CPU oracle, through the C-ABI test hooks."""
import ctypes as C

import numpy as np
import pytest
import torch

from aliparaformerasr_b200 import _lib
from oracle import sanm
from _util import dbg_ffn_chain, dbg_gemm, f, half_round

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,tile", [
    (128, 128, 64, 128), (128, 64, 128, 64), (256, 256, 512, 256), (5312, 1536, 560, 0), (5312, 512, 512, 0),
    (5312, 2048, 512, 0), (5312, 512, 2048, 0), (1344, 8404, 512, 0), (83, 512, 512, 0), (100, 520, 72, 64),
    (300, 8404, 512, 128), (640, 1024, 512, 256), (5312, 512, 2048, 192), (333, 520, 512, 192), (1000, 1536, 512, 192), (5312, 1536, 512, 224), (700, 2048, 512, 160), (300, 8404, 512, 96), (129, 96, 64, 32), (1600, 2048, 512, 192 | (2 << 12)),
    # CTA-pair (cta_group::2) MMA, 256 x tile per pair: tile | (2 << 12)
    (5312, 1536, 512, 256 | (2 << 12)), (640, 1024, 512, 256 | (2 << 12)), (300, 8404, 512, 128 | (2 << 12)),
    (83, 512, 512, 128 | (2 << 12)), (5312, 512, 2048, 128 | (2 << 12)), (1600, 8404, 512, 256 | (2 << 12)),
    (1000, 520, 72, 128 | (2 << 12)), (129, 256, 64, 256 | (2 << 12)), (1600, 2048, 512, 256 | (2 << 12)), (200, 25055, 512, 256),
])
def test_gemm_plain(lib, M, N, K, tile):
    if (tile >> 12) & 0xF == 2 and not lib.pf_build_experiments():
        tile &= 0xFFF                       # CTA pairs exist in PFASR_BUILD_EXPERIMENTS=1 builds only: same shape, one CTA per tile
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    out, _ = dbg_gemm(lib, A, W, tile_n=tile)
    ref = half_round(A).astype(np.float64) @ half_round(W).astype(np.float64).T
    err = np.abs(out - ref).max()
    assert err < 2e-3, f"max abs err {err}"


@pytest.mark.parametrize("M,N,K,tile", [(5344, 1536, 512, 256), (5344, 2048, 512, 0), (300, 520, 192, 128), (129, 96, 64, 0)])
def test_gemm_f16_sixteen_epilogue_warps(lib, M, N, K, tile):
    """fp16 + ReLU output through the 640-thread variant (tile_code bit 22; opt-in on the product path)."""
    if not lib.pf_build_experiments():
        pytest.skip("measured-slower A/B variant: compiled only with PFASR_BUILD_EXPERIMENTS=1 (build.py)")
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    out, _ = dbg_gemm(lib, A, W, bias, relu=1, out_half=1, tile_n=tile | (1 << 22))
    ref = np.maximum(half_round(A).astype(np.float64) @ half_round(W).astype(np.float64).T + bias, 0.0)
    assert np.abs(out - ref).max() < 6e-3


@pytest.mark.parametrize("M,N,K,relu", [(5312, 1536, 512, 0), (5312, 2048, 512, 1), (5312, 1536, 560, 0), (5312, 16384, 512, 0), (1600, 2048, 512, 1),
                                        (83, 1536, 512, 0), (300, 1024, 192, 1), (129, 288, 64, 0), (5344, 2048, 2048, 1)])
def test_gemm_half_sm_kernel(lib, M, N, K, relu):
    """csrc/gemm_half.cu: one 128 x 256 tile per CTA, two CTAs per SM (tile_code bit 23), fp16 output with bias / ReLU;
    ragged M and N edges are clipped by the tensor maps."""
    if not lib.pf_build_experiments():
        pytest.skip("measured-slower A/B variant: compiled only with PFASR_BUILD_EXPERIMENTS=1 (build.py)")
    rng = np.random.default_rng(M + 3 * N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    out, _ = dbg_gemm(lib, A, W, bias, relu=relu, out_half=1, tile_n=256 | (1 << 23))
    ref = half_round(A).astype(np.float64) @ half_round(W).astype(np.float64).T + bias
    if relu:
        ref = np.maximum(ref, 0.0)
    assert np.abs(out - half_round(ref)).max() < 6e-3
    plain, _ = dbg_gemm(lib, A, W, bias, relu=relu, out_half=1, tile_n=256)
    assert np.array_equal(out, plain)                  # same MMAs, same epilogue arithmetic as the persistent kernel


@pytest.mark.parametrize("M,N,relu", [(5312, 1536, 0), (5312, 2048, 1), (8768, 2048, 1), (83, 1536, 0), (129, 288, 1), (1, 256, 0)])
def test_layernorm_gemm_rowtile_kernel(lib, M, N, relu):
    """csrc/gemm_ln.cu: LayerNorm of the fp32 residual rows + fp16-output GEMM in one row-tile-stationary kernel must give
    exactly what the two kernels it replaces give (same LayerNorm arithmetic, same MMA order, same epilogue rounding)."""
    if not lib.pf_build_experiments():
        pytest.skip("gemm_ln.cu is compiled only with PFASR_BUILD_EXPERIMENTS=1 (measured slower than LayerNorm + GEMM)")
    rng = np.random.default_rng(M + N)
    x = (rng.standard_normal((M, 512)) * 3 + 0.7).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(512)).astype(np.float32)
    b = (0.1 * rng.standard_normal(512)).astype(np.float32)
    W = (rng.standard_normal((N, 512)) / np.sqrt(512)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    out = np.zeros((M, N), np.float32)
    ms = C.c_float(0)
    _lib.check(lib.pf_dbg_ln_gemm(M, N, _lib.fptr(x), _lib.fptr(g), _lib.fptr(b), 1e-12, _lib.fptr(f(W)), _lib.fptr(bias), relu, _lib.fptr(out),
                                  C.byref(ms), 0))
    ln = np.zeros((M, 512), np.float32)
    _lib.check(lib.pf_dbg_layernorm(M, 512, _lib.fptr(x), _lib.fptr(g), _lib.fptr(b), 1e-12, _lib.fptr(ln)))     # fp16-rounded LN output
    two, _ = dbg_gemm(lib, ln, W, bias, relu=relu, out_half=1, tile_n=256)
    assert np.array_equal(out, two)
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (512,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
    ref = half_round(ref).astype(np.float64) @ half_round(W).astype(np.float64).T + bias
    if relu:
        ref = np.maximum(ref, 0.0)
    assert np.abs(out - ref).max() < 2e-2


@pytest.mark.parametrize("M,N,K,tile", [(1600, 8404, 512, 0), (300, 25055, 512, 0), (333, 8404, 512, 128), (129, 1000, 64, 256), (5, 40, 64, 0)])
def test_gemm_fused_greedy_pick(lib, M, N, K, tile):
    """K14: the greedy pick in the head GEMM's epilogue + pf_pick_combine == OfflineRecognizer.cs:145-149 on the same logits
    (last maximum wins on ties; a NaN restarts the scan; a NaN in the last column wins outright), for vocabularies that do
    not fill the last tile."""
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    # ties: duplicate weight rows -> bit-identical logits in two columns, made the row maximum through the bias
    W[N - 3] = W[7]
    bias[N - 3] = bias[7] = 9.0
    W[20] = W[N // 2]
    bias[20] = bias[N // 2]
    bias_nan = bias.copy()
    bias_nan[N // 3] = np.nan                          # a NaN column in every row: the scan restarts behind it
    for b in (bias, bias_nan):
        logits, _ = dbg_gemm(lib, A, W, b)
        want = sanm.greedy_pick(logits)
        got = np.zeros(M, np.int32)
        _lib.check(lib.pf_dbg_gemm_pick(M, N, K, _lib.fptr(f(A)), _lib.fptr(f(W)), _lib.fptr(f(b)), tile, _lib.iptr(got)))
        assert np.array_equal(got, want)
        assert (got == N - 3).mean() > 0.5             # the tie really decides rows
    last = bias.copy()
    last[N - 1] = np.nan                               # NaN in the last position wins
    got = np.zeros(M, np.int32)
    _lib.check(lib.pf_dbg_gemm_pick(M, N, K, _lib.fptr(f(A)), _lib.fptr(f(W)), _lib.fptr(f(last)), tile, _lib.iptr(got)))
    assert (got == N - 1).all()


@pytest.mark.parametrize("out_half,relu", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("adds", [0, 1, 2])
@pytest.mark.parametrize("N", [520, 517])          # 517: unaligned pitch -> scalar epilogue path
def test_gemm_epilogue(lib, out_half, relu, adds, N):
    rng = np.random.default_rng(5)
    M, K = 333, 512
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    resid = rng.standard_normal((M, N)).astype(np.float32) if adds >= 1 else None
    addend = rng.standard_normal((M, N)).astype(np.float32) if adds >= 2 else None
    out, _ = dbg_gemm(lib, A, W, bias, resid, addend, relu=relu, out_half=out_half)
    ref = half_round(A).astype(np.float64) @ half_round(W).astype(np.float64).T + bias
    if resid is not None:
        ref = ref + resid
    if addend is not None:
        ref = ref + addend
    if relu:
        ref = np.maximum(ref, 0)
    tol = 2e-2 if out_half else 2e-3
    assert np.abs(out - ref).max() < tol


@pytest.mark.parametrize("M,N,K", [(5312, 512, 512), (5312, 512, 2048), (1600, 512, 512), (333, 512, 512), (83, 512, 64), (700, 1024, 256)])
def test_gemm_fused_layernorm(lib, M, N, K):
    """out-projection / FFN2 epilogue: residual add + LayerNorm of the new rows across the CTAs of a cluster."""
    if not lib.pf_build_experiments():
        pytest.skip("measured-slower A/B variant: compiled only with PFASR_BUILD_EXPERIMENTS=1 (build.py)")
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    resid = (rng.standard_normal((M, N)) * 2 + 0.5).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(N)).astype(np.float32)
    b = (0.1 * rng.standard_normal(N)).astype(np.float32)
    out = np.zeros((M, N), np.float32)
    out_ln = np.zeros((M, N), np.float32)
    _lib.check(lib.pf_dbg_gemm_ln(M, N, K, _lib.fptr(f(A)), _lib.fptr(f(W)), _lib.fptr(bias), _lib.fptr(resid), _lib.fptr(g), _lib.fptr(b),
                                  1e-12, _lib.fptr(out), _lib.fptr(out_ln)))
    ref = half_round(A).astype(np.float64) @ half_round(W).astype(np.float64).T + bias + resid
    assert np.abs(out - ref).max() < 2e-3
    ref_ln = torch.nn.functional.layer_norm(torch.from_numpy(out), (N,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
    assert np.abs(out_ln - ref_ln).max() < 6e-3      # fp16 output of O(1..4) values


@pytest.mark.parametrize("M,D,F", [(5344, 512, 2048), (8768, 512, 2048), (1328, 512, 2048), (320, 512, 2048), (100, 512, 1024), (1, 256, 256),
                                   (20000, 512, 2048)])
def test_ffn_chain(lib, M, D, F):
    """Feed-forward block as one persistent kernel (csrc/ffn_chain.cu): x + relu(a W1^T + b1) W2^T + b2 with the fp16
    hidden activations handed from the first tile set to the second through flags.  Run twice: the flags carry an epoch."""
    if not lib.pf_build_experiments():
        pytest.skip("measured-slower A/B variant: compiled only with PFASR_BUILD_EXPERIMENTS=1 (build.py)")
    rng = np.random.default_rng(M + D + F)
    a = rng.standard_normal((M, D)).astype(np.float32)
    w1 = (rng.standard_normal((F, D)) / np.sqrt(D)).astype(np.float32)
    b1 = (0.3 * rng.standard_normal(F)).astype(np.float32)
    w2 = (rng.standard_normal((D, F)) / np.sqrt(F)).astype(np.float32)
    b2 = (0.3 * rng.standard_normal(D)).astype(np.float32)
    x = (rng.standard_normal((M, D)) * 2).astype(np.float32)
    h = half_round(np.maximum(half_round(a) @ half_round(w1).T + b1, 0.0))
    ref = x + h.astype(np.float64) @ half_round(w2).astype(np.float64).T + b2
    for _ in range(2):
        out, _ms = dbg_ffn_chain(lib, a, w1, b1, w2, b2, x)
        assert np.abs(out - ref).max() < 4e-3


@pytest.mark.parametrize("M,D", [(7, 512), (1000, 512), (333, 2048)])
def test_layernorm(lib, M, D):
    rng = np.random.default_rng(D + M)
    x = (rng.standard_normal((M, D)) * 3 + 1).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    b = (0.1 * rng.standard_normal(D)).astype(np.float32)
    out = np.zeros_like(x)
    _lib.check(lib.pf_dbg_layernorm(M, D, _lib.fptr(x), _lib.fptr(g), _lib.fptr(b), 1e-12, _lib.fptr(out)))
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (D,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
    assert np.abs(out - ref).max() < 1e-5


def test_embed_pe_ln(lib):
    rng = np.random.default_rng(11)
    B, T, D = 3, 50, 560
    x = (rng.standard_normal((B, T, D)) * 0.8 + 2.4).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    b = (0.1 * rng.standard_normal(D)).astype(np.float32)
    out = np.zeros_like(x)
    _lib.check(lib.pf_dbg_embed_pe_ln(B, T, D, _lib.fptr(x), float(np.sqrt(512.0)), _lib.fptr(g), _lib.fptr(b), 1e-12, _lib.fptr(out)))
    xt = torch.from_numpy(x) * (512 ** 0.5) + sanm.sinusoidal_pe(T, D)[None]
    ref = torch.nn.functional.layer_norm(xt, (D,), torch.from_numpy(g), torch.from_numpy(b), 1e-12).numpy()
    assert np.abs(out - ref).max() < 5e-3      # output is fp16


@pytest.mark.parametrize("B,Tq,Tk", [(2, 166, 166), (3, 40, 166), (1, 5, 7), (2, 64, 64), (1, 130, 300), (2, 256, 192), (1, 1, 1), (33, 50, 166)])
def test_attention(lib, B, Tq, Tk):
    rng = np.random.default_rng(Tq * 31 + Tk)
    H, D = 4, 512
    q = rng.standard_normal((B, Tq, D)).astype(np.float32)
    k = rng.standard_normal((B, Tk, D)).astype(np.float32)
    v = rng.standard_normal((B, Tk, D)).astype(np.float32)
    out = np.zeros_like(q)
    _lib.check(lib.pf_dbg_attention(B, H, Tq, Tk, _lib.fptr(q), _lib.fptr(k), _lib.fptr(v), _lib.fptr(out)))
    ref = sanm._mha(torch.from_numpy(half_round(q)), torch.from_numpy(half_round(k)), torch.from_numpy(half_round(v)), H).numpy()
    assert np.abs(out - ref).max() < 5e-3


@pytest.mark.parametrize("B,Tq,Tk", [(1, 100, 2010), (1, 800, 2010), (2, 30, 1000), (1, 64, 513)])
def test_attention_streaming_kernel_split_memory(lib, B, Tq, Tk):
    """Few query tiles against a long memory (the SeACo bias decoder: 800 rows x 2010 hot-word rows): the streaming kernel cuts the
    keys into runs and a combine kernel merges their (max, sum, unnormalised output) - same result as one walk over all keys."""
    rng = np.random.default_rng(Tq + Tk)
    H, D = 4, 512
    q = rng.standard_normal((B, Tq, D)).astype(np.float32)
    k = rng.standard_normal((B, Tk, D)).astype(np.float32)
    v = rng.standard_normal((B, Tk, D)).astype(np.float32)
    k[:, Tk - 7, :] = 0.35 * q[:, 0, :] if Tq > 0 else 0          # one row of the LAST run carries real weight for query 0
    out = np.zeros_like(q)
    _lib.check(lib.pf_dbg_attention(B, H, Tq, Tk, _lib.fptr(q), _lib.fptr(k), _lib.fptr(v), _lib.fptr(out)))
    ref = sanm._mha(torch.from_numpy(half_round(q)), torch.from_numpy(half_round(k)), torch.from_numpy(half_round(v)), H).numpy()
    assert np.abs(out - ref).max() < 5e-3


@pytest.mark.parametrize("B,Tq,Tk", [(2, 166, 166), (3, 40, 166), (2, 200, 192), (1, 1, 1), (2, 129, 16)])
def test_attention_streaming_kernel_matches(lib, B, Tq, Tk):
    """The mma.sync streaming kernel (long sequences) stays covered: force it with PFASR_NO_ATT_TC in a subprocess."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent(f"""
        import sys, numpy as np, torch
        sys.path.insert(0, {repr(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))}); sys.path.insert(0, {repr(os.path.dirname(os.path.abspath(__file__)))})
        from aliparaformerasr_b200 import _lib
        from oracle import sanm
        from _util import half_round
        lib = _lib.load(); rng = np.random.default_rng(7)
        B, Tq, Tk, H, D = {B}, {Tq}, {Tk}, 4, 512
        q = rng.standard_normal((B, Tq, D)).astype(np.float32); k = rng.standard_normal((B, Tk, D)).astype(np.float32); v = rng.standard_normal((B, Tk, D)).astype(np.float32)
        out = np.zeros_like(q)
        _lib.check(lib.pf_dbg_attention(B, H, Tq, Tk, _lib.fptr(q), _lib.fptr(k), _lib.fptr(v), _lib.fptr(out)))
        ref = sanm._mha(torch.from_numpy(half_round(q)), torch.from_numpy(half_round(k)), torch.from_numpy(half_round(v)), H).numpy()
        assert np.abs(out - ref).max() < 5e-3, np.abs(out - ref).max()
    """)
    env = dict(os.environ, PFASR_NO_ATT_TC="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("B,T,K", [(2, 166, 11), (3, 83, 11), (1, 137, 11), (2, 192, 21), (2, 16, 11), (1, 5, 11), (2, 200, 11)])
def test_attention_fsmn_fused(lib, B, T, K):
    """Encoder self-attention + FSMN memory in one launch (tcgen05 path for T <= 192; T = 200 takes the two-kernel path)."""
    rng = np.random.default_rng(T * 13 + K)
    H, D = 4, 512
    qkv = rng.standard_normal((B, T, 3 * D)).astype(np.float32)
    w = (0.1 * rng.standard_normal((D, 1, K))).astype(np.float32)
    ctx = np.zeros((B, T, D), np.float32)
    mem = np.zeros((B, T, D), np.float32)
    _lib.check(lib.pf_dbg_attention_fsmn(B, H, T, K, _lib.fptr(qkv), _lib.fptr(f(w.reshape(D, K))), _lib.fptr(ctx), _lib.fptr(mem)))
    hq = torch.from_numpy(half_round(qkv))
    q, k, v = hq[..., :D], hq[..., D:2 * D], hq[..., 2 * D:]
    ref_ctx = sanm._mha(q.contiguous(), k.contiguous(), v.contiguous(), H).numpy()
    ref_mem = sanm._fsmn(v.contiguous(), torch.from_numpy(w), None).numpy()
    assert np.abs(ctx - ref_ctx).max() < 5e-3
    assert np.abs(mem - ref_mem).max() < 1e-5


@pytest.mark.parametrize("K,half_in,masked", [(11, 1, False), (11, 0, True), (21, 0, True)])
def test_fsmn(lib, K, half_in, masked):
    rng = np.random.default_rng(K)
    B, T, D = 3, 45, 512
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    w = (0.1 * rng.standard_normal((D, 1, K))).astype(np.float32)
    resid = rng.standard_normal((B, T, D)).astype(np.float32)
    lens = np.array([45, 17, 0], dtype=np.int32)
    out = np.zeros_like(x)
    _lib.check(lib.pf_dbg_fsmn(B, T, D, K, _lib.fptr(x), _lib.fptr(f(w.reshape(D, K))), _lib.fptr(resid),
                               _lib.iptr(lens) if masked else None, half_in, _lib.fptr(out)))
    xin = torch.from_numpy(half_round(x) if half_in else x)
    mask = None
    if masked:
        mask = (torch.arange(T)[None, :] < torch.from_numpy(lens.astype(np.int64))[:, None]).float()[:, :, None]
    ref = (torch.from_numpy(resid) + sanm._fsmn(xin, torch.from_numpy(w), mask)).numpy()
    assert np.abs(out - ref).max() < 1e-5


def test_cif_matches_oracle_bit_exact(lib):
    rng = np.random.default_rng(3)
    B, T, D = 4, 166, 512
    hidden = rng.standard_normal((B, T, D)).astype(np.float32)
    alphas = np.concatenate([rng.uniform(0, 0.6, (B, T)).astype(np.float32), np.full((B, 1), 0.45, np.float32)], axis=1)
    alphas[3, :] = 0.0           # silent utterance: only the tail -> zero tokens
    alphas[3, T] = 0.45
    hid1 = np.concatenate([hidden, np.zeros((B, 1, D), np.float32)], axis=1)
    emb_ref, tn_ref, fires_ref, peaks_ref = sanm.cif(hid1, alphas, 1.0)
    lcap = T + 1
    emb = np.zeros((B, lcap, D), np.float32)
    tn = np.zeros(B, np.int32)
    fires = np.zeros(B, np.int32)
    peaks = np.zeros((B, T + 1), np.float32)
    _lib.check(lib.pf_dbg_cif(B, T, D, _lib.fptr(hidden), _lib.fptr(alphas), 1.0, lcap, _lib.fptr(emb), _lib.iptr(tn),
                              _lib.iptr(fires), _lib.fptr(peaks)))
    assert np.array_equal(tn, tn_ref)
    assert np.array_equal(fires, fires_ref)
    assert np.array_equal(peaks, peaks_ref)
    L = emb_ref.shape[1]
    assert np.array_equal(emb[:, :L], emb_ref)
    assert not emb[:, L:].any()


def test_logsoftmax_argmax_tie_and_nan_rules(lib):
    rng = np.random.default_rng(9)
    M, V = 37, 8404
    x = (rng.standard_normal((M, V)) * 4).astype(np.float32)
    x[0, 100] = x[0, 7000] = x[0].max() + 1          # tie -> largest index (Q5)
    x[1, :] = 0.25                                    # all equal -> V-1
    x[2, 50] = np.nan                                 # NaN restarts the scan after it
    x[3, V - 1] = np.nan                              # NaN last -> V-1
    ref_tok = sanm.greedy_pick(x.copy())
    ref_lp = torch.log_softmax(torch.from_numpy(x), dim=-1).numpy()
    y = x.copy()
    tok = np.zeros(M, np.int32)
    _lib.check(lib.pf_dbg_logsoftmax_argmax(M, V, _lib.fptr(y), _lib.iptr(tok)))
    assert np.array_equal(tok, ref_tok)
    assert tok[0] == 7000 and tok[1] == V - 1 and tok[3] == V - 1
    ok = ~np.isnan(ref_lp)
    assert np.abs(y[ok] - ref_lp[ok]).max() < 1e-4
