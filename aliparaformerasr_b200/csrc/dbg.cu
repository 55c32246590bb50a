// Op-level test hooks of the C-ABI (pf_dbg_*): run one kernel on the current device with host fp32 in/out so that
// tests/ can compare each stage with the oracle.  Not on the product path.
#include <string.h>

#include <vector>

#include "../../include/pf_abi.h"
#include "attention.cuh"
#include "common.cuh"
#include "engine.cuh"
#include "gemm_ln.cuh"
#include "gemm.cuh"
#include "ops.cuh"

namespace pf {
void set_last_error(const std::string& msg);

namespace {

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() {
        cudaDeviceSynchronize();
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    T* alloc(size_t n) {
        void* p = nullptr;
        PF_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    float* up(const float* h, size_t n) {
        float* d = alloc<float>(n);
        PF_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
        return d;
    }
    __half* up_half(const float* h, size_t n) {
        float* t = up(h, n);
        __half* d = alloc<__half>(n);
        f32_to_f16_launch(t, d, n, 0);
        return d;
    }
};

__global__ void pf_dbg_f16_to_f32(const __half* in, float* out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __half2float(in[i]);
}

template <typename F>
pf_status guarded(F&& f) {
    try {
        f();
        return PF_OK;
    } catch (const StatusError& e) {
        set_last_error(e.what);
        return e.code;
    } catch (const CudaError& e) {
        set_last_error(e.what);
        cudaGetLastError();
        return PF_ERR_CUDA;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return PF_ERR_BAD_ARG;
    }
}

}  // namespace
}  // namespace pf

using namespace pf;

extern "C" {

pf_status pf_dbg_gemm(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias, const float* resid,
                      const float* addend, int32_t relu, int32_t out_half, int32_t tile_n, float* out, float* elapsed_ms,
                      int32_t iters) {
    return guarded([&] {
        Scratch s;
        __half* dA = s.up_half(A, static_cast<size_t>(M) * K);
        __half* dW = s.up_half(W, static_cast<size_t>(N) * K);
        GemmEpi e;
        if (bias) e.bias = s.up(bias, N);
        float* dres = nullptr;
        if (resid) { dres = s.up(resid, static_cast<size_t>(M) * N); e.resid = dres; e.ld_resid = N; }
        if (addend) { e.addend = s.up(addend, static_cast<size_t>(M) * N); e.ld_addend = N; }
        e.relu = relu;
        e.ld_out = N;
        float* o32 = s.alloc<float>(static_cast<size_t>(M) * N);
        __half* o16 = nullptr;
        if (out_half) { o16 = s.alloc<__half>(static_cast<size_t>(M) * N); e.out_f16 = o16; }
        else e.out_f32 = o32;
        GemmOp op;
        gemm_prepare(op, dA, K, dW, K, M, N, K, e, tile_n);
        gemm_launch(op, 0);
        PF_CUDA(cudaDeviceSynchronize());
        if (out_half) {
            pf_dbg_f16_to_f32<<<256, 256>>>(o16, o32, static_cast<size_t>(M) * N);
            PF_CUDA(cudaGetLastError());
        }
        PF_CUDA(cudaMemcpy(out, o32, static_cast<size_t>(M) * N * sizeof(float), cudaMemcpyDeviceToHost));
        if (elapsed_ms && iters > 0) {
            cudaEvent_t a, b;
            PF_CUDA(cudaEventCreate(&a));
            PF_CUDA(cudaEventCreate(&b));
            for (int i = 0; i < 3; ++i) gemm_launch(op, 0);
            PF_CUDA(cudaEventRecord(a, 0));
            for (int i = 0; i < iters; ++i) gemm_launch(op, 0);
            PF_CUDA(cudaEventRecord(b, 0));
            PF_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            PF_CUDA(cudaEventElapsedTime(&ms, a, b));
            *elapsed_ms = ms / iters;
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    });
}

pf_status pf_dbg_ln_gemm(int32_t M, int32_t N, const float* x, const float* gamma, const float* beta, float eps, const float* W,
                         const float* bias, int32_t relu, float* out, float* elapsed_ms, int32_t iters) {
    return guarded([&] {
        const int K = 512;
        Scratch s;
        float* dx = s.up(x, static_cast<size_t>(M) * K);
        __half* dW = s.up_half(W, static_cast<size_t>(N) * K);
        float* dg = s.up(gamma, K);
        float* db = s.up(beta, K);
        float* dbias = bias ? s.up(bias, N) : nullptr;
        __half* o16 = s.alloc<__half>(static_cast<size_t>(M) * N);
        float* o32 = s.alloc<float>(static_cast<size_t>(M) * N);
        LnGemmOp op;
        ln_gemm_prepare(op, dx, K, dg, db, eps, dW, K, dbias, o16, N, relu, M, N, K);
        ln_gemm_launch(op, 0);
        PF_CUDA(cudaDeviceSynchronize());
        pf_dbg_f16_to_f32<<<256, 256>>>(o16, o32, static_cast<size_t>(M) * N);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(out, o32, static_cast<size_t>(M) * N * sizeof(float), cudaMemcpyDeviceToHost));
        if (elapsed_ms && iters > 0) {
            cudaEvent_t a, b;
            PF_CUDA(cudaEventCreate(&a));
            PF_CUDA(cudaEventCreate(&b));
            for (int i = 0; i < 3; ++i) ln_gemm_launch(op, 0);
            PF_CUDA(cudaEventRecord(a, 0));
            for (int i = 0; i < iters; ++i) ln_gemm_launch(op, 0);
            PF_CUDA(cudaEventRecord(b, 0));
            PF_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            PF_CUDA(cudaEventElapsedTime(&ms, a, b));
            *elapsed_ms = ms / iters;
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    });
}

pf_status pf_dbg_gemm_pick(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias, int32_t tile_n,
                           int32_t* tokens) {
    return guarded([&] {
        if (!A || !W || !tokens) throw StatusError{PF_ERR_BAD_ARG, "null argument"};
        Scratch s;
        __half* dA = s.up_half(A, static_cast<size_t>(M) * K);
        __half* dW = s.up_half(W, static_cast<size_t>(N) * K);
        GemmEpi e;
        if (bias) e.bias = s.up(bias, N);
        const int slots = gemm_pick_slots(N);
        e.pick_out = s.alloc<float>(static_cast<size_t>(M) * slots * 3);
        e.pick_ld = slots;
        int* dtok = s.alloc<int>(M);
        GemmOp op;
        gemm_prepare(op, dA, K, dW, K, M, N, K, e, tile_n);
        gemm_launch(op, 0);
        pick_combine_launch(e.pick_out, M, slots, ceil_div(N, op.bn) * 2, dtok, 0);
        PF_CUDA(cudaDeviceSynchronize());
        PF_CUDA(cudaMemcpy(tokens, dtok, static_cast<size_t>(M) * sizeof(int), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_ffn_chain(int32_t M, int32_t D, int32_t F, const float* a, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* x, float* out, float* elapsed_ms, int32_t iters) {
#ifndef PFASR_EXPERIMENTS
    (void)M; (void)D; (void)F; (void)a; (void)w1; (void)b1; (void)w2; (void)b2; (void)x; (void)out; (void)elapsed_ms; (void)iters;
    return guarded([&] { throw StatusError{PF_ERR_UNSUPPORTED, "the fused feed-forward kernel is an experiment: build with PFASR_BUILD_EXPERIMENTS=1"}; });
#else
    return guarded([&] {
        Scratch s;
        static FfnChainScratch sc;               // test hook: legacy stream of the current device
        if (!sc.flags) ffn_chain_scratch_create(sc);
        if (!ffn_chain_supported(M, D, F, sc)) throw CudaError{"shape cannot run as a feed-forward chain"};
        __half* da = s.up_half(a, static_cast<size_t>(M) * D);
        __half* dw1 = s.up_half(w1, static_cast<size_t>(F) * D);
        __half* dw2 = s.up_half(w2, static_cast<size_t>(D) * F);
        float* db1 = s.up(b1, F);
        float* db2 = s.up(b2, D);
        float* dx = s.up(x, static_cast<size_t>(M) * D);
        __half* h = s.alloc<__half>(static_cast<size_t>(M) * F);
        FfnChainOp op;
        ffn_chain_prepare(op, da, D, dw1, db1, h, F, dw2, db2, dx, D, M, D, F, sc);
        ffn_chain_launch(op, 0);
        PF_CUDA(cudaDeviceSynchronize());
        PF_CUDA(cudaMemcpy(out, dx, static_cast<size_t>(M) * D * sizeof(float), cudaMemcpyDeviceToHost));
        if (elapsed_ms && iters > 0) {
            cudaEvent_t e0, e1;
            PF_CUDA(cudaEventCreate(&e0));
            PF_CUDA(cudaEventCreate(&e1));
            for (int i = 0; i < 3; ++i) ffn_chain_launch(op, 0);
            PF_CUDA(cudaEventRecord(e0, 0));
            for (int i = 0; i < iters; ++i) ffn_chain_launch(op, 0);
            PF_CUDA(cudaEventRecord(e1, 0));
            PF_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            PF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *elapsed_ms = ms / iters;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    });
#endif
}

int32_t pf_build_experiments(void) {
#ifdef PFASR_EXPERIMENTS
    return 1;
#else
    return 0;
#endif
}

pf_status pf_dbg_gemm_ln(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias, const float* resid,
                         const float* gamma, const float* beta, float eps, float* out, float* out_ln) {
    return guarded([&] {
        if (!pf_build_experiments()) throw StatusError{PF_ERR_UNSUPPORTED, "the fused-LayerNorm GEMM epilogue is an experiment: build with PFASR_BUILD_EXPERIMENTS=1"};
        if (!gemm_ln_fusable(M, N)) throw CudaError{"shape cannot carry the fused LayerNorm epilogue"};
        Scratch s;
        __half* dA = s.up_half(A, static_cast<size_t>(M) * K);
        __half* dW = s.up_half(W, static_cast<size_t>(N) * K);
        GemmEpi e;
        if (bias) e.bias = s.up(bias, N);
        float* x = s.up(resid, static_cast<size_t>(M) * N);           // in-place residual, like the encoder's x32
        e.resid = x; e.ld_resid = N; e.out_f32 = x; e.ld_out = N;
        e.ln_gamma = s.up(gamma, N); e.ln_beta = s.up(beta, N); e.ln_eps = eps;
        __half* ln16 = s.alloc<__half>(static_cast<size_t>(M) * N);
        float* ln32 = s.alloc<float>(static_cast<size_t>(M) * N);
        e.ln_out16 = ln16; e.ld_ln16 = N;
        GemmOp op;
        gemm_prepare(op, dA, K, dW, K, M, N, K, e, 0);
        gemm_launch(op, 0);
        PF_CUDA(cudaDeviceSynchronize());
        pf_dbg_f16_to_f32<<<256, 256>>>(ln16, ln32, static_cast<size_t>(M) * N);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(out, x, static_cast<size_t>(M) * N * sizeof(float), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(out_ln, ln32, static_cast<size_t>(M) * N * sizeof(float), cudaMemcpyDeviceToHost));
        if (const char* it = getenv("PFASR_DBG_TIME_ITERS")) {        // tuning aid: back-to-back launches, fused vs GEMM + LayerNorm
            const int iters = atoi(it);
            GemmEpi e2 = e;
            e2.ln_out16 = nullptr; e2.ln_gamma = nullptr; e2.ln_beta = nullptr;
            GemmOp plain;
            gemm_prepare(plain, dA, K, dW, K, M, N, K, e2, 0);
            cudaEvent_t a, b;
            PF_CUDA(cudaEventCreate(&a));
            PF_CUDA(cudaEventCreate(&b));
            float ms[2] = {0, 0};
            for (int variant = 0; variant < 2; ++variant) {
                for (int i = 0; i < iters + 3; ++i) {
                    if (i == 3) PF_CUDA(cudaEventRecord(a, 0));
                    if (variant == 0) gemm_launch(op, 0);
                    else { gemm_launch(plain, 0); layernorm_f32_launch(x, N, M, N, e.ln_gamma, e.ln_beta, eps, ln16, N, nullptr, 0, 0); }
                }
                PF_CUDA(cudaEventRecord(b, 0));
                PF_CUDA(cudaEventSynchronize(b));
                PF_CUDA(cudaEventElapsedTime(&ms[variant], a, b));
            }
            printf("gemm_ln %d x %d x %d: fused %.2f us (tile %d, cluster %d) vs gemm (tile %d) + layernorm %.2f us\n", M, N, K,
                   ms[0] * 1e3f / iters, op.bn, op.ln_cluster, plain.bn, ms[1] * 1e3f / iters);
            fflush(stdout);
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    });
}

pf_status pf_dbg_audio_convert(const pf_audio* a, float* out, int64_t capacity, int64_t* n) {
    return guarded([&] {
        if (!a || !n) throw StatusError{PF_ERR_BAD_ARG, "null argument"};
        const long long ns = audio_num_samples(*a);
        *n = ns;
        if (ns == 0 || !out || capacity <= 0) return;
        Scratch s;
        const size_t nb = static_cast<size_t>(a->n_values) * audio_bytes_per_value(a->format);
        unsigned char* raw = s.alloc<unsigned char>(nb + 16);
        PF_CUDA(cudaMemcpy(raw, a->data, nb, cudaMemcpyHostToDevice));
        const AudioItem item{0, a->n_values, a->format, a->channels, a->sample_rate, 0};
        AudioItem* d_item = s.alloc<AudioItem>(1);
        PF_CUDA(cudaMemcpy(d_item, &item, sizeof(item), cudaMemcpyHostToDevice));
        const long long off = 0;
        const int cnt = static_cast<int>(ns);
        long long* d_off = s.alloc<long long>(1);
        int* d_n = s.alloc<int>(1);
        PF_CUDA(cudaMemcpy(d_off, &off, sizeof(off), cudaMemcpyHostToDevice));
        PF_CUDA(cudaMemcpy(d_n, &cnt, sizeof(cnt), cudaMemcpyHostToDevice));
        float* pcm = s.alloc<float>(static_cast<size_t>(ns));
        audio_convert_launch(raw, d_item, pcm, d_off, d_n, 1, cnt, 0);
        PF_CUDA(cudaDeviceSynchronize());
        PF_CUDA(cudaMemcpy(out, pcm, static_cast<size_t>(std::min<long long>(ns, capacity)) * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_layernorm(int32_t M, int32_t D, const float* x, const float* gamma, const float* beta, float eps, float* out) {
    return guarded([&] {
        Scratch s;
        float* dx = s.up(x, static_cast<size_t>(M) * D);
        float* g = s.up(gamma, D);
        float* b = s.up(beta, D);
        float* o = s.alloc<float>(static_cast<size_t>(M) * D);
        layernorm_f32_launch(dx, D, M, D, g, b, eps, nullptr, 0, o, D, 0);
        PF_CUDA(cudaMemcpy(out, o, static_cast<size_t>(M) * D * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_embed_pe_ln(int32_t B, int32_t T, int32_t D, const float* feats, float scale, const float* gamma, const float* beta,
                             float eps, float* out) {
    return guarded([&] {
        Scratch s;
        const size_t n = static_cast<size_t>(B) * T * D;
        float* dx = s.up(feats, n);
        float* g = s.up(gamma, D);
        float* b = s.up(beta, D);
        const int half = D / 2;
        std::vector<float> inv(half);
        const float inc = static_cast<float>(log(10000.0) / (half - 1));
        for (int i = 0; i < half; ++i) inv[i] = expf(static_cast<float>(i) * -inc);
        float* dinv = s.up(inv.data(), half);
        __half* o16 = s.alloc<__half>(n);
        float* o32 = s.alloc<float>(n);
        float* pe = s.alloc<float>(static_cast<size_t>(T) * D);
        pe_table_launch(pe, T, D, dinv, 0);
        embed_pe_ln_launch(dx, B * T, T, D, scale, pe, g, b, eps, o16, 0);
        pf_dbg_f16_to_f32<<<256, 256>>>(o16, o32, n);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(out, o32, n * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_attention(int32_t B, int32_t H, int32_t Tq, int32_t Tk, const float* q, const float* k, const float* v, float* out) {
    return guarded([&] {
        Scratch s;
        const int D = H * 128;
        const size_t nq = static_cast<size_t>(B) * Tq * D, nk = static_cast<size_t>(B) * Tk * D;
        __half* dq = s.up_half(q, nq);
        __half* dk = s.up_half(k, nk);
        __half* dv = s.up_half(v, nk);
        __half* o16 = s.alloc<__half>(nq);
        float* o32 = s.alloc<float>(nq);
        const size_t ws_bytes = attention_split_workspace_bytes(B, H, Tq);          // lets the streaming kernel split a long memory
        float* ws = s.alloc<float>(ws_bytes / sizeof(float));
        attention_launch(dq, dk, dv, o16, B, H, Tq, Tk, D, D, D, D, 128, 0, ws, ws_bytes);
        pf_dbg_f16_to_f32<<<256, 256>>>(o16, o32, nq);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(out, o32, nq * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_attention_fsmn(int32_t B, int32_t H, int32_t T, int32_t taps, const float* qkv, const float* w, float* ctx, float* mem) {
    return guarded([&] {
        Scratch s;
        const int D = H * 128;
        const size_t n = static_cast<size_t>(B) * T * D;
        __half* dqkv = s.up_half(qkv, 3 * n);                    // [B*T, 3*D] = q | k | v
        float* dw = s.up(w, static_cast<size_t>(D) * taps);
        __half* o16 = s.alloc<__half>(n);
        float* o32 = s.alloc<float>(n);
        float* m32 = s.alloc<float>(n);
        attention_fsmn_launch(dqkv, dqkv + D, dqkv + 2 * D, o16, B, H, T, 3 * D, D, dw, taps, m32, D, false, 0);
        pf_dbg_f16_to_f32<<<256, 256>>>(o16, o32, n);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(ctx, o32, n * sizeof(float), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(mem, m32, n * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_fsmn(int32_t B, int32_t T, int32_t D, int32_t K, const float* x, const float* w, const float* resid,
                      const int32_t* lens, int32_t half_input, float* out) {
    return guarded([&] {
        Scratch s;
        const size_t n = static_cast<size_t>(B) * T * D;
        float* dw = s.up(w, static_cast<size_t>(D) * K);
        float* dres = resid ? s.up(resid, n) : nullptr;
        int* dl = nullptr;
        if (lens) {
            dl = s.alloc<int>(B);
            PF_CUDA(cudaMemcpy(dl, lens, B * sizeof(int), cudaMemcpyHostToDevice));
        }
        float* o = s.alloc<float>(n);
        if (half_input) fsmn_f16_launch(s.up_half(x, n), D, dw, K, o, D, dres, D, dl, B, T, D, 0);
        else fsmn_f32_launch(s.up(x, n), D, dw, K, o, D, dres, D, dl, B, T, D, 0);
        PF_CUDA(cudaMemcpy(out, o, n * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_cif(int32_t B, int32_t T, int32_t D, const float* hidden, const float* alphas_with_tail, float threshold,
                     int32_t lcap, float* embeds, int32_t* token_num, int32_t* fires, float* peaks) {
    return guarded([&] {
        Scratch s;
        const int T1 = T + 1;
        float* dh = s.up(hidden, static_cast<size_t>(B) * T * D);
        float* da = s.up(alphas_with_tail, static_cast<size_t>(B) * T1);
        float* wc = s.alloc<float>(static_cast<size_t>(B) * T1);
        float* wr = s.alloc<float>(static_cast<size_t>(B) * T1);
        float* pk = s.alloc<float>(static_cast<size_t>(B) * T1);
        int* fi = s.alloc<int>(static_cast<size_t>(B) * T1);
        int* tn = s.alloc<int>(B);
        int* fr = s.alloc<int>(B);
        int* meta = s.alloc<int>(4);
        PF_CUDA(cudaMemset(meta, 0, 4 * sizeof(int)));
        float* emb = s.alloc<float>(static_cast<size_t>(B) * lcap * D);
        PF_CUDA(cudaMemset(emb, 0, static_cast<size_t>(B) * lcap * D * sizeof(float)));
        cif_scan_launch(da, B, T1, threshold, wc, wr, fi, pk, tn, fr, meta, 0);
        cif_gather_launch(dh, B, T, D, wc, wr, fi, T1, emb, lcap, 0);
        PF_CUDA(cudaMemcpy(embeds, emb, static_cast<size_t>(B) * lcap * D * sizeof(float), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(token_num, tn, B * sizeof(int), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(fires, fr, B * sizeof(int), cudaMemcpyDeviceToHost));
        if (peaks) PF_CUDA(cudaMemcpy(peaks, pk, static_cast<size_t>(B) * T1 * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

pf_status pf_dbg_logsoftmax_argmax(int32_t M, int32_t V, float* logits_inout, int32_t* tokens) {
    return guarded([&] {
        Scratch s;
        float* d = s.up(logits_inout, static_cast<size_t>(M) * V);
        int* t = s.alloc<int>(M);
        logsoftmax_argmax_launch(d, M, V, V, t, 1, 0);
        PF_CUDA(cudaMemcpy(logits_inout, d, static_cast<size_t>(M) * V * sizeof(float), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(tokens, t, M * sizeof(int), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
