// Streaming (online) paraformer: the device side of OnlineStream / OnlineRecognizer state that the reference keeps in
// managed lists and re-marshals every 600 ms step.  Per stream slot, resident in HBM:
//   fbank FIFO (chunk slots), splice frame, 10-frame feature cache (scaled + position-encoded), CIF carry (alpha, hidden),
//   16 decoder FSMN caches [10, 512].
// Reference semantics (file:line are /root/reference/AliParaformerAsr/...):
//   OnlineStream.GetDecodeChunk      OnlineStream.cs:170-216      splice, LFR, CMVN, x sqrt(512), PE, feature cache (Q13, Q14)
//   OnlineWavFrontend.ApplyLfr / PE  OnlineWavFrontend.cs:73-91, 152-188 (Q12)
//   PadSequence_unittest             OnlineRecognizer.cs:459-471  exact zeros -> -23.025850929940457 (Q4 online)
//   DynamicMask + PredictorProj      OnlineModel.cs:141-165, OnlineRecognizer.cs:125-234 (Q15 host CIF with carry)
//   decoder FSMN caches              OnlineRecognizer.cs:281-293, OnlineModel.cs:207-254 (Q11)
#include "online.cuh"

#include <math.h>

namespace pf {

namespace {

constexpr float kOnlinePad = -23.025850929940457f;

// physical frame address inside a stream's fbank FIFO.  Chunk c lives in chunk slot c % nslot; chunk 0 stores its nf
// frames at rows 1..nf and row 0 aliases row 1 (the first-chunk frame repeat of OnlineStream.cs:141-153).
__device__ __forceinline__ const float* fifo_frame(const float* fifo, int slot, int nslot, int nf, int mel, long long f) {
    long long c, idx;
    if (f < nf + 1) { c = 0; idx = f == 0 ? 1 : f; }
    else { c = 1 + (f - nf - 1) / nf; idx = (f - nf - 1) % nf; }
    return fifo + ((static_cast<size_t>(slot) * nslot + static_cast<size_t>(c % nslot)) * (nf + 1) + static_cast<size_t>(idx)) * mel;
}

// grid (n streams, t_new LFR rows), 256 threads.  tab[b] = {slot, frame_base, has_splice, start_idx}
__global__ void __launch_bounds__(256)
pf_online_assemble(const OnlineDims od, const int* __restrict__ tab, const float* __restrict__ fifo,
                   const float* __restrict__ splice, const float* __restrict__ cache_feats, const float* __restrict__ shift,
                   const float* __restrict__ rescale, const float* __restrict__ inv_ts, float* __restrict__ feats,
                   float* __restrict__ fresh) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, i = blockIdx.y;
    const int slot = tab[b * 4 + 0];
    const long long base = tab[b * 4 + 1];
    const int has_splice = tab[b * 4 + 2];
    const int start_idx = tab[b * 4 + 3];
    const int dim = od.lfr_m * od.mel, half = dim >> 1;
    const int T = od.cache_rows + od.t_new;
    const float pos = static_cast<float>(start_idx + i + 1);
    const double sqrt_d = sqrt(static_cast<double>(od.d_model));          // Math.Pow(512, 0.5)
    for (int e = threadIdx.x; e < dim; e += blockDim.x) {
        const int j = i * od.lfr_n + e / od.mel, m = e % od.mel;          // frame j of the spliced window
        const float* src = (j == 0) ? (has_splice ? splice + static_cast<size_t>(slot) * od.mel
                                                  : fifo_frame(fifo, slot, od.nslot, od.nf, od.mel, base))
                                    : fifo_frame(fifo, slot, od.nslot, od.nf, od.mel, base + j - 1);
        float v = __fmul_rn(__fadd_rn(src[m], shift[e]), rescale[e]);    // OnlineWavFrontend.ApplyCmvn
        v = static_cast<float>(static_cast<double>(v) * sqrt_d);         // OnlineStream.cs:203
        const int k = e < half ? e : e - half;
        const float ang = __fmul_rn(inv_ts[k], pos);
        const float pe = static_cast<float>(e < half ? sin(static_cast<double>(ang)) : cos(static_cast<double>(ang)));
        v = __fadd_rn(v, pe);
        fresh[(static_cast<size_t>(b) * od.t_new + i) * dim + e] = v;
        feats[(static_cast<size_t>(b) * T + od.cache_rows + i) * dim + e] = v == 0.0f ? kOnlinePad : v;
        if (i < od.cache_rows) {
            const float c = cache_feats[(static_cast<size_t>(slot) * od.cache_rows + i) * dim + e];
            feats[(static_cast<size_t>(b) * T + i) * dim + e] = c == 0.0f ? kOnlinePad : c;
        }
    }
}

// after the window is assembled: feature cache <- the fresh rows, splice <- last fbank frame of the window
__global__ void __launch_bounds__(256)
pf_online_commit(const OnlineDims od, const int* __restrict__ tab, const float* __restrict__ fifo, const float* __restrict__ fresh,
                 float* __restrict__ splice, float* __restrict__ cache_feats) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x;
    const int slot = tab[b * 4 + 0];
    const long long base = tab[b * 4 + 1];
    const int dim = od.lfr_m * od.mel;
    const int n = od.cache_rows * dim;
    // cache = last cache_rows rows of [cache | fresh]; t_new == cache_rows in the reference's configuration
    for (int e = threadIdx.x; e < n; e += blockDim.x)
        cache_feats[static_cast<size_t>(slot) * n + e] = fresh[static_cast<size_t>(b) * od.t_new * dim + (od.t_new - od.cache_rows) * dim + e];
    const float* last = fifo_frame(fifo, slot, od.nslot, od.nf, od.mel, base + od.chunk_len - 1);
    for (int m = threadIdx.x; m < od.mel; m += blockDim.x) splice[static_cast<size_t>(slot) * od.mel + m] = last[m];
}

// DynamicMask + PredictorProj for one stream per block; thread = hidden channel.  Sequence = [carry | T encoder frames].
__global__ void __launch_bounds__(512)
pf_online_cif(const int* __restrict__ tab, const float* __restrict__ enc, const float* __restrict__ alphas, int ld_alpha, int T,
              int D, int mask_lo, int mask_hi, float threshold, float* __restrict__ carry_alpha, float* __restrict__ carry_hidden,
              float* __restrict__ frames_out, int lcap, int* __restrict__ counts, int* __restrict__ meta) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, d = threadIdx.x;
    const int slot = tab[b * 4 + 0];
    float integrate = 0.0f, frame = 0.0f;
    int fired = 0;
    for (int j = 0; j <= T && d < D; ++j) {
        float alpha, h;
        if (j == 0) {
            alpha = carry_alpha[slot];
            h = carry_hidden[static_cast<size_t>(slot) * D + d];
        } else {
            const int t = j - 1;
            alpha = (t >= mask_lo && t < mask_hi) ? alphas[static_cast<size_t>(b) * ld_alpha + t] : 0.0f;
            h = enc[(static_cast<size_t>(b) * T + t) * D + d];
        }
        if (__fadd_rn(alpha, integrate) < threshold) {
            integrate = __fadd_rn(integrate, alpha);
            frame = __fadd_rn(frame, __fmul_rn(alpha, h));
        } else {
            frame = __fadd_rn(frame, __fmul_rn(__fsub_rn(threshold, integrate), h));
            if (fired < lcap) frames_out[(static_cast<size_t>(b) * lcap + fired) * D + d] = frame;
            ++fired;
            integrate = __fadd_rn(integrate, alpha);
            integrate = __fsub_rn(integrate, threshold);
            frame = __fmul_rn(integrate, h);
        }
    }
    if (d < D) carry_hidden[static_cast<size_t>(slot) * D + d] = integrate > 0.0f ? __fdiv_rn(frame, integrate) : frame;
    __syncthreads();                       // every thread has read carry_alpha before it is replaced
    if (d == 0) {
        carry_alpha[slot] = integrate;
        counts[b] = fired;
        atomicMax(meta, fired);
    }
}

// acoustic_embeds [n, L, D] from the per-stream frame lists (zero padded past each stream's count)
__global__ void __launch_bounds__(256)
pf_online_compact(const float* __restrict__ frames, int lcap, const int* __restrict__ counts, int L, int D, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, l = blockIdx.x;
    const bool ok = l < counts[b];
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        out[(static_cast<size_t>(b) * L + l) * D + d] = ok ? frames[(static_cast<size_t>(b) * lcap + l) * D + d] : 0.0f;
}

// decoder FSMN with cache: xcat = [cache (K-1 rows) | v = tn * mask]; y[t] = sum_j w[j] xcat[t + j] + v[t]; x += y * mask;
// new cache = last K-1 rows of xcat.  thread = (stream, channel).
template <int K>
__global__ void __launch_bounds__(256)
pf_online_fsmn(const int* __restrict__ tab, const float* __restrict__ tn, const int* __restrict__ counts, int L, int D,
               const float* __restrict__ w, const float* __restrict__ cache_state, size_t state_stride, size_t layer_off,
               float* __restrict__ x, float* __restrict__ cache_out, size_t out_stride, size_t out_layer_off) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LMAX = 32;
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    const int slot = tab[b * 4 + 0];
    const int len = min(counts[b], L);
    float wk[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wk[j] = w[c * K + j];
    float xc[K - 1 + LMAX];
    const float* cin = cache_state + static_cast<size_t>(slot) * state_stride + layer_off;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) xc[j] = cin[static_cast<size_t>(j) * D + c];
#pragma unroll
    for (int t = 0; t < LMAX; ++t) xc[K - 1 + t] = (t < len) ? tn[(static_cast<size_t>(b) * L + t) * D + c] : 0.0f;
#pragma unroll
    for (int t = 0; t < LMAX; ++t) {
        if (t < len) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < K; ++j) acc += wk[j] * xc[t + j];
            const size_t o = (static_cast<size_t>(b) * L + t) * D + c;
            x[o] += acc + xc[K - 1 + t];
        }
    }
    float* cout = cache_out + static_cast<size_t>(b) * out_stride + out_layer_off;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {
        float v = 0.0f;
#pragma unroll
        for (int t = 0; t < K - 1 + LMAX; ++t)
            if (t == L + j) v = xc[t];
        cout[static_cast<size_t>(j) * D + c] = v;
    }
}

__global__ void __launch_bounds__(256)
pf_online_scatter(const int* __restrict__ tab, const float* __restrict__ src, float* __restrict__ state, size_t stride) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int slot = tab[b * 4 + 0];
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < stride; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        state[static_cast<size_t>(slot) * stride + i] = src[static_cast<size_t>(b) * stride + i];
}

}  // namespace

void online_assemble_launch(const OnlineDims& od, const int* tab, int n, const float* fifo, const float* splice,
                            const float* cache_feats, const float* shift, const float* rescale, const float* inv_ts, float* feats,
                            float* fresh, cudaStream_t s) {
    launch_k(pf_online_assemble, dim3(n, od.t_new), dim3(256), 0, s, od, tab, fifo, splice, cache_feats, shift, rescale, inv_ts, feats, fresh);
}

void online_commit_launch(const OnlineDims& od, const int* tab, int n, const float* fifo, const float* fresh, float* splice,
                          float* cache_feats, cudaStream_t s) {
    launch_k(pf_online_commit, dim3(n), dim3(256), 0, s, od, tab, fifo, fresh, splice, cache_feats);
}

void online_cif_launch(const int* tab, int n, const float* enc, const float* alphas, int ld_alpha, int T, int D, int mask_lo,
                       int mask_hi, float threshold, float* carry_alpha, float* carry_hidden, float* frames_out, int lcap,
                       int* counts, int* meta, cudaStream_t s) {
    if (D > 512) throw CudaError{"online cif: d_model > 512 unsupported"};
    launch_k(pf_online_cif, dim3(n), dim3(512), 0, s, tab, enc, alphas, ld_alpha, T, D, mask_lo, mask_hi, threshold, carry_alpha,
             carry_hidden, frames_out, lcap, counts, meta);
}

void online_compact_launch(const float* frames, int lcap, const int* counts, int n, int L, int D, float* out, cudaStream_t s) {
    launch_k(pf_online_compact, dim3(L, n), dim3(256), 0, s, frames, lcap, counts, L, D, out);
}

void online_fsmn_launch(const int* tab, int n, const float* tn, const int* counts, int L, int D, const float* w, int K,
                        const float* cache_state, size_t state_stride, size_t layer_off, float* x, float* cache_out,
                        size_t out_stride, size_t out_layer_off, cudaStream_t s) {
    if (L > 32) throw CudaError{"online fsmn: more than 32 tokens in one chunk"};
    dim3 grid(ceil_div(D, 256), n);
    if (K == 11) launch_k(pf_online_fsmn<11>, grid, dim3(256), 0, s, tab, tn, counts, L, D, w, cache_state, state_stride, layer_off, x, cache_out, out_stride, out_layer_off);
    else if (K == 21) launch_k(pf_online_fsmn<21>, grid, dim3(256), 0, s, tab, tn, counts, L, D, w, cache_state, state_stride, layer_off, x, cache_out, out_stride, out_layer_off);
    else throw CudaError{"online fsmn: unsupported kernel size"};
}

void online_scatter_launch(const int* tab, int n, const float* src, float* state, size_t stride, cudaStream_t s) {
    launch_k(pf_online_scatter, dim3(16, n), dim3(256), 0, s, tab, src, state, stride);
}

}  // namespace pf
