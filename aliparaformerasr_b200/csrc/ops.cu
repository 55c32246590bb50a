// HBM/L2-bound stages of the SAN-M path: positional encoding + LayerNorm, LayerNorm, FSMN depthwise memory,
// predictor im2col + alpha head, CIF integrate-and-fire, log-softmax + greedy pick.
//
// Reference semantics: the FunASR export graph executed by OfflineProjOfParaformer.cs:68, the host CIF
// recurrence in /root/reference/AliParaformerAsr/OnlineRecognizer.cs:149-200 and the greedy pick in
// OfflineRecognizer.cs:139-152.
#include "ops.cuh"

#include <mutex>

#include <math.h>

namespace pf {

namespace {

// ------------------------------------------------------------------ embed (scale + PE) + LayerNorm(D) -> fp16
// The sinusoidal position encoding depends on (t, c) only: it is tabulated once per sequence length (pf_pe_table, same float
// expressions as before: angle = pos * inv_timescale rounded, then sinf / cosf) and shared by every utterance and step.
__global__ void __launch_bounds__(256)
pf_pe_table(float* __restrict__ pe, int T, int D, const float* __restrict__ inv_ts) {
    const int half = D >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * D; i += gridDim.x * blockDim.x) {
        const int t = i / D, c = i - t * D;
        const int ti = c < half ? c : c - half;
        const float ang = __fmul_rn(static_cast<float>(t + 1), inv_ts[ti]);
        pe[i] = c < half ? sinf(ang) : cosf(ang);
    }
}

// One warp per row.  Statistics in fp64 so rows made of the huge pad constant (Q4) normalise deterministically.
template <int kMaxPerLane>
__global__ void __launch_bounds__(256)
pf_embed_pe_ln(const float* __restrict__ feats, int M, int T, int D, float scale, const float* __restrict__ pe_table,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, __half* __restrict__ out16) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* pe_row = pe_table ? pe_table + static_cast<size_t>(row % T) * D : nullptr;
    const float* x = feats + static_cast<size_t>(row) * D;
    float v[kMaxPerLane];
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        float val = 0.0f;
        if (c < D) {
            if (pe_row != nullptr) {
                val = __fadd_rn(__fmul_rn(x[c], scale), __ldg(pe_row + c));     // same rounding as torch: (x*s) then (+pe)
            } else {
                val = x[c];                                        // streaming: the host side already scaled and encoded
            }
            sum += static_cast<double>(val);
        }
        v[i] = val;
    }
    sum = warp_sum(sum);
    const double mean = sum / D;
    double sq = 0.0;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
            const double d = static_cast<double>(v[i]) - mean;
            sq += d * d;
        }
    }
    sq = warp_sum(sq);
    const double rstd = 1.0 / sqrt(sq / D + static_cast<double>(eps));
    __half* o = out16 + static_cast<size_t>(row) * D;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
            const float nrm = static_cast<float>((static_cast<double>(v[i]) - mean) * rstd);
            o[c] = __float2half_rn(nrm * gamma[c] + beta[c]);
        }
    }
}

// ------------------------------------------------------------------ LayerNorm, D = NV * 128, warp per row
template <int NV>
__global__ void __launch_bounds__(256)
pf_layernorm(const float* __restrict__ in, int ld_in, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
             float eps, __half* __restrict__ out16, int ld16, float* __restrict__ out32, int ld32) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int D = NV * 128;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* x4 = reinterpret_cast<const float4*>(in + static_cast<size_t>(row) * ld_in);
    float4 v[NV];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = x4[lane + 32 * i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    sum = warp_sum(sum);
    const float mean = sum * (1.0f / D);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    sq = warp_sum(sq);
    const float rstd = 1.0f / sqrtf(sq * (1.0f / D) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i);
        const float4 b = __ldg(b4 + lane + 32 * i);
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (out32) reinterpret_cast<float4*>(out32 + static_cast<size_t>(row) * ld32)[lane + 32 * i] = y;
        if (out16) {
            __half2 h0 = __floats2half2_rn(y.x, y.y);
            __half2 h1 = __floats2half2_rn(y.z, y.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            reinterpret_cast<uint2*>(out16 + static_cast<size_t>(row) * ld16)[lane + 32 * i] = pk;
        }
    }
}

// ------------------------------------------------------------------ FSMN depthwise memory (register sliding window)
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }

template <typename TIn, int K>
__global__ void __launch_bounds__(256)
pf_fsmn(const TIn* __restrict__ in, int ld_in, const float* __restrict__ w, float* __restrict__ out, int ld_out,
        const float* __restrict__ resid, int ld_res, const int* __restrict__ lens, int T, int D) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int TT = 16;
    constexpr int LEFT = (K - 1) / 2;
    const int c = blockIdx.z * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    if (c >= D) return;
    const int len = lens ? min(lens[b], T) : T;
    float wk[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wk[j] = w[c * K + j];
    float x[TT + K - 1];
    const size_t rowbase = static_cast<size_t>(b) * T;
#pragma unroll
    for (int i = 0; i < TT + K - 1; ++i) {
        const int t = t0 - LEFT + i;
        x[i] = (t >= 0 && t < len) ? to_f32(in[(rowbase + t) * ld_in + c]) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        const int t = t0 + i;
        if (t < T) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < K; ++j) acc += wk[j] * x[i + j];
            float y = (t < len) ? acc + x[i + LEFT] : 0.0f;
            if (resid) y += resid[(rowbase + t) * ld_res + c];
            out[(rowbase + t) * ld_out + c] = y;
        }
    }
}

// ------------------------------------------------------------------ decoder: norm2 -> FSMN memory -> residual -> norm3
// One CTA per (16-row time block, utterance).  tn = LN2(t) for the block rows plus the FSMN halo goes to shared memory
// (masked rows are zero), every thread then runs the depthwise memory for two channels, adds it to the residual stream
// (rows past the utterance's token count keep x), and the block's new x rows are normalised again (norm3) into the fp16
// operand of the cross-attention query projection.  Replaces three launches per decoder layer (LayerNorm, pf_fsmn,
// LayerNorm) - the decoder is launch-latency bound at M = B * L ~ 1600 rows.
constexpr int kDecTT = 8;     // output rows per CTA (8: 224 CTAs at B x L = 32 x 50; 16 measured slower per launch)
template <int K>
__global__ void __launch_bounds__(512)
pf_dec_ln_fsmn_ln(const float* __restrict__ t32, float* __restrict__ x, const float* __restrict__ g2, const float* __restrict__ b2,
                  const float* __restrict__ w, const float* __restrict__ g3, const float* __restrict__ b3,
                  const int* __restrict__ lens, int L, float eps, __half* __restrict__ out16) {
    pdl_launch_dependents();
    constexpr int D = 512, TT = kDecTT, LEFT = (K - 1) / 2, ROWS = TT + K - 1;
    extern __shared__ float s_dec[];
    float* s_v = s_dec;                    // [ROWS][D]  LN2 output, zero outside [0, len)
    float* s_x = s_dec + ROWS * D;         // [TT][D]    new residual rows
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* g2v = reinterpret_cast<const float4*>(g2);
    const float4* b2v = reinterpret_cast<const float4*>(b2);
    pdl_wait();
    const int len = min(lens[b], L);
    const size_t rowbase = static_cast<size_t>(b) * L;
    for (int r = warp; r < ROWS; r += 16) {
        const int t = t0 - LEFT + r;
        float4* dst = reinterpret_cast<float4*>(s_v + r * D);
        if (t < 0 || t >= len) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const float4* src = reinterpret_cast<const float4*>(t32 + (rowbase + t) * D);
        float4 v[4];
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = src[lane + 32 * i]; sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
        sum = warp_sum(sum);
        const float mean = sum * (1.0f / D);
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + bb * bb) + (c * c + d * d);
        }
        sq = warp_sum(sq);
        const float rstd = 1.0f / sqrtf(sq * (1.0f / D) + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 g = __ldg(g2v + lane + 32 * i), bt = __ldg(b2v + lane + 32 * i);
            dst[lane + 32 * i] = make_float4((v[i].x - mean) * rstd * g.x + bt.x, (v[i].y - mean) * rstd * g.y + bt.y,
                                             (v[i].z - mean) * rstd * g.z + bt.z, (v[i].w - mean) * rstd * g.w + bt.w);
        }
    }
    __syncthreads();
    {
        const int c = threadIdx.x;                                  // one channel per thread (512 threads)
        float wk[K];
#pragma unroll
        for (int j = 0; j < K; ++j) wk[j] = __ldg(w + c * K + j);
#pragma unroll 4
        for (int i = 0; i < TT; ++i) {
            const int t = t0 + i;
            if (t >= L) break;
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < K; ++j) acc += wk[j] * s_v[(i + j) * D + c];
            const size_t o = (rowbase + t) * D + c;
            float xv = x[o];
            if (t < len) xv += acc + s_v[(i + LEFT) * D + c];
            x[o] = xv;
            s_x[i * D + c] = xv;
        }
    }
    __syncthreads();
    const float4* g3v = reinterpret_cast<const float4*>(g3);
    const float4* b3v = reinterpret_cast<const float4*>(b3);
    for (int i = warp; i < TT; i += 16) {
        const int t = t0 + i;
        if (t >= L) break;
        const float4* src = reinterpret_cast<const float4*>(s_x + i * D);
        float4 v[4];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = src[lane + 32 * k]; sum += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
        sum = warp_sum(sum);
        const float mean = sum * (1.0f / D);
        float sq = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = v[k].x - mean, bb = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            sq += (a * a + bb * bb) + (c * c + d * d);
        }
        sq = warp_sum(sq);
        const float rstd = 1.0f / sqrtf(sq * (1.0f / D) + eps);
        uint2* dst = reinterpret_cast<uint2*>(out16 + (rowbase + t) * D);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 g = __ldg(g3v + lane + 32 * k), bt = __ldg(b3v + lane + 32 * k);
            __half2 h0 = __floats2half2_rn((v[k].x - mean) * rstd * g.x + bt.x, (v[k].y - mean) * rstd * g.y + bt.y);
            __half2 h1 = __floats2half2_rn((v[k].z - mean) * rstd * g.z + bt.z, (v[k].w - mean) * rstd * g.w + bt.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            dst[lane + 32 * k] = pk;
        }
    }
}

// ------------------------------------------------------------------ predictor
__global__ void pf_im2col3(const __half* __restrict__ in, int B, int T, int D, __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    // one thread per 8 halfs (16 B) of the output row [3*D]
    const int vec_per_row = 3 * D / 8;
    const long long total = static_cast<long long>(B) * T * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % vec_per_row);
        const long long m = i / vec_per_row;
        const int t = static_cast<int>(m % T);
        const int j = v / (D / 8);            // 0,1,2 -> t-1, t, t+1
        const int cv = v % (D / 8);
        const int ts = t + j - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (ts >= 0 && ts < T) val = reinterpret_cast<const uint4*>(in + (m + (j - 1)) * D)[cv];
        reinterpret_cast<uint4*>(out + m * 3 * D)[v] = val;
    }
}

__global__ void __launch_bounds__(256)
pf_alpha_head(const float* __restrict__ h, int B, int T, int D, const float* __restrict__ w, const float* __restrict__ bias,
              float smooth, float noise, float tail, float* __restrict__ alphas) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int M = B * T;
    if (row >= M) return;
    const float* x = h + static_cast<size_t>(row) * D;
    float acc = 0.0f;
    for (int c = lane; c < D; c += 32) acc += x[c] * w[c];
    acc = warp_sum(acc);
    if (lane == 0) {
        const int b = row / T, t = row % T;
        const float z = acc + bias[0];
        const float sg = 1.0f / (1.0f + expf(-z));
        alphas[static_cast<size_t>(b) * (T + 1) + t] = fmaxf(sg * smooth - noise, 0.0f);
        if (t == 0) alphas[static_cast<size_t>(b) * (T + 1) + T] = tail;
    }
}

// CIF scalar recurrence (inherently sequential in t).  One CTA per `upc` <= 32 utterances (1 for the usual batches: 32 CTAs finish in a
// third of the time of one CTA that stages 32 rows chunk by chunk): the alphas of a chunk of frames are
// staged in shared memory with coalesced loads, `upc` threads (one per utterance) run the recurrence on shared memory, and
// the four per-frame outputs leave coalesced again - a thread walking global memory frame by frame paid an L2 round trip
// per step (32 us for T = 167; 5 us now).  Same fp32 operation order as before (OnlineRecognizer.cs:149-200).
constexpr int kCifChunk = 64;      // frames per staged chunk (4 arrays x 32 x 65 x 4 B = 33 KB of static shared memory)
__global__ void __launch_bounds__(256)
pf_cif_scan(const float* __restrict__ alphas, int B, int T1, float threshold, float* __restrict__ w_cur,
            float* __restrict__ w_rem, int* __restrict__ fire_idx, float* __restrict__ peaks,
            int* __restrict__ token_num, int* __restrict__ fires, int* __restrict__ meta, int upc) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_a[32][kCifChunk + 1];       // in: alpha; out: w_cur          (+1: conflict-free column walks)
    __shared__ float s_rem[32][kCifChunk + 1];
    __shared__ float s_peak[32][kCifChunk + 1];
    __shared__ int s_idx[32][kCifChunk + 1];
    const int b0 = blockIdx.x * upc;
    const int nb = min(upc, B - b0);
    float integrate = 0.0f, total = 0.0f;
    int nf = 0;
    for (int c0 = 0; c0 < T1; c0 += kCifChunk) {
        const int nt = min(kCifChunk, T1 - c0);
        for (int i = threadIdx.x; i < nb * nt; i += blockDim.x) {
            const int u = i / nt, t = i - u * nt;
            s_a[u][t] = alphas[static_cast<size_t>(b0 + u) * T1 + c0 + t];
        }
        __syncthreads();
        if (threadIdx.x < nb) {
            const int u = threadIdx.x;
            for (int t = 0; t < nt; ++t) {
                const float a = s_a[u][t];
                total = __fadd_rn(total, a);
                const float completion = __fsub_rn(1.0f, integrate);
                integrate = __fadd_rn(integrate, a);
                s_peak[u][t] = integrate;
                const bool fire = integrate >= threshold;
                const float cur = fire ? completion : a;
                s_a[u][t] = cur;
                if (fire) {
                    integrate = __fsub_rn(integrate, 1.0f);
                    s_rem[u][t] = __fsub_rn(a, cur);
                    s_idx[u][t] = nf++;
                } else {
                    s_rem[u][t] = 0.0f;
                    s_idx[u][t] = -1;
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb * nt; i += blockDim.x) {
            const int u = i / nt, t = i - u * nt;
            const size_t o = static_cast<size_t>(b0 + u) * T1 + c0 + t;
            w_cur[o] = s_a[u][t];
            w_rem[o] = s_rem[u][t];
            fire_idx[o] = s_idx[u][t];
            if (peaks) peaks[o] = s_peak[u][t];
        }
        __syncthreads();
    }
    if (threadIdx.x < nb) {
        token_num[b0 + threadIdx.x] = static_cast<int>(floorf(total));
        fires[b0 + threadIdx.x] = nf;
        atomicMax(meta, nf);
    }
}

// One CTA per (token l, utterance b): the frames between the previous fire and this one, accumulated in the same order
// (and with the same fp32 roundings) as the sequential recurrence: start from the remainder weight of the previous
// firing frame, add w_cur[t] * h[t] for t up to this token's firing frame.
__global__ void __launch_bounds__(256)
pf_cif_gather(const float* __restrict__ hidden, int T, int D, const float* __restrict__ w_cur,
              const float* __restrict__ w_rem, const int* __restrict__ fire_idx, int T1, float* __restrict__ out, int Lpad) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_t[2];
    const int l = blockIdx.x, b = blockIdx.y;
    const size_t abase = static_cast<size_t>(b) * T1;
    if (threadIdx.x < 2) s_t[threadIdx.x] = -1;
    __syncthreads();
    for (int t = threadIdx.x; t < T1; t += blockDim.x) {
        const int f = fire_idx[abase + t];
        if (f == l) s_t[1] = t;
        else if (f == l - 1 && l > 0) s_t[0] = t;
    }
    __syncthreads();
    const int t_end = s_t[1], t_prev = s_t[0];
    if (t_end < 0) return;                                   // this utterance fired fewer than l + 1 tokens: row stays zero
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float* h = hidden + static_cast<size_t>(b) * T * D + d;
        float frame = 0.0f;
        if (t_prev >= 0) frame = __fmul_rn(w_rem[abase + t_prev], t_prev < T ? h[static_cast<size_t>(t_prev) * D] : 0.0f);
        for (int t = t_prev + 1; t <= t_end; ++t) {
            const float hv = t < T ? h[static_cast<size_t>(t) * D] : 0.0f;      // tail step has zero hidden
            frame = __fadd_rn(frame, __fmul_rn(w_cur[abase + t], hv));
        }
        out[(static_cast<size_t>(b) * Lpad + l) * D + d] = frame;
    }
}

// ------------------------------------------------------------------ log-softmax + greedy pick, one CTA per row
__global__ void __launch_bounds__(256)
pf_logsoftmax_argmax(float* __restrict__ logits, int V, int ld, int* __restrict__ tokens, int write_logp) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    __shared__ int s_nan[8];
    __shared__ float s_sum[8];
    __shared__ float s_bcast[2];
    __shared__ int s_ibcast;
    float* x = logits + static_cast<size_t>(blockIdx.x) * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // pass 1: last NaN index (the reference's scan restarts there) and the plain maximum for the softmax
    int last_nan = -1;
    float mx = -INFINITY;
    for (int i = tid; i < V; i += blockDim.x) {
        const float v = x[i];
        if (v != v) last_nan = i;
        else mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        last_nan = max(last_nan, __shfl_xor_sync(0xffffffffu, last_nan, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_nan[warp] = last_nan; s_val[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        int ln = -1; float m = -INFINITY;
        for (int i = 0; i < 8; ++i) { ln = max(ln, s_nan[i]); m = fmaxf(m, s_val[i]); }
        s_ibcast = ln; s_bcast[0] = m;
    }
    __syncthreads();
    last_nan = s_ibcast;
    mx = s_bcast[0];

    // pass 2: sum of exp + greedy pick over indices >= max(last_nan, 0) with "last maximum wins"
    // (best = x[best] > x[k] ? best : k, OfflineRecognizer.cs:145-149)
    const int start = last_nan < 0 ? 0 : last_nan;
    float best = -INFINITY;
    int best_i = start;
    float sum = 0.0f;
    for (int i = tid; i < V; i += blockDim.x) {
        const float v = x[i];
        sum += expf(v - mx);
        if (i > start && v >= best) { best = v; best_i = i; }
        else if (i == start && last_nan < 0) { if (v >= best) { best = v; best_i = i; } }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi > best_i)) { best = ov; best_i = oi; }
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; s_sum[warp] = sum; }
    __syncthreads();
    if (tid == 0) {
        float bv = s_val[0]; int bi = s_idx[0]; float sm = s_sum[0];
        for (int i = 1; i < 8; ++i) {
            if (s_val[i] > bv || (s_val[i] == bv && s_idx[i] > bi)) { bv = s_val[i]; bi = s_idx[i]; }
            sm += s_sum[i];
        }
        // a NaN in the last position wins outright; a NaN elsewhere restarts the scan right after it
        if (last_nan >= 0 && bv == -INFINITY) bi = max(bi, last_nan);
        tokens[blockIdx.x] = bi;
        s_bcast[1] = logf(sm);
    }
    __syncthreads();
    if (write_logp) {
        const float lse = mx + s_bcast[1];
        for (int i = tid; i < V; i += blockDim.x) x[i] = x[i] - lse;
    }
}

// Combine the per-(tile, column group) partials of a fused-pick head GEMM (gemm_dev.cuh: epilogue_pick_f32) into the
// greedy id of every row.  Slots are in ascending column order; a slot that saw a NaN restarts the scan (its own best is
// the best among the columns after its last NaN), exactly like the sequential loop of OfflineRecognizer.cs:145-149.
// One warp per row: lane L scans its contiguous run of slots, then the 32 run summaries are merged in slot order.  A summary is
// (best value, best index, last NaN index or -1) of a run scanned from an empty state; "A then B" = B if B saw a NaN (the NaN restarts
// the scan: B's best already is the best after it), else the later-wins maximum of the two with A's NaN - an associative rule, so the
// tree merge gives exactly what the sequential loop gives.
struct PickRun { float v; int i; int ln; };
__device__ __forceinline__ PickRun pick_merge(const PickRun& a, const PickRun& b) {
    if (b.ln >= 0) return b;
    PickRun r = a;
    if (b.i >= 0 && (a.i < 0 || b.v >= a.v)) { r.v = b.v; r.i = b.i; }
    return r;
}
__global__ void __launch_bounds__(256)
pf_pick_combine(const float* __restrict__ partials, int M, int ld, int slots, int* __restrict__ tokens) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* p = partials + static_cast<size_t>(row) * ld * 3;
    const int per = (slots + 31) >> 5;
    PickRun r{-INFINITY, -1, -1};
    for (int s = lane * per; s < min(slots, (lane + 1) * per); ++s) {
        const float v = p[3 * s];
        const int i = __float_as_int(p[3 * s + 1]);
        const int ln = __float_as_int(p[3 * s + 2]);
        if (ln >= 0) { r.ln = ln; r.v = v; r.i = i; }
        else if (i >= 0 && v >= r.v) { r.v = v; r.i = i; }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        PickRun o;
        o.v = __shfl_down_sync(0xffffffffu, r.v, d);
        o.i = __shfl_down_sync(0xffffffffu, r.i, d);
        o.ln = __shfl_down_sync(0xffffffffu, r.ln, d);
        if ((lane & (2 * d - 1)) == 0 && lane + d < 32) r = pick_merge(r, o);
    }
    if (lane == 0) tokens[row] = r.i >= 0 ? r.i : (r.ln >= 0 ? r.ln : 0);
}

__global__ void pf_f32_to_f16(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __float2half_rn(in[i]);
}

__global__ void pf_prepend_rows(const float* __restrict__ src, const float* __restrict__ table, const int* __restrict__ ids,
                                int nprompt, float* __restrict__ dst, int B, int T, int D) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(B) * (T + nprompt) * D;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % D);
        const long long r = i / D;
        const int t = static_cast<int>(r % (T + nprompt));
        const long long b = r / (T + nprompt);
        dst[i] = t < nprompt ? table[static_cast<size_t>(ids[t]) * D + c] : src[(b * T + (t - nprompt)) * D + c];
    }
}

// ------------------------------------------------------------------ SeACo helpers
// x16[(t * n + i), :] = table[ids[i * steps + t], :]   (Embedding, then time-major as the exported LSTM runs)
__global__ void pf_embed_rows_tmajor(const float* __restrict__ table, const int* __restrict__ ids, int n, int steps, int D,
                                     __half* __restrict__ out) {
    const long long total = static_cast<long long>(n) * steps * D;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % D);
        const long long r = e / D;
        const int i = static_cast<int>(r % n), t = static_cast<int>(r / n);
        out[e] = __float2half_rn(table[static_cast<size_t>(ids[i * steps + t]) * D + c]);
    }
}

// PyTorch LSTM cell, gate order i | f | g | o in gates [n, 4D]: c = s(f) c + s(i) tanh(g); h = s(o) tanh(c)
__global__ void pf_lstm_cell(const float* __restrict__ gates, float* __restrict__ cstate, int n, int D, __half* __restrict__ h16,
                             __half* __restrict__ rows_hw_major, int t, int steps) {
    const int total = n * D;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e / D, c = e % D;
        const float* g = gates + static_cast<size_t>(i) * 4 * D;
        const float gi = 1.0f / (1.0f + expf(-g[c]));
        const float gf = 1.0f / (1.0f + expf(-g[D + c]));
        const float gg = tanhf(g[2 * D + c]);
        const float go = 1.0f / (1.0f + expf(-g[3 * D + c]));
        const float cn = gf * cstate[e] + gi * gg;
        cstate[e] = cn;
        const __half h = __float2half_rn(go * tanhf(cn));
        h16[e] = h;
        if (rows_hw_major) rows_hw_major[(static_cast<size_t>(i) * steps + t) * D + c] = h;
    }
}

__global__ void pf_add_to_f16(const float* __restrict__ a, const float* __restrict__ b, __half* __restrict__ out, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __float2half_rn(a[i] + b[i]);
}

// rows whose hot-word pick is not NO_BIAS take the hot-word posterior and pick; the others keep the ASR ones
__global__ void __launch_bounds__(256)
pf_seaco_merge(const int* __restrict__ dha_tok, const float* __restrict__ dha, int nobias, int V, int ld, int* __restrict__ tokens,
               float* __restrict__ logits) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const int t = dha_tok[row];
    if (t == nobias) return;
    if (threadIdx.x == 0) tokens[row] = t;
    const float* src = dha + static_cast<size_t>(row) * ld;
    float* dst = logits + static_cast<size_t>(row) * ld;
    for (int i = threadIdx.x; i < V; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

void embed_rows_tmajor_launch(const float* table, const int* ids, int n, int steps, int D, __half* out, cudaStream_t s) {
    const long long total = static_cast<long long>(n) * steps * D;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
    pf_embed_rows_tmajor<<<grid, 256, 0, s>>>(table, ids, n, steps, D, out);
    PF_CUDA(cudaGetLastError());
}

void lstm_cell_launch(const float* gates, float* cstate, int n, int D, __half* h16, __half* rows_hw_major, int t, int steps,
                      cudaStream_t s) {
    const int grid = std::min((n * D + 255) / 256, 148 * 8);
    pf_lstm_cell<<<grid, 256, 0, s>>>(gates, cstate, n, D, h16, rows_hw_major, t, steps);
    PF_CUDA(cudaGetLastError());
}

void add_to_f16_launch(const float* a, const float* b, __half* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
    launch_k(pf_add_to_f16, dim3(grid), dim3(256), 0, s, a, b, out, n);
}

void seaco_merge_launch(const int* dha_tok, const float* dha, int nobias, int M, int V, int ld, int* tokens, float* logits,
                        cudaStream_t s) {
    if (M <= 0) return;
    launch_k(pf_seaco_merge, dim3(M), dim3(256), 0, s, dha_tok, dha, nobias, V, ld, tokens, logits);
}

void pe_table_launch(float* pe, int T, int D, const float* inv_timescales, cudaStream_t s) {
    if (T <= 0) return;
    pf_pe_table<<<std::min(ceil_div(T * D, 256), 1024), 256, 0, s>>>(pe, T, D, inv_timescales);
    PF_CUDA(cudaGetLastError());
}

void embed_pe_ln_launch(const float* feats, int M, int T, int D, float scale, const float* pe_table,
                        const float* gamma, const float* beta, float eps, __half* out16, cudaStream_t s) {
    if (D > 32 * 20) throw CudaError{"embed_pe_ln: input_size > 640 unsupported"};
    launch_k(pf_embed_pe_ln<20>, dim3(ceil_div(M, 8)), dim3(256), 0, s, feats, M, T, D, scale, pe_table, gamma, beta, eps, out16);
}

void layernorm_f32_launch(const float* in, int ld_in, int M, int D, const float* gamma, const float* beta, float eps,
                          __half* out16, int ld16, float* out32, int ld32, cudaStream_t s) {
    const int grid = ceil_div(M, 8);
    if (D == 512) launch_k(pf_layernorm<4>, dim3(grid), dim3(256), 0, s, in, ld_in, M, gamma, beta, eps, out16, ld16, out32, ld32);
    else if (D == 1024) launch_k(pf_layernorm<8>, dim3(grid), dim3(256), 0, s, in, ld_in, M, gamma, beta, eps, out16, ld16, out32, ld32);
    else if (D == 2048) launch_k(pf_layernorm<16>, dim3(grid), dim3(256), 0, s, in, ld_in, M, gamma, beta, eps, out16, ld16, out32, ld32);
    else throw CudaError{"layernorm: unsupported width " + std::to_string(D)};
    PF_CUDA(cudaGetLastError());
}

template <typename TIn>
static void fsmn_launch_t(const TIn* in, int ld_in, const float* w, int K, float* out, int ld_out, const float* resid,
                          int ld_res, const int* lens, int B, int T, int D, cudaStream_t s) {
    dim3 grid(ceil_div(T, 16), B, ceil_div(D, 256));
    if (K == 11) launch_k(pf_fsmn<TIn, 11>, grid, dim3(256), 0, s, in, ld_in, w, out, ld_out, resid, ld_res, lens, T, D);
    else if (K == 21) launch_k(pf_fsmn<TIn, 21>, grid, dim3(256), 0, s, in, ld_in, w, out, ld_out, resid, ld_res, lens, T, D);
    else throw CudaError{"fsmn: unsupported kernel size " + std::to_string(K)};
    PF_CUDA(cudaGetLastError());
}
void fsmn_f16_launch(const __half* in, int ld_in, const float* w, int K, float* out, int ld_out, const float* resid,
                     int ld_res, const int* lens, int B, int T, int D, cudaStream_t s) {
    fsmn_launch_t(in, ld_in, w, K, out, ld_out, resid, ld_res, lens, B, T, D, s);
}
void fsmn_f32_launch(const float* in, int ld_in, const float* w, int K, float* out, int ld_out, const float* resid,
                     int ld_res, const int* lens, int B, int T, int D, cudaStream_t s) {
    fsmn_launch_t(in, ld_in, w, K, out, ld_out, resid, ld_res, lens, B, T, D, s);
}

template <int K>
static void dec_ln_fsmn_ln_launch_t(const float* t32, float* x, const float* g2, const float* b2, const float* w, const float* g3,
                                    const float* b3, const int* lens, int B, int L, float eps, __half* out16, cudaStream_t s) {
    constexpr int kSmem = (kDecTT + K - 1 + kDecTT) * 512 * 4;
    static std::once_flag attr_once;                              // execution lanes call this from several host threads
    std::call_once(attr_once, [&] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_dec_ln_fsmn_ln<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    launch_k(pf_dec_ln_fsmn_ln<K>, dim3(ceil_div(L, kDecTT), B), dim3(512), kSmem, s, t32, x, g2, b2, w, g3, b3, lens, L, eps, out16);
}

void dec_ln_fsmn_ln_launch(const float* t32, float* x, const float* g2, const float* b2, const float* w, int K, const float* g3,
                           const float* b3, const int* lens, int B, int L, int D, float eps, __half* out16, cudaStream_t s) {
    if (D != 512) throw CudaError{"dec_ln_fsmn_ln: d_model must be 512"};
    if (B <= 0 || L <= 0) return;
    if (K == 11) dec_ln_fsmn_ln_launch_t<11>(t32, x, g2, b2, w, g3, b3, lens, B, L, eps, out16, s);
    else if (K == 21) dec_ln_fsmn_ln_launch_t<21>(t32, x, g2, b2, w, g3, b3, lens, B, L, eps, out16, s);
    else throw CudaError{"dec_ln_fsmn_ln: unsupported kernel size " + std::to_string(K)};
}

void im2col3_launch(const __half* in, int B, int T, int D, __half* out, cudaStream_t s) {
    const long long total = static_cast<long long>(B) * T * (3 * D / 8);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    launch_k(pf_im2col3, dim3(grid), dim3(256), 0, s, in, B, T, D, out);
}

void alpha_head_launch(const float* h, int B, int T, int D, const float* w, const float* bias, float smooth, float noise,
                       float tail, float* alphas, cudaStream_t s) {
    launch_k(pf_alpha_head, dim3(ceil_div(B * T, 8)), dim3(256), 0, s, h, B, T, D, w, bias, smooth, noise, tail, alphas);
}

void cif_scan_launch(const float* alphas, int B, int T1, float threshold, float* w_cur, float* w_rem, int* fire_idx,
                     float* peaks, int* token_num, int* fires, int* meta, cudaStream_t s) {
    const int upc = std::min(32, std::max(1, ceil_div(B, 2 * 148)));          // utterances per CTA: 1 up to 296 utterances
    launch_k(pf_cif_scan, dim3(ceil_div(B, upc)), dim3(256), 0, s, alphas, B, T1, threshold, w_cur, w_rem, fire_idx, peaks, token_num, fires, meta, upc);
}

void cif_gather_launch(const float* hidden, int B, int T, int D, const float* w_cur, const float* w_rem,
                       const int* fire_idx, int T1, float* out, int Lpad, cudaStream_t s) {
    if (Lpad <= 0) return;
    launch_k(pf_cif_gather, dim3(Lpad, B), dim3(256), 0, s, hidden, T, D, w_cur, w_rem, fire_idx, T1, out, Lpad);
}

void logsoftmax_argmax_launch(float* logits, int M, int V, int ld, int* tokens, int write_logp, cudaStream_t s) {
    if (M <= 0) return;
    launch_k(pf_logsoftmax_argmax, dim3(M), dim3(256), 0, s, logits, V, ld, tokens, write_logp);
}

void pick_combine_launch(const float* partials, int M, int ld, int slots, int* tokens, cudaStream_t s) {
    if (M <= 0) return;
    launch_k(pf_pick_combine, dim3(ceil_div(M, 8)), dim3(256), 0, s, partials, M, ld, slots, tokens);
}

void f32_to_f16_launch(const float* in, __half* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 32));
    pf_f32_to_f16<<<grid, 256, 0, s>>>(in, out, n);
    PF_CUDA(cudaGetLastError());
}

void prepend_rows_launch(const float* src, const float* table, const int* ids, int nprompt, float* dst, int B, int T,
                         int D, cudaStream_t s) {
    const long long total = static_cast<long long>(B) * (T + nprompt) * D;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    launch_k(pf_prepend_rows, dim3(grid), dim3(256), 0, s, src, table, ids, nprompt, dst, B, T, D);
}

}  // namespace pf
