// CifPredictorV3 timestamp branch: the graph outputs us_alphas / us_cif_peak that the reference reads as
// results[3] (OfflineProjOfParaformer.cs:75-79, OfflineProjOfSeacoParaformer.cs:122-126) and turns into per-token
// timestamps in OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302).  Graph semantics: FunASR
// CifPredictorV3.get_upsample_timestmap [EXT]: ConvTranspose1d x3 (a GEMM here) -> BiLSTM -> Linear(1024,1) -> sigmoid ->
// relu(a*0.25 - 0.01) -> rescale to token_num -> cif_wo_hidden(threshold 1 - 1e-4).
//
// The BiLSTM is a 3T-step (498 at 10 s) sequential recurrence over a [B,512] state: as per-step GEMM launches it would cost
// more than the whole recogniser, so it runs as ONE persistent cooperative kernel - W_hh stays in shared memory, the
// CTAs of a direction synchronise through a global counter once per step.
#include "timestamp.cuh"

#include <math.h>

namespace pf {

namespace {

constexpr int kWPitch = 512 + 8;      // halfs per staged W_hh row (+16 B: rows of one warp hit different banks)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
pf_bilstm_persistent(const float* __restrict__ gin, const __half* __restrict__ w_hh, int B, int T3, float* __restrict__ y,
                     float* hbuf, unsigned int* bar) {
    constexpr int H = 512, U = kLstmUnitsPerCta, R = 4 * U;      // 32 gate rows per CTA
    extern __shared__ __align__(16) uint8_t smem_l[];
    __half* s_w = reinterpret_cast<__half*>(smem_l);                         // [R][kWPitch]
    float* s_h = reinterpret_cast<float*>(smem_l + R * kWPitch * 2);          // [B][H]
    float* s_g = s_h + kLstmMaxBatch * H;                                     // [B][R] gate pre-activations
    const int ctas_per_dir = H / U;
    const int dir = blockIdx.x / ctas_per_dir;
    const int u0 = (blockIdx.x % ctas_per_dir) * U;
    const int tid = threadIdx.x;
    // stage this CTA's rows of W_hh: row r = gate (r / U), unit u0 + r % U
    for (int i = tid; i < R * (H / 8); i += blockDim.x) {
        const int r = i / (H / 8), c8 = i % (H / 8);
        const int grow = (r / U) * H + u0 + (r % U);
        *reinterpret_cast<uint4*>(s_w + r * kWPitch + c8 * 8) =
            *reinterpret_cast<const uint4*>(w_hh + (static_cast<size_t>(dir) * 4 * H + grow) * H + c8 * 8);
    }
    // thread roles: dot products (row r = tid % 32, batch group tid / 32), cell update (b = tid / U, unit tid % U)
    const int r = tid & 31, bg = tid >> 5;
    const int nb_per = (B + 7) / 8;                                          // batches per dot-product thread (<= 4)
    const int cb = tid / U, cu = tid % U;
    float cstate = 0.0f;
    float* hb = hbuf + static_cast<size_t>(dir) * 2 * B * H;
    unsigned int* my_bar = bar + dir;
    __syncthreads();
    for (int step = 0; step < T3; ++step) {
        const int t = dir == 0 ? step : T3 - 1 - step;
        const float* hprev = hb + static_cast<size_t>(step & 1) * B * H;
        float* hnext = hb + static_cast<size_t>((step & 1) ^ 1) * B * H;
        // input projections of this step for my (row, batches): issued first, consumed after the dot products
        float gi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = bg * nb_per + j;
            gi[j] = (j < nb_per && b < B) ? gin[(static_cast<size_t>(b) * T3 + t) * (8 * H) + dir * 4 * H + (r / U) * H + u0 + (r % U)] : 0.0f;
        }
        if (step == 0) {
            for (int i = tid; i < B * H; i += blockDim.x) s_h[i] = 0.0f;
        } else {
            for (int i = tid; i < B * H / 4; i += blockDim.x)
                reinterpret_cast<float4*>(s_h)[i] = __ldcg(reinterpret_cast<const float4*>(hprev) + i);    // L2: written by other SMs
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const __half* wr = s_w + r * kWPitch;
#pragma unroll 4
        for (int k = 0; k < H; k += 8) {
            const uint4 wv = *reinterpret_cast<const uint4*>(wr + k);
            const __half2* w2 = reinterpret_cast<const __half2*>(&wv);
            float wf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(w2[i]); wf[2 * i] = f.x; wf[2 * i + 1] = f.y; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = bg * nb_per + j;
                if (j < nb_per && b < B) {
                    const float4 h0 = *reinterpret_cast<const float4*>(s_h + b * H + k);
                    const float4 h1 = *reinterpret_cast<const float4*>(s_h + b * H + k + 4);
                    acc[j] += wf[0] * h0.x + wf[1] * h0.y + wf[2] * h0.z + wf[3] * h0.w + wf[4] * h1.x + wf[5] * h1.y + wf[6] * h1.z + wf[7] * h1.w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = bg * nb_per + j;
            if (j < nb_per && b < B) s_g[b * R + r] = acc[j] + gi[j];
        }
        __syncthreads();
        if (cb < B) {
            const float* g = s_g + cb * R;
            const float ig = 1.0f / (1.0f + expf(-g[cu]));
            const float fg = 1.0f / (1.0f + expf(-g[U + cu]));
            const float gg = tanhf(g[2 * U + cu]);
            const float og = 1.0f / (1.0f + expf(-g[3 * U + cu]));
            cstate = fg * cstate + ig * gg;
            const float h = og * tanhf(cstate);
            hnext[cb * H + u0 + cu] = h;
            y[(static_cast<size_t>(cb) * T3 + t) * (2 * H) + dir * H + u0 + cu] = h;
        }
        // all CTAs of this direction have published h_t before anyone reads it
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(my_bar, 1u);
            const unsigned int target = static_cast<unsigned int>(ctas_per_dir) * (step + 1);
            unsigned int spins = 0;
            while (ld_acquire_u32(my_bar) < target) {
                if (++spins > (1u << 28)) { printf("pfasr: bilstm grid barrier timeout\n"); __trap(); }
            }
        }
        __syncthreads();
    }
}

// one CTA per utterance: alphas2 -> rescale -> integrate-and-fire trace
__global__ void __launch_bounds__(256)
pf_us_alphas_peaks(const float* __restrict__ y, int T3, int D2, const float* __restrict__ w2, const float* __restrict__ b2,
                   float smooth, float noise, const int* __restrict__ token_num, float thr, float* __restrict__ us_alphas,
                   float* __restrict__ us_peaks) {
    extern __shared__ float s_a[];              // [T3]
    __shared__ float s_red[8];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float part = 0.0f;
    for (int t = warp; t < T3; t += 8) {
        const float* row = y + (static_cast<size_t>(b) * T3 + t) * D2;
        float acc = 0.0f;
        for (int c = lane; c < D2; c += 32) acc += row[c] * w2[c];
        acc = warp_sum(acc);
        const float sg = 1.0f / (1.0f + expf(-(acc + b2[0])));
        const float a = fmaxf(sg * smooth - noise, 0.0f);
        if (lane == 0) { s_a[t] = a; part += a; }
    }
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int i = 0; i < 8; ++i) tot += s_red[i];
        const float ratio = static_cast<float>(token_num[b]) / tot;
        float integrate = 0.0f;
        for (int t = 0; t < T3; ++t) {
            const float a = s_a[t] * ratio;
            us_alphas[static_cast<size_t>(b) * T3 + t] = a;
            integrate = __fadd_rn(integrate, a);
            us_peaks[static_cast<size_t>(b) * T3 + t] = integrate;
            if (integrate >= thr) integrate = __fsub_rn(integrate, thr);
        }
    }
}

}  // namespace

void bilstm_launch(const float* gin, const __half* w_hh, int B, int T3, int H, float* y, float* hbuf, unsigned int* bar, cudaStream_t s) {
    if (H != 512) throw CudaError{"bilstm: hidden size must be 512"};
    if (B < 1 || B > kLstmMaxBatch) throw CudaError{"bilstm: 1..32 utterances per launch"};
    const int smem = 4 * kLstmUnitsPerCta * kWPitch * 2 + kLstmMaxBatch * H * 4 + kLstmMaxBatch * 4 * kLstmUnitsPerCta * 4;
    static bool attr_set = false;
    if (!attr_set) {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_bilstm_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        }
        PF_CUDA(cudaSetDevice(cur));
        attr_set = true;
    }
    PF_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), s));
    const int grid = 2 * (H / kLstmUnitsPerCta);
    void* args[] = {(void*)&gin, (void*)&w_hh, (void*)&B, (void*)&T3, (void*)&y, (void*)&hbuf, (void*)&bar};
    // cooperative launch: the step barrier needs every CTA resident (128 CTAs, one per SM)
    PF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(pf_bilstm_persistent), dim3(grid), dim3(256), args, static_cast<size_t>(smem), s));
}

void us_alphas_peaks_launch(const float* y, int B, int T3, int D2, const float* w2, const float* b2, float smooth, float noise,
                            const int* token_num, float thr, float* us_alphas, float* us_peaks, cudaStream_t s) {
    if (B <= 0 || T3 <= 0) return;
    pf_us_alphas_peaks<<<B, 256, static_cast<size_t>(T3) * sizeof(float), s>>>(y, T3, D2, w2, b2, smooth, noise, token_num, thr, us_alphas, us_peaks);
    PF_CUDA(cudaGetLastError());
}

}  // namespace pf
