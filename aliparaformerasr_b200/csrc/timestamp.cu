// CifPredictorV3 timestamp branch: the graph outputs us_alphas / us_cif_peak that the reference reads as
// results[3] (OfflineProjOfParaformer.cs:75-79, OfflineProjOfSeacoParaformer.cs:122-126) and turns into per-token
// timestamps in OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302).  Graph semantics: FunASR
// CifPredictorV3.get_upsample_timestmap [EXT]: ConvTranspose1d x3 (a GEMM here) -> BiLSTM -> Linear(1024,1) -> sigmoid ->
// relu(a*0.25 - 0.01) -> rescale to token_num -> cif_wo_hidden(threshold 1 - 1e-4).
//
// The BiLSTM is a 3T-step (498 at 10 s) sequential recurrence over a [B,512] state: as per-step GEMM launches it would cost
// more than the whole recogniser, so it runs as ONE persistent kernel - a 16-CTA cluster per direction keeps W_hh in
// shared memory, multiplies with mma.sync, exchanges h_t through distributed shared memory and meets at the hardware
// cluster barrier once per step (a first version with a global-memory barrier and scalar dot products took 16 us / step).
#include "timestamp.cuh"

#include <math.h>

namespace pf {

namespace {

constexpr int kPitch = 512 + 8;       // halfs per staged row of W_hh / h (+16 B: ldmatrix rows hit different banks)
constexpr int kClusterCtas = 16;      // CTAs per direction = one thread-block cluster; each owns 32 hidden units
constexpr int kUnits = 32;            // hidden units per CTA -> 128 gate rows of W_hh resident in shared memory
constexpr int kRows = 4 * kUnits;
constexpr int kMb = 16;               // utterances per launch = M of mma.m16n8k16

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_cluster_u128(uint32_t cluster_addr, const uint4& v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One thread-block cluster (16 CTAs) per direction.  CTA c keeps the 128 rows of W_hh that produce the four gates of
// hidden units [32c, 32c+32) in shared memory for all T3 steps.  Per step: gates = h_{t-1} W_hh^T (mma.sync, M = 16
// utterances, N = 128 gate rows, K = 512) + the precomputed input projections; cell update; the CTA's slice of h_t is
// pushed (fp16) into every peer's shared memory through DSMEM; one hardware cluster barrier.  No global-memory
// synchronisation, no re-reading of weights.
__global__ void __launch_bounds__(256, 1)
pf_bilstm_cluster(const float* __restrict__ gin, const __half* __restrict__ w_hh, int B, int T3, float* __restrict__ y) {
    constexpr int H = 512;
    extern __shared__ __align__(16) uint8_t smem_l[];
    __half* s_w = reinterpret_cast<__half*>(smem_l);                                  // [kRows][kPitch]
    __half* s_h = s_w + kRows * kPitch;                                               // [2][kMb][kPitch]  h_{t-1} (all units), double buffered
    float* s_g = reinterpret_cast<float*>(s_h + 2 * kMb * kPitch);                    // [kMb][kRows] gate pre-activations (recurrent part)
    __half* s_o = reinterpret_cast<__half*>(s_g + kMb * kRows);                       // [kMb][kUnits] this CTA's h_t slice
    const int rank = static_cast<int>(cluster_ctarank());
    const int dir = blockIdx.x / kClusterCtas;
    const int u0 = rank * kUnits;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // stage W_hh rows: local row r = gate (r / 32) * 32 + unit offset (r % 32)
    for (int i = tid; i < kRows * (H / 8); i += blockDim.x) {
        const int r = i / (H / 8), c8 = i % (H / 8);
        const int grow = (r / kUnits) * H + u0 + (r % kUnits);
        *reinterpret_cast<uint4*>(s_w + r * kPitch + c8 * 8) =
            *reinterpret_cast<const uint4*>(w_hh + (static_cast<size_t>(dir) * 4 * H + grow) * H + c8 * 8);
    }
    for (int i = tid; i < 2 * kMb * kPitch / 8; i += blockDim.x) reinterpret_cast<uint4*>(s_h)[i] = make_uint4(0, 0, 0, 0);   // h_0 = 0
    // cell roles: thread -> (utterance cb, units cu, cu + 16)
    const int cb = tid >> 4, cu = tid & 15;
    float cstate[2] = {0.0f, 0.0f};
    cluster_sync();                                                   // every CTA's buffers are initialised before remote writes
    const uint32_t s_w_u32 = smem_u32(s_w), s_h_u32 = smem_u32(s_h);
    // input projections of my two cells (4 gates each) are fetched ONE STEP AHEAD: gin (130 MB at 16 x 10 s) streams from
    // HBM, and a step is far shorter than a DRAM round trip
    auto load_gi = [&](int step_, float (&dst)[2][4]) {
        const int t_ = dir == 0 ? step_ : T3 - 1 - step_;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int g = 0; g < 4; ++g)
                dst[k][g] = (cb < B && step_ < T3) ? __ldg(gin + (static_cast<size_t>(cb) * T3 + t_) * (8 * H) + dir * 4 * H + g * H + u0 + cu + 16 * k) : 0.0f;
    };
    float gi[2][4], gi_next[2][4];
    load_gi(0, gi_next);
    for (int step = 0; step < T3; ++step) {
        const int t = dir == 0 ? step : T3 - 1 - step;
        const int cur = step & 1;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int g = 0; g < 4; ++g) gi[k][g] = gi_next[k][g];
        load_gi(step + 1, gi_next);
        // gates[16 x 128] = h_{t-1}[16 x 512] * W^T: warp w owns gate rows [16w, 16w + 16) = two n-tiles
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const uint32_t a_base = s_h_u32 + static_cast<uint32_t>(((cur * kMb + (lane & 15)) * kPitch + (lane >> 4) * 8) * 2);
        const uint32_t b_base = s_w_u32 + static_cast<uint32_t>(((warp * 16 + (lane >> 4) * 8 + (lane & 7)) * kPitch + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll 8
        for (int k0 = 0; k0 < H; k0 += 16) {
            uint32_t a[4], b0, b1, b2, b3;
            ldsm_x4(a_base + k0 * 2, a[0], a[1], a[2], a[3]);
            ldsm_x4(b_base + k0 * 2, b0, b1, b2, b3);
            mma16816(acc[0], a, b0, b1);
            mma16816(acc[1], a, b2, b3);
        }
        {
            const int g = lane >> 2, c2 = (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int col = warp * 16 + nt * 8 + c2;
                s_g[g * kRows + col] = acc[nt][0];
                s_g[g * kRows + col + 1] = acc[nt][1];
                s_g[(g + 8) * kRows + col] = acc[nt][2];
                s_g[(g + 8) * kRows + col + 1] = acc[nt][3];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int u = cu + 16 * k;
            const float* g = s_g + cb * kRows;
            const float ig = 1.0f / (1.0f + expf(-(g[u] + gi[k][0])));
            const float fg = 1.0f / (1.0f + expf(-(g[kUnits + u] + gi[k][1])));
            const float gg = tanhf(g[2 * kUnits + u] + gi[k][2]);
            const float og = 1.0f / (1.0f + expf(-(g[3 * kUnits + u] + gi[k][3])));
            cstate[k] = fg * cstate[k] + ig * gg;
            const float h = og * tanhf(cstate[k]);
            s_o[cb * kUnits + u] = __float2half_rn(cb < B ? h : 0.0f);
            if (cb < B) y[(static_cast<size_t>(cb) * T3 + t) * (2 * H) + dir * H + u0 + u] = h;
        }
        __syncthreads();
        // push my [16 x 32] fp16 slice of h_t into the next buffer of every CTA of the cluster (64 x 16-byte pieces x 16 peers)
        for (int i = tid; i < kMb * (kUnits / 8) * kClusterCtas; i += blockDim.x) {
            const int peer = i / (kMb * (kUnits / 8)), p = i % (kMb * (kUnits / 8));
            const int row = p / (kUnits / 8), c8 = p % (kUnits / 8);
            const uint4 v = *reinterpret_cast<const uint4*>(s_o + row * kUnits + c8 * 8);
            const uint32_t dst = s_h_u32 + static_cast<uint32_t>((((cur ^ 1) * kMb + row) * kPitch + u0 + c8 * 8) * 2);
            st_cluster_u128(mapa_shared(dst, peer), v);
        }
        cluster_sync();                                               // release / acquire: h_t is complete everywhere
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
    (void)hbuf; (void)bar;
    if (H != 512) throw CudaError{"bilstm: hidden size must be 512"};
    if (B < 1 || B > kLstmMaxBatch) throw CudaError{"bilstm: 1..16 utterances per launch"};
    const int smem = kRows * kPitch * 2 + 2 * kMb * kPitch * 2 + kMb * kRows * 4 + kMb * kUnits * 2;
    static bool attr_set = false;
    if (!attr_set) {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_bilstm_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            PF_CUDA(cudaFuncSetAttribute(pf_bilstm_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // 16-CTA clusters
        }
        PF_CUDA(cudaSetDevice(cur));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * kClusterCtas);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kClusterCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PF_CUDA(cudaLaunchKernelEx(&cfg, pf_bilstm_cluster, gin, w_hh, B, T3, y));
}

void us_alphas_peaks_launch(const float* y, int B, int T3, int D2, const float* w2, const float* b2, float smooth, float noise,
                            const int* token_num, float thr, float* us_alphas, float* us_peaks, cudaStream_t s) {
    if (B <= 0 || T3 <= 0) return;
    pf_us_alphas_peaks<<<B, 256, static_cast<size_t>(T3) * sizeof(float), s>>>(y, T3, D2, w2, b2, smooth, noise, token_num, thr, us_alphas, us_peaks);
    PF_CUDA(cudaGetLastError());
}

}  // namespace pf
