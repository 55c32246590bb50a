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
// 16 bytes into a peer CTA's shared memory, completing 16 bytes of the transaction count of THAT CTA's mbarrier: data and
// "it arrived" travel together, so the receiver needs no barrier round trip
__device__ __forceinline__ void st_async_u128(uint32_t cluster_addr, const uint4& v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_acq(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) { printf("pfasr: bilstm h exchange timeout (block %d)\n", blockIdx.x); __trap(); }
    }
}

// One thread-block cluster (16 CTAs) per direction.  CTA c keeps the 128 rows of W_hh that produce the four gates of
// hidden units [32c, 32c+32) in shared memory for all T3 steps.  Per step: gates = h_{t-1} W_hh^T (mma.sync, M = 16
// utterances, N = 128 gate rows, K = 512) + the precomputed input projections; cell update; the CTA's slice of h_t is
// pushed (fp16) into every peer's shared memory with st.async, each 16-byte piece completing the transaction count of the
// RECEIVER's mbarrier for that step: a CTA starts step t+1 as soon as the 16 slices of h_t have landed in its own buffer - no
// cluster-wide barrier (2600 of the 6900 cycles of a step with barrier.cluster), and no WAR hazard either: a peer can only run
// one step ahead, and it writes the buffer this CTA finished reading before it pushed the slice the peer waited for.  No
// global-memory synchronisation, no re-reading of weights.
__global__ void __launch_bounds__(256, 1)
pf_bilstm_cluster(const float* __restrict__ gin, const __half* __restrict__ w_hh, int B, int T3, float* __restrict__ y) {
    constexpr int H = 512;
    extern __shared__ __align__(16) uint8_t smem_l[];
    __half* s_w = reinterpret_cast<__half*>(smem_l);                                  // [kRows][kPitch]
    __half* s_h = s_w + kRows * kPitch;                                               // [2][kMb][kPitch]  h_{t-1} (all units), double buffered
    float* s_g = reinterpret_cast<float*>(s_h + 2 * kMb * kPitch);                    // [kMb][kRows] gate pre-activations (recurrent part)
    __half* s_o = reinterpret_cast<__half*>(s_g + kMb * kRows);                       // [kMb][kUnits] this CTA's h_t slice
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_o + kMb * kUnits);                // [2] "h buffer b is complete" (16 KiB of transactions)
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
    const uint32_t bar_u32 = smem_u32(s_bar);
    if (tid == 0) {
        mbar_init(bar_u32, 1);
        mbar_init(bar_u32 + 8, 1);
        fence_barrier_init();
    }
    cluster_sync();                                                   // every CTA's buffers and barriers are initialised before remote writes
    const uint32_t s_w_u32 = smem_u32(s_w), s_h_u32 = smem_u32(s_h);
    constexpr uint32_t kStepBytes = kClusterCtas * kMb * kUnits * 2;                 // 16 peers x [16 x 32] halfs = 16 KiB per step
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
        if (tid == 0) mbar_arrive_expect_tx(bar_u32 + 8 * (cur ^ 1), kStepBytes);    // the buffer this step's h_t will land in
        // four independent accumulator sets: the K = 512 reduction is a chain of 8 dependent MMAs instead of 32
        float acc4[4][2][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc4[a][nt][e] = 0.0f;
        const uint32_t a_base = s_h_u32 + static_cast<uint32_t>(((cur * kMb + (lane & 15)) * kPitch + (lane >> 4) * 8) * 2);
        const uint32_t b_base = s_w_u32 + static_cast<uint32_t>(((warp * 16 + (lane >> 4) * 8 + (lane & 7)) * kPitch + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
        for (int k0 = 0; k0 < H; k0 += 16) {
            uint32_t a[4], b0, b1, b2, b3;
            ldsm_x4(a_base + k0 * 2, a[0], a[1], a[2], a[3]);
            ldsm_x4(b_base + k0 * 2, b0, b1, b2, b3);
            mma16816(acc4[(k0 >> 4) & 3][0], a, b0, b1);
            mma16816(acc4[(k0 >> 4) & 3][1], a, b2, b3);
        }
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = (acc4[0][nt][e] + acc4[1][nt][e]) + (acc4[2][nt][e] + acc4[3][nt][e]);
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
        // push my [16 x 32] fp16 slice of h_t into the next buffer of every CTA of the cluster (64 x 16-byte pieces x 16 peers,
        // myself included), then wait until all 16 slices have landed in MY next buffer
        for (int i = tid; i < kMb * (kUnits / 8) * kClusterCtas; i += blockDim.x) {
            const int peer = i / (kMb * (kUnits / 8)), p = i % (kMb * (kUnits / 8));
            const int row = p / (kUnits / 8), c8 = p % (kUnits / 8);
            const uint4 v = *reinterpret_cast<const uint4*>(s_o + row * kUnits + c8 * 8);
            const uint32_t dst = s_h_u32 + static_cast<uint32_t>((((cur ^ 1) * kMb + row) * kPitch + u0 + c8 * 8) * 2);
            st_async_u128(mapa_shared(dst, peer), v, mapa_shared(bar_u32 + 8 * (cur ^ 1), peer));
        }
        mbar_wait_cluster_acq(bar_u32 + 8 * (cur ^ 1), (step >> 1) & 1);
    }
    cluster_sync();                                                   // nobody leaves while a peer's last slices may still be in flight to it
}

// alphas2 of the upsampled frames, one warp per (utterance, frame): Linear(1024, 1) -> sigmoid -> relu(a * smooth - noise).
// (The first version did this inside the per-utterance scan kernel below - 16 CTAs for 7968 dot products of 1024: 480 us.)
__global__ void __launch_bounds__(256)
pf_us_alphas_raw(const float* __restrict__ y, int rows, int D2, const float* __restrict__ w2, const float* __restrict__ b2,
                 float smooth, float noise, float* __restrict__ raw) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = y + static_cast<size_t>(row) * D2;
    float acc = 0.0f;
    for (int c = lane; c < D2; c += 32) acc += r[c] * w2[c];
    acc = warp_sum(acc);
    const float sg = 1.0f / (1.0f + expf(-(acc + b2[0])));
    if (lane == 0) raw[row] = fmaxf(sg * smooth - noise, 0.0f);
}

// one CTA per utterance: rescale to token_num -> integrate-and-fire trace (a sequential recurrence over 3T frames: one thread, in
// shared memory; everything around it is parallel).  Summation order of the total as in the first version: eight strided partial
// sums, then left to right.
__global__ void __launch_bounds__(256)
pf_us_alphas_peaks(const float* raw, int T3, const int* __restrict__ token_num, float thr, float* __restrict__ us_alphas,
                   float* us_peaks) {      // raw may alias us_peaks (read into shared memory before anything is written)
    extern __shared__ float s_a[];              // [T3] alphas, then [T3] peaks
    float* s_p = s_a + T3;
    __shared__ float s_red[8];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < T3; t += blockDim.x) s_a[t] = raw[static_cast<size_t>(b) * T3 + t];
    __syncthreads();
    if (lane == 0) {
        float part = 0.0f;
        for (int t = warp; t < T3; t += 8) part += s_a[t];
        s_red[warp] = part;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int i = 0; i < 8; ++i) tot += s_red[i];
        const float ratio = static_cast<float>(token_num[b]) / tot;
        float integrate = 0.0f;
        for (int t = 0; t < T3; ++t) {
            const float a = s_a[t] * ratio;
            s_a[t] = a;
            integrate = __fadd_rn(integrate, a);
            s_p[t] = integrate;
            if (integrate >= thr) integrate = __fsub_rn(integrate, thr);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T3; t += blockDim.x) {
        us_alphas[static_cast<size_t>(b) * T3 + t] = s_a[t];
        us_peaks[static_cast<size_t>(b) * T3 + t] = s_p[t];
    }
}

}  // namespace

void bilstm_launch(const float* gin, const __half* w_hh, int B, int T3, int H, float* y, float* hbuf, unsigned int* bar, cudaStream_t s) {
    (void)hbuf; (void)bar;
    if (H != 512) throw CudaError{"bilstm: hidden size must be 512"};
    if (B < 1 || B > kLstmMaxBatch) throw CudaError{"bilstm: 1..16 utterances per launch"};
    const int smem = kRows * kPitch * 2 + 2 * kMb * kPitch * 2 + kMb * kRows * 4 + kMb * kUnits * 2 + 16;
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
    // us_peaks doubles as the scratch for the un-scaled alphas: the scan kernel reads its row into shared memory before it writes
    const int rows = B * T3;
    pf_us_alphas_raw<<<ceil_div(rows, 8), 256, 0, s>>>(y, rows, D2, w2, b2, smooth, noise, us_peaks);
    PF_CUDA(cudaGetLastError());
    pf_us_alphas_peaks<<<B, 256, static_cast<size_t>(2 * T3) * sizeof(float), s>>>(us_peaks, T3, token_num, thr, us_alphas, us_peaks);
    PF_CUDA(cudaGetLastError());
}

}  // namespace pf
