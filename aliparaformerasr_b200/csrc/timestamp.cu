// CifPredictorV3 timestamp branch: the graph outputs us_alphas / us_cif_peak that the reference reads as
// results[3] (OfflineProjOfParaformer.cs:75-79, OfflineProjOfSeacoParaformer.cs:122-126) and turns into per-token
// timestamps in OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302).  Graph semantics: FunASR
// CifPredictorV3.get_upsample_timestmap [EXT]: ConvTranspose1d x3 (a GEMM here) -> BiLSTM -> Linear(1024,1) -> sigmoid ->
// relu(a*0.25 - 0.01) -> rescale to token_num -> cif_wo_hidden(threshold 1 - 1e-4).
//
// The BiLSTM is a 3T-step (498 at 10 s) sequential recurrence over a [B,512] state: as per-step GEMM launches it would cost
// more than the whole recogniser, so it runs as ONE persistent kernel - a 16-CTA cluster per direction keeps W_hh in
// shared memory, multiplies with tcgen05.mma and exchanges h_t through distributed shared memory (st.async + mbarrier
// transaction counts).  History: global-memory barrier + scalar dot products 16 us / step; mma.sync + barrier.cluster
// 5.8 us; this version ~2 us.
#include "timestamp.cuh"

#include <mutex>

#include "gemm_dev.cuh"

#include <math.h>

namespace pf {

namespace {

constexpr int kClusterCtas = 16;      // CTAs per direction = one thread-block cluster; each owns 32 hidden units
constexpr int kUnits = 32;            // hidden units per CTA -> 128 gate rows of W_hh resident in shared memory
constexpr int kRows = 4 * kUnits;

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
// hidden units [32c, 32c+32) in shared memory for all T3 steps, as the K-major 128B-swizzled A operand of tcgen05.mma.
// Per step: gates^T [128 gate rows x 16 utterances] = W_hh[128 x 512] h_{t-1}^T - 32 tcgen05.mma (M = 128, N = 16, K = 16)
// issued by one thread, fp32 accumulator in TMEM (mma.sync is throttled on sm_100: the same product took 2460 of a step's
// 5700 cycles) - read back by four warps (thread = gate row), transposed through shared memory, + the precomputed input
// projections; cell update; the CTA's slice of h_t is pushed (fp16) into every peer's h buffer (the B operand, same swizzled
// layout) with st.async, each 16-byte piece completing the transaction count of the RECEIVER's mbarrier for that step: a CTA
// starts step t+1 as soon as the 16 slices of h_t have landed in its own buffer - no cluster-wide barrier (2600 cycles per
// step with barrier.cluster), and no WAR hazard either: a peer can only run one step ahead, and it writes the buffer this
// CTA finished reading before it pushed the slice the peer waited for.  No global-memory synchronisation, no re-reading of
// weights.
constexpr int kWBytes = kRows * 512 * 2;              // 128 KiB: 8 k-blocks x [128 rows x 128 B]
// MB = utterances per launch = N of the tcgen05 MMA (16 or 32: the instruction costs the same ~55 cycles either way)
template <int MB> constexpr int lstm_hbuf_bytes() { return MB * 512 * 2; }            // 8 k-blocks x [MB rows x 128 B]
template <int MB> constexpr int lstm_smem_bytes() { return kWBytes + 2 * lstm_hbuf_bytes<MB>() + MB * kRows * 4 + MB * kUnits * 2 + 64; }

// byte offset of element (row, k) inside a K-major SWIZZLE_128B operand with `rows` rows per 64-wide k-block
__device__ __forceinline__ uint32_t sw128_elem_off(int rows, int row, int k) {
    return static_cast<uint32_t>((k >> 6) * (rows * 128) + (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int MB>
__global__ void __launch_bounds__(256, 1)
pf_bilstm_cluster(const float* __restrict__ gin, const __half* __restrict__ w_hh, int B, int T3, float* __restrict__ y) {
    constexpr int H = 512;
    constexpr int kMb = MB;
    constexpr int kHBufBytes = lstm_hbuf_bytes<MB>();
    constexpr int kNh = MB / 16;                                                      // utterances per thread in the cell update
    extern __shared__ __align__(1024) uint8_t smem_l[];
    uint8_t* s_w = smem_l;                                                            // A operand: W_hh rows of this CTA
    uint8_t* s_h = s_w + kWBytes;                                                     // [2] B operand: h_{t-1} of all 512 units, 16 utterances
    float* s_g = reinterpret_cast<float*>(s_h + 2 * kHBufBytes);                      // [kMb][kRows] gate pre-activations (recurrent part)
    __half* s_o = reinterpret_cast<__half*>(s_g + kMb * kRows);                       // [kMb][kUnits] this CTA's h_t slice
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_o + kMb * kUnits);                // [0,1] "h buffer b complete" (16 KiB of transactions), [2] MMAs done
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(s_bar + 3);
    const int rank = static_cast<int>(cluster_ctarank());
    const int dir = blockIdx.x / kClusterCtas;
    const int u0 = rank * kUnits;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0 && (smem_u32(smem_l) & 1023u) != 0) { printf("pfasr: bilstm needs 1024-byte aligned dynamic shared memory\n"); __trap(); }
    // stage W_hh rows: local row r = gate (r / 32) * 32 + unit offset (r % 32), 16-byte pieces into the swizzled operand layout
    for (int i = tid; i < kRows * (H / 8); i += blockDim.x) {
        const int r = i / (H / 8), c8 = i % (H / 8);
        const int grow = (r / kUnits) * H + u0 + (r % kUnits);
        *reinterpret_cast<uint4*>(s_w + sw128_elem_off(kRows, r, c8 * 8)) =
            *reinterpret_cast<const uint4*>(w_hh + (static_cast<size_t>(dir) * 4 * H + grow) * H + c8 * 8);
    }
    for (int i = tid; i < 2 * kHBufBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_h)[i] = make_uint4(0, 0, 0, 0);   // h_0 = 0
    // cell roles: thread -> (utterances cb0 + 16 j, units cu, cu + 16)
    const int cb0 = tid >> 4, cu = tid & 15;
    float cstate[kNh][2];
#pragma unroll
    for (int j = 0; j < kNh; ++j) cstate[j][0] = cstate[j][1] = 0.0f;
    const uint32_t bar_u32 = smem_u32(s_bar);
    const uint32_t mma_bar = bar_u32 + 16;
    if (tid == 0) {
        mbar_init(bar_u32, 1);
        mbar_init(bar_u32 + 8, 1);
        mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 32);
        tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // W_hh and h_0 were written through the generic proxy
    tc_fence_before_sync();
    cluster_sync();                                                   // every CTA's buffers and barriers are initialised before remote writes
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t s_w_u32 = smem_u32(s_w), s_h_u32 = smem_u32(s_h);
    constexpr uint32_t kStepBytes = kClusterCtas * kMb * kUnits * 2;                 // 16 peers x [16 x 32] halfs = 16 KiB per step
    const uint32_t idesc = gemm_dev::make_idesc(kRows, kMb);
    // input projections of my two cells (4 gates each) are fetched ONE STEP AHEAD: gin (130 MB at 16 x 10 s) streams from
    // HBM, and a step is far shorter than a DRAM round trip
    auto load_gi = [&](int step_, float (&dst)[kNh][2][4]) {
        const int t_ = dir == 0 ? step_ : T3 - 1 - step_;
#pragma unroll
        for (int j = 0; j < kNh; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    dst[j][k][g] = (cb0 + 16 * j < B && step_ < T3)
                                       ? __ldg(gin + (static_cast<size_t>(cb0 + 16 * j) * T3 + t_) * (8 * H) + dir * 4 * H + g * H + u0 + cu + 16 * k) : 0.0f;
    };
    float gi[kNh][2][4], gi_next[kNh][2][4];
    load_gi(0, gi_next);
    for (int step = 0; step < T3; ++step) {
        const int t = dir == 0 ? step : T3 - 1 - step;
        const int cur = step & 1;
#pragma unroll
        for (int j = 0; j < kNh; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int g = 0; g < 4; ++g) gi[j][k][g] = gi_next[j][k][g];
        load_gi(step + 1, gi_next);
        if (tid == 0) mbar_arrive_expect_tx(bar_u32 + 8 * (cur ^ 1), kStepBytes);    // the buffer this step's h_t will land in
        // gates^T[128 x 16] = W_hh[128 x 512] * h_{t-1}[16 x 512]^T: one thread issues the 32 K = 16 steps
        if (tid == 128) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the peers' slices arrived through the generic proxy
            tc_fence_after_sync();
#pragma unroll
            for (int kb = 0; kb < H / 64; ++kb) {
                const uint64_t adesc0 = gemm_dev::make_sw128_kmajor_desc(s_w_u32 + kb * (kRows * 128));
                const uint64_t bdesc0 = gemm_dev::make_sw128_kmajor_desc(s_h_u32 + cur * kHBufBytes + kb * (kMb * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(mma_bar);
        }
        if (warp < 4) {                                               // thread = gate row (TMEM lane), 16 utterances in 16 columns
            mbar_wait(mma_bar, step & 1);
            tc_fence_after_sync();
            uint32_t v[32];
            if constexpr (MB == 16) {
                uint32_t v16[16];
                tmem_ld_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), v16);
                tmem_ld_wait();
#pragma unroll
                for (int b = 0; b < 16; ++b) v[b] = v16[b];
            } else {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), v);
                tmem_ld_wait();
            }
#pragma unroll
            for (int b = 0; b < kMb; ++b) s_g[b * kRows + tid] = __uint_as_float(v[b]);
            tc_fence_before_sync();
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kNh; ++j) {
            const int cb = cb0 + 16 * j;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int u = cu + 16 * k;
                const float* g = s_g + cb * kRows;
                const float ig = 1.0f / (1.0f + expf(-(g[u] + gi[j][k][0])));
                const float fg = 1.0f / (1.0f + expf(-(g[kUnits + u] + gi[j][k][1])));
                const float gg = tanhf(g[2 * kUnits + u] + gi[j][k][2]);
                const float og = 1.0f / (1.0f + expf(-(g[3 * kUnits + u] + gi[j][k][3])));
                cstate[j][k] = fg * cstate[j][k] + ig * gg;
                const float h = og * tanhf(cstate[j][k]);
                s_o[cb * kUnits + u] = __float2half_rn(cb < B ? h : 0.0f);
                if (cb < B) y[(static_cast<size_t>(cb) * T3 + t) * (2 * H) + dir * H + u0 + u] = h;
            }
        }
        __syncthreads();
        // push my [MB x 32] fp16 slice of h_t into the next buffer of every CTA of the cluster (4 MB x 16-byte pieces x 16 peers,
        // myself included), then wait until all 16 slices have landed in MY next buffer
        for (int i = tid; i < kMb * (kUnits / 8) * kClusterCtas; i += blockDim.x) {
            const int peer = i / (kMb * (kUnits / 8)), p = i % (kMb * (kUnits / 8));
            const int row = p / (kUnits / 8), c8 = p % (kUnits / 8);
            const uint4 v = *reinterpret_cast<const uint4*>(s_o + row * kUnits + c8 * 8);
            const uint32_t dst = s_h_u32 + (cur ^ 1) * kHBufBytes + sw128_elem_off(kMb, row, u0 + c8 * 8);
            st_async_u128(mapa_shared(dst, peer), v, mapa_shared(bar_u32 + 8 * (cur ^ 1), peer));
        }
        mbar_wait_cluster_acq(bar_u32 + 8 * (cur ^ 1), (step >> 1) & 1);
    }
    tc_fence_before_sync();
    cluster_sync();                                                   // nobody leaves while a peer's last slices may still be in flight to it
    if (warp == 5) tmem_dealloc(tmem_base, 32);
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

template <int MB>
static void bilstm_launch_t(const float* gin, const __half* w_hh, int B, int T3, float* y, cudaStream_t s) {
    constexpr int smem = lstm_smem_bytes<MB>();
    static std::once_flag attr_once;                              // execution lanes call this from several host threads
    std::call_once(attr_once, [&] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_bilstm_cluster<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            PF_CUDA(cudaFuncSetAttribute(pf_bilstm_cluster<MB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // 16-CTA clusters
        }
        PF_CUDA(cudaSetDevice(cur));
    });
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
    PF_CUDA(cudaLaunchKernelEx(&cfg, pf_bilstm_cluster<MB>, gin, w_hh, B, T3, y));
}

void bilstm_launch(const float* gin, const __half* w_hh, int B, int T3, int H, float* y, float* hbuf, unsigned int* bar, cudaStream_t s) {
    (void)hbuf; (void)bar;
    if (H != 512) throw CudaError{"bilstm: hidden size must be 512"};
    if (B < 1 || B > kLstmMaxBatch) throw CudaError{"bilstm: 1..32 utterances per launch"};
    if (B <= 16) bilstm_launch_t<16>(gin, w_hh, B, T3, y, s);
    else bilstm_launch_t<32>(gin, w_hh, B, T3, y, s);
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
