// LayerNorm + GEMM in one kernel for the two fp16-output projections of an encoder layer whose input is the residual
// stream:   out[M,N] = relu?( LN(x[M,512]) W[N,512]^T + bias )  ->  fp16      (QKV: N = 1536, FFN1: N = 2048)
//
// Row-tile stationary: one CTA owns 128 rows.  All twelve warps first normalise those rows straight from the fp32
// residual stream into shared memory as the COMPLETE fp16 A operand (8 k-blocks x 16 KiB, 128B-swizzled K-major, the layout
// tcgen05.mma expects) - same arithmetic, lane ownership and summation order as pf_layernorm (ops.cu), so the operand is
// bit-identical to what the stand-alone kernel would have written.  Then the CTA walks the N tiles of its row tile: warp 0
// streams 128-row W tiles through a 6-stage TMA ring, warp 1 issues the MMAs against the resident A (fp32 accumulators
// double-buffered in TMEM), warps 4..11 drain tile i (bias, ReLU, fp16, 16-byte stores straight from registers) while the
// MMAs of tile i+1 run.
//
// Why: with several execution lanes the step is the sum of the kernels' SM time.  LayerNorm + GEMM as two
// launches hold 148 SMs for ~3.2 us and 112-126 SMs for 10-13 us; this kernel holds 42 SMs for ~20-30 us - about a third
// less SM time - re-reads nothing (A is loaded once per row tile instead of once per N tile: L2->SM bytes 97 MB instead of
// 148 MB for FFN1) and removes a launch.  It is slower for a lone batch (42 of 148 SMs), so only multi-lane handles use it
// (throughput objective; PFASR_LN_GEMM=0 off, 2 always).
//
// Replaces MLAS LayerNormalization + MatMul behind InferenceSession.Run (OfflineProjOfParaformer.cs:68), like gemm.cu.
#include "gemm_ln.cuh"

#include <stdlib.h>

#include <mutex>

#include "gemm_dev.cuh"

namespace pf {

using namespace gemm_dev;

namespace {

constexpr int kK = 512;                               // d_model: the whole row is one A operand
constexpr int kKb = kK / BK;                          // 8 k-blocks
constexpr int kBN = 128;                              // W tile rows: 16 KiB stages, six of them in flight cover the L2 latency
constexpr int kStages = 6;
constexpr int kThreadsLn = 384;
constexpr int kABytesAll = kKb * kABytes;             // 128 KiB
constexpr int kBStage = kBN * BK * 2;                 // 16 KiB
constexpr int kBiasBytes = 2 * kBN * 4;
constexpr int kBarBytes = 256;
constexpr int kSmemLn = kABytesAll + kStages * kBStage + kBiasBytes + kBarBytes;   // no alignment slack: the base is checked
static_assert(kSmemLn <= kSmemMax, "A operand + W ring must fit one SM");

// thread = row: 32 fp32 accumulator columns -> bias / ReLU -> 32 halfs = four 16-byte stores
__device__ __forceinline__ void store_chunk_f16(const uint32_t (&r)[32], const float* bias_c, float lo, __half* dst) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __half2 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = 8 * j + 2 * i;
            h[i] = __floats2half2_rn(fmaxf(__uint_as_float(r[c]) + bias_c[c], lo), fmaxf(__uint_as_float(r[c + 1]) + bias_c[c + 1], lo));
        }
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h[0]);
        pk.y = *reinterpret_cast<uint32_t*>(&h[1]);
        pk.z = *reinterpret_cast<uint32_t*>(&h[2]);
        pk.w = *reinterpret_cast<uint32_t*>(&h[3]);
        *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
    }
}

__global__ void __launch_bounds__(kThreadsLn, 1)
pf_ln_gemm_f16_rowtile(const float* __restrict__ x, const int ld_x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float eps, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ bias, __half* __restrict__ out,
                       const int ld_out, const int relu, const int M, const int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);                 // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
    constexpr uint32_t kBOff = kABytesAll;
    constexpr uint32_t kBiasOff = kBOff + kStages * kBStage;
    constexpr uint32_t kBarOff = kBiasOff + kBiasBytes;
    const uint32_t bar_base = base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * kStages + 4));
    float* bias_s = reinterpret_cast<float*>(smem + kBiasOff);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int tiles_n = (N + kBN - 1) / kBN;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((base & 1023u) != 0) { printf("pfasr: ln_gemm needs 1024-byte aligned dynamic shared memory\n"); __trap(); }
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(tmem_slot)), 2 * kBN);
        tmem_relinquish<1>();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // the W tiles of the first stages do not depend on the previous kernel: requested before the programmatic-launch wait
    int npre = 0;
    if (warp == 0 && lane == 0) {
        const int total = tiles_n * kKb;
        for (; npre < kStages && npre < total; ++npre) {
            mbar_arrive_expect_tx(full_bar(npre), kBStage);
            tma_load_2d(base + kBOff + npre * kBStage, &tmB, full_bar(npre), (npre % kKb) * BK, (npre / kKb) * kBN);
        }
    }
    pdl_wait();

    // ------------------------------------------------ LayerNorm of the CTA's 128 rows -> resident A operand (all warps)
    {
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* b4 = reinterpret_cast<const float4*>(beta);
        float4 g[4], bt[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { g[i] = __ldg(g4 + lane + 32 * i); bt[i] = __ldg(b4 + lane + 32 * i); }
        constexpr int kWarpsAll = kThreadsLn / 32;
#pragma unroll 1
        for (int r0 = warp; r0 < BM; r0 += 2 * kWarpsAll) {
            // two rows per iteration: their loads are in flight together
            float4 v[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + u * kWarpsAll;
                const int row = m0 + r;
                const bool ok = r < BM && row < M;
                const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<size_t>(ok ? row : 0) * ld_x);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][i] = ok ? x4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + u * kWarpsAll;
                if (r >= BM) continue;                                    // warp-uniform
                const bool ok = m0 + r < M;
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) sum += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
                sum = warp_sum(sum);
                const float mean = sum * (1.0f / kK);
                float sq = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
                    sq += (a * a + b * b) + (c * c + d * d);
                }
                sq = warp_sum(sq);
                const float rstd = 1.0f / sqrtf(sq * (1.0f / kK) + eps);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 y;
                    y.x = (v[u][i].x - mean) * rstd * g[i].x + bt[i].x;
                    y.y = (v[u][i].y - mean) * rstd * g[i].y + bt[i].y;
                    y.z = (v[u][i].z - mean) * rstd * g[i].z + bt[i].z;
                    y.w = (v[u][i].w - mean) * rstd * g[i].w + bt[i].w;
                    if (!ok) y = make_float4(0.f, 0.f, 0.f, 0.f);         // rows past M: a zero operand row
                    __half2 h0 = __floats2half2_rn(y.x, y.y);
                    __half2 h1 = __floats2half2_rn(y.z, y.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    // element k = 4 * (lane + 32 i): k-block k / 64, 16-byte chunk (k % 64) / 8 (XOR-swizzled with the row), half of it
                    const int k = 4 * (lane + 32 * i);
                    const uint32_t off = static_cast<uint32_t>((k >> 6) * kABytes + (r >> 3) * 1024 + (r & 7) * 128 +
                                                               ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
                    *reinterpret_cast<uint2*>(smem + off) = pk;
                }
            }
        }
        fence_proxy_async_smem();                                         // generic-proxy stores -> visible to tcgen05.mma
    }
    __syncthreads();

    if (warp == 0) {
        // ------------------------------------------------ W producer
        if (lane == 0) {
            const int total = tiles_n * kKb;
            for (int it = npre; it < total; ++it) {
                const int s = it % kStages;
                mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), kBStage);
                tma_load_2d(base + kBOff + s * kBStage, &tmB, full_bar(s), (it % kKb) * BK, (it / kKb) * kBN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, kBN);
            int it = 0;
            for (int n = 0; n < tiles_n; ++n) {
                const uint32_t acc = n & 1u;
                mbar_wait(tmem_empty_bar(acc), ((n >> 1) & 1u) ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * kBN;
                for (int kb = 0; kb < kKb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(full_bar(s), (it / kStages) & 1);
                    tc_fence_after_sync();
                    const uint64_t adesc0 = make_sw128_kmajor_desc(base + kb * kABytes);
                    const uint64_t bdesc0 = make_sw128_kmajor_desc(base + kBOff + s * kBStage);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_f16<1>(d_tmem, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(s));
                }
                umma_commit(tmem_full_bar(acc));
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: TMEM row -> bias / ReLU -> fp16 -> global
        const int ew = warp - 4;
        const int q = warp & 3;
        const int grp = ew >> 2;                                          // left / right half of the tile's columns
        constexpr int kHalfCols = kBN / 2;
        constexpr int kChunks = kHalfCols / 32;
        const int row = m0 + q * 32 + lane;
        const float lo = relu ? 0.0f : -INFINITY;
        for (int n = 0; n < tiles_n; ++n) {
            const uint32_t acc = n & 1u;
            const int n0 = n * kBN;
            float* bias_t = bias_s + acc * kBN;
            for (int i = ew * 32 + lane; i < kBN; i += 8 * 32) bias_t[i] = (bias != nullptr && n0 + i < N) ? __ldg(bias + n0 + i) : 0.0f;
            mbar_wait(tmem_full_bar(acc), (n >> 1) & 1u);
            tc_fence_after_sync();
            asm volatile("bar.sync 1, 256;" ::: "memory");               // bias tile visible to the eight epilogue warps
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBN + grp * kHalfCols;
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t (&cur)[32] = (c & 1) ? rb : ra;
                uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
                tmem_ld_wait();
                if (c + 1 < kChunks) tmem_ld_32x32(t_acc + (c + 1) * 32, nxt);
                const int col = n0 + grp * kHalfCols + c * 32;
                if (row < M && col < N)                                   // N is a multiple of 32 (checked on the host)
                    store_chunk_f16(cur, bias_t + grp * kHalfCols + c * 32, lo, out + static_cast<size_t>(row) * ld_out + col);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 2 * kBN);
}

}  // namespace

int ln_gemm_mode() {
    // 0 = off (default until an A/B shows a win), 1 = multi-lane handles only (throughput objective), 2 = always
    static const int mode = [] { const char* e = getenv("PFASR_LN_GEMM"); return e ? atoi(e) : 0; }();
    return mode;
}

bool ln_gemm_supported(int M, int N, int K, const void* x, int ld_x, const void* out, int ld_out, const void* bias) {
    return M > 0 && K == kK && N % 32 == 0 && N >= kBN && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld_x % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ld_out % 8 == 0 && (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 3) == 0);
}

void ln_gemm_prepare(LnGemmOp& op, const float* x, int ld_x, const float* gamma, const float* beta, float eps, const __half* W, int ldw,
                     const float* bias, __half* out, int ld_out, int relu, int M, int N, int K) {
    if (!ln_gemm_supported(M, N, K, x, ld_x, out, ld_out, bias)) throw CudaError{"ln_gemm: unsupported shape or alignment"};
    op.x = x; op.ld_x = ld_x; op.gamma = gamma; op.beta = beta; op.eps = eps; op.bias = bias; op.out = out; op.ld_out = ld_out;
    op.relu = relu; op.M = M; op.N = N; op.K = K;
    gemm_make_tmap(&op.tmB, W, false, N, K, ldw, kBN);
    op.valid = true;
}

void ln_gemm_launch(const LnGemmOp& op, cudaStream_t stream) {
    static std::once_flag once;
    std::call_once(once, [] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_ln_gemm_f16_rowtile, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLn));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    launch_k(pf_ln_gemm_f16_rowtile, dim3(ceil_div(op.M, BM)), dim3(kThreadsLn), static_cast<size_t>(kSmemLn), stream, op.x, op.ld_x, op.gamma,
             op.beta, op.eps, op.tmB, op.bias, op.out, op.ld_out, op.relu, op.M, op.N);
}

double ln_gemm_flops(const LnGemmOp& op) { return 2.0 * op.M * static_cast<double>(op.N) * op.K; }

}  // namespace pf
