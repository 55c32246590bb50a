// Half-SM variant of the tcgen05 GEMM for the fp16-output projections (QKV, FFN1, cross-attention K/V):
// C[M,N] = relu?(A[M,K] W[N,K]^T + bias) -> fp16, one 128 x 256 tile per CTA, sized so that TWO CTAs are resident per SM
// (<= 113 KB shared memory, 256 TMEM columns, 256 threads).
//
// Why: the persistent kernel of gemm.cu owns a whole SM (200 KB of operand ring, 512 TMEM columns), so nothing else can
// become resident while it runs its prologue, waits for its first loads or drains its last epilogue - ~3 us of every
// launch in which the SM's tensor pipe idles.  Here those phases of one CTA overlap the main loop of the
// CTA that shares the SM - a tile of the same launch, the first tile of the NEXT launch (programmatic dependent launch:
// its CTAs become resident, allocate TMEM and prefetch their weight tiles while this kernel still computes), or another
// execution lane's tile.  The bet: two stages per CTA are enough because the two co-resident CTAs together keep four
// operand stages in flight per SM, and their MMAs interleave on the one tensor pipe.
//
// MEASURED (round 2, A/B on one B200, scripts/quick_bench.py): parity-green and bit-identical to the persistent kernel,
// but slower - GEMM replay on one stream 484 vs 552 TFLOP/s (256-wide tiles), three lanes concurrently 706 vs 768 TFLOP/s,
// cfg2 step 3.94 vs 3.79 ms.  Two stages per CTA do not cover the L2 latency: each CTA's pipe stalls after every second
// k-block and the partner CTA does not fill the gap, so the main loop loses more than the overlapped prologue / epilogue
// returns.  Kept as an experiment (PFASR_BUILD_EXPERIMENTS=1 builds; PFASR_GEMM_HALFSM=1|2 or tile_code bit 23).
//
// Same warp roles and mbarrier protocol as gemm.cu, minus the tile loop: warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM row -> bias / ReLU -> fp16 -> swizzled 32 x 32 boxes -> TMA store).
// Replaces the MLAS GEMMs behind InferenceSession.Run (OfflineProjOfParaformer.cs:68), like gemm.cu.
#include <stdlib.h>

#include <mutex>

#include "gemm.cuh"
#include "gemm_dev.cuh"

namespace pf {

using namespace gemm_dev;

namespace {

constexpr int kHalfBN = 256;
constexpr int kHalfStages = 2;
constexpr int kHalfThreads = 256;
constexpr int kHalfStageBytes = kABytes + kHalfBN * BK * 2;          // 16 KiB A + 32 KiB W
constexpr int kHalfEpiBytes = 4 * 4096;                              // four epilogue warps x two 2 KiB boxes
constexpr int kHalfBarBytes = 128;
constexpr int kHalfSmemBytes = kHalfStages * kHalfStageBytes + kHalfEpiBytes + kHalfBarBytes;
static_assert(2 * (kHalfSmemBytes + 1024) <= 228 * 1024, "two CTAs (plus their 1 KiB system slices) must fit one SM");

__global__ void __launch_bounds__(kHalfThreads, 2)
pf_gemm_f16_tn_tcgen05_half(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const int relu, const int M,
                            const int N, const int K, const int tiles_n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    constexpr uint32_t kEpiOff = kHalfStages * kHalfStageBytes;
    constexpr uint32_t kBarOff = kEpiOff + kHalfEpiBytes;
    const uint32_t bar_base = base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kHalfStages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * kHalfStages);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * kHalfStages + 1));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;
    const int t = blockIdx.x;
    const int m0 = (t / tiles_n) * BM;
    const int n0 = (t % tiles_n) * kHalfBN;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((base & 1023u) != 0) { printf("pfasr: half-SM GEMM needs 1024-byte aligned dynamic shared memory\n"); __trap(); }
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < kHalfStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(tmem_slot)), kHalfBN);
        tmem_relinquish<1>();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kStageTx = kHalfStageBytes;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: the weight halves of the first stages do not depend on
        // the previous kernel, so they are requested BEFORE the programmatic-launch wait; activations after it
        if (lane == 0) {
            const int npre = num_kb < kHalfStages ? num_kb : kHalfStages;
            for (int kb = 0; kb < npre; ++kb) {
                mbar_arrive_expect_tx(full_bar(kb), kStageTx);
                tma_load_2d(base + kb * kHalfStageBytes + kABytes, &tmB, full_bar(kb), kb * BK, n0);
            }
            pdl_wait();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kHalfStages;
                const uint32_t a_s = base + s * kHalfStageBytes;
                if (kb < npre) {
                    tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                    continue;
                }
                mbar_wait(empty_bar(s), ((kb / kHalfStages) & 1) ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), kStageTx);
                tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                tma_load_2d(a_s + kABytes, &tmB, full_bar(s), kb * BK, n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, kHalfBN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kHalfStages;
                mbar_wait(full_bar(s), (kb / kHalfStages) & 1);
                tc_fence_after_sync();
                const uint32_t a_s = base + s * kHalfStageBytes;
                const uint64_t adesc0 = make_sw128_kmajor_desc(a_s);
                const uint64_t bdesc0 = make_sw128_kmajor_desc(a_s + kABytes);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_f16<1>(tmem_base, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(empty_bar(s));                       // frees the stage once the MMAs have read it
            }
            umma_commit(tmem_full_bar);                          // accumulator complete
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: one warp per TMEM lane quadrant, all 8 column chunks
        const int q = warp & 3;
        const int row0 = m0 + q * 32;
        uint8_t* wstage = smem + kEpiOff + (warp - 4) * 4096;
        int nchunks = 0;
        for (int c = 0; c < kHalfBN / 32; ++c)
            if (n0 + c * 32 < N) ++nchunks;
        pdl_wait();                                              // the output buffer may still be read by the previous kernel
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after_sync();
        if (row0 < M && nchunks > 0) {
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
            epilogue_tma_f16<kHalfBN / 32>(t_acc, nchunks, wstage, &tmC, bias + n0, relu ? 0.0f : -INFINITY, row0, n0, lane);
        }
        if (lane == 0) tma_store_wait_read();                    // staging boxes stay valid until the stores have read them
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, kHalfBN);
}

}  // namespace

bool gemm_half_eligible(const GemmOp& op, bool force) {
    // 0 = off (default: measured slower), 1 = multi-lane handles only (throughput objective), 2 = always
    static const int mode = [] { const char* e = getenv("PFASR_GEMM_HALFSM"); return e ? atoi(e) : 0; }();
    if (!force && (mode <= 0 || (mode == 1 && !op.throughput))) return false;
    return op.epi.out_f16 != nullptr && op.n_adds == 0 && (op.vec_ok & 2) != 0 && op.bn == kHalfBN && op.cm == 1 && op.cn == 1 &&
           op.ln_cluster == 0 && op.epi.bias != nullptr && (reinterpret_cast<uintptr_t>(op.epi.bias) & 15) == 0 && op.N % 32 == 0 &&
           (op.vec_ok & 0x1f00) == 0;
}

void gemm_half_launch(const GemmOp& op, cudaStream_t stream) {
    static std::once_flag once;
    std::call_once(once, [] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_gemm_f16_tn_tcgen05_half, cudaFuncAttributeMaxDynamicSharedMemorySize, kHalfSmemBytes));
            PF_CUDA(cudaFuncSetAttribute(pf_gemm_f16_tn_tcgen05_half, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    const int tiles_n = ceil_div(op.N, kHalfBN);
    const int num_tiles = tiles_n * ceil_div(op.M, BM);
    launch_k(pf_gemm_f16_tn_tcgen05_half, dim3(num_tiles), dim3(kHalfThreads), static_cast<size_t>(kHalfSmemBytes), stream, op.tmA, op.tmB,
             op.tmC, op.epi.bias, op.epi.relu, op.M, op.N, op.K, tiles_n);
}

}  // namespace pf
