// Hand-written sm_100a GEMM (v2): persistent CTAs, TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
// tcgen05.mma (single issuing thread, fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue that
// transposes through shared memory so every global load/store of the fused epilogue is a coalesced row segment.
//
// Replaces the MLAS GEMMs that OnnxRuntime runs for the reference at
// /root/reference/AliParaformerAsr/OfflineProjOfParaformer.cs:68 (InferenceSession.Run).
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4..11 = epilogue (two warpgroups; a warp owns the 32 TMEM lanes of quadrant warp % 4 and every second
// 32-column chunk).  The epilogue of tile i overlaps the main loop of tile i+1 (two TMEM accumulator stages).
#include "gemm.cuh"

#include <algorithm>
#include <mutex>

namespace pf {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                  // 64 fp16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int kABytes = BM * BK * 2;    // 16 KiB
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kStageTileBytes = 32 * 32 * 4;   // per-warp transpose buffer: 32 rows x 32 fp32

template <int BN>
struct Cfg {
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes = kEpiWarps * kStageTileBytes;   // 32 KiB
    static constexpr int kBarBytes = 256;
    static constexpr int kTmemCols = 2 * BN;                         // two accumulator stages (128 / 256 / 512)
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;  // +1024: alignment slack
};

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 (Blackwell) | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4)                                  // c_format = F32
           | (0u << 7) | (0u << 10)                   // a_format = b_format = F16
           | (0u << 15) | (0u << 16)                  // a_major = b_major = K
           | (static_cast<uint32_t>(n >> 3) << 17)    // n_dim
           | (static_cast<uint32_t>(m >> 4) << 24);   // m_dim
}

// One 32x32 accumulator chunk of one warp: registers (thread = row) -> XOR-swizzled smem (conflict-free both ways)
// -> row-segment layout (fp32 out: 8 lanes x float4 per row, 4 rows per instruction; fp16 out: 4 lanes x 8 halfs per
// row, 8 rows per instruction) -> bias / addends / ReLU -> 16-byte coalesced stores.  kAdds = number of fp32 tensors
// added to the product (FSMN memory, residual); the adds of a chunk are all issued before the first use.
template <bool kOutHalf, int kAdds>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], float* stage, int lane, int row0, int col0,
                                               int M, int N, const GemmEpi& e, bool vec_ok) {
    float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    const float lo = e.relu ? 0.0f : -INFINITY;
    if (vec_ok && col0 + 32 <= N) {
        if (kOutHalf) {
            const int g = lane & 3, rsub = lane >> 2;
            const int col = col0 + g * 8;
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (e.bias) {
                b0 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
                b1 = __ldg(reinterpret_cast<const float4*>(e.bias + col + 4));
            }
            __half* optr = e.out_f16 + static_cast<size_t>(row0 + rsub) * e.ld_out + col;
            const float* a0p = kAdds > 0 ? e.add0 + static_cast<size_t>(row0 + rsub) * e.ld_add0 + col : nullptr;
            const float* a1p = kAdds > 1 ? e.add1 + static_cast<size_t>(row0 + rsub) * e.ld_add1 + col : nullptr;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rl = it * 8 + rsub;
                const bool ok = row0 + rl < M;
                float4 v0 = st4[rl * 8 + ((2 * g) ^ (rl & 7))];
                float4 v1 = st4[rl * 8 + ((2 * g + 1) ^ (rl & 7))];
                v0.x += b0.x; v0.y += b0.y; v0.z += b0.z; v0.w += b0.w;
                v1.x += b1.x; v1.y += b1.y; v1.z += b1.z; v1.w += b1.w;
                if (kAdds > 0 && ok) {
                    const float4 x0 = *reinterpret_cast<const float4*>(a0p + static_cast<size_t>(it) * 8 * e.ld_add0);
                    const float4 x1 = *reinterpret_cast<const float4*>(a0p + static_cast<size_t>(it) * 8 * e.ld_add0 + 4);
                    v0.x += x0.x; v0.y += x0.y; v0.z += x0.z; v0.w += x0.w;
                    v1.x += x1.x; v1.y += x1.y; v1.z += x1.z; v1.w += x1.w;
                }
                if (kAdds > 1 && ok) {
                    const float4 x0 = *reinterpret_cast<const float4*>(a1p + static_cast<size_t>(it) * 8 * e.ld_add1);
                    const float4 x1 = *reinterpret_cast<const float4*>(a1p + static_cast<size_t>(it) * 8 * e.ld_add1 + 4);
                    v0.x += x0.x; v0.y += x0.y; v0.z += x0.z; v0.w += x0.w;
                    v1.x += x1.x; v1.y += x1.y; v1.z += x1.z; v1.w += x1.w;
                }
                __half2 h0 = __floats2half2_rn(fmaxf(v0.x, lo), fmaxf(v0.y, lo));
                __half2 h1 = __floats2half2_rn(fmaxf(v0.z, lo), fmaxf(v0.w, lo));
                __half2 h2 = __floats2half2_rn(fmaxf(v1.x, lo), fmaxf(v1.y, lo));
                __half2 h3 = __floats2half2_rn(fmaxf(v1.z, lo), fmaxf(v1.w, lo));
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                if (ok) *reinterpret_cast<uint4*>(optr + static_cast<size_t>(it) * 8 * e.ld_out) = pk;
            }
        } else {
            const int jj = lane & 7, rsub = lane >> 3;
            const int col = col0 + jj * 4;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + col));
            float* optr = e.out_f32 + static_cast<size_t>(row0 + rsub) * e.ld_out + col;
            float4 x0[8], x1[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const bool ok = row0 + it * 4 + rsub < M;
                x0[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                x1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kAdds > 0 && ok) x0[it] = *reinterpret_cast<const float4*>(e.add0 + static_cast<size_t>(row0 + it * 4 + rsub) * e.ld_add0 + col);
                if (kAdds > 1 && ok) x1[it] = *reinterpret_cast<const float4*>(e.add1 + static_cast<size_t>(row0 + it * 4 + rsub) * e.ld_add1 + col);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rl = it * 4 + rsub;
                float4 v = st4[rl * 8 + (jj ^ (rl & 7))];
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                if (kAdds > 0) { v.x += x0[it].x; v.y += x0[it].y; v.z += x0[it].z; v.w += x0[it].w; }
                if (kAdds > 1) { v.x += x1[it].x; v.y += x1[it].y; v.z += x1[it].z; v.w += x1[it].w; }
                v.x = fmaxf(v.x, lo); v.y = fmaxf(v.y, lo); v.z = fmaxf(v.z, lo); v.w = fmaxf(v.w, lo);
                if (row0 + rl < M) *reinterpret_cast<float4*>(optr + static_cast<size_t>(it) * 4 * e.ld_out) = v;
            }
        }
    } else {
        // ragged N edge (e.g. vocab 25055) or unaligned pitches: scalar, bounds-checked
        const int jj = lane & 7, rsub = lane >> 3;
        const int col = col0 + jj * 4;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rsub;
            const int row = row0 + rl;
            const float4 v4 = st4[rl * 8 + (jj ^ (rl & 7))];
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            if (row < M) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = col + i;
                    if (c < N) {
                        float x = v[i];
                        if (e.bias) x += __ldg(e.bias + c);
                        if (kAdds > 0) x += e.add0[static_cast<size_t>(row) * e.ld_add0 + c];
                        if (kAdds > 1) x += e.add1[static_cast<size_t>(row) * e.ld_add1 + c];
                        x = fmaxf(x, lo);
                        if (kOutHalf) e.out_f16[static_cast<size_t>(row) * e.ld_out + c] = __float2half_rn(x);
                        else e.out_f32[static_cast<size_t>(row) * e.ld_out + c] = x;
                    }
                }
            }
        }
    }
    __syncwarp();   // the transpose buffer is reused by the next chunk
}

// Thread-block cluster of CM x CN CTAs computes a (CM*128) x (CN*BN) super-tile.  CTA (rm, rn) owns the 128 x BN
// sub-tile and issues its own 1-CTA MMAs, but operand tiles are fetched from L2 once per cluster: the CTA loads
// 1/CN of its A tile and 1/CM of its B tile and TMA-multicasts each slice to the CTAs that share it (row peers share
// A, column peers share B).  L2->SM bytes per k-block drop from 16K + BN*128 to 16K/CN + BN*128/CM, which is what
// bounds this kernel (see pick_config).  A stage may be refilled only when every CTA that receives my slices has
// consumed it, so the MMA commit arrives on the empty barrier of all row and column peers (count CM + CN - 1).
template <int BN, int CM, int CN, bool kOutHalf, int kAdds>
__global__ void __launch_bounds__(kThreads, 1)
pf_gemm_f16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmEpi epi, const int M, const int N, const int K, const int stiles_n, const int num_stiles,
                       const int vec_ok) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::kStages;
    constexpr int CSIZE = CM * CN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    constexpr uint32_t kEpiOff = STAGES * C::kStageBytes;
    constexpr uint32_t kBarOff = kEpiOff + C::kEpiBytes;
    const uint32_t bar_base = base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;
    const int crank = CSIZE > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int rm = crank % CM, rn = crank / CM;
    const int cluster_id = blockIdx.x / CSIZE;
    const int num_clusters = gridDim.x / CSIZE;
    // CTAs exchanging operand slices with me: row peers (same rm) share A, column peers (same rn) share B
    uint16_t mask_a = 0, mask_b = 0;
#pragma unroll
    for (int j = 0; j < CN; ++j) mask_a |= static_cast<uint16_t>(1u << (rm + CM * j));
#pragma unroll
    for (int i = 0; i < CM; ++i) mask_b |= static_cast<uint16_t>(1u << (i + CM * rn));
    const uint16_t mask_peers = mask_a | mask_b;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CM + CN - 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    if (CSIZE > 1) cluster_sync(); else __syncthreads();        // peers' barriers are initialised before any remote arrive
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // prologue above overlaps the previous kernel's tail; operands / residuals are read below

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;                                     // global k-block counter across tiles
            for (int st = cluster_id; st < num_stiles; st += num_clusters) {
                const int m0 = ((st / stiles_n) * CM + rm) * BM;
                const int n0 = ((st % stiles_n) * CN + rn) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), C::kStageBytes);
                    const uint32_t a_s = base + s * C::kStageBytes;
                    if (CN == 1) tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                    else tma_load_2d_mc(a_s + rn * (kABytes / CN), &tmA, full_bar(s), kb * BK, m0 + rn * (BM / CN), mask_a);
                    if (CM == 1) tma_load_2d(a_s + kABytes, &tmB, full_bar(s), kb * BK, n0);
                    else tma_load_2d_mc(a_s + kABytes + rm * (C::kBBytes / CM), &tmB, full_bar(s), kb * BK, n0 + rm * (BN / CM), mask_b);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0;
            uint32_t local = 0;                                  // tiles processed by this CTA
            for (int st = cluster_id; st < num_stiles; st += num_clusters, ++local) {
                const uint32_t acc = local & 1u;
                const uint32_t acc_ph = (local >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);     // epilogue has drained this accumulator stage
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after_sync();
                    const uint32_t a_s = base + s * C::kStageBytes;
                    const uint64_t adesc0 = make_sw128_kmajor_desc(a_s);
                    const uint64_t bdesc0 = make_sw128_kmajor_desc(a_s + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K inside the swizzle atom = +32 bytes on the start address (>>4 => +2)
                        umma_f16(d_tmem, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // frees this smem stage (here and in every peer that multicasts into it) once the MMAs have read it
                    if (CSIZE == 1) umma_commit(empty_bar(s));
                    else umma_commit_mc(empty_bar(s), mask_peers);
                }
                umma_commit(tmem_full_bar(acc));                 // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------ epilogue: TMEM -> registers -> smem transpose -> global
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
        const int grp = ew >> 2;                                 // warpgroup: chunks grp, grp+2, ...
        float* stage = reinterpret_cast<float*>(smem + kEpiOff + ew * kStageTileBytes);
        uint32_t local = 0;
        for (int st = cluster_id; st < num_stiles; st += num_clusters, ++local) {
            const uint32_t acc = local & 1u;
            const uint32_t acc_ph = (local >> 1) & 1u;
            const int m0 = ((st / stiles_n) * CM + rm) * BM;
            const int n0 = ((st % stiles_n) * CN + rn) * BN;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after_sync();
            if (m0 < M) {
#pragma unroll 1
                for (int c = grp; c < BN / 32; c += 2) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= N) break;                        // warp-uniform
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + static_cast<uint32_t>(c * 32), r);
                    tmem_ld_wait();
                    epilogue_chunk<kOutHalf, kAdds>(r, stage, lane, m0 + q * 32, col0, M, N, epi, vec_ok != 0);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    tc_fence_before_sync();
    if (CSIZE > 1) cluster_sync(); else __syncthreads();        // no CTA exits while a peer may still signal its barriers
    if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) throw CudaError{"cuTensorMapEncodeTiled entry point not available (driver too old?)"};
    return fn;
}

// 2-D fp16 row-major [rows, cols] with row pitch ld elements; box = [box_rows, 64 cols], 128B swizzle, OOB -> 0.
void make_tmap(CUtensorMap* tm, const __half* ptr, int rows, int cols, int ld, int box_rows) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0)
        throw CudaError{"gemm operand must be 16-byte aligned with a row pitch that is a multiple of 8 elements"};
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(__half)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))};
}

int num_sms() {
    static int n = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) n = prop.multiProcessorCount;
        if (n <= 0) n = 148;
    });
    return n;
}

template <int BN, int CM, int CN, bool kOutHalf, int kAdds>
struct Launcher {
    static int max_clusters;   // co-resident clusters of this shape (GPC boundaries make it < sms / cluster size)

    static void run(const GemmOp& op, cudaStream_t stream) {
        using C = Cfg<BN>;
        constexpr int CSIZE = CM * CN;
        auto kern = pf_gemm_f16_tn_tcgen05<BN, CM, CN, kOutHalf, kAdds>;
        static std::once_flag once;
        std::call_once(once, [&] {
            int ndev = 0, cur = 0;
            PF_CUDA(cudaGetDeviceCount(&ndev));
            PF_CUDA(cudaGetDevice(&cur));
            for (int d = 0; d < ndev; ++d) {
                PF_CUDA(cudaSetDevice(d));
                PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
            }
            PF_CUDA(cudaSetDevice(cur));
            max_clusters = num_sms() / CSIZE;
            if (CSIZE > 1) {
                cudaLaunchConfig_t q{};
                q.gridDim = dim3(num_sms() / CSIZE * CSIZE);
                q.blockDim = dim3(kThreads);
                q.dynamicSmemBytes = C::kSmemBytes;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = CSIZE; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                q.attrs = at; q.numAttrs = 1;
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, kern, &q) == cudaSuccess && nc > 0) max_clusters = std::min(max_clusters, nc);
                else cudaGetLastError();
            }
        });
        const int stiles_n = ceil_div(ceil_div(op.N, BN), CN);
        const int num_stiles = stiles_n * ceil_div(ceil_div(op.M, BM), CM);
        const int grid = std::min(num_stiles, max_clusters) * CSIZE;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = C::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        int na = 0;
        if (pdl_enabled()) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (CSIZE > 1) {
            at[na].id = cudaLaunchAttributeClusterDimension;
            at[na].val.clusterDim.x = CSIZE; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = at;
        cfg.numAttrs = na;
        PF_CUDA(cudaLaunchKernelEx(&cfg, kern, op.tmA, op.tmB, op.epi, op.M, op.N, op.K, stiles_n, num_stiles, op.vec_ok));
    }
};
template <int BN, int CM, int CN, bool kOutHalf, int kAdds>
int Launcher<BN, CM, CN, kOutHalf, kAdds>::max_clusters = 1;

template <int BN, bool kOutHalf, int kAdds>
void launch_cl(const GemmOp& op, cudaStream_t stream) {
    if (op.cm == 2 && op.cn == 1) Launcher<BN, 2, 1, kOutHalf, kAdds>::run(op, stream);
    else if (op.cm == 1 && op.cn == 1) Launcher<BN, 1, 1, kOutHalf, kAdds>::run(op, stream);
    else throw CudaError{"gemm: cluster shape not instantiated"};
}

template <int BN>
void launch_bn(const GemmOp& op, cudaStream_t stream) {
    const bool h = op.epi.out_f16 != nullptr;
    switch (op.n_adds) {
        case 0:  h ? launch_cl<BN, true, 0>(op, stream) : launch_cl<BN, false, 0>(op, stream); break;
        case 1:  h ? launch_cl<BN, true, 1>(op, stream) : launch_cl<BN, false, 1>(op, stream); break;
        default: h ? launch_cl<BN, true, 2>(op, stream) : launch_cl<BN, false, 2>(op, stream); break;
    }
}

// Tile width and cluster shape from a cycle model of the persistent kernel fitted to scripts/gemm_sweep.py on B200:
// a tile's k-block costs max(MMA floor = 2*BN cycles for four K=16 steps, operand bytes / L2->SM bandwidth).  The L2
// fabric delivers ~6300 B/cycle chip-wide (B300_MICROARCH.md, TMA chip throughput), at most ~80 B/cycle to one SM,
// and is what bounds the 1-CTA 128xBN tile (85 FLOP/B at BN=256); multicast clusters cut the bytes per CTA.
void pick_config(int M, int N, int K, int& bn_out, int& cm_out, int& cn_out) {
    const int mt = ceil_div(M, BM), kb = ceil_div(K, BK), sms = num_sms();
    double best_cost = 1e30;
    for (int bn : {256, 128, 64}) {
        for (int cfg = 0; cfg < 1; ++cfg) {     // 2x1 multicast clusters measured slower on every shape of the path
            const int cm = cfg >= 1 ? 2 : 1, cn = 1;
            const int csize = cm * cn;
            const int stiles = ceil_div(mt, cm) * ceil_div(ceil_div(N, bn), cn);
            const int slots = csize == 4 ? sms / 4 - 1 : sms / csize;      // GPC boundaries cost a 4-cluster slot
            const int waves = ceil_div(stiles, slots);
            const int active = std::min(stiles, slots) * csize;
            const double bw = std::min(80.0, 6300.0 / active);
            const double t_kb = std::max(2.0 * bn, (16384.0 / cn + 128.0 * bn / cm) / bw) + (csize > 1 ? 20.0 : 0.0);
            const double cost = waves * kb * t_kb + 150.0 * (bn / 32);
            if (cost < best_cost) { best_cost = cost; bn_out = bn; cm_out = cm; cn_out = cn; }
        }
    }
}

}  // namespace

void* tensormap_encode_fn() { return reinterpret_cast<void*>(get_encode_fn()); }

void gemm_prepare(GemmOp& op, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  const GemmEpi& epi, int tile_code) {
    if ((epi.out_f32 != nullptr) == (epi.out_f16 != nullptr)) throw CudaError{"gemm: exactly one output pointer must be set"};
    if (M <= 0 || N <= 0 || K <= 0) throw CudaError{"gemm: empty problem"};
    int bn = tile_code & 0xFFF, cm = (tile_code >> 12) & 0xF, cn = (tile_code >> 16) & 0xF;
    if (tile_code == 0) pick_config(M, N, K, bn, cm, cn);
    cm = std::max(cm, 1);
    cn = std::max(cn, 1);
    if (bn != 64 && bn != 128 && bn != 256) throw CudaError{"gemm: unsupported N tile"};
    if (cm > 2 || cn != 1) throw CudaError{"gemm: unsupported cluster shape (1x1 and 2x1 are instantiated)"};
    op.M = M; op.N = N; op.K = K; op.bn = bn; op.cm = cm; op.cn = cn; op.epi = epi;
    // the kernel adds up to two fp32 tensors in a fixed order: FSMN memory first, then the residual
    op.n_adds = 0;
    op.epi.add0 = op.epi.add1 = nullptr;
    if (epi.addend) { op.epi.add0 = epi.addend; op.epi.ld_add0 = epi.ld_addend; op.n_adds = 1; }
    if (epi.resid) {
        if (op.n_adds == 0) { op.epi.add0 = epi.resid; op.epi.ld_add0 = epi.ld_resid; }
        else { op.epi.add1 = epi.resid; op.epi.ld_add1 = epi.ld_resid; }
        ++op.n_adds;
    }
    // vector epilogue: 16-byte row segments of every fp32 tensor (8-byte for the fp16 output) must be aligned
    auto al = [](const void* p, int ld, int bytes) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % bytes) == 0 && ld % 4 == 0); };
    op.vec_ok = (al(epi.bias, 0, 16) && al(epi.resid, epi.ld_resid, 16) && al(epi.addend, epi.ld_addend, 16) &&
                 al(epi.out_f32, epi.ld_out, 16) && al(epi.out_f16, epi.ld_out, 8)) ? 1 : 0;
    make_tmap(&op.tmA, A, M, K, lda, BM / cn);      // each CTA loads (and multicasts) 1/cn of its A tile
    make_tmap(&op.tmB, W, N, K, ldw, bn / cm);      // ... and 1/cm of its B tile
}

void gemm_launch(const GemmOp& op, cudaStream_t stream) {
    switch (op.bn) {
        case 64:  launch_bn<64>(op, stream);  break;
        case 128: launch_bn<128>(op, stream); break;
        default:  launch_bn<256>(op, stream); break;
    }
}

double gemm_flops(const GemmOp& op) { return 2.0 * op.M * static_cast<double>(op.N) * op.K; }

}  // namespace pf
