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

// One 32x32 accumulator chunk of one warp: registers (thread = row) -> swizzled smem -> (thread = 4 columns of a
// row, 8 lanes per row) -> bias / addend / residual / ReLU -> coalesced store.
template <bool kOutHalf>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], float* stage, int lane, int row0, int col0,
                                               int M, int N, const GemmEpi& e, bool vec_ok) {
    // ---- transpose through shared memory (XOR swizzle on the float4 index: conflict-free both ways)
    float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    const int jj = lane & 7;
    const int rsub = lane >> 3;
    const int col = col0 + jj * 4;
    if (vec_ok && col + 4 <= N) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + col));
        float4 ad[8], rs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int row = row0 + it * 4 + rsub;
            ad[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < M) {
                if (e.addend) ad[it] = *reinterpret_cast<const float4*>(e.addend + static_cast<size_t>(row) * e.ld_addend + col);
                if (e.resid) rs[it] = *reinterpret_cast<const float4*>(e.resid + static_cast<size_t>(row) * e.ld_resid + col);
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rsub;
            const int row = row0 + rl;
            float4 v = st4[rl * 8 + (jj ^ (rl & 7))];
            v.x += b.x + ad[it].x + rs[it].x;
            v.y += b.y + ad[it].y + rs[it].y;
            v.z += b.z + ad[it].z + rs[it].z;
            v.w += b.w + ad[it].w + rs[it].w;
            if (e.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (row < M) {
                if (kOutHalf) {
                    __half2 h0 = __floats2half2_rn(v.x, v.y);
                    __half2 h1 = __floats2half2_rn(v.z, v.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(e.out_f16 + static_cast<size_t>(row) * e.ld_out + col) = pk;
                } else {
                    *reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(row) * e.ld_out + col) = v;
                }
            }
        }
    } else {
        // ragged N edge (e.g. vocab 25055) or unaligned pitches: scalar, bounds-checked
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
                        if (e.addend) x += e.addend[static_cast<size_t>(row) * e.ld_addend + c];
                        if (e.resid) x += e.resid[static_cast<size_t>(row) * e.ld_resid + c];
                        if (e.relu) x = fmaxf(x, 0.0f);
                        if (kOutHalf) e.out_f16[static_cast<size_t>(row) * e.ld_out + c] = __float2half_rn(x);
                        else e.out_f32[static_cast<size_t>(row) * e.ld_out + c] = x;
                    }
                }
            }
        }
    }
    __syncwarp();   // the transpose buffer is reused by the next chunk
}

template <int BN, bool kOutHalf>
__global__ void __launch_bounds__(kThreads, 1)
pf_gemm_f16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmEpi epi, const int M, const int N, const int K, const int tiles_n, const int num_tiles,
                       const int vec_ok) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::kStages;
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

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
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
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // prologue above overlaps the previous kernel's tail; operands / residuals are read below

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;                                     // global k-block counter across tiles
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM;
                const int n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), C::kStageBytes);
                    const uint32_t a_s = base + s * C::kStageBytes;
                    tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                    tma_load_2d(a_s + kABytes, &tmB, full_bar(s), kb * BK, n0);
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
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
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
                    umma_commit(empty_bar(s));                   // frees this smem stage once the MMAs above have read it
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
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const uint32_t acc = local & 1u;
            const uint32_t acc_ph = (local >> 1) & 1u;
            const int m0 = (tile / tiles_n) * BM;
            const int n0 = (tile % tiles_n) * BN;
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = grp; c < BN / 32; c += 2) {
                const int col0 = n0 + c * 32;
                if (col0 >= N) break;                            // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + static_cast<uint32_t>(c * 32), r);
                tmem_ld_wait();
                epilogue_chunk<kOutHalf>(r, stage, lane, m0 + q * 32, col0, M, N, epi, vec_ok != 0);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
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

// Tile width from a cycle model of the persistent kernel fitted to scripts/gemm_sweep.py on B200: a tile's k-block
// costs max(MMA floor = 2*BN cycles for four K=16 steps, operand bytes / L2->SM bandwidth).  The L2 fabric delivers
// ~6300 B/cycle chip-wide (B300_MICROARCH.md, TMA chip throughput), at most ~80 B/cycle to one SM, and is what
// bounds the 1-CTA 128xBN tile (85 FLOP/B at BN=256); plus the un-overlapped epilogue of the last tile.
int pick_bn(int M, int N, int K) {
    const int mt = ceil_div(M, BM), kb = ceil_div(K, BK), sms = num_sms();
    int best = 128;
    double best_cost = 1e30;
    for (int bn : {256, 128, 64}) {
        const int tiles = mt * ceil_div(N, bn);
        const int waves = ceil_div(tiles, sms);
        const double bw = std::min(80.0, 6300.0 / std::min(tiles, sms));
        const double t_kb = std::max(2.0 * bn, (16384.0 + 128.0 * bn) / bw);
        const double cost = waves * kb * t_kb + 150.0 * (bn / 32);
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

template <int BN, bool kOutHalf>
void launch_impl(const GemmOp& op, cudaStream_t stream) {
    using C = Cfg<BN>;
    auto kern = pf_gemm_f16_tn_tcgen05<BN, kOutHalf>;
    static bool attr_set = false;   // per instantiation; attribute is per-device but identical everywhere we run
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> g(mu);
        if (!attr_set) {
            int ndev = 0;
            PF_CUDA(cudaGetDeviceCount(&ndev));
            int cur = 0;
            PF_CUDA(cudaGetDevice(&cur));
            for (int d = 0; d < ndev; ++d) {
                PF_CUDA(cudaSetDevice(d));
                PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
            }
            PF_CUDA(cudaSetDevice(cur));
            attr_set = true;
        }
    }
    const int tiles_n = ceil_div(op.N, BN);
    const int num_tiles = tiles_n * ceil_div(op.M, BM);
    const int grid = std::min(num_tiles, num_sms());
    launch_k(kern, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, op.tmA, op.tmB, op.epi, op.M, op.N, op.K, tiles_n, num_tiles, op.vec_ok);
}

}  // namespace

void gemm_prepare(GemmOp& op, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  const GemmEpi& epi, int bn) {
    if ((epi.out_f32 != nullptr) == (epi.out_f16 != nullptr)) throw CudaError{"gemm: exactly one output pointer must be set"};
    if (M <= 0 || N <= 0 || K <= 0) throw CudaError{"gemm: empty problem"};
    if (bn == 0) bn = pick_bn(M, N, K);
    if (bn != 64 && bn != 128 && bn != 256) throw CudaError{"gemm: unsupported N tile"};
    op.M = M; op.N = N; op.K = K; op.bn = bn; op.epi = epi;
    // vector epilogue: 16-byte row segments of every fp32 tensor (8-byte for the fp16 output) must be aligned
    auto al = [](const void* p, int ld, int bytes) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % bytes) == 0 && ld % 4 == 0); };
    op.vec_ok = (al(epi.bias, 0, 16) && al(epi.resid, epi.ld_resid, 16) && al(epi.addend, epi.ld_addend, 16) &&
                 al(epi.out_f32, epi.ld_out, 16) && al(epi.out_f16, epi.ld_out, 8)) ? 1 : 0;
    make_tmap(&op.tmA, A, M, K, lda, BM);
    make_tmap(&op.tmB, W, N, K, ldw, bn);
}

void gemm_launch(const GemmOp& op, cudaStream_t stream) {
    const bool h = op.epi.out_f16 != nullptr;
    switch (op.bn) {
        case 64:  h ? launch_impl<64, true>(op, stream)  : launch_impl<64, false>(op, stream);  break;
        case 128: h ? launch_impl<128, true>(op, stream) : launch_impl<128, false>(op, stream); break;
        default:  h ? launch_impl<256, true>(op, stream) : launch_impl<256, false>(op, stream); break;
    }
}

double gemm_flops(const GemmOp& op) { return 2.0 * op.M * static_cast<double>(op.N) * op.K; }

}  // namespace pf
