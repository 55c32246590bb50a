// Hand-written sm_100a GEMM: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
// tcgen05.mma (single issuing thread, fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
// Replaces the MLAS GEMMs that OnnxRuntime runs for the reference at
// /root/reference/AliParaformerAsr/OfflineProjOfParaformer.cs:68 (InferenceSession.Run).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (each owns the 32 TMEM lanes of its quadrant = warp_idx % 4).
#include "gemm.cuh"

#include <mutex>

namespace pf {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                  // 64 fp16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int kABytes = BM * BK * 2;    // 16 KiB

template <int BN>
struct Cfg {
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarBytes = 256;
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment slack
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

template <bool kOutHalf>
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[32], int row, int col0, int N, const GemmEpi& e) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    const bool full = (col0 + 32 <= N);
    if (full) {
        if (e.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = __ldg(b4 + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
        }
        if (e.addend) {
            const float4* a4 = reinterpret_cast<const float4*>(e.addend + static_cast<size_t>(row) * e.ld_addend + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 a = a4[j];
                v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
            }
        }
        if (e.resid) {
            const float4* r4 = reinterpret_cast<const float4*>(e.resid + static_cast<size_t>(row) * e.ld_resid + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 a = r4[j];
                v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
            }
        }
        if (e.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (kOutHalf) {
            uint4* o = reinterpret_cast<uint4*>(e.out_f16 + static_cast<size_t>(row) * e.ld_out + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]);
                __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
                __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                o[j] = pk;
            }
        } else {
            float4* o = reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(row) * e.ld_out + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    } else {
        // ragged last N tile (e.g. vocab 8404): scalar, bounds-checked
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < N) {
                float x = v[j];
                if (e.bias) x += __ldg(e.bias + col);
                if (e.addend) x += e.addend[static_cast<size_t>(row) * e.ld_addend + col];
                if (e.resid) x += e.resid[static_cast<size_t>(row) * e.ld_resid + col];
                if (e.relu) x = fmaxf(x, 0.0f);
                if (kOutHalf) e.out_f16[static_cast<size_t>(row) * e.ld_out + col] = __float2half_rn(x);
                else e.out_f32[static_cast<size_t>(row) * e.ld_out + col] = x;
            }
        }
    }
}

template <int BN, bool kOutHalf>
__global__ void __launch_bounds__(192, 1)
pf_gemm_f16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmEpi epi, const int M, const int N, const int K) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    const uint32_t bar_base = base + STAGES * C::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + STAGES * C::kStageBytes + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), BN);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), C::kStageBytes);
                const uint32_t a_s = base + s * C::kStageBytes;
                const uint32_t b_s = a_s + kABytes;
                tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                tma_load_2d(b_s, &tmB, full_bar(s), kb * BK, n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after_sync();
                const uint32_t a_s = base + s * C::kStageBytes;
                const uint32_t b_s = a_s + kABytes;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adesc = make_sw128_kmajor_desc(a_s + k * UMMA_K * 2);
                    const uint64_t bdesc = make_sw128_kmajor_desc(b_s + k * UMMA_K * 2);
                    umma_f16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty_bar(s));      // frees this smem stage once the MMAs above have read it
            }
            umma_commit(tmem_full_bar);         // accumulator complete
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue: TMEM -> registers -> global
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after_sync();
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= N) break;               // warp-uniform
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
            tmem_ld_wait();
            if (row < M) epilogue_store<kOutHalf>(r, row, col0, N, epi);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
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
    dim3 grid(ceil_div(op.N, BN), ceil_div(op.M, BM));
    kern<<<grid, 192, C::kSmemBytes, stream>>>(op.tmA, op.tmB, op.epi, op.M, op.N, op.K);
    PF_CUDA(cudaGetLastError());
}

}  // namespace

void gemm_prepare(GemmOp& op, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  const GemmEpi& epi, int bn) {
    if ((epi.out_f32 != nullptr) == (epi.out_f16 != nullptr)) throw CudaError{"gemm: exactly one output pointer must be set"};
    if (M <= 0 || N <= 0 || K <= 0) throw CudaError{"gemm: empty problem"};
    if (bn == 0) {
        const int mt = ceil_div(M, BM);
        if (mt * ceil_div(N, 256) >= 132) bn = 256;
        else if (mt * ceil_div(N, 128) >= 100) bn = 128;
        else bn = 64;
    }
    if (bn != 64 && bn != 128 && bn != 256) throw CudaError{"gemm: unsupported N tile"};
    op.M = M; op.N = N; op.K = K; op.bn = bn; op.epi = epi;
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
