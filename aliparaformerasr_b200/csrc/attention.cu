// Multi-head scaled-dot-product attention for SAN-M self-attention (encoder) and cross-attention (decoder):
//   O[b, :, h] = softmax(Q[b,:,h] K[b,:,h]^T / sqrt(128)) V[b,:,h],  head dim 128, fp16 operands, fp32 softmax state.
// Single pass over K/V in 64-row chunks with an online (running max / running sum) softmax reduced with warp
// shuffles; QK^T and PV run on mma.sync m16n8k16 (the north-star keeps tcgen05 for the projection GEMMs).
// Masks: the reference feeds speech_lengths = T for every item (OfflineProjOfParaformer.cs:53-63, Q3), so key
// masks are all ones; only the chunk tail (kv >= Tk) is masked here.
#include "attention.cuh"

#include <mutex>

#include "ops.cuh"
#include "gemm.cuh"

#include <stdlib.h>

#include <algorithm>

namespace pf {

namespace {

constexpr int HD = 128;
constexpr int BQ = 64;
constexpr int BKV = 64;
constexpr int LDS = HD + 8;            // padded row: 272 B, conflict-free for ldmatrix
constexpr int kThreads = 128;
constexpr int kMaxSplits = 4;
constexpr int kSmemBytes = (BQ + 4 * BKV) * LDS * 2;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Stage `rows` rows x 128 halfs from global (row pitch ld) into padded smem; rows >= valid are zero-filled.
__device__ __forceinline__ void stage_tile(uint32_t smem_base, const __half* g, int ld, int valid_rows, int rows) {
    for (int i = threadIdx.x; i < rows * (HD / 8); i += kThreads) {
        const int r = i >> 4;
        const int c = i & 15;
        const bool ok = r < valid_rows;
        const __half* src = g + static_cast<size_t>(ok ? r : 0) * ld + c * 8;
        cp_async_16(smem_base + (r * LDS + c * 8) * 2, src, ok);
    }
}

// splits > 1 (few query tiles against a long memory, e.g. 800 decoder rows x 2010 SeACo hot-word rows = 52 CTAs): the key range is cut
// into `splits` runs of whole chunks, CTA (q tile, run) leaves its UNNORMALISED fp32 output and (running max, running sum) per row in
// part_o / part_ml, and pf_sanm_attention_combine merges the runs - the usual flash-decoding split.
__global__ void __launch_bounds__(kThreads)
pf_sanm_attention(const __half* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V,
                  __half* __restrict__ O, int Tq, int Tk, int ldq, int ldk, int ldv, int ldo, float scale_log2e,
                  int splits, int chunks_per_split, float* __restrict__ part_o, float* __restrict__ part_ml) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + BQ * LDS * 2;
    const uint32_t sV = sK + 2 * BKV * LDS * 2;

    const int b = blockIdx.z, h = blockIdx.y;
    const int split = blockIdx.x % splits;
    const int q0 = (blockIdx.x / splits) * BQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;

    const __half* Qg = Q + (static_cast<size_t>(b) * Tq + q0) * ldq + h * HD;
    const __half* Kg = K + static_cast<size_t>(b) * Tk * ldk + h * HD;
    const __half* Vg = V + static_cast<size_t>(b) * Tk * ldv + h * HD;

    const int nchunks_all = (Tk + BKV - 1) / BKV;
    const int j_begin = split * chunks_per_split;
    const int nchunks = min(nchunks_all, j_begin + chunks_per_split);      // this CTA walks chunks [j_begin, nchunks)
    stage_tile(sQ, Qg, ldq, min(BQ, Tq - q0), BQ);
    stage_tile(sK, Kg + static_cast<size_t>(j_begin) * BKV * ldk, ldk, min(BKV, Tk - j_begin * BKV), BKV);
    stage_tile(sV, Vg + static_cast<size_t>(j_begin) * BKV * ldv, ldv, min(BKV, Tk - j_begin * BKV), BKV);
    cp_async_commit();

    uint32_t qf[HD / 16][4];
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

    for (int j = j_begin; j < nchunks; ++j) {
        const int buf = (j - j_begin) & 1;
        if (j + 1 < nchunks) {
            const int kv1 = (j + 1) * BKV;
            stage_tile(sK + (buf ^ 1) * BKV * LDS * 2, Kg + static_cast<size_t>(kv1) * ldk, ldk, min(BKV, Tk - kv1), BKV);
            stage_tile(sV + (buf ^ 1) * BKV * LDS * 2, Vg + static_cast<size_t>(kv1) * ldv, ldv, min(BKV, Tk - kv1), BKV);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == j_begin) {
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) {
                const uint32_t addr = sQ + ((warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8) * 2;
                ldmatrix_x4(addr, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
        }
        const uint32_t kb = sK + buf * BKV * LDS * 2;
        const uint32_t vb = sV + buf * BKV * LDS * 2;

        // S = Q K^T for this chunk: 16 x 64 per warp
        float s[BKV / 8][4];
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f; }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < BKV / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = kb + ((np * 16 + ((lane >> 4) & 1) * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8) * 2;
                ldmatrix_x4(addr, b0, b1, b2, b3);
                mma_16816(s[2 * np], qf[kk], b0, b1);
                mma_16816(s[2 * np + 1], qf[kk], b2, b3);
            }
        }
        // scale into the log2 domain, mask the chunk tail, running max
        const int kv0 = j * BKV;
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < BKV / 8; ++nt) {
            const int col = kv0 + nt * 8 + 2 * t4;
            const bool v0 = col < Tk, v1 = col + 1 < Tk;
            s[nt][0] = v0 ? s[nt][0] * scale_log2e : -INFINITY;
            s[nt][1] = v1 ? s[nt][1] * scale_log2e : -INFINITY;
            s[nt][2] = v0 ? s[nt][2] * scale_log2e : -INFINITY;
            s[nt][3] = v1 ? s[nt][3] * scale_log2e : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);   // first chunk: exp2(-inf) = 0
        m0 = mx0; m1 = mx1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
        for (int nt = 0; nt < BKV / 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - m0);
            s[nt][1] = exp2f(s[nt][1] - m0);
            s[nt][2] = exp2f(s[nt][2] - m1);
            s[nt][3] = exp2f(s[nt][3] - m1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = vb + ((kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + dp * 16 + ((lane >> 4) & 1) * 8) * 2;
                ldmatrix_x4_trans(addr, b0, b1, b2, b3);
                mma_16816(o[2 * dp], pa, b0, b1);
                mma_16816(o[2 * dp + 1], pa, b2, b3);
            }
        }
        __syncthreads();   // everyone done with this K/V buffer before it is refilled
    }

    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int r0 = q0 + warp * 16 + g;
    const int r1 = r0 + 8;
    if (splits > 1) {
        const size_t base = ((static_cast<size_t>(split) * gridDim.z + b) * gridDim.y + h) * Tq;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            const int col = nt * 8 + 2 * t4;
            if (r0 < Tq) *reinterpret_cast<float2*>(part_o + (base + r0) * HD + col) = make_float2(o[nt][0], o[nt][1]);
            if (r1 < Tq) *reinterpret_cast<float2*>(part_o + (base + r1) * HD + col) = make_float2(o[nt][2], o[nt][3]);
        }
        if (t4 == 0) {
            if (r0 < Tq) *reinterpret_cast<float2*>(part_ml + (base + r0) * 2) = make_float2(m0, l0);
            if (r1 < Tq) *reinterpret_cast<float2*>(part_ml + (base + r1) * 2) = make_float2(m1, l1);
        }
        return;
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __half* Og = O + static_cast<size_t>(b) * Tq * ldo + h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
        const int col = nt * 8 + 2 * t4;
        if (r0 < Tq) *reinterpret_cast<uint32_t*>(Og + static_cast<size_t>(r0) * ldo + col) = pack_half2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < Tq) *reinterpret_cast<uint32_t*>(Og + static_cast<size_t>(r1) * ldo + col) = pack_half2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

// One warp per (b, h, query row): merge the runs' (max, sum, unnormalised output) and write the fp16 context row.
__global__ void __launch_bounds__(256)
pf_sanm_attention_combine(const float* __restrict__ part_o, const float* __restrict__ part_ml, int splits, int rows_total, int H, int Tq,
                          __half* __restrict__ O, int ldo) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // row index over (b, h, t)
    if (row >= rows_total) return;
    const int t = row % Tq, h = (row / Tq) % H, b = row / (Tq * H);
    float M = -INFINITY;
    for (int s = 0; s < splits; ++s) M = fmaxf(M, part_ml[(static_cast<size_t>(s) * rows_total + row) * 2]);
    float L = 0.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float2 ml = *reinterpret_cast<const float2*>(part_ml + (static_cast<size_t>(s) * rows_total + row) * 2);
        const float w = exp2f(ml.x - M);
        L += w * ml.y;
        const float4 v = *reinterpret_cast<const float4*>(part_o + (static_cast<size_t>(s) * rows_total + row) * HD + lane * 4);
        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    const float inv = 1.0f / L;
    uint2 pk;
    pk.x = pack_half2(acc.x * inv, acc.y * inv);
    pk.y = pack_half2(acc.z * inv, acc.w * inv);
    *reinterpret_cast<uint2*>(O + (static_cast<size_t>(b) * Tq + t) * ldo + h * HD + lane * 4) = pk;
}

}  // namespace

size_t attention_split_workspace_bytes(int B, int H, int Tq) {
    return static_cast<size_t>(kMaxSplits) * B * H * Tq * (HD + 2) * sizeof(float);
}

void attention_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int Tq, int Tk, int ldq,
                      int ldk, int ldv, int ldo, int head_dim, cudaStream_t s, float* split_ws, size_t split_ws_bytes) {
    if (head_dim != HD) throw CudaError{"attention: head dim must be 128"};
    if (Tq <= 0 || Tk <= 0 || B <= 0) return;
    // Up to 64 query rows (the decoder's cross-attention: one row per token) go to the streaming mma.sync kernel below: one 64-row
    // query tile per CTA, 87 KB of shared memory and 128 threads, so two CTAs share an SM and leave room for other lanes' kernels,
    // where the tcgen05 kernel holds a whole SM per (b, h) for a quarter-filled 128-row tile (0.239 -> 0.205 ms per 16 layers,
    // three-lane step 3.69 -> 3.65 ms; PFASR_XATT_SMALLQ=0 restores the tcgen05 path for every length it covers)
    static const int small_q = [] { const char* e = getenv("PFASR_XATT_SMALLQ"); return e ? atoi(e) : 64; }();
    if (attention_tc_eligible(Tq, Tk, head_dim) && !(small_q > 0 && Tq <= small_q)) {
        attention_tc_launch(Q, K, V, O, B, H, Tq, Tk, ldq, ldk, ldv, ldo, nullptr, 0, nullptr, 0, false, s);
        return;
    }
    static std::once_flag attr_once;                              // execution lanes call this from several host threads
    std::call_once(attr_once, [&] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_sanm_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    const int qtiles = ceil_div(Tq, BQ), nchunks = ceil_div(Tk, BKV);
    // few CTAs against a long memory: cut the keys into runs so that the grid fills the chip (needs the caller's workspace)
    int splits = 1;
    static const bool no_split = [] { const char* e = getenv("PFASR_ATT_NO_SPLIT"); return e && *e && *e != '0'; }();
    if (!no_split && split_ws != nullptr && qtiles * H * B * 2 <= gemm_num_sms() && nchunks >= 8) {
        splits = std::min(kMaxSplits, std::max(1, gemm_num_sms() / (qtiles * H * B)));
        splits = std::min(splits, nchunks / 4);
        if (attention_split_workspace_bytes(B, H, Tq) > split_ws_bytes) splits = 1;
    }
    int per = ceil_div(nchunks, splits);
    splits = ceil_div(nchunks, per);                                      // every run holds at least one chunk
    float* part_o = split_ws;
    float* part_ml = split_ws ? split_ws + static_cast<size_t>(kMaxSplits) * B * H * Tq * HD : nullptr;
    dim3 grid(qtiles * splits, H, B);
    launch_k(pf_sanm_attention, grid, dim3(kThreads), kSmemBytes, s, Q, K, V, O, Tq, Tk, ldq, ldk, ldv, ldo, scale_log2e, splits, per, part_o, part_ml);
    if (splits > 1) {
        const int rows_total = B * H * Tq;
        launch_k(pf_sanm_attention_combine, dim3(ceil_div(rows_total, 8)), dim3(256), 0, s, static_cast<const float*>(part_o),
                 static_cast<const float*>(part_ml), splits, rows_total, H, Tq, O, ldo);
    }
}

int attention_fsmn_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int T, int ldqkv, int ldo,
                          const float* fsmn_w, int taps, float* mem, int ld_mem, bool mem_accum, cudaStream_t s) {
    if (T <= 0 || B <= 0) return 0;
    if (attention_tc_eligible(T, T, HD)) {
        attention_tc_launch(Q, K, V, O, B, H, T, T, ldqkv, ldqkv, ldqkv, ldo, fsmn_w, taps, mem, ld_mem, mem_accum, s);
        return 1;
    }
    fsmn_f16_launch(V, ldqkv, fsmn_w, taps, mem, ld_mem, mem_accum ? mem : nullptr, ld_mem, nullptr, B, T, H * HD, s);
    attention_launch(Q, K, V, O, B, H, T, T, ldqkv, ldqkv, ldqkv, ldo, HD, s);
    return 2;
}

}  // namespace pf
