// tcgen05 attention for sequences that fit one pass (Tk <= 192 key frames = 11.5 s of audio, Tq <= 256), with the
// SAN-M FSMN memory fused in:   O[b,:,h] = softmax(Q K^T / sqrt(128)) V        (fp16 operands, fp32 in TMEM)
//                               mem[b,t,c] = sum_j w[c,j] v[b,t+j-(k-1)/2,c] + v[b,t,c]   (fp32, zero padded)
// One CTA per (batch, head): K and V of the head are TMA-loaded once (3-D tensor maps clip at the utterance end and
// zero-fill), both 128-row query tiles run concurrently on two softmax warpgroups:
//   control warp : TMA loads, S = Q K^T (tcgen05.mma, K-major operands), O = P V (V is the MN-major B operand)
//   softmax group: tcgen05.ld row of S (thread = query row, so max / sum are thread-local), exp2, P -> smem as the
//                  fp16 A operand (128B-swizzled), then O / sum -> smem -> TMA store.
//   FSMN warps   : depthwise memory of this head from the V tile in shared memory, concurrently with the softmax
//                  (thread = channel x half of the time axis, 16 outputs per register block).
// Reference semantics: FunASR MultiHeadedAttentionSANM inside the graph run by OfflineProjOfParaformer.cs:68
// (SURVEY.md 2.5); key masks are all ones because the reference feeds speech_lengths = T (Q3).
#include "attention.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "gemm.cuh"

namespace pf {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int kMaxTk = 192;
constexpr int kGroups = 2;
constexpr int kFsmnWarps = 8;                        // dedicated FSMN warps: run while the softmax groups work
constexpr int kThreads = (kGroups * 4 + 1 + kFsmnWarps) * 32;   // 2 softmax warpgroups + control warp + FSMN warps
constexpr int kQBytes = BQ * HD * 2;                 // 32 KiB: two 64-column boxes of 128 rows

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttParams {
    int Tq, Tk, Tkp;          // Tkp = Tk rounded up to 16 (MMA N of S, K extent of P V)
    int tile_bytes;           // per query tile region: Q, later P, later the O staging
    int kv_bytes;             // 2 * Tkp * 128
    float scale_log2e;
    const float* fsmn_w;      // [H*128, taps] or null
    float* mem;               // [B*Tk, ld_mem] fp32
    int ld_mem, taps;
    int mem_accum;            // 1: mem += FSMN memory (the residual stream already holds x), 0: mem = FSMN memory
    uint32_t v_lbo, v_sbo;    // MN-major descriptor strides of the V tile (bytes)
};

__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous, 8 K-rows per 1024-B atom;
// LBO = byte distance between 64-element MN groups, SBO = between 8-row K groups.
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int b_mn_major) {
    return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 128B-swizzled K-major tile of 128-byte rows: byte offset of the 16-byte chunk `chunk` (0..7) of row r
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

__device__ __forceinline__ void tma_store_3d_f32(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1, int c2, bool reduce_add) {
    if (reduce_add)
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// mem[b, t, h*128 + c] (+)= sum_j w[c, j] v[t + j - (TAPS-1)/2, c] + v[t, c], v read from the 128B-swizzled V tile (two
// 64-channel boxes of Tkp rows).  Thread ft: channel ft % 128, time part ft / 128; a warp owns 32 channels and stages
// 16 time steps x 32 channels (2 KiB, two alternating boxes) that one lane hands to the TMA engine: a plain store, or
// an fp32 reduce-add into the residual stream (mem_accum) - no global loads, so no L2 latency on the FSMN warps' chain.
template <int TAPS>
__device__ __forceinline__ void fsmn_from_smem(const AttParams& p, const CUtensorMap* tmX, const uint8_t* sV, float* stage, int ft,
                                               int h, int b) {
    constexpr int TB = 16;                                         // outputs per register block = rows of a staging box
    constexpr int LEFT = (TAPS - 1) / 2;
    const int c = ft & 127, lane = ft & 31;
    const int parts = kFsmnWarps * 32 / 128;
    const int per = (((p.Tk + parts - 1) / parts) + TB - 1) / TB * TB;   // whole boxes per part: parts never share a box
    const int t_begin = (ft >> 7) * per;
    const int t_end = min(p.Tk, t_begin + per);
    float w[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) w[j] = __ldg(p.fsmn_w + (h * HD + c) * TAPS + j);
    const uint8_t* vb = sV + (c >> 6) * (p.Tkp * 128) + (c & 7) * 2;
    const int cchunk = (c & 63) >> 3;
    const uint32_t stage_u32 = smem_u32(stage);
    int blk = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TB, ++blk) {
        float x[TB + TAPS - 1];
#pragma unroll
        for (int i = 0; i < TB + TAPS - 1; ++i) {
            const int t = t0 - LEFT + i;
            x[i] = (t >= 0 && t < p.Tk) ? __half2float(*reinterpret_cast<const __half*>(vb + sw128_off(t, cchunk))) : 0.0f;
        }
        const int buf = blk & 1;
        if (blk >= 2) {                                            // the box is reused: its store from two blocks ago has read it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        float* box = stage + buf * (TB * 32);
#pragma unroll
        for (int i = 0; i < TB; ++i) {
            float acc = x[i + LEFT];
#pragma unroll
            for (int j = 0; j < TAPS; ++j) acc = fmaf(w[j], x[i + j], acc);
            box[i * 32 + lane] = acc;                              // rows past the utterance end are clipped by the map
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d_f32(tmX, stage_u32 + buf * (TB * 32 * 4), h * HD + (c & ~31), t0, b, p.mem_accum != 0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
pf_sanm_attention_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmX, const AttParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const uint32_t sK = base;
    const uint32_t sV = sK + p.kv_bytes;
    const uint32_t sT0 = sV + p.kv_bytes;                        // tile regions (Q -> P -> O staging)
    const uint32_t bar_off = 2 * p.kv_bytes + kGroups * p.tile_bytes;
    const uint32_t bars = base + bar_off;
    const uint32_t bar_k = bars, bar_v = bars + 8;
    auto bar_q = [&](int g) { return bars + 16 + 8 * g; };
    auto bar_s = [&](int g) { return bars + 32 + 8 * g; };       // S accumulator complete
    auto bar_p = [&](int g) { return bars + 48 + 8 * g; };       // P written by the 128 threads of the group
    auto bar_o = [&](int g) { return bars + 64 + 8 * g; };       // O accumulator complete
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 80);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int nq = (p.Tq + BQ - 1) / BQ;                          // 1 or 2 query tiles

    pdl_launch_dependents();
    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
            if (p.fsmn_w != nullptr) tma_prefetch_desc(&tmX);
            mbar_init(bar_k, 1);
            mbar_init(bar_v, 1);
            for (int g = 0; g < kGroups; ++g) {
                mbar_init(bar_q(g), 1);
                mbar_init(bar_s(g), 1);
                mbar_init(bar_p(g), 128);
                mbar_init(bar_o(g), 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 8) {
        // ------------------------------------------------ control: TMA loads + MMA issue
        if (lane == 0) {
            const int kb_bytes = p.Tkp * 128;                    // one 64-column box of K or V
            for (int g = 0; g < nq; ++g) {
                mbar_arrive_expect_tx(bar_q(g), kQBytes);
                tma_load_3d(sT0 + g * p.tile_bytes, &tmQ, bar_q(g), h * HD, g * BQ, b);
                tma_load_3d(sT0 + g * p.tile_bytes + kQBytes / 2, &tmQ, bar_q(g), h * HD + 64, g * BQ, b);
                if (g == 0) {
                    mbar_arrive_expect_tx(bar_k, p.kv_bytes);
                    tma_load_3d(sK, &tmK, bar_k, h * HD, 0, b);
                    tma_load_3d(sK + kb_bytes, &tmK, bar_k, h * HD + 64, 0, b);
                }
            }
            mbar_arrive_expect_tx(bar_v, p.kv_bytes);
            tma_load_3d(sV, &tmV, bar_v, h * HD, 0, b);
            tma_load_3d(sV + kb_bytes, &tmV, bar_v, h * HD + 64, 0, b);

            // S_g = Q_g K^T : M = 128, N = Tkp, K = 128 (two 64-wide swizzle atoms x four K=16 steps)
            mbar_wait(bar_k, 0);
            const uint32_t idesc_s = idesc_f16(BQ, p.Tkp, 0);
            for (int g = 0; g < nq; ++g) {
                mbar_wait(bar_q(g), 0);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + g * 256;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t adesc0 = desc_kmajor_sw128(sT0 + g * p.tile_bytes + kb * (kQBytes / 2));
                    const uint64_t bdesc0 = desc_kmajor_sw128(sK + kb * kb_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc0 + 2u * k, bdesc0 + 2u * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(bar_s(g));
            }
            // O_g = P_g V : M = 128, N = 128, K = Tkp; P is K-major (16 KiB per 64 keys), V is MN-major
            mbar_wait(bar_v, 0);
            const uint32_t idesc_o = idesc_f16(BQ, HD, 1);
            for (int g = 0; g < nq; ++g) {
                mbar_wait(bar_p(g), 0);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + g * 256;      // O aliases the first 128 columns of S_g
                const uint32_t sP = sT0 + g * p.tile_bytes;
                for (int j = 0; j < p.Tkp / 16; ++j) {
                    const uint64_t adesc = desc_kmajor_sw128(sP + (j >> 2) * 16384 + (j & 3) * 32);
                    const uint64_t bdesc = desc_mnmajor_sw128(sV + j * 2048, p.v_lbo, p.v_sbo);
                    umma_f16(d_tmem, adesc, bdesc, idesc_o, j != 0 ? 1u : 0u);
                }
                umma_commit(bar_o(g));
            }
        }
        __syncwarp();
    } else if (warp > 8) {
        // ------------------------------------------------ FSMN memory of this head from the V tile in shared memory
        if (p.fsmn_w != nullptr) {
            mbar_wait(bar_v, 0);
            const int ft = threadIdx.x - 9 * 32;                  // 0 .. kFsmnWarps*32-1
            float* stage = reinterpret_cast<float*>(smem + bar_off + 128) + (ft >> 5) * (2 * 16 * 32);   // 4 KiB per FSMN warp
            if (p.taps == 11) fsmn_from_smem<11>(p, &tmX, smem + p.kv_bytes, stage, ft, h, b);
            else fsmn_from_smem<21>(p, &tmX, smem + p.kv_bytes, stage, ft, h, b);
        }
    } else {
        // ------------------------------------------------ softmax / epilogue warpgroups
        const int g = warp >> 2;                                  // query tile of this warpgroup
        const int q = warp & 3;                                   // TMEM lane quadrant
        const int r = q * 32 + lane;                              // query row inside the tile = TMEM lane
        const int tid = threadIdx.x;                              // 0..255
        const bool active = g < nq;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
        uint8_t* tile = smem + 2 * p.kv_bytes + g * p.tile_bytes;
        float inv_sum = 0.0f;
        if (active) {
            mbar_wait(bar_s(g), 0);
            tc_fence_after_sync();
            // pass 1: row maximum over the valid keys (only the last chunk can hold padded keys)
            float mx = -INFINITY;
            for (int c0 = 0; c0 < p.Tkp; c0 += 32) {
                uint32_t v[32];
                const int n = min(32, p.Tkp - c0);                // 32 or 16 (Tkp % 16 == 0)
                if (n == 32) tmem_ld_32x32(trow + c0, v); else tmem_ld_32x32b_x16(trow + c0, v);
                tmem_ld_wait();
                if (c0 + 32 <= p.Tk) {
                    float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        m0 = fmaxf(m0, __uint_as_float(v[i]));
                        m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                        m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                        m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
                    }
                    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < n && c0 + i < p.Tk) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float mxs = mx * p.scale_log2e;
            // pass 2: p = exp2(s * scale - max * scale); fp32 row sum; fp16 P into the A-operand layout
            float sum = 0.0f;
            for (int c0 = 0; c0 < p.Tkp; c0 += 32) {
                uint32_t v[32];
                const int n = min(32, p.Tkp - c0);
                if (n == 32) tmem_ld_32x32(trow + c0, v); else tmem_ld_32x32b_x16(trow + c0, v);
                tmem_ld_wait();
                float e[32];
                if (c0 + 32 <= p.Tk) {
                    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        e[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2e, -mxs));
                        e[i + 1] = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, -mxs));
                        e[i + 2] = ex2_approx(fmaf(__uint_as_float(v[i + 2]), p.scale_log2e, -mxs));
                        e[i + 3] = ex2_approx(fmaf(__uint_as_float(v[i + 3]), p.scale_log2e, -mxs));
                        s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3];
                    }
                    sum += (s0 + s1) + (s2 + s3);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const bool ok = i < n && c0 + i < p.Tk;
                        e[i] = ok ? ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2e, -mxs)) : 0.0f;
                        sum += e[i];
                    }
                }
                uint8_t* kblk = tile + (c0 >> 6) * 16384;         // 64 keys per 16 KiB k-block
                const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j * 8 < n) {
                        uint4 pk;
                        __half2 h0 = __floats2half2_rn(e[8 * j], e[8 * j + 1]);
                        __half2 h1 = __floats2half2_rn(e[8 * j + 2], e[8 * j + 3]);
                        __half2 h2 = __floats2half2_rn(e[8 * j + 4], e[8 * j + 5]);
                        __half2 h3 = __floats2half2_rn(e[8 * j + 6], e[8 * j + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0);
                        pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2);
                        pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(kblk + sw128_off(r, chunk0 + j)) = pk;
                    }
                }
            }
            inv_sum = 1.0f / sum;
            fence_proxy_async();                                  // generic-proxy smem writes -> visible to tcgen05.mma
            tc_fence_before_sync();
            mbar_arrive(bar_p(g));
        }
        if (active) {
            mbar_wait(bar_o(g), 0);
            tc_fence_after_sync();
            // O row / sum -> fp16 -> swizzled staging (two 64-column boxes) -> TMA store (clipped at Tq by the map)
#pragma unroll 1
            for (int c0 = 0; c0 < HD; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(trow + c0, v);
                tmem_ld_wait();
                uint8_t* box = tile + (c0 >> 6) * (BQ * 128);
                const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 pk;
                    __half2 h0 = __floats2half2_rn(__uint_as_float(v[8 * j]) * inv_sum, __uint_as_float(v[8 * j + 1]) * inv_sum);
                    __half2 h1 = __floats2half2_rn(__uint_as_float(v[8 * j + 2]) * inv_sum, __uint_as_float(v[8 * j + 3]) * inv_sum);
                    __half2 h2 = __floats2half2_rn(__uint_as_float(v[8 * j + 4]) * inv_sum, __uint_as_float(v[8 * j + 5]) * inv_sum);
                    __half2 h3 = __floats2half2_rn(__uint_as_float(v[8 * j + 6]) * inv_sum, __uint_as_float(v[8 * j + 7]) * inv_sum);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2);
                    pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(box + sw128_off(r, chunk0 + j)) = pk;
                }
            }
            fence_proxy_async();
            named_bar_sync(1 + g, 128);
            if ((tid & 127) == 0) {
                const uint32_t st = sT0 + g * p.tile_bytes;
                tma_store_3d(&tmO, st, h * HD, g * BQ, b);
                tma_store_3d(&tmO, st + BQ * 128, h * HD + 64, g * BQ, b);
                tma_store_commit_wait();
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp16 [B, T, ld] viewed as (cols, t, b); box = 64 columns x box_rows x 1, 128B swizzle, OOB rows read as zero / are
// not written, so tiles never cross into the next utterance.
void make_tmap3(CUtensorMap* tm, const __half* ptr, int cols, int T, int B, int ld, int box_rows) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0)
        throw CudaError{"attention operand must be 16-byte aligned with a row pitch that is a multiple of 8 elements"};
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(T) * ld * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(tensormap_encode_fn())(
        tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled (attention) failed with CUresult " + std::to_string(static_cast<int>(r))};
}

// fp32 [B, T, ld] viewed as (cols, t, b); box = 32 columns x 16 rows x 1, no swizzle: the FSMN staging boxes
void make_tmap3_f32(CUtensorMap* tm, const float* ptr, int cols, int T, int B, int ld) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 4) != 0)
        throw CudaError{"attention: FSMN memory must be 16-byte aligned with a row pitch that is a multiple of 4 floats"};
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(T) * ld * 4};
    cuuint32_t box[3] = {32, 16, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(tensormap_encode_fn())(
        tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled (FSMN memory) failed with CUresult " + std::to_string(static_cast<int>(r))};
}

}  // namespace

bool attention_tc_eligible(int Tq, int Tk, int head_dim) {
    static const bool off = [] { const char* e = getenv("PFASR_NO_ATT_TC"); return e && *e && *e != '0'; }();
    return !off && head_dim == HD && Tk >= 1 && Tk <= kMaxTk && Tq >= 1 && Tq <= kGroups * BQ;
}

void attention_tc_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int Tq, int Tk, int ldq,
                         int ldk, int ldv, int ldo, const float* fsmn_w, int taps, float* mem, int ld_mem, bool mem_accum, cudaStream_t s) {
    if (fsmn_w && (taps != 11 && taps != 21)) throw CudaError{"attention: FSMN kernel must be 11 or 21"};
    if (fsmn_w && Tq != Tk) throw CudaError{"attention: fused FSMN needs self-attention (Tq == Tk)"};
    AttParams p{};
    p.Tq = Tq; p.Tk = Tk; p.Tkp = (Tk + 15) & ~15;
    p.kv_bytes = 2 * p.Tkp * 128;
    p.tile_bytes = std::max(kQBytes, ((p.Tkp + 63) / 64) * 16384);
    p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    p.fsmn_w = fsmn_w; p.mem = mem; p.ld_mem = ld_mem; p.taps = taps; p.mem_accum = mem_accum ? 1 : 0;
    p.v_lbo = static_cast<uint32_t>(p.Tkp * 128);
    p.v_sbo = 1024;
    if (const char* e = getenv("PFASR_ATT_VDESC")) {             // experiment hook: "lbo,sbo" in bytes
        unsigned a = 0, b2 = 0;
        if (sscanf(e, "%u,%u", &a, &b2) == 2) { p.v_lbo = a; p.v_sbo = b2; }
    }
    CUtensorMap tq, tk, tv, to, tx{};
    if (fsmn_w) make_tmap3_f32(&tx, mem, H * HD, Tk, B, ld_mem);
    make_tmap3(&tq, Q, H * HD, Tq, B, ldq, BQ);
    make_tmap3(&tk, K, H * HD, Tk, B, ldk, p.Tkp);
    make_tmap3(&tv, V, H * HD, Tk, B, ldv, p.Tkp);
    make_tmap3(&to, O, H * HD, Tq, B, ldo, BQ);
    const int smem_bytes = 2 * p.kv_bytes + kGroups * p.tile_bytes + 128 + kFsmnWarps * 4096 + 1024;   // + FSMN staging boxes
    static std::once_flag once;
    std::call_once(once, [] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_sanm_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    launch_k(pf_sanm_attention_tc, dim3(H, B), dim3(kThreads), static_cast<size_t>(smem_bytes), s, tq, tk, tv, to, tx, p);
}

}  // namespace pf
