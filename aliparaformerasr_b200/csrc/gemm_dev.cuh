// Device-side building blocks shared by the tcgen05 GEMM kernels (gemm.cu, ffn_chain.cu): tile constants, UMMA
// descriptors, and the asynchronous TMA-store epilogues.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace pf {
namespace gemm_dev {

constexpr int BM = 128;
constexpr int BK = 64;                  // 64 fp16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int kABytes = BM * BK * 2;    // 16 KiB
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kStageTileBytes = 32 * 32 * 4;   // per-warp transpose buffer: 32 rows x 32 fp32

constexpr int kSmemMax = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100


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

// ---- asynchronous (TMA) epilogue --------------------------------------------------------------------------------
// The register -> smem transpose + per-thread global stores above cost ~6000 cycles per 128 x 256 tile (measured
// with scripts/gemm_probe.py: the epilogue alone ran longer than the MMAs of a K = 512 tile) and kept the accumulator
// stage busy.  Here every warp writes its TMEM rows (thread = row) straight into a 128B-swizzled staging box and one
// lane hands the box to the TMA store engine; the fp32 residual is TMA-loaded into the same box ahead of time and
// updated in place.  No ld.shared of other threads' data, no per-thread global traffic, edges clipped by the maps.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// fp32 tile += staging box, performed by the L2 reduction units (cp.reduce.async.bulk): the in-place residual update
// x += A W^T + b needs no residual load at all
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
template <int kWarps>
__device__ __forceinline__ void epi_bar_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory"); }
__device__ __forceinline__ void epi_bar_sync() { epi_bar_sync_n<kEpiWarps>(); }

// fp16 output: this warp's rows [row0, row0+32) x columns [colw, colw + 32*nchunks) of the tile.  One 32 x 32 box
// (64-byte rows, SWIZZLE_64B) per chunk, two alternating 2 KiB staging boxes, so a box is rewritten two chunks after
// its store was issued.
// wide: stg holds one box per chunk (the CTA's last tile borrows the idle operand ring), so no box is ever reused and
// the chunks never wait for the store engine.
template <int kMaxChunks>
__device__ __forceinline__ void epilogue_tma_f16(uint32_t t_acc, int nchunks, uint8_t* stg, const CUtensorMap* tmC,
                                                 const float* bias_w, float lo, int row0, int colw, int lane, int dbg = 0, bool wide = false) {
    const uint32_t stg_u32 = smem_u32(stg);
    uint32_t ra[32], rb[32];
    tmem_ld_issue(t_acc, ra);
    if (lane == 0) tma_store_wait_read();                          // the previous tile's stores have read both boxes
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
        if (k < nchunks) {
            uint32_t (&cur)[32] = (k & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
            const int buf = wide ? k : (k & 1);
            tmem_ld_wait();
            if (k + 1 < nchunks) tmem_ld_issue(t_acc + static_cast<uint32_t>((k + 1) * 32), nxt);
            if (k >= 2 && !wide) {                                // box reuse: at most the previous chunk's store may be pending
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            uint8_t* box = stg + buf * 2048 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias_w + k * 32 + 8 * j);       // smem broadcast
                const float4 b1 = *reinterpret_cast<const float4*>(bias_w + k * 32 + 8 * j + 4);
                __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(cur[8 * j]) + b0.x, lo), fmaxf(__uint_as_float(cur[8 * j + 1]) + b0.y, lo));
                __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(cur[8 * j + 2]) + b0.z, lo), fmaxf(__uint_as_float(cur[8 * j + 3]) + b0.w, lo));
                __half2 h2 = __floats2half2_rn(fmaxf(__uint_as_float(cur[8 * j + 4]) + b1.x, lo), fmaxf(__uint_as_float(cur[8 * j + 5]) + b1.y, lo));
                __half2 h3 = __floats2half2_rn(fmaxf(__uint_as_float(cur[8 * j + 6]) + b1.z, lo), fmaxf(__uint_as_float(cur[8 * j + 7]) + b1.w, lo));
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                if (!(dbg & 2) || pk.x == 0x7fc07fc1u)                                     // probe: no smem traffic
                    *reinterpret_cast<uint4*>(box + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;      // SWIZZLE_64B: chunk ^= (row / 2) % 4
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(dbg & 3)) {
                tma_store_2d(tmC, stg_u32 + buf * 2048, colw + k * 32, row0);
                tma_store_commit();
            }
        }
    }
}

// Fused greedy pick (head GEMMs): this warp's rows x columns [colw, colw + 32*nchunks) are scanned in ascending column
// order, thread = row, and reduced to one partial per row: the last NaN column (or -1) and the last maximum among the
// columns after it ("v >= best" keeps the LAST index attaining the maximum: OfflineRecognizer.cs:145-149 walks k upwards
// with best = x[best] > x[k] ? best : k, and a NaN makes every comparison false, restarting the scan there).
template <int kMaxChunks>
__device__ __forceinline__ void epilogue_pick_f32(uint32_t t_acc, int nchunks, const float* bias_w, int colw, int N, float* dst) {
    uint32_t ra[32], rb[32];
    float bv = -INFINITY;
    int bi = -1, ln = -1;
    if (nchunks > 0) tmem_ld_issue(t_acc, ra);
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
        if (k < nchunks) {
            uint32_t (&cur)[32] = (k & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
            tmem_ld_wait();
            if (k + 1 < nchunks) tmem_ld_issue(t_acc + static_cast<uint32_t>((k + 1) * 32), nxt);
            const int col = colw + k * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = __uint_as_float(cur[j]) + bias_w[k * 32 + j];
                const int c = col + j;
                if (c < N) {
                    if (x != x) { ln = c; bv = -INFINITY; bi = -1; }
                    else if (x >= bv) { bv = x; bi = c; }
                }
            }
        }
    }
    if (dst != nullptr) {
        dst[0] = bv;
        dst[1] = __int_as_float(bi);
        dst[2] = __int_as_float(ln);
    }
}

// fp32 output (+ optional fp32 residual, TMA-prefetched into the staging boxes): 32 columns per box, two boxes.
// rbar: the two "residual landed" mbarriers of this warp; rcount: loads issued so far per buffer (phase tracking).
struct ResidPipe {
    uint32_t bar[2];
    uint32_t count[2];
};
template <bool kResid>
__device__ __forceinline__ void resid_issue(ResidPipe& rp, int buf, uint32_t stg_u32, const CUtensorMap* tmR, int col, int row0, int lane) {
    if (kResid && lane == 0) {
        mbar_arrive_expect_tx(rp.bar[buf], 4096);
        tma_load_2d(stg_u32 + buf * 4096, tmR, rp.bar[buf], col, row0);
    }
}
// kBoxes = 1 (no residual load only): a single staging box, so every chunk waits until the previous store has read it.
template <int kMaxChunks, bool kResid, bool kLn = false, int kBoxes = 2>
__device__ __forceinline__ void epilogue_tma_f32(uint32_t t_acc, int nchunks, uint8_t* stg, ResidPipe& rp, const CUtensorMap* tmC,
                                                 const CUtensorMap* tmR, const float* bias_w, float lo, int row0,
                                                 int colw, int lane, int prefetched, float* ln_s1 = nullptr, float* ln_s2 = nullptr,
                                                 bool red_add = false, bool wide = false) {
    static_assert(kBoxes == 2 || (kBoxes == 1 && !kResid && !kLn), "the single-box epilogue has no residual pipeline");
    float s1 = 0.0f, s2 = 0.0f;                                    // kLn: this row's sum / sum of squares over the warp's columns
    const uint32_t stg_u32 = smem_u32(stg);
    uint32_t ra[32], rb[32];
    tmem_ld_issue(t_acc, ra);
    if (!kResid) {                                                 // (with a residual the caller waited before prefetching)
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
    }
    for (int k = prefetched; k < 2 && k < nchunks; ++k) resid_issue<kResid>(rp, k, stg_u32, tmR, colw + k * 32, row0, lane);
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
        if (k < nchunks) {
            uint32_t (&cur)[32] = (k & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
            const int buf = (!kResid && wide) ? k : (kBoxes == 2 ? (k & 1) : 0);
            tmem_ld_wait();
            if (k + 1 < nchunks) tmem_ld_issue(t_acc + static_cast<uint32_t>((k + 1) * 32), nxt);
            if (kResid) {
                mbar_wait(rp.bar[buf], rp.count[buf] & 1u);
                rp.count[buf]++;
            } else if (wide) {                                     // one box per chunk (the idle operand ring): nothing to wait for
            } else if (kBoxes == 1) {
                if (k >= 1) {                                      // the one box: the previous chunk's store must have read it
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
            } else if (k >= 2) {                                   // buffer reuse without a residual load in between:
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // only the previous chunk's store may be pending
                __syncwarp();
            }
            uint8_t* box = stg + buf * 4096 + lane * 128;
            const int col = colw + k * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4* slot = reinterpret_cast<float4*>(box + ((j ^ (lane & 7)) << 4));
                float4 v = make_float4(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]),
                                       __uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3]));
                const float4 b = *reinterpret_cast<const float4*>(bias_w + k * 32 + 4 * j);            // smem broadcast
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                if (kResid) {
                    const float4 x = *slot;
                    v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
                }
                v.x = fmaxf(v.x, lo); v.y = fmaxf(v.y, lo); v.z = fmaxf(v.z, lo); v.w = fmaxf(v.w, lo);
                *slot = v;
                if (kLn) {                                         // keep the final values for the normalisation pass
                    s1 += (v.x + v.y) + (v.z + v.w);
                    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                    cur[4 * j] = __float_as_uint(v.x); cur[4 * j + 1] = __float_as_uint(v.y);
                    cur[4 * j + 2] = __float_as_uint(v.z); cur[4 * j + 3] = __float_as_uint(v.w);
                }
            }
            if (kLn) tmem_st_32x32(t_acc + static_cast<uint32_t>(k * 32), cur);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (!kResid && red_add) tma_reduce_add_2d(tmC, stg_u32 + buf * 4096, col, row0);
                else tma_store_2d(tmC, stg_u32 + buf * 4096, col, row0);
                tma_store_commit();
                if (kResid && k + 2 < nchunks) tma_store_wait_read();     // box k is read before chunk k+2 lands in it
            }
            if (kResid && k + 2 < nchunks) {
                __syncwarp();
                resid_issue<kResid>(rp, buf, stg_u32, tmR, colw + (k + 2) * 32, row0, lane);
            }
        }
    }
    if (kLn) { *ln_s1 = s1; *ln_s2 = s2; }
}

}  // namespace gemm_dev
}  // namespace pf
