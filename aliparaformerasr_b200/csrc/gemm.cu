// Hand-written sm_100a GEMM: persistent CTAs, TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
// tcgen05.mma (single issuing thread, fp32 accumulators double-buffered in TMEM) -> asynchronous epilogue: tcgen05.ld
// rows -> bias / ReLU / residual in registers -> swizzled staging boxes -> TMA store (the fp32 residual is TMA-prefetched
// into the same boxes).  A smem-transposing epilogue with per-thread global accesses remains for the cases the TMA path
// does not cover (two fp32 addends, unaligned pitches).  Optional variants: CTA-pair MMA (cta_group::2) and a fused
// LayerNorm second pass across the CTAs of a cluster - both parity-green, both measured slower on this path.
//
// Replaces the MLAS GEMMs that OnnxRuntime runs for the reference at
// /root/reference/AliParaformerAsr/OfflineProjOfParaformer.cs:68 (InferenceSession.Run).
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4..11 = epilogue (two warpgroups; a warp owns the 32 TMEM lanes of quadrant warp % 4 and every second
// 32-column chunk).  The epilogue of tile i overlaps the main loop of tile i+1 (two TMEM accumulator stages).
#include "gemm.cuh"
#include "gemm_dev.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>

namespace pf {

using namespace gemm_dev;

namespace {

// BNMAX = widest N tile the instantiation can hold (128 or 256); the actual width bn <= BNMAX (a multiple of 32) is a
// launch parameter, so one instantiation serves every width that shares its shared-memory / TMEM layout.
constexpr int kLnMaxCluster = 4;       // column tiles (CTAs of one cluster) a fused-LayerNorm row may span

// kBoxes = staging boxes per epilogue warp: two (a box is rewritten two chunks after its store was issued) or, for the
// 256-wide fp32 tiles without a residual load, one - the 32 KiB that frees are the fourth operand stage.
template <int BNMAX, bool kPair, bool kOutHalf, bool kLn = false, int kEW = kEpiWarps, int kBoxes = 2>
struct Cfg {
    static constexpr int kThreadsCta = kEpiWarp0 * 32 + kEW * 32;    // producer, MMA, TMEM, spare warp + kEW epilogue warps
    static constexpr int kBRows = kPair ? BNMAX / 2 : BNMAX;        // W-tile rows a stage slot can hold
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // per-warp staging: fp16 output = two 32 x 32 boxes (2 KiB each); fp32 output = two 32 x 32 boxes (4 KiB each)
    static constexpr int kWarpStageBytes = kOutHalf ? 4096 : 4096 * kBoxes;
    static constexpr int kEpiBytes = kEW * kWarpStageBytes;
    static constexpr int kBarBytes = 512;
    static constexpr int kBiasBytes = 2 * BNMAX * 4;                 // two tiles' bias columns
    // fused LayerNorm: gamma | beta tiles and the per-row partial sums of every CTA of the cluster
    static constexpr int kLnBytes = kLn ? 2 * BNMAX * 4 + kLnMaxCluster * 2 * BM * 2 * 4 : 0;
    // the dynamic shared memory base is 1024-byte aligned (declared so, checked in the kernel): no alignment slack, which
    // is what lets the 256-wide fp16-output tiles keep FOUR 48 KiB operand stages instead of three
    static constexpr int kFit = (kSmemMax - kBarBytes - kBiasBytes - kLnBytes - kEpiBytes) / kStageBytes;
    static constexpr int kStages = kFit > 8 ? 8 : kFit;
    static constexpr int kTmemCols = 2 * BNMAX;                      // two accumulator stages (256 / 512 columns)
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + kBiasBytes + kLnBytes;
    static_assert(kSmemBytes <= kSmemMax, "shared memory budget");
    static_assert(kStages >= 3, "operand ring too shallow");
};

template <int BNMAX, bool kPair, bool kOutHalf, int kAdds, bool kLn, int kEW>
constexpr int epi_boxes() { return (BNMAX == 256 && !kPair && !kOutHalf && kAdds == 0 && !kLn && kEW == kEpiWarps) ? 1 : 2; }

// One 32x32 accumulator chunk of one warp: registers (thread = row) -> XOR-swizzled smem (conflict-free both ways)
// -> row-segment layout (fp32 out: 8 lanes x float4 per row, 4 rows per instruction; fp16 out: 4 lanes x 8 halfs per
// row, 8 rows per instruction) -> bias / addends / ReLU -> 16-byte coalesced stores.  kAdds = number of fp32 tensors
// added to the product (FSMN memory, residual); the adds of a chunk are all issued before the first use.
template <bool kOutHalf, int kAdds>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], float* stage, int lane, int row0, int col0,
                                               int M, int N, const GemmEpi& e, bool vec_ok, const float* bias_t, int n0) {
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
            const float4 b0 = *reinterpret_cast<const float4*>(bias_t + (col - n0));
            const float4 b1 = *reinterpret_cast<const float4*>(bias_t + (col - n0) + 4);
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
            const float4 b = *reinterpret_cast<const float4*>(bias_t + (col - n0));
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
                        x += bias_t[c - n0];
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

// Fused LayerNorm of the rows this GEMM just produced (x = acc + bias + residual): a row spans the CTAs of one
// cluster (one column tile each).  Every epilogue thread owns a row (TMEM lane): it publishes the row's partial sums
// to all CTAs of the cluster through distributed shared memory, waits for everybody's, and normalises the values it
// parked in TMEM into the fp16 operand of the next GEMM (TMA store).  Replaces a LayerNorm launch and its 16 MB round trip.
struct LnFuse {
    const CUtensorMap* tmL;      // fp16 [M, N] output, 32 x 32 boxes, SWIZZLE_64B
    const float* gamma_t;        // smem: gamma | beta of this tile's columns ([2][BNMAX])
    int bnmax;
    float* part;                 // smem [kLnMaxCluster][2][BM][2] partial sums landed here by every CTA of the cluster
    uint32_t part_u32, bar_u32;
    int cs, rank;
    float eps, inv_n;
};
__device__ __forceinline__ void ln_publish(const LnFuse& f, int grp, int rowq, float s1, float s2) {
    const uint32_t off = static_cast<uint32_t>((((f.rank * 2 + grp) * BM) + rowq) * 8);
    for (int j = 0; j < f.cs; ++j) {
        st_cluster_f32x2(mapa_shared(f.part_u32 + off, j), s1, s2);
        mbar_arrive_remote(mapa_shared(f.bar_u32, j));                 // release.cluster: orders the store above
    }
}
template <int kMaxChunks>
__device__ __forceinline__ void epilogue_ln_pass2(const LnFuse& f, uint32_t t_acc, int nchunks, uint8_t* stg, int cbase, int rowq,
                                                  int row0, int colw, int lane) {
    mbar_wait_cluster(f.bar_u32, 0);
    float S1 = 0.0f, S2 = 0.0f;
    for (int j = 0; j < f.cs; ++j)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float2 p = *reinterpret_cast<const float2*>(f.part + (((j * 2 + g) * BM) + rowq) * 2);
            S1 += p.x; S2 += p.y;
        }
    const float mean = S1 * f.inv_n;
    const float rstd = 1.0f / sqrtf(fmaxf(S2 * f.inv_n - mean * mean, 0.0f) + f.eps);
    if (nchunks <= 0) return;
    const uint32_t stg_u32 = smem_u32(stg);
    tmem_st_wait();
    if (lane == 0) tma_store_wait_read();                          // the fp32 boxes have been read: reuse them for fp16
    __syncwarp();
    uint32_t r[32];
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
        if (k < nchunks) {
            tmem_ld_32x32(t_acc + static_cast<uint32_t>(k * 32), r);
            tmem_ld_wait();
            if (k >= 2) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            uint8_t* box = stg + (k & 1) * 2048 + lane * 64;
            const float* g_t = f.gamma_t + (cbase + k) * 32;
            const float* b_t = g_t + f.bnmax;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = (__uint_as_float(r[8 * j + i]) - mean) * rstd * g_t[8 * j + i] + b_t[8 * j + i];
                __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
                __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(box + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(f.tmL, stg_u32 + (k & 1) * 2048, colw + k * 32, row0);
                tma_store_commit();
            }
        }
    }
}

// kPair = false: one CTA computes a 128 x bn tile with cta_group::1 MMAs.
// kPair = true : a CTA pair (thread-block cluster of 2) computes a 256 x bn tile with cta_group::2 MMAs issued by the
// leader (cluster rank 0).  CTA r stages its own 128 rows of A and only rows [r*bn/2, (r+1)*bn/2) of the W tile, and
// owns the accumulator rows m0 + r*128 .. +127 (all bn columns) in its own TMEM.  L2->SM bytes per k-block drop from
// 16K + bn*128 to 16K + bn*64 per SM, which is what bounds the 1-CTA kernel on this path (TMA chip throughput
// ~6300 B/cycle => ~42 B/cycle/SM with 148 CTAs pulling operands; B300_MICROARCH.md "TMA chip-throughput").  TMA
// multicast does not help at cluster size 2 (the L2 dedup window only pays from ~8 CTAs), the 2-SM MMA does.
// Protocol (same shape as CUTLASS's 2-SM pipelines): both CTAs' TMA loads complete_tx on the LEADER's full barrier
// (the leader alone arms it with the pair's total bytes); the leader's tcgen05.commit multicasts the "stage free"
// and "accumulator complete" arrivals to both CTAs; epilogue warps of both CTAs arrive on the leader's
// tmem_empty barrier (the peer through a mapa-translated shared::cluster address).
// kEW = epilogue warps: 8 (two column groups per TMEM quadrant) or 16 (four groups: half the chunks per warp; 640
// threads cap the kernel at 96 registers per thread; opt-in, see launch_cl).
template <int BNMAX, bool kPair, bool kOutHalf, int kAdds, bool kLn = false, int kEW = kEpiWarps, bool kPick = false>
__global__ void __launch_bounds__(kEpiWarp0 * 32 + kEW * 32, 1)
pf_gemm_f16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                       const __grid_constant__ CUtensorMap tmL, const GemmEpi epi, const int M, const int N, const int K,
                       const int bn, const int tiles_n, const int num_tiles, const int vec_ok_flags) {
    static_assert(!kLn || (!kPair && !kOutHalf && kAdds == 1), "fused LayerNorm rides on the fp32 + residual epilogue");
    static_assert(!kPick || (!kPair && !kOutHalf && kAdds == 0 && !kLn), "the fused greedy pick replaces the plain fp32 epilogue");
    const int vec_ok = vec_ok_flags & 1;
    const bool tma_epi = (vec_ok_flags & 2) != 0;                // asynchronous epilogue (tmC / tmR are valid)
    constexpr int kBoxes = epi_boxes<BNMAX, kPair, kOutHalf, kAdds, kLn, kEW>();
    using C = Cfg<BNMAX, kPair, kOutHalf, kLn, kEW, kBoxes>;
    constexpr int STAGES = C::kStages;
    constexpr int kGroups = kEW / 4;                             // column groups (warps per TMEM quadrant)
    constexpr int CS = kPair ? 2 : 1;
    const int brows = bn / CS;                                   // W rows this CTA stages per k-block
    const uint32_t stage_tx = kABytes + static_cast<uint32_t>(brows) * BK * 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);                    // SWIZZLE_128B tiles need 1024 B alignment (checked below)
    uint8_t* smem = smem_raw;

    constexpr uint32_t kEpiOff = STAGES * C::kStageBytes;
    constexpr uint32_t kBarOff = kEpiOff + C::kEpiBytes;
    const uint32_t bar_base = base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * STAGES + 4));
    auto resid_bar = [&](int ew, int b) { return bar_base + 8u * (2 * STAGES + 5 + 2 * ew + b); };
    float* bias_s = reinterpret_cast<float*>(smem + kBarOff + C::kBarBytes);       // [2][BNMAX] bias of the current tiles
    float* ln_gb = bias_s + 2 * BNMAX;                                             // kLn: [2][BNMAX] gamma | beta of this tile
    float* ln_part = ln_gb + 2 * BNMAX;                                            // kLn: [cluster][2][BM][2] partial sums
    const uint32_t ln_bar = bar_base + 8u * (2 * STAGES + 5 + 2 * kEW);
    const int ln_cs = kLn ? static_cast<int>(cluster_nctarank()) : 1;
    const int ln_rank = kLn ? static_cast<int>(cluster_ctarank()) : 0;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;
    const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;
    const bool leader = rank == 0;
    const bool dbg_no_epi = (vec_ok_flags & 0x100) != 0;               // PFASR_GEMM_DBG bottleneck probes (results are garbage)
    const bool dbg_no_tma = (vec_ok_flags & 0x200) != 0;
    const bool dbg_no_mma = (vec_ok_flags & 0x400) != 0;
    const int unit = blockIdx.x / CS;                            // CTA (or CTA pair) index = tile scheduler slot
    const int num_units = gridDim.x / CS;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((base & 1023u) != 0) { printf("pfasr: gemm needs 1024-byte aligned dynamic shared memory\n"); __trap(); }
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (tma_epi) {
            tma_prefetch_desc(&tmC);
            if (kAdds == 1) tma_prefetch_desc(&tmR);
        }
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
            mbar_init(tmem_empty_bar(a), kEW * CS);        // pair: the leader's barrier collects both CTAs' warps
        }
        for (int w = 0; w < kEW; ++w) {
            mbar_init(resid_bar(w, 0), 1);
            mbar_init(resid_bar(w, 1), 1);
        }
        if (kLn) mbar_init(ln_bar, ln_cs * kEW * 32);      // every epilogue thread of every CTA of the cluster arrives
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<CS>(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::kTmemCols);
        tmem_relinquish<CS>();
    }
    tc_fence_before_sync();
    // peer barriers must be initialised before any remote arrive.  Pair: right away (the first TMA already signals the
    // leader).  Fused LayerNorm: the first remote access is the statistics exchange at the end of the epilogue, so the
    // CTA only ARRIVES here and waits much later - it does not stall on its cluster peers becoming resident.
    if (kPair) cluster_sync();
    else { if (kLn) cluster_arrive(); __syncthreads(); }
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  The weight tiles do not
    // depend on that kernel either, so the producer requests the W halves of its first STAGES k-blocks BEFORE waiting
    // for it; activations (A), residuals and outputs are only touched after the wait.
    uint32_t npre = 0;
    if (!kPair && warp == 0 && lane == 0 && !dbg_no_tma) {
        uint32_t it = 0;
        for (int t = unit; t < num_tiles && it < STAGES; t += num_units) {
            const int n0 = (t % tiles_n) * bn;
            for (int kb = 0; kb < num_kb && it < STAGES; ++kb, ++it) {
                mbar_arrive_expect_tx(full_bar(it), stage_tx);
                tma_load_2d(base + it * C::kStageBytes + kABytes, &tmB, full_bar(it), kb * BK, n0);
            }
        }
        npre = it;
    }
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (one thread in every CTA)
        if (lane == 0) {
            uint32_t it = 0;                                     // global k-block counter across tiles
            for (int t = unit; t < num_tiles; t += num_units) {
                const int m0 = (t / tiles_n) * (BM * CS) + rank * BM;
                const int n0 = (t % tiles_n) * bn + rank * brows;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const uint32_t a_s = base + s * C::kStageBytes;
                    if (it < npre) {                             // stage armed and its W half already in flight
                        tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                        continue;
                    }
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (dbg_no_tma) {
                        if (leader) mbar_arrive(full_bar(s));
                    } else if (!kPair) {
                        mbar_arrive_expect_tx(full_bar(s), stage_tx);
                        tma_load_2d(a_s, &tmA, full_bar(s), kb * BK, m0);
                        tma_load_2d(a_s + kABytes, &tmB, full_bar(s), kb * BK, n0);
                    } else {
                        if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * stage_tx);
                        const uint32_t lead_full = mapa_shared(full_bar(s), 0);
                        tma_load_2d_pair(a_s, &tmA, lead_full, kb * BK, m0);
                        tma_load_2d_pair(a_s + kABytes, &tmB, lead_full, kb * BK, n0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread; of the leader CTA in pair mode)
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc(BM * CS, bn);
            uint32_t it = 0;
            uint32_t local = 0;                                  // tiles processed by this CTA (pair)
            for (int t = unit; t < num_tiles; t += num_units, ++local) {
                const uint32_t acc = local & 1u;
                const uint32_t acc_ph = (local >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);     // epilogue has drained this accumulator stage
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BNMAX;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after_sync();
                    const uint32_t a_s = base + s * C::kStageBytes;
                    const uint64_t adesc0 = make_sw128_kmajor_desc(a_s);
                    const uint64_t bdesc0 = make_sw128_kmajor_desc(a_s + kABytes);
                    if (!dbg_no_mma) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advancing K inside the swizzle atom = +32 bytes on the start address (>>4 => +2)
                            umma_f16<CS>(d_tmem, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
                    if (!kPair) umma_commit(empty_bar(s));
                    else umma_commit_pair(empty_bar(s));
                }
                if (!kPair) umma_commit(tmem_full_bar(acc));     // accumulator complete
                else umma_commit_pair(tmem_full_bar(acc));
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------ epilogue: TMEM -> registers -> smem transpose -> global
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
        const int grp = ew >> 2;                                 // warpgroup: left / right half of the tile's columns
        uint8_t* wstage = smem + kEpiOff + ew * C::kWarpStageBytes;
        float* stage = reinterpret_cast<float*>(wstage);         // legacy path: 32 x 32 fp32 transpose buffer
        constexpr bool kTmaOk = kOutHalf ? (kAdds == 0) : (kAdds <= 1);     // shapes the asynchronous epilogue covers
        constexpr int kCpwMax = (BNMAX / 32 + kGroups - 1) / kGroups;        // 32-column chunks per warp (TMA epilogue), at most
        const int nch = bn / 32;                                             // chunks of this tile width
        const int cpw = (nch + kGroups - 1) / kGroups;                       // share of every column group
        const int cbase = grp * cpw;
        ResidPipe rp;
        rp.bar[0] = resid_bar(ew, 0); rp.bar[1] = resid_bar(ew, 1);
        rp.count[0] = rp.count[1] = 0;
        const float lo = epi.relu ? 0.0f : -INFINITY;
        uint32_t local = 0;
        for (int t = unit; t < num_tiles; t += num_units, ++local) {
            const uint32_t acc = local & 1u;
            const uint32_t acc_ph = (local >> 1) & 1u;
            const int m0 = (t / tiles_n) * (BM * CS) + rank * BM;
            const int n0 = (t % tiles_n) * bn;
            const int row0 = m0 + q * 32;
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BNMAX;
            // bias of this tile's columns -> shared memory (zeros past N / without bias), one L2 round trip per tile
            // taken while the MMAs still run; two buffers so a fast warp may already fill the next tile's
            float* bias_t = bias_s + (local & 1u) * BNMAX;
            for (int i = ew * 32 + lane; i < bn; i += kEW * 32)
                bias_t[i] = (epi.bias != nullptr && n0 + i < N) ? __ldg(epi.bias + n0 + i) : 0.0f;
            if (kLn) {                                            // gamma | beta of this tile's columns (fused LayerNorm)
                for (int i = ew * 32 + lane; i < bn; i += kEW * 32) {
                    const bool ok = n0 + i < N;
                    ln_gb[i] = ok ? __ldg(epi.ln_gamma + n0 + i) : 0.0f;
                    ln_gb[BNMAX + i] = ok ? __ldg(epi.ln_beta + n0 + i) : 0.0f;
                }
            }
            if constexpr (kPick) {
                int nchunks = 0;
                for (int c = cbase; c < cbase + cpw && c < nch; ++c)
                    if (n0 + c * 32 < N) ++nchunks;
                mbar_wait(tmem_full_bar(acc), acc_ph);
                tc_fence_after_sync();
                epi_bar_sync_n<kEW>();                                   // bias tile visible to all epilogue warps
                if (m0 < M && !dbg_no_epi) {
                    const int row = row0 + lane;
                    float* dst = row < M ? epi.pick_out + (static_cast<size_t>(row) * epi.pick_ld + (t % tiles_n) * kGroups + grp) * 3 : nullptr;
                    epilogue_pick_f32<kCpwMax>(t_acc + cbase * 32, nchunks, bias_t + cbase * 32, n0 + cbase * 32, N, dst);
                }
            } else if (kTmaOk && tma_epi) {
                int nchunks = 0;
                for (int c = cbase; c < cbase + cpw && c < nch; ++c)
                    if (n0 + c * 32 < N) ++nchunks;
                const bool work = row0 < M && nchunks > 0 && !dbg_no_epi;
                const int colw = n0 + cbase * 32;
                // the CTA's last tile: its MMAs have consumed every operand stage, so the ring serves as staging with one
                // box per chunk and the chunks never wait for the store engine to release a box (PFASR_GEMM_DBG bit 32 off)
                const bool wide = !kPair && !kLn && kAdds == 0 && t + num_units >= num_tiles && (vec_ok_flags & 0x8000) == 0;
                uint8_t* wide_stage = smem + ew * (kCpwMax * (kOutHalf ? 2048 : 4096));
                static_assert(kPair || kLn || kAdds != 0 || kEW * kCpwMax * (kOutHalf ? 2048 : 4096) <= STAGES * C::kStageBytes, "ring too small for the last tile's boxes");
                int prefetched = 0;
                if constexpr (!kOutHalf && kAdds == 1) {
                    // residual boxes of the first two chunks travel while the MMAs of this tile still run
                    if (work) {
                        const uint32_t stg_u32 = smem_u32(wstage);
                        if (lane == 0) tma_store_wait_read();    // the previous tile's stores have read the boxes
                        __syncwarp();
                        for (; prefetched < 2 && prefetched < nchunks; ++prefetched)
                            resid_issue<true>(rp, prefetched, stg_u32, &tmR, colw + prefetched * 32, row0, lane);
                    }
                }
                mbar_wait(tmem_full_bar(acc), acc_ph);
                tc_fence_after_sync();
                epi_bar_sync_n<kEW>();                                   // bias tile visible to all epilogue warps
                if (work) {
                    if constexpr (kOutHalf && kAdds == 0)
                        epilogue_tma_f16<kCpwMax>(t_acc + cbase * 32, nchunks, wide ? wide_stage : wstage, &tmC, bias_t + cbase * 32, lo, row0, colw, lane,
                                                  (vec_ok_flags >> 11) & 3, wide);
                    else if constexpr (!kOutHalf && kAdds == 0)
                        epilogue_tma_f32<kCpwMax, false, false, kBoxes>(t_acc + cbase * 32, nchunks, wide ? wide_stage : wstage, rp, &tmC, &tmR,
                                                                        bias_t + cbase * 32, lo, row0, colw, lane, 0, nullptr, nullptr,
                                                                        (vec_ok_flags & 0x2000) != 0, wide);
                    else if constexpr (!kOutHalf && kAdds == 1 && !kLn)
                        epilogue_tma_f32<kCpwMax, true>(t_acc + cbase * 32, nchunks, wstage, rp, &tmC, &tmR, bias_t + cbase * 32, lo, row0, colw, lane, prefetched);
                }
                if constexpr (kLn) {
                    float s1 = 0.0f, s2 = 0.0f;
                    if (work)
                        epilogue_tma_f32<kCpwMax, true, true>(t_acc + cbase * 32, nchunks, wstage, rp, &tmC, &tmR, bias_t + cbase * 32, lo, row0, colw,
                                                              lane, prefetched, &s1, &s2);
                    LnFuse f;
                    f.tmL = &tmL; f.gamma_t = ln_gb; f.bnmax = BNMAX; f.part = ln_part; f.part_u32 = smem_u32(ln_part); f.bar_u32 = ln_bar;
                    f.cs = ln_cs; f.rank = ln_rank; f.eps = epi.ln_eps; f.inv_n = 1.0f / static_cast<float>(N);
                    cluster_wait();                               // (arrived in the prologue) the peers' ln barriers exist
                    ln_publish(f, grp, q * 32 + lane, s1, s2);   // every epilogue thread reports, also the ones without columns
                    if (work) epilogue_ln_pass2<kCpwMax>(f, t_acc + cbase * 32, nchunks, wstage, cbase, q * 32 + lane, row0, colw, lane);
                }
            } else {
                mbar_wait(tmem_full_bar(acc), acc_ph);
                tc_fence_after_sync();
                epi_bar_sync_n<kEW>();
                if (m0 < M && !dbg_no_epi) {
#pragma unroll 1
                    for (int c = grp; c < nch; c += kGroups) {
                        const int col0 = n0 + c * 32;
                        if (col0 >= N) break;                    // warp-uniform
                        uint32_t r[32];
                        tmem_ld_32x32(t_acc + static_cast<uint32_t>(c * 32), r);
                        tmem_ld_wait();
                        epilogue_chunk<kOutHalf, kAdds>(r, stage, lane, row0, col0, M, N, epi, vec_ok != 0, bias_t, n0);
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (!kPair || leader) mbar_arrive(tmem_empty_bar(acc));
                else mbar_arrive_remote(mapa_shared(tmem_empty_bar(acc), 0));
            }
        }
        if (lane == 0) tma_store_wait_read();                    // staging boxes stay valid until the stores have read them
    }
    tc_fence_before_sync();
    // no CTA exits while a peer may still signal / write to it
    if (kLn) { if (warp < kEpiWarp0) cluster_wait(); cluster_sync(); }
    else if (kPair) cluster_sync();
    else __syncthreads();
    if (warp == 2) tmem_dealloc<CS>(tmem_base, C::kTmemCols);
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

// 2-D row-major [rows, cols] (fp16 or fp32) with row pitch ld elements; box = [box_rows, 128 bytes of columns],
// 128B swizzle; loads read out-of-bounds elements as 0, stores clip them.
void make_tmap_any(CUtensorMap* tm, const void* ptr, bool f32, int rows, int cols, int ld, int box_rows, int box_bytes = 128) {
    const int esz = f32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (static_cast<size_t>(ld) * esz) % 16 != 0)
        throw CudaError{"gemm tensor must be 16-byte aligned with a row pitch that is a multiple of 16 bytes"};
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / esz), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr),
                                 dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))};
}
void make_tmap(CUtensorMap* tm, const __half* ptr, int rows, int cols, int ld, int box_rows) {
    make_tmap_any(tm, ptr, false, rows, cols, ld, box_rows);
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

template <int BNMAX, bool kPair, bool kOutHalf, int kAdds, bool kLn = false, int kEW = kEpiWarps, bool kPick = false>
struct Launcher {
    static int max_units;   // co-resident CTAs (or CTA pairs: GPC boundaries can make it < sms / 2)

    static void run(const GemmOp& op, cudaStream_t stream) {
        using C = Cfg<BNMAX, kPair, kOutHalf, kLn, kEW, epi_boxes<BNMAX, kPair, kOutHalf, kAdds, kLn, kEW>()>;
        constexpr int CS = kPair ? 2 : 1;
        auto kern = pf_gemm_f16_tn_tcgen05<BNMAX, kPair, kOutHalf, kAdds, kLn, kEW, kPick>;
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
            max_units = num_sms() / CS;
            if (kPair) {
                cudaLaunchConfig_t q{};
                q.gridDim = dim3(num_sms() / CS * CS);
                q.blockDim = dim3(C::kThreadsCta);
                q.dynamicSmemBytes = C::kSmemBytes;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                q.attrs = at; q.numAttrs = 1;
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, kern, &q) == cudaSuccess && nc > 0) max_units = std::min(max_units, nc);
                else cudaGetLastError();
            }
        });
        const int tiles_n = ceil_div(op.N, op.bn);
        const int num_tiles = tiles_n * ceil_div(op.M, BM * CS);
        // fused LayerNorm: one tile per CTA, the column tiles of a row tile form a cluster
        // persistent grid, balanced: with r = ceil(tiles / SMs) rounds, ceil(tiles / r) CTAs finish at the same time as a
        // full grid would, and the SMs left over start the next kernel (PDL) or another lane's CTAs
        const int rounds = ceil_div(num_tiles, max_units);     // (more rounds on fewer CTAs measured slower, also with lanes)
        const int grid = kLn ? num_tiles : ceil_div(num_tiles, rounds) * CS;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(C::kThreadsCta);
        cfg.dynamicSmemBytes = C::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        int na = 0;
        if (pdl_enabled()) {      // (GEMMs without PDL: 4.18 vs 3.82 ms/step with three lanes - early residency is worth it)
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (kPair || kLn) {
            at[na].id = cudaLaunchAttributeClusterDimension;
            at[na].val.clusterDim.x = kLn ? tiles_n : CS; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = at;
        cfg.numAttrs = na;
        PF_CUDA(cudaLaunchKernelEx(&cfg, kern, op.tmA, op.tmB, op.tmC, op.tmR, op.tmL, op.epi, op.M, op.N, op.K, op.bn, tiles_n, num_tiles, op.vec_ok));
    }
};
template <int BNMAX, bool kPair, bool kOutHalf, int kAdds, bool kLn, int kEW, bool kPick>
int Launcher<BNMAX, kPair, kOutHalf, kAdds, kLn, kEW, kPick>::max_units = 1;

template <int BNMAX, bool kOutHalf, int kAdds>
void launch_cl(const GemmOp& op, cudaStream_t stream) {
#ifdef PFASR_EXPERIMENTS
    if constexpr (kOutHalf && kAdds == 0) {
        // fp16 output without addends (QKV, FFN1, K/V projections) with sixteen epilogue warps: opt-in (PFASR_GEMM_EPI16=1).
        // Measured slower - 5.22 vs 5.12 ms/step (one lane), 3.92 vs 3.87 (three lanes), GEMM replay 587 vs 606 TFLOP/s:
        // the extra staging costs the fourth operand stage and 640 threads cap the kernel at 96 registers.
        static const bool epi16 = [] { const char* e = getenv("PFASR_GEMM_EPI16"); return e && *e == '1'; }();
        if ((epi16 || op.epi16) && op.cm == 1 && op.cn == 1 && (op.vec_ok & 2)) { Launcher<BNMAX, false, true, 0, false, 16>::run(op, stream); return; }
    }
    if (op.cm == 2 && op.cn == 1) { Launcher<BNMAX, true, kOutHalf, kAdds>::run(op, stream); return; }
#endif
    if (op.cm == 1 && op.cn == 1) Launcher<BNMAX, false, kOutHalf, kAdds>::run(op, stream);
    else throw CudaError{"gemm: cluster shape not instantiated (CTA pairs need a PFASR_BUILD_EXPERIMENTS=1 build)"};
}

template <int BNMAX>
void launch_bn(const GemmOp& op, cudaStream_t stream) {
    if (op.pick) { Launcher<BNMAX, false, false, 0, false, kEpiWarps, true>::run(op, stream); return; }
    if (op.ln_cluster > 0) {
#ifdef PFASR_EXPERIMENTS
        Launcher<BNMAX, false, false, 1, true>::run(op, stream);
        return;
#else
        throw CudaError{"gemm: the fused-LayerNorm epilogue needs a PFASR_BUILD_EXPERIMENTS=1 build"};
#endif
    }
    const bool h = op.epi.out_f16 != nullptr;
    switch (op.n_adds) {
        case 0:  h ? launch_cl<BNMAX, true, 0>(op, stream) : launch_cl<BNMAX, false, 0>(op, stream); break;
        case 1:  h ? launch_cl<BNMAX, true, 1>(op, stream) : launch_cl<BNMAX, false, 1>(op, stream); break;
        default: h ? launch_cl<BNMAX, true, 2>(op, stream) : launch_cl<BNMAX, false, 2>(op, stream); break;
    }
}

int pair_mode() {                      // 0 off, 1 the tile picker may choose CTA pairs, 2 pairs wherever the shape allows
#ifdef PFASR_EXPERIMENTS
    static const int mode = [] { const char* e = getenv("PFASR_GEMM_PAIR"); return e ? atoi(e) : 0; }();
    return mode;
#else
    return 0;
#endif
}
bool pair_enabled() { return pair_mode() != 0; }

// Tile width (any multiple of 32 up to 256) and pairing from a wave model fitted to scripts/gemm_probe.py on B200
// (K = 512, fp16 epilogue, one wave of 128 x bn tiles: 2.96 us @128, 3.63 us @192, 4.56 us @256):
//   t_wave(bn, kb) = 0.3 + 0.004 bn  +  kb (0.13 + 0.0011 bn)  us     (tile switch + epilogue, then the k-blocks)
// and the kernel costs waves(bn) * t_wave.  What it captures: M = 5312 is 41.5 row tiles, so the number of waves jumps
// with the column-tile count - 224-wide tiles cover N = 1536 / 2048 in 294 / 420 tiles (2 / 3 waves of 0.875-size tiles)
// where 256-wide ones need 252 / 336 tiles (2 / 3 waves of full-size tiles), and N = 512 fits one wave at 192.
// Two objectives.  Latency (one batch at a time): kernel time = waves x t_wave.  Throughput (several batches in flight,
// execution lanes): other work fills idle SMs, so what counts is the SM time a launch occupies = CTAs x fixed cost
// (launch, prologue, first loads, exposed last epilogue, ~3 us) + tiles x t_tile - e.g. N = 512, K = 2048 prefers 84 full
// 256-wide tiles (1460 SM-us) over one wave of 126 192-wide ones (1890 SM-us).
thread_local int g_policy = 0;      // 0 latency, 1 throughput

void pick_config(int M, int N, int K, int& bn_out, int& cm_out, int& cn_out) {
    const int mt = ceil_div(M, BM), kb = ceil_div(K, BK), sms = num_sms();
    cn_out = 1;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 32; bn -= 32) {
        if (bn > 32 && ceil_div(N, bn) == ceil_div(N, bn - 32)) continue;      // a narrower tile covers N with as many columns
        const bool pair_ok = pair_enabled() && bn % 64 == 0;
        for (int cm = (pair_ok && pair_mode() == 2 && mt >= 2) ? 2 : 1; cm <= (pair_ok ? 2 : 1); ++cm) {
            const int tiles = ceil_div(mt, cm) * ceil_div(N, bn);
            const int waves = ceil_div(tiles, sms / cm);
            const double t_wave = 0.3 + 0.004 * bn + kb * (0.13 + 0.0011 * bn) + (cm > 1 ? 0.1 * kb : 0.0);
            const double cost = g_policy == 1 ? cm * (std::min(tiles, sms / cm) * 3.0 + tiles * t_wave) : waves * t_wave;
            if (cost < best_cost) { best_cost = cost; bn_out = bn; cm_out = cm; }
        }
    }
}

}  // namespace

void* tensormap_encode_fn() { return reinterpret_cast<void*>(get_encode_fn()); }
void gemm_set_policy(int throughput) {
    static const int forced = [] { const char* e = getenv("PFASR_GEMM_POLICY"); return e ? (*e == 't' ? 1 : 0) : -1; }();
    g_policy = forced >= 0 ? forced : (throughput ? 1 : 0);
}
void gemm_make_tmap(CUtensorMap* tm, const void* ptr, bool f32, int rows, int cols, int ld, int box_rows, int box_bytes) {
    make_tmap_any(tm, ptr, f32, rows, cols, ld, box_rows, box_bytes);
}
int gemm_num_sms() { return num_sms(); }

void gemm_prepare(GemmOp& op, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  const GemmEpi& epi, int tile_code) {
    const bool pick = epi.pick_out != nullptr;
    if (!pick && (epi.out_f32 != nullptr) == (epi.out_f16 != nullptr)) throw CudaError{"gemm: exactly one output pointer must be set"};
    if (pick && (epi.out_f32 || epi.out_f16 || epi.resid || epi.addend || epi.ln_out16 || epi.relu))
        throw CudaError{"gemm: the fused greedy pick takes bias only and stores nothing else"};
    if (M <= 0 || N <= 0 || K <= 0) throw CudaError{"gemm: empty problem"};
    int bn = tile_code & 0xFFF, cm = (tile_code >> 12) & 0xF, cn = (tile_code >> 16) & 0xF;
    if ((tile_code & 0xFFFFF) == 0) pick_config(M, N, K, bn, cm, cn);
    op.ln_cluster = 0;
    if (epi.ln_out16 != nullptr) {
        // fused LayerNorm: narrowest tile whose column tiles (<= 4) cover a row inside one cluster and the grid in one wave
        if (!epi.out_f32 || !epi.resid || epi.addend || !epi.ln_gamma || !epi.ln_beta || !gemm_ln_fusable(M, N))
            throw CudaError{"gemm: fused LayerNorm needs fp32 output + residual and a row that fits one cluster in one wave"};
        for (int w = 128; w <= 256; w += 32)
            if (ceil_div(N, w) <= kLnMaxCluster && ceil_div(M, BM) * ceil_div(N, w) <= num_sms()) { bn = w; break; }
        cm = 1; cn = 1;
        op.ln_cluster = ceil_div(N, bn);
    }
    cm = std::max(cm, 1);
    cn = std::max(cn, 1);
    if (bn < 32 || bn > 256 || bn % 32 != 0) throw CudaError{"gemm: the N tile must be a multiple of 32 in [32, 256]"};
    if (cm > 2 || cn != 1) throw CudaError{"gemm: unsupported cluster shape (1x1 and the 2x1 CTA pair are instantiated)"};
#ifndef PFASR_EXPERIMENTS
    if (cm == 2 || ((tile_code >> 22) & 1)) throw CudaError{"gemm: CTA pairs / sixteen epilogue warps need a PFASR_BUILD_EXPERIMENTS=1 build"};
#endif
    if (cm == 2 && bn % 64 != 0) throw CudaError{"gemm: the CTA-pair MMA needs an N tile that is a multiple of 64"};
    if (pick) {
        if (bn < 128) bn = 128;                                      // bounds the partial slots per row (gemm_pick_slots)
        cm = 1;
        if (ceil_div(N, bn) * (kEpiWarps / 4) > epi.pick_ld) throw CudaError{"gemm: pick_ld too small for this tile width"};
    }
    op.pick = pick ? 1 : 0;
    op.M = M; op.N = N; op.K = K; op.bn = bn; op.cm = cm; op.cn = cn; op.epi = epi;
    op.throughput = g_policy;
    op.epi16 = (tile_code >> 22) & 1;
    // the kernel adds up to two fp32 tensors in a fixed order: FSMN memory first, then the residual
    op.n_adds = 0;
    op.epi.add0 = op.epi.add1 = nullptr;
    if (epi.addend) { op.epi.add0 = epi.addend; op.epi.ld_add0 = epi.ld_addend; op.n_adds = 1; }
    if (epi.resid) {
        if (op.n_adds == 0) { op.epi.add0 = epi.resid; op.epi.ld_add0 = epi.ld_resid; }
        else { op.epi.add1 = epi.resid; op.epi.ld_add1 = epi.ld_resid; }
        ++op.n_adds;
    }
    // in-place residual update (x += A W^T + b: out-projection, FFN2, decoder out-projection): the epilogue hands
    // acc + bias to the TMA engine as an fp32 REDUCE-ADD into x instead of loading x, adding and storing - one rounding
    // for acc + bias and one for the sum either way, so the result is bit-identical, but the residual never travels
    // to the SM and the staging boxes are never waited on for a load.  PFASR_GEMM_NO_REDADD=1 keeps the load-add-store path.
    static const bool no_redadd = [] { const char* e = getenv("PFASR_GEMM_NO_REDADD"); return e && *e && *e != '0'; }();
    const bool red_add = !no_redadd && op.ln_cluster == 0 && epi.resid != nullptr && epi.addend == nullptr && epi.out_f32 != nullptr &&
                         epi.resid == epi.out_f32 && epi.ld_resid == epi.ld_out && !epi.relu;
    if (red_add) { op.n_adds = 0; op.epi.add0 = nullptr; }
    // vector epilogue: 16-byte row segments of every fp32 tensor (8-byte for the fp16 output) must be aligned
    auto al = [](const void* p, int ld, int bytes) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % bytes) == 0 && ld % 4 == 0); };
    op.vec_ok = (al(epi.bias, 0, 16) && al(epi.resid, epi.ld_resid, 16) && al(epi.addend, epi.ld_addend, 16) &&
                 al(epi.out_f32, epi.ld_out, 16) && al(epi.out_f16, epi.ld_out, 8)) ? 1 : 0;
    // asynchronous epilogue (TMA store, TMA-prefetched residual): fp16 output without addends, fp32 output with at
    // most one; everything 16-byte aligned.  PFASR_GEMM_LEGACY_EPI=1 keeps the smem-transposed epilogue (A/B switch).
    static const bool legacy_epi = [] { const char* e = getenv("PFASR_GEMM_LEGACY_EPI"); return e && *e && *e != '0'; }();
    auto row_ok = [](const void* p, int ld, int esz) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (static_cast<size_t>(ld) * esz) % 16 == 0; };
    const bool tma_epi = !pick && !legacy_epi && op.vec_ok &&
                         (epi.out_f16 ? (op.n_adds == 0 && row_ok(epi.out_f16, epi.ld_out, 2))
                                      : (op.n_adds <= 1 && row_ok(epi.out_f32, epi.ld_out, 4) &&
                                         (op.n_adds == 0 || row_ok(op.epi.add0, op.epi.ld_add0, 4))));
    op.tmC = CUtensorMap{};
    op.tmR = CUtensorMap{};
    if (red_add && !tma_epi) {                                    // the reduce needs the TMA epilogue: fall back to load-add-store
        op.epi.add0 = epi.resid; op.epi.ld_add0 = epi.ld_resid; op.n_adds = 1;
    }
    op.red_add = red_add && tma_epi ? 1 : 0;
    if (op.red_add) op.vec_ok |= 0x2000;
    if (tma_epi) {
        op.vec_ok |= 2;
        if (epi.out_f16) make_tmap_any(&op.tmC, epi.out_f16, false, M, N, epi.ld_out, 32, 64);   // 32 x 32 fp16 boxes
        else make_tmap_any(&op.tmC, epi.out_f32, true, M, N, epi.ld_out, 32);
        if (op.n_adds == 1) make_tmap_any(&op.tmR, op.epi.add0, true, M, N, op.epi.ld_add0, 32);
    }
    op.tmL = CUtensorMap{};
    if (op.ln_cluster > 0) {
        if (!tma_epi || !row_ok(epi.ln_out16, epi.ld_ln16, 2) || (reinterpret_cast<uintptr_t>(epi.ln_gamma) & 3) != 0)
            throw CudaError{"gemm: fused LayerNorm needs 16-byte aligned tensors"};
        make_tmap_any(&op.tmL, epi.ln_out16, false, M, N, epi.ld_ln16, 32, 64);
    }
    if (const char* e = getenv("PFASR_GEMM_DBG")) op.vec_ok |= (atoi(e) & 31) << 8;   // 1 no epilogue, 2 no TMA, 4 no MMA, 8 no TMA store, 16 no staging either
    static const bool no_wide = [] { const char* e = getenv("PFASR_GEMM_NO_WIDE_EPI"); return e && *e && *e != '0'; }();
    if (no_wide) op.vec_ok |= 0x8000;                               // A/B switch: the last tile keeps the two / one staging boxes
#ifdef PFASR_EXPERIMENTS
    op.half_sm = gemm_half_eligible(op, ((tile_code >> 23) & 1) != 0) ? 1 : 0;
#else
    op.half_sm = 0;
    if ((tile_code >> 23) & 1) throw CudaError{"gemm: the half-SM kernel needs a PFASR_BUILD_EXPERIMENTS=1 build"};
#endif
    make_tmap(&op.tmA, A, M, K, lda, BM);           // every CTA stages its own 128 rows of A
    make_tmap(&op.tmB, W, N, K, ldw, bn / cm);      // ... and (in a CTA pair) half of the W tile
}

void gemm_launch(const GemmOp& op, cudaStream_t stream) {
#ifdef PFASR_EXPERIMENTS
    if (op.half_sm) { gemm_half_launch(op, stream); return; }
#endif
    if (op.bn <= 128) launch_bn<128>(op, stream);      // 32 KiB stage slots, 6 stages
    else launch_bn<256>(op, stream);                   // 48 KiB stage slots
}

double gemm_flops(const GemmOp& op) { return 2.0 * op.M * static_cast<double>(op.N) * op.K; }
int gemm_pick_slots(int N) { return ceil_div(N, 128) * (kEpiWarps / 4); }

bool gemm_ln_fusable(int M, int N) {
#ifndef PFASR_EXPERIMENTS
    (void)M; (void)N;
    return false;
#endif
    if (N % 32 != 0 || N > kLnMaxCluster * 256) return false;
    for (int w = 128; w <= 256; w += 32)
        if (ceil_div(N, w) <= kLnMaxCluster && ceil_div(M, BM) * ceil_div(N, w) <= num_sms()) return true;
    return false;
}

}  // namespace pf
