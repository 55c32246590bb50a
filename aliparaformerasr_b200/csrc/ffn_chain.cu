// Feed-forward block of an encoder layer as one persistent tcgen05 kernel.
//
// The two GEMMs of the block (relu(a W1^T + b1) -> h, then x += h W2^T + b2) used to be two launches.  Each paid its
// own launch / prologue / first-load latency and its own exposed last epilogue (~5 us of a ~15 us kernel), the first
// one ran 2.27 waves of tiles as 3, the second one left 22 of 148 SMs idle.  Here both tile sets are work items of one
// grid.  Every CTA runs a static list: first its share of the 128 x 256 tiles of h ("a" items, n-major, so the first
// wave already covers the low columns of every row block), then "b" items = a 128 x 256 tile of x restricted to one
// K range (the second GEMM is split along K so that its work units are small enough to balance over 148 SMs).
//
// Dependencies:
//  * a b item reads h.  Warp 3 of its CTA (the "watcher") polls the flags of that row block - one per tile of h, set by
//    the a tile's epilogue once its TMA stores have COMPLETED - and mirrors them into shared memory; the TMA producer
//    only looks at shared memory, per group of 4 k-blocks (= one tile of h), so a published tile costs it nothing.
//  * the K parts of one tile of x accumulate in a fixed order (bit-reproducible): part 0 writes x + b2 + p0, part j
//    waits - in its epilogue, never in its main loop - for part j-1's flag and adds its own product on top.
//  a items never wait; every CTA runs its a items first, then its part-0 items, then later parts: whatever an item
//  waits for sits earlier in every list, so the kernel cannot deadlock (the host builds the lists that way).
//
// Replaces two of the MLAS GEMMs per encoder layer that OnnxRuntime runs for the reference
// (/root/reference/AliParaformerAsr/OfflineProjOfParaformer.cs:68, InferenceSession.Run).
#include "ffn_chain.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "gemm_dev.cuh"

namespace pf {

using namespace gemm_dev;

namespace {

constexpr int BN = 256;                               // tile width of both GEMMs
constexpr int kStages = 3;
constexpr int kStageBytes = kABytes + BN * BK * 2;    // 48 KiB
constexpr int kWarpStageBytes = 8192;                 // two 32 x 32 fp32 boxes (the fp16 epilogue uses the first 4 KiB)
constexpr int kEpiBytes = kEpiWarps * kWarpStageBytes;
constexpr int kBarBytes = 512;
constexpr int kBiasBytes = 2 * BN * 4;
constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + kBiasBytes + 1024;
constexpr int kTmemCols = 2 * BN;
constexpr int kKbPerATile = BN / BK;                  // k-blocks of the second GEMM that one tile of h feeds
constexpr int kItemB = 1 << 30;                       // item code: bit 30 = b item, bits 24..29 = K part, low 24 bits = tile
constexpr int kMaxSplits = 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct ChainParams {
    const float* bias1;
    const float* bias2;
    const int* sched;
    int* flags;
    int max_items;
    int epoch;
    int M, D, F;
    int mt, n1t, n2t;
    int splits;                       // K parts of a tile of x
    int dbg;                          // PFASR_CHAIN_DBG probes: 1 no flag waits, 2 no publication, 4 no proxy fence after a wait
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr int kMaxN1T = 16;
// One look at all flags of a row block (they are contiguous): bit n = tile (m, n) of h is published.  The loads are
// independent, so the whole look costs one L2 round trip.
__device__ __forceinline__ uint32_t poll_flags(const int* row_flags, int n1t, int epoch) {
    int v[kMaxN1T];
#pragma unroll
    for (int n = 0; n < kMaxN1T; ++n) v[n] = n < n1t ? ld_relaxed_gpu(row_flags + n) : 0;
    uint32_t mask = 0;
#pragma unroll
    for (int n = 0; n < kMaxN1T; ++n) mask |= (v[n] == epoch ? 1u : 0u) << n;
    return mask;
}
// after a successful look: acquire (the relaxed loads above + this fence) and order the TMA loads that follow after it
__device__ __forceinline__ void acquire_for_tma(bool proxy_fence) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    if (proxy_fence) fence_proxy_async_all();
}

struct Item {
    bool is_b;
    int part;          // b: K part index
    int m, nt;         // row tile, column tile
    int kb_lo, kb_hi;  // k-block range
};
__device__ __forceinline__ Item decode_item(int code, const ChainParams& p) {
    Item it;
    it.is_b = (code & kItemB) != 0;
    if (!it.is_b) {
        it.part = 0;
        it.m = code % p.mt;
        it.nt = code / p.mt;
        it.kb_lo = 0;
        it.kb_hi = p.D / BK;
    } else {
        const int idx = code & 0xFFFFFF;
        it.part = (code >> 24) & 0x3F;
        it.m = idx / p.n2t;
        it.nt = idx % p.n2t;
        const int kb2 = p.F / BK;
        it.kb_lo = it.part * kb2 / p.splits;
        it.kb_hi = (it.part + 1) * kb2 / p.splits;
    }
    return it;
}
// bits of the tiles of h that k-blocks [kb_lo, kb_hi) of the second GEMM read
__device__ __forceinline__ uint32_t need_mask(int kb_lo, int kb_hi) {
    const int g0 = kb_lo / kKbPerATile, g1 = (kb_hi + kKbPerATile - 1) / kKbPerATile;
    return ((1u << g1) - 1u) & ~((1u << g0) - 1u);
}

__global__ void __launch_bounds__(kThreads, 1)
pf_ffn_chain_tcgen05(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                     const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2, const ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    constexpr uint32_t kEpiOff = kStages * kStageBytes;
    constexpr uint32_t kBarOff = kEpiOff + kEpiBytes;
    const uint32_t bar_base = base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * kStages + 4));
    auto resid_bar = [&](int ew, int b) { return bar_base + 8u * (2 * kStages + 5 + 2 * ew + b); };
    float* bias_s = reinterpret_cast<float*>(smem + kBarOff + kBarBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int* my_items = p.sched + static_cast<size_t>(blockIdx.x) * p.max_items;
    constexpr uint32_t stage_tx = kStageBytes;
    volatile uint32_t* wtag = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 400);   // watcher -> producer: (b seq << 16) | ready bits
    int* flags_b = p.flags + p.mt * p.n1t;                       // per tile of x: (epoch << 3) | parts written

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmC1);
        tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmC2);
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
            mbar_init(tmem_empty_bar(a), kEpiWarps);
        }
        for (int w = 0; w < kEpiWarps; ++w) {
            mbar_init(resid_bar(w, 0), 1);
            mbar_init(resid_bar(w, 1), 1);
        }
        *wtag = 0;
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
        tmem_relinquish<1>();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // the schedule and the weights do not depend on the previous kernel: the producer requests the weight halves of its
    // first k-blocks before the programmatic-dependent-launch wait
    uint32_t npre = 0;
    if (warp == 0 && lane == 0 && my_items[0] >= 0) {
        const Item it0 = decode_item(my_items[0], p);
        for (int kb = it0.kb_lo; kb < it0.kb_hi && npre < kStages; ++kb, ++npre) {
            mbar_arrive_expect_tx(full_bar(npre), stage_tx);
            tma_load_2d(base + npre * kStageBytes + kABytes, it0.is_b ? &tmB2 : &tmB1, full_bar(npre), kb * BK, it0.nt * BN);
        }
    }
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            uint32_t my_seq = 0;                                 // b items seen so far (1-based id of the current one)
            for (int i = 0; i < p.max_items; ++i) {
                const int code = my_items[i];
                if (code < 0) break;
                const Item w = decode_item(code, p);
                const int m0 = w.m * BM, n0 = w.nt * BN;
                const CUtensorMap* ta = w.is_b ? &tmA2 : &tmA1;
                const CUtensorMap* tb = w.is_b ? &tmB2 : &tmB1;
                uint32_t ready = 0;                              // tiles of h (this row block) known to be published
                if (w.is_b) ++my_seq;
                for (int kb = w.kb_lo; kb < w.kb_hi; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    const uint32_t a_s = base + s * kStageBytes;
                    // rows m of h, columns [kb * 64, +64) come from the a tile (m, kb / 4): wait until the watcher has seen it
                    if (w.is_b && !(p.dbg & 1)) {
                        const uint32_t need = 1u << (kb / kKbPerATile);
                        if (!(ready & need)) {
                            uint32_t spins = 0;
                            for (;;) {
                                const uint32_t tag = *wtag;
                                const uint32_t ts = tag >> 16;
                                ready = ts > my_seq ? 0xFFFFu : (ts == my_seq ? (tag & 0xFFFFu) : 0u);
                                if (ready & need) break;
                                if (++spins > (1u << 26)) {
                                    printf("pfasr: ffn chain producer wait timeout (block %d)\n", blockIdx.x);
                                    __trap();
                                }
                            }
                            asm volatile("fence.acq_rel.cta;" ::: "memory");
                            if (!(p.dbg & 4)) fence_proxy_async_all();   // the TMA loads below read what other SMs stored through TMA
                        }
                    }
                    if (it < npre) {
                        tma_load_2d(a_s, ta, full_bar(s), kb * BK, m0);
                        continue;
                    }
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), stage_tx);
                    tma_load_2d(a_s, ta, full_bar(s), kb * BK, m0);
                    tma_load_2d(a_s + kABytes, tb, full_bar(s), kb * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0;
            uint32_t local = 0;
            for (int i = 0; i < p.max_items; ++i, ++local) {
                const int code = my_items[i];
                if (code < 0) break;
                const Item w = decode_item(code, p);
                const uint32_t acc = local & 1u;
                const uint32_t acc_ph = (local >> 1) & 1u;
                mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = w.kb_lo; kb < w.kb_hi; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after_sync();
                    const uint32_t a_s = base + s * kStageBytes;
                    const uint64_t adesc0 = make_sw128_kmajor_desc(a_s);
                    const uint64_t bdesc0 = make_sw128_kmajor_desc(a_s + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_f16<1>(d_tmem, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb > w.kb_lo || k != 0) ? 1u : 0u);
                    umma_commit(empty_bar(s));
                }
                umma_commit(tmem_full_bar(acc));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ------------------------------------------------ watcher: global flags of h -> shared memory
        if (lane == 0 && !(p.dbg & 1)) {
            uint32_t seq = 0;
            for (int i = 0; i < p.max_items; ++i) {
                const int code = my_items[i];
                if (code < 0) break;
                const Item w = decode_item(code, p);
                if (!w.is_b) continue;
                ++seq;
                const uint32_t need = need_mask(w.kb_lo, w.kb_hi);
                uint32_t seen = 0, spins = 0;
                for (;;) {
                    const uint32_t mask = poll_flags(p.flags + w.m * p.n1t, p.n1t, p.epoch);
                    if ((mask & need) != (seen & need)) {
                        asm volatile("fence.acq_rel.gpu;" ::: "memory");      // acquire: the relaxed looks + this fence
                        seen = mask;
                        *wtag = (seq << 16) | (mask & 0xFFFFu);
                    }
                    if ((seen & need) == need) break;
                    __nanosleep(20);
                    if (++spins > (1u << 22)) {
                        printf("pfasr: ffn chain watcher timeout (block %d)\n", blockIdx.x);
                        __trap();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------ epilogue
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;
        const int grp = ew >> 2;
        uint8_t* wstage = smem + kEpiOff + ew * kWarpStageBytes;
        constexpr int nch = BN / 32, cpw = nch / 2;
        const int cbase = grp * cpw;
        ResidPipe rp;
        rp.bar[0] = resid_bar(ew, 0); rp.bar[1] = resid_bar(ew, 1);
        rp.count[0] = rp.count[1] = 0;
        uint32_t local = 0;
        for (int i = 0; i < p.max_items; ++i, ++local) {
            const int code = my_items[i];
            if (code < 0) break;
            const Item w = decode_item(code, p);
            const bool is_b = w.is_b;
            const int m0 = w.m * BM, n0 = w.nt * BN;
            const int N = is_b ? p.D : p.F;
            const uint32_t acc = local & 1u;
            const uint32_t acc_ph = (local >> 1) & 1u;
            const int row0 = m0 + q * 32;
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
            float* bias_t = bias_s + (local & 1u) * BN;
            const float* bias = is_b ? (w.part == 0 ? p.bias2 : nullptr) : p.bias1;       // b2 is added once, by part 0
            for (int j = ew * 32 + lane; j < BN; j += kEpiWarps * 32)
                bias_t[j] = (bias != nullptr && n0 + j < N) ? __ldg(bias + n0 + j) : 0.0f;
            int nchunks = 0;
            for (int c = cbase; c < cbase + cpw; ++c)
                if (n0 + c * 32 < N) ++nchunks;
            const bool work = row0 < p.M && nchunks > 0;
            const int colw = n0 + cbase * 32;
            int* flag_b = flags_b + w.m * p.n2t + w.nt;
            int prefetched = 0;
            auto prefetch_resid = [&]() {
                const uint32_t stg_u32 = smem_u32(wstage);
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                for (; prefetched < 2 && prefetched < nchunks; ++prefetched)
                    resid_issue<true>(rp, prefetched, stg_u32, &tmC2, colw + prefetched * 32, row0, lane);
            };
            // part 0 of a tile of x: its residual boxes travel while the MMAs of this item still run
            if (is_b && w.part == 0 && work) prefetch_resid();
            mbar_wait(tmem_full_bar(acc), acc_ph);
            tc_fence_after_sync();
            if (is_b && w.part > 0 && !(p.dbg & 1)) {
                // later K parts add on top of what the previous part stored: wait for it (fixed order = reproducible sums)
                if (ew == 0 && lane == 0) {
                    const int want = (p.epoch << 3) | w.part;
                    uint32_t spins = 0;
                    while (ld_acquire_gpu(flag_b) != want) {
                        __nanosleep(32);
                        if (++spins > (1u << 22)) {
                            printf("pfasr: ffn chain K-part wait timeout (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                }
            }
            epi_bar_sync();                                       // bias tile visible; previous K part published
            if (is_b && w.part > 0 && work) {
                if (lane == 0) fence_proxy_async_all();
                __syncwarp();
                prefetch_resid();
            }
            if (work) {
                if (is_b)
                    epilogue_tma_f32<cpw, true>(t_acc + cbase * 32, nchunks, wstage, rp, &tmC2, &tmC2, bias_t + cbase * 32, -INFINITY, row0,
                                                colw, lane, prefetched);
                else
                    epilogue_tma_f16<cpw>(t_acc + cbase * 32, nchunks, wstage, &tmC1, bias_t + cbase * 32, 0.0f, row0, colw, lane);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));     // the accumulator stage is free for the item after next
            const bool publish = !is_b || w.part + 1 < p.splits;
            if (publish && !(p.dbg & 2)) {
                // every warp's TMA stores have completed (not merely been read), then one release
                if (lane == 0) tma_store_wait_all();
                __syncwarp();
                epi_bar_sync();
                if (ew == 0 && lane == 0) {
                    fence_proxy_async_all();
                    __threadfence();
                    if (is_b) st_release_gpu(flag_b, (p.epoch << 3) | (w.part + 1));
                    else st_release_gpu(p.flags + w.m * p.n1t + w.nt, p.epoch);
                }
            }
        }
        if (lane == 0) tma_store_wait_read();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, kTmemCols);
}

double tile_us(int kb) { return 0.3 + 0.004 * BN + kb * (0.13 + 0.0011 * BN); }     // gemm.cu wave model

}  // namespace

void ffn_chain_scratch_create(FfnChainScratch& s) {
    s.capacity = 1 << 16;
    PF_CUDA(cudaMalloc(&s.flags, static_cast<size_t>(s.capacity) * sizeof(int)));
    PF_CUDA(cudaMemset(s.flags, 0, static_cast<size_t>(s.capacity) * sizeof(int)));
    s.epoch = 0;
}
void ffn_chain_scratch_destroy(FfnChainScratch& s) {
    if (s.flags) cudaFree(s.flags);
    for (void* p : s.owned) cudaFree(p);
    s = FfnChainScratch{};
}

// Opt-in (PFASR_FFN_CHAIN=1).  Measured on B200 at M = 5344 (profiles/ffn_chain_probe_r01.md): 32 us against 30 us for
// the two separate launches.  The main loops alone would win (24 us with the dependency waits disabled), but a tile of
// h is only visible ~4 us after its MMAs end (epilogue + TMA store completion + release), so the second GEMM's items
// stall on the last columns of h, and its 13 us work units balance poorly over 148 SMs; K-splitting them adds a
// read-modify-write epilogue per part that costs more than it balances.  With several batches in flight (execution
// lanes, abi.cu) the tails this kernel removes are filled by other batches anyway.
bool ffn_chain_enabled() {
    static const bool on = [] { const char* e = getenv("PFASR_FFN_CHAIN"); return e && *e == '1'; }();
    return on;
}

bool ffn_chain_supported(int M, int D, int F, const FfnChainScratch& s) {
    if (!s.flags || M <= 0) return false;
    if (D % BN != 0 || F % BN != 0 || F / BN > kMaxN1T) return false;
    return ceil_div(M, BM) * (F / BN + D / BN) <= s.capacity;
}

void ffn_chain_prepare(FfnChainOp& op, const __half* a16, int lda, const __half* w1, const float* b1, __half* h16, int ldh,
                       const __half* w2, const float* b2, float* x32, int ldx, int M, int D, int F, FfnChainScratch& s) {
    if (!ffn_chain_supported(M, D, F, s)) throw CudaError{"ffn chain: unsupported shape"};
    op = FfnChainOp{};
    op.M = M; op.D = D; op.F = F;
    op.mt = ceil_div(M, BM); op.n1t = F / BN; op.n2t = D / BN;
    op.bias1 = b1; op.bias2 = b2;
    op.scratch = &s;
    gemm_make_tmap(&op.tmA1, a16, false, M, D, lda, BM);
    gemm_make_tmap(&op.tmB1, w1, false, F, D, D, BN);
    gemm_make_tmap(&op.tmC1, h16, false, M, F, ldh, 32, 64);
    gemm_make_tmap(&op.tmA2, h16, false, M, F, ldh, BM);
    gemm_make_tmap(&op.tmB2, w2, false, D, F, F, BN);
    gemm_make_tmap(&op.tmC2, x32, true, M, D, ldx, 32);

    // ---- static schedule (times from the GEMM wave model).  b parts: all part-0 items first, then part 1, ... dealt
    // round-robin from the highest CTA down, so every list holds its parts in ascending part order.  a items: n-major,
    // dealt in waves to the CTAs that still have room; CTAs that own more b work get less room.
    static const int env_splits = [] { const char* e = getenv("PFASR_CHAIN_SPLITK"); return e ? atoi(e) : 0; }();
    const int kb2 = F / BK;
    int splits = env_splits > 0 ? env_splits : 2;
    splits = std::max(1, std::min({splits, kMaxSplits, kb2 / kKbPerATile}));
    op.splits = splits;
    const int na = op.mt * op.n1t, nb = op.mt * op.n2t;
    const int G = std::min(gemm_num_sms(), std::max(na, 1));
    const double ta = tile_us(D / BK), tb = tile_us(kb2 / splits);
    std::vector<std::vector<int>> bl(G), al(G);
    int q = 0;
    for (int part = 0; part < splits; ++part)
        for (int j = 0; j < nb; ++j, ++q) bl[G - 1 - (q % G)].push_back(j | (part << 24) | kItemB);
    const double target = (na * ta + static_cast<double>(nb) * splits * tb) / G;
    static const int a_first = [] { const char* e = getenv("PFASR_CHAIN_A_FIRST"); return e ? atoi(e) : 1; }();
    std::vector<int> room(G);
    int total_room = 0;
    for (int c = 0; c < G; ++c) {
        room[c] = std::max(0, static_cast<int>((target - bl[c].size() * tb) / ta + 0.5));
        if (!bl[c].empty()) room[c] = std::max(room[c], std::min(a_first, ceil_div(na, G)));   // b owners run the first a waves too
        total_room += room[c];
    }
    std::vector<int> spare;                                                 // rounding left tiles over: least b work first
    for (int c = 0; c < G; ++c) spare.push_back(c);
    std::stable_sort(spare.begin(), spare.end(), [&](int x, int y) { return bl[x].size() < bl[y].size(); });
    for (size_t k = 0; total_room < na; ++k) { ++room[spare[k % spare.size()]]; ++total_room; }
    int next = 0;
    for (int w = 0; next < na; ++w)
        for (int c = 0; c < G && next < na; ++c)
            if (room[c] > w) al[c].push_back(next++);
    size_t max_items = 1;
    for (int c = 0; c < G; ++c) max_items = std::max(max_items, al[c].size() + bl[c].size());
    max_items += 1;                                                          // -1 terminator
    std::vector<int> sched(static_cast<size_t>(G) * max_items, -1);
    for (int c = 0; c < G; ++c) {
        size_t k = 0;
        for (int v : al[c]) sched[c * max_items + k++] = v;
        for (int v : bl[c]) sched[c * max_items + k++] = v;
    }
    int* d_sched = nullptr;
    PF_CUDA(cudaMalloc(&d_sched, sched.size() * sizeof(int)));
    PF_CUDA(cudaMemcpy(d_sched, sched.data(), sched.size() * sizeof(int), cudaMemcpyHostToDevice));
    s.owned.push_back(d_sched);
    op.sched = d_sched;
    op.max_items = static_cast<int>(max_items);
    op.grid = G;
    op.valid = true;
}

void ffn_chain_launch(const FfnChainOp& op, cudaStream_t stream) {
    static std::once_flag once;
    std::call_once(once, [] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_ffn_chain_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    ChainParams p;
    p.bias1 = op.bias1; p.bias2 = op.bias2; p.sched = op.sched; p.flags = op.scratch->flags; p.max_items = op.max_items;
    p.epoch = ++op.scratch->epoch;
    p.M = op.M; p.D = op.D; p.F = op.F; p.mt = op.mt; p.n1t = op.n1t; p.n2t = op.n2t; p.splits = op.splits;
    static const int dbg = [] { const char* e = getenv("PFASR_CHAIN_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(op.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    PF_CUDA(cudaLaunchKernelEx(&cfg, pf_ffn_chain_tcgen05, op.tmA1, op.tmB1, op.tmC1, op.tmA2, op.tmB2, op.tmC2, p));
}

}  // namespace pf
