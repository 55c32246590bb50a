// tcgen05 / TMEM / TMA GEMM used for every dense projection on the path:
//   C[M,N] = A[M,K] (fp16, K-major) x W[N,K]^T (fp16, K-major = torch Linear layout), fp32 accumulate in TMEM,
// with a fused epilogue (bias, ReLU, fp32 residual, fp32 addend, fp16 or fp32 output).
#pragma once
#include "common.cuh"

namespace pf {

struct GemmEpi {
    const float* bias = nullptr;     // [N] fp32 or null
    const float* resid = nullptr;    // fp32 [M, ld_resid] added to the result (may alias out_f32: in-place residual)
    const float* addend = nullptr;   // fp32 [M, ld_addend] second addend (FSMN memory) or null
    float* out_f32 = nullptr;        // exactly one of out_f32 / out_f16 is set
    __half* out_f16 = nullptr;
    int ld_out = 0;
    int ld_resid = 0;
    int ld_addend = 0;
    int relu = 0;
    // fused LayerNorm of the produced rows (fp32 output + residual only): ln_out16 = fp16(LN(out) * gamma + beta), the A
    // operand of the next GEMM.  The row must fit one cluster of at most 4 column tiles in a single wave (gemm_prepare
    // picks the tile width; see gemm_ln_fusable).
    const float* ln_gamma = nullptr;
    const float* ln_beta = nullptr;
    float ln_eps = 0.0f;
    __half* ln_out16 = nullptr;
    int ld_ln16 = 0;
    // greedy pick fused into the epilogue (head GEMMs when the caller does not ask for the log-probs): nothing is stored
    // to out_f32; instead every (row, column group of a tile) writes one partial {best value, best index, last NaN index}
    // to pick_out [M, pick_ld, 3] in ascending column order, combined by pick_combine_launch (ops.cu) with the reference's
    // rule "best = x[best] > x[k] ? best : k" (OfflineRecognizer.cs:145-149, Q5)
    float* pick_out = nullptr;
    int pick_ld = 0;
    // filled by gemm_prepare: the (up to two) fp32 addends in the order the kernel applies them
    const float* add0 = nullptr;
    const float* add1 = nullptr;
    int ld_add0 = 0, ld_add1 = 0;
};

struct GemmOp {
    CUtensorMap tmA;
    CUtensorMap tmB;
    CUtensorMap tmC;      // output (TMA store epilogue), valid when vec_ok & 2
    CUtensorMap tmR;      // fp32 residual (TMA-prefetched into the epilogue staging), valid when vec_ok & 2 and n_adds == 1
    CUtensorMap tmL;      // fp16 fused-LayerNorm output, valid when ln_cluster > 0
    int ln_cluster = 0;   // > 0: fused LayerNorm epilogue, cluster of this many column-tile CTAs per row tile
    GemmEpi epi;
    int M = 0, N = 0, K = 0;
    int bn = 128;    // N tile: 64, 128 or 256
    int cm = 1, cn = 1;   // cm = 2: CTA pair (cluster of 2 along M) driving cta_group::2 MMAs on a 256 x bn tile; cn is always 1
    int throughput = 0;   // prepared under the throughput objective (gemm_set_policy)
    int epi16 = 0;        // sixteen epilogue warps (fp16 output without addends; tile_code bit 22 or PFASR_GEMM_EPI16=1)
    int n_adds = 0;  // fp32 tensors added in the epilogue (0..2)
    int half_sm = 0; // launched as one 128 x 256 tile per CTA, two CTAs per SM (gemm_half.cu)
    int pick = 0;    // epilogue = fused greedy pick (epi.pick_out), no output store
    int red_add = 0; // the residual aliases the output: added by a TMA fp32 reduce-add instead of load + add + store
    int vec_ok = 0;  // bit 0: all epilogue tensors 16-byte aligned with pitches % 4 == 0; bit 1: asynchronous (TMA) epilogue
};

// Build the TMA descriptors for one GEMM.  lda / ldw are in elements and must be multiples of 8 (16 B).
// tile_code = 0 picks tile width and pairing from the problem shape; otherwise bn | (cm << 12); bit 22 asks for the
// sixteen-epilogue-warp variant (fp16 output without addends), bit 23 for the half-SM kernel (256-wide tiles, fp16 output).
void gemm_prepare(GemmOp& op, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  const GemmEpi& epi, int tile_code = 0);
void gemm_launch(const GemmOp& op, cudaStream_t stream);
// half-SM variant (csrc/gemm_half.cu, PFASR_BUILD_EXPERIMENTS=1 builds only): fp16 output, 128 x 256 tile per CTA, two CTAs
// per SM.  PFASR_GEMM_HALFSM = 0 off (default: measured slower), 1 for multi-lane handles, 2 always
bool gemm_half_eligible(const GemmOp& op, bool force = false);
void gemm_half_launch(const GemmOp& op, cudaStream_t stream);
// tile-selection objective of the ops this thread prepares from now on: 0 = shortest kernel (one batch at a time),
// 1 = least SM time (several batches in flight); PFASR_GEMM_POLICY=latency|throughput overrides
void gemm_set_policy(int throughput);
double gemm_flops(const GemmOp& op);
// partial slots per row a fused-pick GEMM over N columns writes at most (tiles are at least 128 wide, two column groups each)
int gemm_pick_slots(int N);
// can a [M, N] fp32 + residual GEMM carry the fused LayerNorm epilogue (row inside one cluster, one wave)?
bool gemm_ln_fusable(int M, int N);
// cuTensorMapEncodeTiled entry point (resolved through the runtime, no libcuda link dependency); throws if missing
void* tensormap_encode_fn();
// 2-D row-major tensor map (fp16 or fp32), box = [box_rows, box_bytes of columns], 128B (or 64B) swizzle
void gemm_make_tmap(CUtensorMap* tm, const void* ptr, bool f32, int rows, int cols, int ld, int box_rows, int box_bytes = 128);
int gemm_num_sms();

}  // namespace pf
