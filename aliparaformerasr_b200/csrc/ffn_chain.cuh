// Position-wise feed-forward block of a SAN-M encoder layer as ONE persistent tcgen05 kernel:
//   h = relu(a W1^T + b1) (fp16)   ->   x += h W2^T + b2 (fp32, in place)
// See ffn_chain.cu.
#pragma once
#include <vector>

#include "gemm.cuh"

namespace pf {

// per compute stream: dependency flags (one per 128 x 256 tile of h) and the launch epoch that validates them
struct FfnChainScratch {
    int* flags = nullptr;
    int capacity = 0;
    int epoch = 0;
    std::vector<void*> owned;      // device schedules of the prepared ops
};
void ffn_chain_scratch_create(FfnChainScratch& s);      // on the current device
void ffn_chain_scratch_destroy(FfnChainScratch& s);

struct FfnChainOp {
    CUtensorMap tmA1, tmB1, tmC1;   // a [M, D] fp16, W1 [F, D] fp16, h [M, F] fp16 (store boxes)
    CUtensorMap tmA2, tmB2, tmC2;   // h [M, F] fp16 (operand boxes), W2 [D, F] fp16, x [M, D] fp32 (residual in, result out)
    const float* bias1 = nullptr;
    const float* bias2 = nullptr;
    int M = 0, D = 0, F = 0;
    int mt = 0, n1t = 0, n2t = 0;   // row tiles, column tiles of h, column tiles of x
    int splits = 1;                 // K parts of a tile of x (accumulated in order)
    const int* sched = nullptr;     // [grid][max_items] work items of every CTA (device)
    int max_items = 0;
    int grid = 0;
    FfnChainScratch* scratch = nullptr;
    bool valid = false;
};

// engine policy: opt-in through PFASR_FFN_CHAIN=1 (measured slower than two launches on the path's shapes, see ffn_chain.cu)
bool ffn_chain_enabled();
// can this shape run as a chain (D, F multiples of 256, flags fit)?
bool ffn_chain_supported(int M, int D, int F, const FfnChainScratch& s);
void ffn_chain_prepare(FfnChainOp& op, const __half* a16, int lda, const __half* w1, const float* b1, __half* h16, int ldh,
                       const __half* w2, const float* b2, float* x32, int ldx, int M, int D, int F, FfnChainScratch& s);
void ffn_chain_launch(const FfnChainOp& op, cudaStream_t stream);
inline double ffn_chain_flops(const FfnChainOp& op) { return 4.0 * op.M * static_cast<double>(op.D) * op.F; }

}  // namespace pf
