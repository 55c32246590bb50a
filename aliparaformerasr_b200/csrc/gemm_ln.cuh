// LayerNorm + fp16-output GEMM as one row-tile-stationary kernel (csrc/gemm_ln.cu): out = relu?(LN(x) W^T + bias).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace pf {

struct LnGemmOp {
    bool valid = false;
    const float* x = nullptr;        // fp32 residual stream [M, ld_x], K = 512 columns
    int ld_x = 0;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    float eps = 0.0f;
    CUtensorMap tmB;                 // W [N, K] fp16, 256-row boxes
    const float* bias = nullptr;
    __half* out = nullptr;
    int ld_out = 0;
    int relu = 0;
    int M = 0, N = 0, K = 0;
};

#ifdef PFASR_EXPERIMENTS
int ln_gemm_mode();                  // PFASR_LN_GEMM: 0 off (default), 1 multi-lane handles, 2 always
bool ln_gemm_supported(int M, int N, int K, const void* x, int ld_x, const void* out, int ld_out, const void* bias);
void ln_gemm_prepare(LnGemmOp& op, const float* x, int ld_x, const float* gamma, const float* beta, float eps, const __half* W, int ldw,
                     const float* bias, __half* out, int ld_out, int relu, int M, int N, int K);
void ln_gemm_launch(const LnGemmOp& op, cudaStream_t stream);
double ln_gemm_flops(const LnGemmOp& op);
#else
// Product build: the kernel was measured slower than LayerNorm + GEMM (three lanes 4.14 - 4.40 vs 3.74 - 3.78 ms per step: one CTA per
// row tile is 42 CTAs) and is compiled only with PFASR_BUILD_EXPERIMENTS=1, like the other rejected variants.
inline int ln_gemm_mode() { return 0; }
inline bool ln_gemm_supported(int, int, int, const void*, int, const void*, int, const void*) { return false; }
inline void ln_gemm_prepare(LnGemmOp&, const float*, int, const float*, const float*, float, const __half*, int, const float*, __half*, int, int,
                            int, int, int) {
    throw CudaError{"ln_gemm: the LayerNorm + GEMM kernel needs a PFASR_BUILD_EXPERIMENTS=1 build"};
}
inline void ln_gemm_launch(const LnGemmOp&, cudaStream_t) { throw CudaError{"ln_gemm: the LayerNorm + GEMM kernel needs a PFASR_BUILD_EXPERIMENTS=1 build"}; }
inline double ln_gemm_flops(const LnGemmOp&) { return 0.0; }
#endif

}  // namespace pf
