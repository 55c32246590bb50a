// SAN-M / cross attention launcher. See attention.cu.
#pragma once
#include "common.cuh"

namespace pf {

// Q: rows (b*Tq + i), K/V: rows (b*Tk + j); head h occupies columns [h*head_dim, (h+1)*head_dim) of each row.
// split_ws (optional, attention_split_workspace_bytes(B, H, Tq) bytes): lets the streaming kernel cut a long memory into runs when there
// are too few query tiles to fill the chip (flash-decoding split + combine kernel); without it every (b, h, query tile) walks all keys.
void attention_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int Tq, int Tk, int ldq,
                      int ldk, int ldv, int ldo, int head_dim, cudaStream_t s, float* split_ws = nullptr, size_t split_ws_bytes = 0);
size_t attention_split_workspace_bytes(int B, int H, int Tq);

// SAN-M self-attention of one encoder layer: context O plus the FSMN memory of V (mem = dwconv_taps(v) + v, fp32).
// One fused tcgen05 kernel when the sequence fits (see below), else the FSMN kernel + the streaming attention kernel.
// mem_accum: add the memory into `mem` (the fp32 residual stream, so the out-projection has one addend) instead of
// overwriting it.  Returns the number of kernels launched.
int attention_fsmn_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int T, int ldqkv, int ldo,
                          const float* fsmn_w, int taps, float* mem, int ld_mem, bool mem_accum, cudaStream_t s);

// One-pass tcgen05 path (attention_tc.cu) for Tk <= 192, Tq <= 256, head_dim 128; optionally also writes the SAN-M
// FSMN memory  mem[b,t,:] = dwconv_taps(v)[b,t,:] + v[b,t,:]  (fp32, row pitch ld_mem) of the same V (self-attention).
bool attention_tc_eligible(int Tq, int Tk, int head_dim);
void attention_tc_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int Tq, int Tk, int ldq,
                         int ldk, int ldv, int ldo, const float* fsmn_w, int taps, float* mem, int ld_mem, bool mem_accum, cudaStream_t s);

}  // namespace pf
