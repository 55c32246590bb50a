// Memory-bound stages of the SAN-M path (everything that is not a dense contraction). See ops.cu.
#pragma once
#include "common.cuh"

namespace pf {

// pe[t, c] = sin / cos((t + 1) * inv_timescales[c mod D/2]) for t < T: the position encoding of every row index, tabulated once
void pe_table_launch(float* pe, int T, int D, const float* inv_timescales, cudaStream_t s);
// x = feats * scale + pe_table[row % T]  ->  LayerNorm(D = input_size)  -> fp16     (encoder input + encoders0.norm1);
// pe_table == nullptr: x = feats (streaming windows arrive scaled and position-encoded)
void embed_pe_ln_launch(const float* feats, int M, int T, int D, float scale, const float* pe_table,
                        const float* gamma, const float* beta, float eps, __half* out16, cudaStream_t s);

// LayerNorm over the last dim (D in {512, 2048}); input fp32 or fp16; writes fp16 and/or fp32.
void layernorm_f32_launch(const float* in, int ld_in, int M, int D, const float* gamma, const float* beta, float eps,
                          __half* out16, int ld16, float* out32, int ld32, cudaStream_t s);

// Depthwise FSMN memory over time: out = resid + mask * (conv_k(mask*in) + mask*in); per-utterance zero padding.
// in: [B*T, ld_in] (fp16 or fp32), w: [D, K] fp32 (K = 11 or 21), lens: [B] valid rows or null (= T).
void fsmn_f16_launch(const __half* in, int ld_in, const float* w, int K, float* out, int ld_out, const float* resid,
                     int ld_res, const int* lens, int B, int T, int D, cudaStream_t s);
void fsmn_f32_launch(const float* in, int ld_in, const float* w, int K, float* out, int ld_out, const float* resid,
                     int ld_res, const int* lens, int B, int T, int D, cudaStream_t s);

// Decoder layer middle, fused: x += mask * (FSMN(LN2(t) * mask) + LN2(t) * mask); out16 = fp16(LN3(x)).  t32 / x: [B*L, 512]
// fp32 (x updated in place), lens: [B] valid token rows, w: [512, K] (K = 11 or 21).
void dec_ln_fsmn_ln_launch(const float* t32, float* x, const float* g2, const float* b2, const float* w, int K, const float* g3,
                           const float* b3, const int* lens, int B, int L, int D, float eps, __half* out16, cudaStream_t s);

// Predictor conv1d(k=3, pad 1|1) as a GEMM: rows [h(t-1) | h(t) | h(t+1)] fp16, zero outside the utterance.
void im2col3_launch(const __half* in, int B, int T, int D, __half* out, cudaStream_t s);

// alpha[b,t] = relu(sigmoid(<h[b,t,:], w> + bias) * smooth - noise); alpha[b,T] = tail.  alphas: [B, T+1].
void alpha_head_launch(const float* h, int B, int T, int D, const float* w, const float* bias, float smooth, float noise,
                       float tail, float* alphas, cudaStream_t s);

// CIF scalar scan over alphas [B, T1]: per step the weight that completes the current token (w_cur), the weight
// carried into the next one (w_rem) and the token index fired at that step (or -1); totals per utterance.
// meta: int[2] = {Lmax over the batch (atomicMax; zero it first), unused}.
void cif_scan_launch(const float* alphas, int B, int T1, float threshold, float* w_cur, float* w_rem, int* fire_idx,
                     float* peaks, int* token_num, int* fires, int* meta, cudaStream_t s);
// acoustic_embeds[b, l, :] from hidden [B, T, D] (the tail step t = T has zero hidden): out [B, Lpad, D] (pre-zeroed).
void cif_gather_launch(const float* hidden, int B, int T, int D, const float* w_cur, const float* w_rem,
                       const int* fire_idx, int T1, float* out, int Lpad, cudaStream_t s);

// In-place log-softmax over [M, V] fp32 (row pitch ld) + greedy pick with the reference's tie rule (last max wins,
// scan restarts after a NaN; OfflineRecognizer.cs:139-152).  tokens: [M] int32.  write_logp = 0 leaves logits raw.
void logsoftmax_argmax_launch(float* logits, int M, int V, int ld, int* tokens, int write_logp, cudaStream_t s);
// greedy ids from the partials of a fused-pick head GEMM ([M, ld, 3] floats, `slots` used per row)
void pick_combine_launch(const float* partials, int M, int ld, int slots, int* tokens, cudaStream_t s);

// Gather rows: dst[b, l, :] = src[b, l, :] for l < L (compacts [B, Lsrc, W] -> [B, L, W]); int32 / fp32 payloads.
void compact_rows_launch(const void* src, void* dst, int B, int Lsrc, int L, int width_bytes, cudaStream_t s);

// SeACo hot-word encoder / merge helpers (EmbedSeacoModel.cs:70-108, OfflineProjOfSeacoParaformer.cs:85-108)
void embed_rows_tmajor_launch(const float* table, const int* ids, int n, int steps, int D, __half* out, cudaStream_t s);
void lstm_cell_launch(const float* gates, float* cstate, int n, int D, __half* h16, __half* rows_hw_major, int t, int steps,
                      cudaStream_t s);
void add_to_f16_launch(const float* a, const float* b, __half* out, size_t n, cudaStream_t s);
void seaco_merge_launch(const int* dha_tok, const float* dha, int nobias, int M, int V, int ld, int* tokens, float* logits,
                        cudaStream_t s);

void f32_to_f16_launch(const float* in, __half* out, size_t n, cudaStream_t s);
// SenseVoice prompt prepend (Q6/Q7): dst[b] = [table[ids[0..3]] ; src[b]]  ([B,T,D] -> [B,T+4,D])
void prepend_rows_launch(const float* src, const float* table, const int* ids, int nprompt, float* dst, int B, int T,
                         int D, cudaStream_t s);

}  // namespace pf
