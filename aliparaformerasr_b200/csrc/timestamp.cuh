// Timestamp branch of CifPredictorV3 (row f1): persistent bidirectional LSTM + upsampled alphas / CIF peaks. See timestamp.cu.
#pragma once
#include "common.cuh"

namespace pf {

constexpr int kLstmMaxBatch = 32;        // utterances per launch (= N of the tcgen05 product: 16- and 32-wide variants)

// y[b, t, dir*H + u] for t in [0, T3): h_t of a single-layer bidirectional LSTM (PyTorch gate order i|f|g|o).
//   gin  [B*T3, 2*4H] fp32 : W_ih x_t + b_ih + b_hh for both directions (forward gates first)
//   w_hh [2][4H, H] fp16   : recurrent weights, forward then reverse
//   hbuf, bar              : unused (kept for ABI stability of the launcher)
// One launch of two 16-CTA clusters (forward / reverse): every CTA keeps its 128 rows of W_hh in shared memory for all
// T3 steps, h_t travels through distributed shared memory, one cluster barrier per step.
void bilstm_launch(const float* gin, const __half* w_hh, int B, int T3, int H, float* y, float* hbuf, unsigned int* bar, cudaStream_t s);

// alphas2 = relu(sigmoid(<y[b,t,:], w2> + b2) * smooth - noise); rescaled so that every utterance sums to token_num[b];
// us_cif_peak = running integrate of cif_wo_hidden with threshold thr (integrate -= thr on fire).  Outputs [B, T3].
void us_alphas_peaks_launch(const float* y, int B, int T3, int D2, const float* w2, const float* b2, float smooth, float noise,
                            const int* token_num, float thr, float* us_alphas, float* us_peaks, cudaStream_t s);

}  // namespace pf
