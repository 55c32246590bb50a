// Streaming (online) paraformer kernels: chunk assembly, CIF with carry, cached decoder FSMN. See online.cu.
#pragma once
#include "common.cuh"

namespace pf {

// geometry of one 600 ms step (OnlineModel.cs:15-16,30-31; OnlineStream.cs:61,269-273)
struct OnlineDims {
    int mel = 80;
    int lfr_m = 7, lfr_n = 6;
    int d_model = 512;
    int chunk_len = 60;       // fbank frames consumed per decode chunk (_lfr * _chunkSize + 10)
    int nf = 60;              // fbank frames one 9600-sample chunk yields (60 with snip_edges=false, 58 with true)
    int nslot = 16;           // chunk slots in a stream's fbank FIFO
    int t_new = 10;           // LFR rows of a decode chunk (streaming LFR rule on chunk_len + 1 frames)
    int cache_rows = 10;      // feature-cache rows prepended to every window (InitCacheFeats)
};

// tab: [n, 4] int32 = {slot, first logical fbank frame of the window, has_splice, start_idx}
void online_assemble_launch(const OnlineDims& od, const int* tab, int n, const float* fifo, const float* splice,
                            const float* cache_feats, const float* shift, const float* rescale, const float* inv_ts, float* feats,
                            float* fresh, cudaStream_t s);
void online_commit_launch(const OnlineDims& od, const int* tab, int n, const float* fifo, const float* fresh, float* splice,
                          float* cache_feats, cudaStream_t s);
// alphas zeroed outside [mask_lo, mask_hi) (DynamicMask), recurrence of OnlineRecognizer.cs:149-200 over
// [carry | enc frames]; frames_out [n, lcap, D]; counts [n]; meta[0] = max count (atomicMax; zero it first)
void online_cif_launch(const int* tab, int n, const float* enc, const float* alphas, int ld_alpha, int T, int D, int mask_lo,
                       int mask_hi, float threshold, float* carry_alpha, float* carry_hidden, float* frames_out, int lcap,
                       int* counts, int* meta, cudaStream_t s);
void online_compact_launch(const float* frames, int lcap, const int* counts, int n, int L, int D, float* out, cudaStream_t s);
// cache_state: persistent [slots][state_stride] floats, the layer's in-cache at + layer_off ([K-1, D] row-major);
// cache_out: per-step [n][out_stride], this layer's out-cache at + out_layer_off
void online_fsmn_launch(const int* tab, int n, const float* tn, const int* counts, int L, int D, const float* w, int K,
                        const float* cache_state, size_t state_stride, size_t layer_off, float* x, float* cache_out,
                        size_t out_stride, size_t out_layer_off, cudaStream_t s);
void online_scatter_launch(const int* tab, int n, const float* src, float* state, size_t stride, cudaStream_t s);

}  // namespace pf
