// Fused fbank + LFR + CMVN (+ pad quirk) front-end launcher. See frontend.cu.
#pragma once
#include "common.cuh"

namespace pf {

struct FrontendLaunch {
    const void* tables = nullptr;        // frontend_tables_create()
    const float* pcm = nullptr;          // packed device PCM, utterance b at pcm + pcm_off[b]
    const long long* pcm_off = nullptr;  // [B] device
    const int* nsamp = nullptr;          // [B] device
    const int* nframes = nullptr;        // [B] device, fbank frames per utterance
    const int* nlfr = nullptr;           // [B] device, LFR frames per utterance
    const float* add_shift = nullptr;    // [lfr_m*80] device (am.mvn <AddShift>)
    const float* rescale = nullptr;      // [lfr_m*80] device (am.mvn <Rescale>)
    float* fbank_out = nullptr;          // optional raw fbank, utterance b at fbank_out + fbank_off[b]
    const long long* fbank_off = nullptr;
    float* feats_out = nullptr;          // optional LFR+CMVN features, utterance b at feats_out + feats_off[b]
    const long long* feats_off = nullptr;
    int batch = 0;
    int max_frames = 0;                  // max over b of nframes[b]
    int tmax_lfr = 0;                    // rows per utterance in feats_out when pad_fill
    int lfr_m = 7, lfr_n = 6;
    bool snip_edges = false;
    bool pad_quirk = false;              // Q4: exact zeros -> pad_value
    bool pad_fill = false;               // fill rows [nlfr[b], tmax_lfr) with pad_value
    float pad_value = 0.0f;
};

int frontend_num_frames(int nsamp, bool snip_edges);
void* frontend_tables_create();          // on the current device
void frontend_tables_destroy(void* tables);
void frontend_launch(const FrontendLaunch& a, cudaStream_t stream);

}  // namespace pf
