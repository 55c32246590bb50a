// Caller-side audio ingestion on the device: WAV container walk (host), sample-format
// conversion, stereo -> mono, linear resampling to 16 kHz.  See audio.cu.
#pragma once
#include "common.cuh"
#include "../../include/pf_abi.h"

namespace pf {

// one utterance of raw interleaved audio already resident in HBM
struct AudioItem {
    long long raw_off;     // byte offset of the first value inside the raw buffer
    long long n_values;    // interleaved values (frames * channels)
    int format;            // PF_AUDIO_*
    int channels;
    int rate;
    int pad;
};

int audio_bytes_per_value(int format);                   // 0 for an unknown format
// samples AudioHelper.GetFileSample would hand to AddSamples; throws StatusError for what the reference rejects
long long audio_num_samples(const pf_audio& a);
void audio_convert_launch(const unsigned char* raw, const AudioItem* items, float* pcm, const long long* pcm_off,
                          const int* nsamp, int batch, int max_nsamp, cudaStream_t stream);

}  // namespace pf
