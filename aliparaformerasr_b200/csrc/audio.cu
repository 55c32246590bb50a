// Audio ingestion in front of the fbank kernel.
//
// Replaces AliParaformerAsr.Examples/Utils/AudioHelper.cs:12-32 (GetFileSample) and :223-279 (Resample) together with
// the sample providers of the un-vendored NuGet NAudio 2.2.1 that AudioFileReader chains for PCM WAV data
// (Pcm8/16/24/32BitToSampleProvider: b/128f-1, s16/32768f, s24/8388608f, s32/(Int32.MaxValue+1f); IEEE float as is).
// The caller ships the file's raw sample bytes (half the PCIe traffic of float PCM for 16-bit audio); one kernel
// writes the float PCM the fbank kernel reads.
//
// Reference behaviour kept as is:
//  * a 16 kHz file is NOT down-mixed: GetFileSample only calls Resample when the rate differs, so stereo 16 kHz audio
//    reaches AddSamples interleaved (AudioHelper.cs:27-30);
//  * other rates: (L + R) * 0.5f in float, then linear interpolation in double at position i * (src / 16000), the last
//    input sample repeated from index >= len - 1, length = Math.Round(len / ratio) (round half to even).
#include <math.h>
#include <string.h>

#include "audio.cuh"
#include "engine.cuh"

namespace pf {

namespace {

__device__ __forceinline__ float audio_value(const unsigned char* __restrict__ p, long long j, int format) {
    switch (format) {
        case PF_AUDIO_U8: return __fsub_rn(__fdiv_rn(static_cast<float>(p[j]), 128.0f), 1.0f);
        case PF_AUDIO_S16: return __fdiv_rn(static_cast<float>(reinterpret_cast<const short*>(p)[j]), 32768.0f);
        case PF_AUDIO_S24: {
            const unsigned char* q = p + 3 * j;
            const int v = (static_cast<int>(static_cast<signed char>(q[2])) << 16) | (static_cast<int>(q[1]) << 8) | static_cast<int>(q[0]);
            return __fdiv_rn(static_cast<float>(v), 8388608.0f);
        }
        case PF_AUDIO_S32: return __fdiv_rn(__int2float_rn(reinterpret_cast<const int*>(p)[j]), 2147483648.0f);
        default: return reinterpret_cast<const float*>(p)[j];
    }
}

__global__ void __launch_bounds__(256)
pf_audio_convert(const unsigned char* __restrict__ raw, const AudioItem* __restrict__ items, float* __restrict__ pcm,
                 const long long* __restrict__ pcm_off, const int* __restrict__ nsamp) {
    pdl_launch_dependents();
    pdl_wait();
    const AudioItem it = items[blockIdx.y];
    const int n_out = nsamp[blockIdx.y];
    const unsigned char* src = raw + it.raw_off;
    float* dst = pcm + pcm_off[blockIdx.y];
    const int stride = gridDim.x * blockDim.x;
    if (it.rate == 16000) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) dst[i] = audio_value(src, i, it.format);
        return;
    }
    const bool stereo = it.channels == 2;
    const long long mono_len = stereo ? it.n_values / 2 : it.n_values;
    const double ratio = static_cast<double>(it.rate) / 16000.0;
    auto mono = [&](long long j) -> float {
        if (!stereo) return audio_value(src, j, it.format);
        return __fmul_rn(__fadd_rn(audio_value(src, 2 * j, it.format), audio_value(src, 2 * j + 1, it.format)), 0.5f);
    };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        const double pos = __dmul_rn(static_cast<double>(i), ratio);
        const int idx = static_cast<int>(pos);
        const double frac = __dsub_rn(pos, static_cast<double>(idx));
        float v;
        if (idx >= mono_len - 1) {
            v = mono(mono_len - 1);
        } else {
            // (1 - fraction) * x0 + fraction * x1 in double, no contraction (the C# JIT emits separate mul / add)
            const double a = __dmul_rn(__dsub_rn(1.0, frac), static_cast<double>(mono(idx)));
            const double b = __dmul_rn(frac, static_cast<double>(mono(idx + 1)));
            v = __double2float_rn(__dadd_rn(a, b));
        }
        dst[i] = v;
    }
}

unsigned rd16(const unsigned char* p) { return p[0] | (p[1] << 8); }
unsigned rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<unsigned>(p[3]) << 24); }

}  // namespace

int audio_bytes_per_value(int format) {
    switch (format) {
        case PF_AUDIO_U8: return 1;
        case PF_AUDIO_S16: return 2;
        case PF_AUDIO_S24: return 3;
        case PF_AUDIO_S32: return 4;
        case PF_AUDIO_F32: return 4;
        default: return 0;
    }
}

long long audio_num_samples(const pf_audio& a) {
    if (audio_bytes_per_value(a.format) == 0) throw StatusError{PF_ERR_UNSUPPORTED, "unknown pf_audio.format"};
    if (a.n_values < 0 || (a.n_values > 0 && !a.data)) throw StatusError{PF_ERR_BAD_ARG, "pf_audio: null data or negative length"};
    if (a.sample_rate <= 0) throw StatusError{PF_ERR_BAD_ARG, "sample rate must be positive (ArgumentException in AudioHelper.Resample)"};
    if (a.channels < 1) throw StatusError{PF_ERR_BAD_ARG, "pf_audio.channels must be >= 1"};
    long long n = a.n_values;
    if (a.sample_rate != 16000) {
        // AudioHelper.cs:231-234: only mono or stereo input may be resampled
        if (a.channels != 1 && a.channels != 2) throw StatusError{PF_ERR_BAD_ARG, "only 1 or 2 channels can be resampled (ArgumentException in AudioHelper.Resample)"};
        if (n == 0) return 0;
        const long long mono = a.channels == 2 ? n / 2 : n;
        const double ratio = static_cast<double>(a.sample_rate) / 16000.0;
        const double rounded = nearbyint(static_cast<double>(mono) / ratio);                          // Math.Round: half to even
        if (rounded > 2147483647.0) throw StatusError{PF_ERR_SHAPE, "utterance longer than Int32.MaxValue samples"};
        n = static_cast<long long>(rounded);
    }
    if (n > 0x7fffffffLL) throw StatusError{PF_ERR_SHAPE, "utterance longer than Int32.MaxValue samples"};
    return n;
}

void audio_convert_launch(const unsigned char* raw, const AudioItem* items, float* pcm, const long long* pcm_off,
                          const int* nsamp, int batch, int max_nsamp, cudaStream_t stream) {
    if (batch <= 0 || max_nsamp <= 0) return;
    const int per_block = 256 * 8;
    const int gx = std::max(1, std::min((max_nsamp + per_block - 1) / per_block, 1024));
    launch_k(pf_audio_convert, dim3(gx, batch), dim3(256), 0, stream, raw, items, pcm, pcm_off, nsamp);
}

}  // namespace pf

extern "C" {

// RIFF/WAVE walk: "fmt " (PCM = 1, IEEE float = 3, EXTENSIBLE = 0xFFFE with the tag in the sub-format GUID) and "data"
pf_status pf_wav_parse(const void* file, size_t bytes, pf_audio* out) {
    auto fail = [](pf_status st, const char* msg) {
        pf::set_last_error(msg);
        return st;
    };
    if (!file || !out) return fail(PF_ERR_BAD_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    const unsigned char* p = static_cast<const unsigned char*>(file);
    if (bytes < 12 || memcmp(p, "RIFF", 4) != 0 || memcmp(p + 8, "WAVE", 4) != 0) return fail(PF_ERR_UNSUPPORTED, "not a RIFF/WAVE file");
    size_t pos = 12;
    unsigned tag = 0, channels = 0, rate = 0, bits = 0, align = 0;
    bool have_fmt = false;
    while (pos + 8 <= bytes) {
        const unsigned size = pf::rd32(p + pos + 4);
        const unsigned char* body = p + pos + 8;
        const size_t avail = bytes - (pos + 8);
        if (memcmp(p + pos, "fmt ", 4) == 0) {
            if (size < 16 || avail < 16) return fail(PF_ERR_UNSUPPORTED, "truncated fmt chunk");
            tag = pf::rd16(body);
            channels = pf::rd16(body + 2);
            rate = pf::rd32(body + 4);
            align = pf::rd16(body + 12);
            bits = pf::rd16(body + 14);
            if (tag == 0xFFFE) {
                if (size < 40 || avail < 40) return fail(PF_ERR_UNSUPPORTED, "truncated WAVE_FORMAT_EXTENSIBLE chunk");
                tag = pf::rd16(body + 24);
            }
            have_fmt = true;
        } else if (memcmp(p + pos, "data", 4) == 0) {
            if (!have_fmt) return fail(PF_ERR_UNSUPPORTED, "data chunk before fmt chunk");
            int format = -1;
            if (tag == 1 && bits == 8) format = PF_AUDIO_U8;
            else if (tag == 1 && bits == 16) format = PF_AUDIO_S16;
            else if (tag == 1 && bits == 24) format = PF_AUDIO_S24;
            else if (tag == 1 && bits == 32) format = PF_AUDIO_S32;
            else if (tag == 3 && bits == 32) format = PF_AUDIO_F32;
            if (format < 0) return fail(PF_ERR_UNSUPPORTED, "only PCM 8/16/24/32-bit and 32-bit IEEE float WAV data is decoded on the device");
            if (channels == 0 || rate == 0 || rate > 0x7fffffffu) return fail(PF_ERR_UNSUPPORTED, "fmt chunk without a usable channel count / sample rate");
            const size_t data_bytes = size < avail ? size : avail;        // a streamed file may overstate the length
            const size_t frame = static_cast<size_t>(channels) * (bits / 8);
            (void)align;
            out->data = body;
            out->n_values = static_cast<int64_t>(data_bytes / frame * channels);   // whole frames only
            out->format = format;
            out->channels = static_cast<int32_t>(channels);
            out->sample_rate = static_cast<int32_t>(rate);
            return PF_OK;
        }
        pos += 8 + static_cast<size_t>(size) + (size & 1);                // chunks are word aligned
    }
    return fail(PF_ERR_UNSUPPORTED, "no data chunk");
}

int64_t pf_audio_num_samples(const pf_audio* a) {
    if (!a) return -1;
    try {
        return pf::audio_num_samples(*a);
    } catch (const pf::StatusError& e) {
        pf::set_last_error(e.what);
        return -1;
    }
}

}  // extern "C"
