// Fused offline front-end: PCM -> (x32768) Kaldi fbank -> LFR stack -> CMVN -> padded batch tensor.
//
// Replaces, for the offline path of the reference:
//   WavFrontend.GetFbank   /root/reference/AliParaformerAsr/WavFrontend.cs:31-37  (+ SpeechFeatures.OnlineFbank)
//   WavFrontend.ApplyLfr   WavFrontend.cs:73-111   (Q1: three ZERO left-pad frames; Q2: T_lfr = floor(T/n))
//   WavFrontend.ApplyCmvn  WavFrontend.cs:53-71
//   PadHelper.PadSequence  Utils/PadHelper.cs:23-65 (Q4: every exact 0.0 -> -23.0258509f*32768)
//
// One warp per fbank frame: reflect-indexed window load, DC removal, pre-emphasis, Hamming, 512-point real FFT
// (256-point complex radix-2 in shared memory + split post-pass), 80 triangular mel bins, log, then the frame is
// scattered straight into every LFR slot that references it, with CMVN applied on the way out.  HBM traffic =
// read PCM once + write features once.
#include "frontend.cuh"

#include <math.h>

#include <vector>

namespace pf {

namespace {

constexpr int kFrameLen = 400;
constexpr int kFrameShift = 160;
constexpr int kFft = 512;
constexpr int kMel = 80;
constexpr int kWarps = 8;
constexpr int kMaxMelNnz = 1024;

struct FrontendTables {          // device-resident, built once per handle
    float2 tw[256];              // exp(-2*pi*i*k/512), k = 0..255
    float window[kFrameLen];     // Hamming
    int mel_start[kMel];
    int mel_len[kMel];
    int mel_off[kMel];
    float mel_w[kMaxMelNnz];
    unsigned char bitrev[256];
};

__device__ __forceinline__ int reflect_index(int i, int n) {
    // Kaldi ExtractWindow for snip_edges=false: mirror until inside [0, n)
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        else i = 2 * n - 1 - i;
    }
    return i;
}

__global__ void __launch_bounds__(kWarps * 32)
pf_frontend_fbank_lfr_cmvn(const FrontendTables* __restrict__ tab, const float* __restrict__ pcm,
                           const long long* __restrict__ pcm_off, const int* __restrict__ nsamp,
                           const int* __restrict__ nframes, const int* __restrict__ nlfr,
                           const float* __restrict__ add_shift, const float* __restrict__ rescale,
                           float* __restrict__ fbank_out, const long long* __restrict__ fbank_off,
                           float* __restrict__ feats_out, const long long* __restrict__ feats_off,
                           int lfr_m, int lfr_n, int snip_edges, int pad_quirk, float pad_value) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float2 s_tw[256];
    __shared__ float s_win[kFrameLen];
    __shared__ float2 s_z[kWarps][256];
    __shared__ float s_tmp[kWarps][kFft];
    __shared__ unsigned char s_rev[256];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_tw[i] = tab->tw[i];
        s_rev[i] = tab->bitrev[i];
    }
    for (int i = threadIdx.x; i < kFrameLen; i += blockDim.x) s_win[i] = tab->window[i];
    __syncthreads();

    const int f = blockIdx.x * kWarps + warp;
    const int T = nframes[b];
    if (f >= T) return;                       // whole warp exits together
    const int n = nsamp[b];
    const float* x = pcm + pcm_off[b];
    const int start = snip_edges ? f * kFrameShift : f * kFrameShift + (kFrameShift / 2 - kFrameLen / 2);

    float* tmp = s_tmp[warp];
    float2* z = s_z[warp];

    // 1. load + scale (WavFrontend.cs:34 multiplies by 32768f in float32), accumulate the frame mean
    float sum = 0.0f;
    for (int j = lane; j < kFrameLen; j += 32) {
        int idx = start + j;
        if (!snip_edges) idx = reflect_index(idx, n);
        const float v = x[idx] * 32768.0f;
        tmp[j] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    const float mean = sum / static_cast<float>(kFrameLen);
    __syncwarp();

    // 2. remove DC, pre-emphasis 0.97 (x[-1] := x[0]), Hamming; pack even/odd samples into a complex sequence
    //    stored bit-reversed for the in-place radix-2 DIT FFT.
    for (int nn = lane; nn < 256; nn += 32) {
        float re = 0.0f, im = 0.0f;
        const int j0 = 2 * nn, j1 = 2 * nn + 1;
        if (j0 < kFrameLen) {
            const float cur = tmp[j0] - mean;
            const float prev = (j0 == 0 ? tmp[0] : tmp[j0 - 1]) - mean;
            re = (cur - 0.97f * prev) * s_win[j0];
        }
        if (j1 < kFrameLen) {
            const float cur = tmp[j1] - mean;
            const float prev = tmp[j1 - 1] - mean;
            im = (cur - 0.97f * prev) * s_win[j1];
        }
        z[s_rev[nn]] = make_float2(re, im);
    }
    __syncwarp();

    // 3. 256-point complex FFT, 8 radix-2 stages, 128 butterflies per stage = 4 per lane
#pragma unroll 1
    for (int s = 1; s <= 8; ++s) {
        const int m = 1 << s;
        const int half = m >> 1;
        const int tw_stride = 512 >> s;        // index into the 512-th roots table: 2 * pos * (256 / m)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = lane + 32 * i;
            const int grp = j >> (s - 1);
            const int pos = j & (half - 1);
            const int i0 = grp * m + pos;
            const int i1 = i0 + half;
            const float2 w = s_tw[pos * tw_stride];
            const float2 a0 = z[i0];
            const float2 a1 = z[i1];
            const float tr = w.x * a1.x - w.y * a1.y;
            const float ti = w.x * a1.y + w.y * a1.x;
            z[i0] = make_float2(a0.x + tr, a0.y + ti);
            z[i1] = make_float2(a0.x - tr, a0.y - ti);
        }
        __syncwarp();
    }

    // 4. split post-pass -> power spectrum bins 0..255 (Kaldi's mel banks never read the Nyquist bin)
    float* power = tmp;                         // tmp is free again
    for (int k = lane; k < 256; k += 32) {
        const float2 zk = z[k];
        const float2 zc = z[(256 - k) & 255];   // conj applied below
        const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);      // E = (Z[k] + conj(Z[N-k])) / 2
        const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);      // D = (Z[k] - conj(Z[N-k])) / 2
        const float orr = di, oi = -dr;                                        // O = D / i
        const float2 w = s_tw[k];
        const float xr = er + (w.x * orr - w.y * oi);
        const float xi = ei + (w.x * oi + w.y * orr);
        power[k] = xr * xr + xi * xi;
    }
    __syncwarp();

    // 5. mel filterbank + log, then scatter into the LFR slots (CMVN on the way out)
    const int Tl = nlfr[b];
    const int left = (lfr_m - 1) / 2;
    const int p = f + left;                     // index in the left-padded frame sequence
    // LFR frame i reads padded frames [i*n, i*n + m): i ranges over ceil((p-m+1)/n) .. floor(p/n)
    int i_hi = p / lfr_n;
    int i_lo = (p - lfr_m + 1 + lfr_n - 1);
    i_lo = i_lo <= 0 ? 0 : i_lo / lfr_n;
    if (i_hi > Tl - 1) i_hi = Tl - 1;

#pragma unroll 1
    for (int mb = lane; mb < kMel; mb += 32) {
        const int st = tab->mel_start[mb];
        const int ln = tab->mel_len[mb];
        const float* w = tab->mel_w + tab->mel_off[mb];
        float e = 0.0f;
        for (int k = 0; k < ln; ++k) e += w[k] * power[st + k];
        const float v = logf(fmaxf(e, 1.1920928955078125e-07f));
        if (fbank_out) fbank_out[fbank_off[b] + static_cast<long long>(f) * kMel + mb] = v;
        if (feats_out) {
            for (int i = i_lo; i <= i_hi; ++i) {
                const int slot = p - i * lfr_n;
                const int col = slot * kMel + mb;
                float o = (v + add_shift[col]) * rescale[col];
                if (pad_quirk && o == 0.0f) o = pad_value;
                feats_out[feats_off[b] + static_cast<long long>(i) * (lfr_m * kMel) + col] = o;
            }
        }
    }
    // The (m-1)/2 left-pad frames are zeros in the reference (Q1): after CMVN they read shift*scale.
    if (feats_out && f == 0 && Tl > 0) {
        for (int pp = 0; pp < left; ++pp) {
            // padded frame pp belongs to LFR frames i with i*n <= pp < i*n + m
            for (int i = 0; i * lfr_n <= pp && i < Tl; ++i) {
                const int slot = pp - i * lfr_n;
                if (slot >= lfr_m) continue;
                for (int mb = lane; mb < kMel; mb += 32) {
                    const int col = slot * kMel + mb;
                    float o = (0.0f + add_shift[col]) * rescale[col];
                    if (pad_quirk && o == 0.0f) o = pad_value;
                    feats_out[feats_off[b] + static_cast<long long>(i) * (lfr_m * kMel) + col] = o;
                }
            }
        }
    }
}

// Right padding of short utterances: PadSequence pads with 0 and then maps 0 -> pad_value (Q4).
__global__ void pf_frontend_pad_fill(float* __restrict__ feats, const long long* __restrict__ feats_off,
                                     const int* __restrict__ nlfr, int tmax, int dim, float pad_value) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int t0 = nlfr[b];
    const long long total = static_cast<long long>(tmax - t0) * dim;
    float* dst = feats + feats_off[b] + static_cast<long long>(t0) * dim;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = pad_value;
}

double mel_of(double f) { return 1127.0 * log(1.0 + f / 700.0); }

}  // namespace

int frontend_num_frames(int nsamp, bool snip_edges) {
    if (snip_edges) return nsamp < kFrameLen ? 0 : 1 + (nsamp - kFrameLen) / kFrameShift;
    return (nsamp + kFrameShift / 2) / kFrameShift;
}

void* frontend_tables_create() {
    std::vector<unsigned char> raw(sizeof(FrontendTables), 0);
    FrontendTables& t = *reinterpret_cast<FrontendTables*>(raw.data());
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < 256; ++k) {
        t.tw[k].x = static_cast<float>(cos(-2.0 * pi * k / 512.0));
        t.tw[k].y = static_cast<float>(sin(-2.0 * pi * k / 512.0));
        int r = 0;
        for (int bit = 0; bit < 8; ++bit) r |= ((k >> bit) & 1) << (7 - bit);
        t.bitrev[k] = static_cast<unsigned char>(r);
    }
    for (int i = 0; i < kFrameLen; ++i) t.window[i] = static_cast<float>(0.54 - 0.46 * cos(2.0 * pi * i / (kFrameLen - 1)));
    // Kaldi MelBanks: low 20 Hz, high = Nyquist, triangles in the mel domain evaluated at bin centres 0..255
    const double mel_low = mel_of(20.0), mel_high = mel_of(8000.0);
    const double delta = (mel_high - mel_low) / (kMel + 1);
    const double bin_width = 16000.0 / kFft;
    int off = 0;
    for (int b = 0; b < kMel; ++b) {
        const double left = mel_low + b * delta, center = mel_low + (b + 1) * delta, right = mel_low + (b + 2) * delta;
        int first = -1, last = -1;
        std::vector<float> w;
        for (int i = 0; i < kFft / 2; ++i) {
            const double mel = mel_of(bin_width * i);
            if (mel > left && mel < right) {
                const double wt = mel <= center ? (mel - left) / (center - left) : (right - mel) / (right - center);
                if (first < 0) first = i;
                last = i;
                w.push_back(static_cast<float>(wt));
            }
        }
        t.mel_start[b] = first < 0 ? 0 : first;
        t.mel_len[b] = first < 0 ? 0 : last - first + 1;
        t.mel_off[b] = off;
        if (off + static_cast<int>(w.size()) > kMaxMelNnz) throw CudaError{"mel table overflow"};
        for (size_t i = 0; i < w.size(); ++i) t.mel_w[off + i] = w[i];
        off += static_cast<int>(w.size());
    }
    void* d = nullptr;
    PF_CUDA(cudaMalloc(&d, sizeof(FrontendTables)));
    PF_CUDA(cudaMemcpy(d, raw.data(), sizeof(FrontendTables), cudaMemcpyHostToDevice));
    return d;
}

void frontend_tables_destroy(void* tables) {
    if (tables) cudaFree(tables);
}

void frontend_launch(const FrontendLaunch& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_frames <= 0) return;
    dim3 grid(ceil_div(a.max_frames, kWarps), a.batch);
    launch_k(pf_frontend_fbank_lfr_cmvn, grid, dim3(kWarps * 32), 0, stream,
             static_cast<const FrontendTables*>(a.tables), a.pcm, a.pcm_off, a.nsamp, a.nframes, a.nlfr, a.add_shift, a.rescale,
             a.fbank_out, a.fbank_off, a.feats_out, a.feats_off, a.lfr_m, a.lfr_n, a.snip_edges ? 1 : 0, a.pad_quirk ? 1 : 0,
             a.pad_value);
    if (a.feats_out && a.pad_fill && a.tmax_lfr > 0) {
        dim3 g2(32, a.batch);
        launch_k(pf_frontend_pad_fill, g2, dim3(256), 0, stream, a.feats_out, a.feats_off, a.nlfr, a.tmax_lfr, a.lfr_m * kMel, a.pad_value);
    }
}

}  // namespace pf
