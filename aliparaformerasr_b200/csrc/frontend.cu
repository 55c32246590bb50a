// Fused offline front-end: PCM -> (x32768) Kaldi fbank -> LFR stack -> CMVN -> padded batch tensor.
//
// Replaces, for the offline path of the reference:
//   WavFrontend.GetFbank   /root/reference/AliParaformerAsr/WavFrontend.cs:31-37  (+ SpeechFeatures.OnlineFbank)
//   WavFrontend.ApplyLfr   WavFrontend.cs:73-111   (Q1: three ZERO left-pad frames; Q2: T_lfr = floor(T/n))
//   WavFrontend.ApplyCmvn  WavFrontend.cs:53-71
//   PadHelper.PadSequence  Utils/PadHelper.cs:23-65 (Q4: every exact 0.0 -> -23.0258509f*32768)
//
// One CTA per chunk of 32 consecutive frames of one utterance; the chunk's samples (reflect-indexed at the utterance
// edges, x32768) are staged ONCE in shared memory - frames overlap by 240 of 400 samples, so this removes the 2.5x
// re-read of the PCM - together with the window, twiddle, mel and CMVN tables.  A warp then takes frames of the chunk in
// turn and keeps the whole 256-point complex FFT (the 512-point real transform, even/odd packed) in REGISTERS: lane L
// holds positions r*32 + L (r = 0..7), radix-2 DIT stages 1..5 pair lanes (shuffle-xor butterflies), stages 6..8 pair
// registers, the split post-pass fetches Z[256-k] with one more shuffle.  Shared memory is touched for the 17 input
// samples a lane needs (consecutive: bit-reversed position r*32+L <-> samples 16*rev5(L) .. +15), for the power spectrum
// handed to the mel filters and for nothing else; the one-pad-word-per-32 layout makes the 64-byte-strided sample and
// window reads conflict-free.  The frame leaves through every LFR slot that references it, CMVN applied on the way out.
// HBM traffic = read PCM once + write features once.
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

constexpr int kChunkFrames = 32;                                             // frames per CTA
constexpr int kChunkSamples = kFrameShift * (kChunkFrames - 1) + kFrameLen;  // 5360 staged samples
constexpr int kMelNnzSmem = 512;                                             // Kaldi's 80 triangles over 256 bins: 501 taps
constexpr int kMaxCmvnDim = 640;

__device__ __forceinline__ int padi(int a) { return a + (a >> 5); }          // one pad word per 32: stride-16 reads hit 32 banks

__device__ __forceinline__ float2 cmul_tw(const float2 w, const float2 a) {  // same expression as the butterfly of the first kernel
    return make_float2(w.x * a.x - w.y * a.y, w.x * a.y + w.y * a.x);
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
    __shared__ float s_x[kChunkSamples + kChunkSamples / 32 + 2];
    __shared__ float s_win[kFrameLen + kFrameLen / 32 + 2];
    __shared__ float2 s_tw[256];
    __shared__ float s_p[kWarps][256];
    __shared__ float s_melw[kMelNnzSmem];
    __shared__ short s_mst[kMel], s_mln[kMel], s_mof[kMel];
    __shared__ float s_shift[kMaxCmvnDim], s_scale[kMaxCmvnDim];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int T = nframes[b];
    const int f0 = blockIdx.x * kChunkFrames;
    if (f0 >= T) return;                                   // the whole CTA leaves before any barrier
    const int n = nsamp[b];
    const float* x = pcm + pcm_off[b];
    const int nfr = min(kChunkFrames, T - f0);
    const int start0 = snip_edges ? f0 * kFrameShift : f0 * kFrameShift + (kFrameShift / 2 - kFrameLen / 2);
    const int ns = kFrameShift * (nfr - 1) + kFrameLen;

    // stage the chunk: scale by 32768 in float32 like WavFrontend.cs:34; Kaldi mirrors indices outside the utterance
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        int idx = start0 + i;
        if (!snip_edges) idx = reflect_index(idx, n);
        s_x[padi(i)] = (idx >= 0 && idx < n) ? x[idx] * 32768.0f : 0.0f;
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tw[i] = tab->tw[i];
    for (int i = threadIdx.x; i < kFrameLen; i += blockDim.x) s_win[padi(i)] = tab->window[i];
    for (int i = threadIdx.x; i < kMelNnzSmem; i += blockDim.x) s_melw[i] = tab->mel_w[i];
    for (int i = threadIdx.x; i < kMel; i += blockDim.x) {
        s_mst[i] = static_cast<short>(tab->mel_start[i]);
        s_mln[i] = static_cast<short>(tab->mel_len[i]);
        s_mof[i] = static_cast<short>(tab->mel_off[i]);
    }
    const int dim = lfr_m * kMel;
    if (feats_out)
        for (int i = threadIdx.x; i < dim; i += blockDim.x) { s_shift[i] = add_shift[i]; s_scale[i] = rescale[i]; }
    __syncthreads();

    // per-lane constants: the 17 consecutive samples behind the lane's 8 bit-reversed positions, and its twiddles
    constexpr unsigned kFull = 0xffffffffu;
    const int brl = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);       // rev5(lane)
    const int jbase = 16 * brl;                                                          // first sample index of the lane
    float2 tws[5];
#pragma unroll
    for (int st = 1; st <= 5; ++st) tws[st - 1] = s_tw[(lane & ((1 << (st - 1)) - 1)) * (512 >> st)];
    const float2 tw6 = s_tw[lane * 8];
    const float2 tw7[2] = {s_tw[lane * 4], s_tw[(32 + lane) * 4]};
    const float2 tw8[4] = {s_tw[lane * 2], s_tw[(32 + lane) * 2], s_tw[(64 + lane) * 2], s_tw[(96 + lane) * 2]};
    float win[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) win[t] = jbase + t < kFrameLen ? s_win[padi(jbase + t)] : 0.0f;

    const int Tl = nlfr[b];
    const int left = (lfr_m - 1) / 2;
    float* power = s_p[warp];

#pragma unroll 1
    for (int fl = warp; fl < nfr; fl += kWarps) {
        const int f = f0 + fl;
        const int foff = fl * kFrameShift;
        // 1. the lane's samples j = jbase - 1 .. jbase + 15 (pre-emphasis needs the previous one; x[-1] := x[0])
        float xs[17];
#pragma unroll
        for (int t = 0; t < 17; ++t) {
            const int jj = jbase - 1 + t;
            xs[t] = (jj >= 0 && jj < kFrameLen) ? s_x[padi(foff + jj)] : 0.0f;
        }
        if (brl == 0) xs[0] = xs[1];
        float sum = 0.0f;
#pragma unroll
        for (int t = 1; t < 17; ++t) sum += xs[t];             // samples past 400 read as 0
        sum = warp_sum(sum);
        const float mean = sum / static_cast<float>(kFrameLen);
        // 2. remove DC, pre-emphasis 0.97, Hamming; even / odd samples packed as re / im at the bit-reversed position
        float y[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const float cur = xs[t + 1] - mean;
            const float prev = xs[t] - mean;
            y[t] = jbase + t < kFrameLen ? (cur - 0.97f * prev) * win[t] : 0.0f;
        }
        float2 z[8];                                           // position r*32 + lane holds input n = rev8(position) = 8*brl + rev3(r)
        z[0] = make_float2(y[0], y[1]);   z[1] = make_float2(y[8], y[9]);
        z[2] = make_float2(y[4], y[5]);   z[3] = make_float2(y[12], y[13]);
        z[4] = make_float2(y[2], y[3]);   z[5] = make_float2(y[10], y[11]);
        z[6] = make_float2(y[6], y[7]);   z[7] = make_float2(y[14], y[15]);
        // 3. stages 1..5: the partner position differs in a lane bit
#pragma unroll
        for (int st = 1; st <= 5; ++st) {
            const int half = 1 << (st - 1);
            const bool upper = (lane & half) != 0;
            const float2 w = tws[st - 1];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float2 o;
                o.x = __shfl_xor_sync(kFull, z[r].x, half);
                o.y = __shfl_xor_sync(kFull, z[r].y, half);
                const float2 a1 = upper ? z[r] : o;
                const float2 a0 = upper ? o : z[r];
                const float2 tt = cmul_tw(w, a1);
                z[r] = upper ? make_float2(a0.x - tt.x, a0.y - tt.y) : make_float2(a0.x + tt.x, a0.y + tt.y);
            }
        }
        // stages 6..8: the partner is another register of the same lane
#pragma unroll
        for (int r = 0; r < 8; r += 2) {
            const float2 tt = cmul_tw(tw6, z[r + 1]);
            const float2 a0 = z[r];
            z[r] = make_float2(a0.x + tt.x, a0.y + tt.y);
            z[r + 1] = make_float2(a0.x - tt.x, a0.y - tt.y);
        }
#pragma unroll
        for (int g = 0; g < 8; g += 4)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = g + q;
                const float2 tt = cmul_tw(tw7[q], z[r + 2]);
                const float2 a0 = z[r];
                z[r] = make_float2(a0.x + tt.x, a0.y + tt.y);
                z[r + 2] = make_float2(a0.x - tt.x, a0.y - tt.y);
            }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 tt = cmul_tw(tw8[r], z[r + 4]);
            const float2 a0 = z[r];
            z[r] = make_float2(a0.x + tt.x, a0.y + tt.y);
            z[r + 4] = make_float2(a0.x - tt.x, a0.y - tt.y);
        }
        // 4. split post-pass -> power spectrum bins k = r*32 + lane (Kaldi's mel banks never read the Nyquist bin);
        //    Z[256 - k] sits in register 7 - r of lane 32 - L (register (8 - r) & 7 of lane 0 when L = 0)
        __syncwarp();                                          // the previous frame's mel reads of `power` are done
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 mine = lane == 0 ? z[(8 - r) & 7] : z[7 - r];
            float2 zc;
            zc.x = __shfl_sync(kFull, mine.x, (32 - lane) & 31);
            zc.y = __shfl_sync(kFull, mine.y, (32 - lane) & 31);
            const float2 zk = z[r];
            const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);      // E = (Z[k] + conj(Z[N-k])) / 2
            const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);      // D = (Z[k] - conj(Z[N-k])) / 2
            const float orr = di, oi = -dr;                                        // O = D / i
            const float2 w = s_tw[r * 32 + lane];
            const float xr = er + (w.x * orr - w.y * oi);
            const float xi = ei + (w.x * oi + w.y * orr);
            power[r * 32 + lane] = xr * xr + xi * xi;
        }
        __syncwarp();

        // 5. mel filterbank + log, then scatter into the LFR slots (CMVN on the way out)
        const int p = f + left;                     // index in the left-padded frame sequence
        // LFR frame i reads padded frames [i*n, i*n + m): i ranges over ceil((p-m+1)/n) .. floor(p/n)
        int i_hi = p / lfr_n;
        int i_lo = (p - lfr_m + 1 + lfr_n - 1);
        i_lo = i_lo <= 0 ? 0 : i_lo / lfr_n;
        if (i_hi > Tl - 1) i_hi = Tl - 1;
#pragma unroll 1
        for (int mb = lane; mb < kMel; mb += 32) {
            const int st = s_mst[mb];
            const int ln = s_mln[mb];
            const float* w = s_melw + s_mof[mb];
            float e = 0.0f;
            for (int k = 0; k < ln; ++k) e += w[k] * power[st + k];
            const float v = logf(fmaxf(e, 1.1920928955078125e-07f));
            if (fbank_out) fbank_out[fbank_off[b] + static_cast<long long>(f) * kMel + mb] = v;
            if (feats_out) {
                for (int i = i_lo; i <= i_hi; ++i) {
                    const int slot = p - i * lfr_n;
                    const int col = slot * kMel + mb;
                    float o = (v + s_shift[col]) * s_scale[col];
                    if (pad_quirk && o == 0.0f) o = pad_value;
                    feats_out[feats_off[b] + static_cast<long long>(i) * dim + col] = o;
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
                        float o = (0.0f + s_shift[col]) * s_scale[col];
                        if (pad_quirk && o == 0.0f) o = pad_value;
                        feats_out[feats_off[b] + static_cast<long long>(i) * dim + col] = o;
                    }
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
        if (off + static_cast<int>(w.size()) > kMelNnzSmem) throw CudaError{"mel table overflow"};
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
    if (a.feats_out && a.lfr_m * kMel > kMaxCmvnDim) throw CudaError{"front-end: lfr_m * 80 exceeds the staged CMVN table"};
    dim3 grid(ceil_div(a.max_frames, kChunkFrames), a.batch);
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
