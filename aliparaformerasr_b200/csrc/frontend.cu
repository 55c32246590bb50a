// Fused offline front-end: PCM -> (x32768) Kaldi fbank -> LFR stack -> CMVN -> padded batch tensor.
//
// Replaces, for the offline path of the reference:
//   WavFrontend.GetFbank   /root/reference/AliParaformerAsr/WavFrontend.cs:31-37  (+ SpeechFeatures.OnlineFbank)
//   WavFrontend.ApplyLfr   WavFrontend.cs:73-111   (Q1: three ZERO left-pad frames; Q2: T_lfr = floor(T/n))
//   WavFrontend.ApplyCmvn  WavFrontend.cs:53-71
//   PadHelper.PadSequence  Utils/PadHelper.cs:23-65 (Q4: every exact 0.0 -> -23.0258509f*32768)
//
// One CTA per chunk of 32 consecutive frames of one utterance, three phases:
//  A. the chunk's samples (reflect-indexed at the utterance edges, x32768) are staged ONCE in shared memory - frames overlap by
//     240 of 400 samples - with 128-bit loads that are all in flight together (interior chunks), next to the window, twiddle, mel
//     and CMVN tables;
//  B. a warp takes frames of the chunk in turn and keeps the whole 256-point complex FFT (the 512-point real transform, even/odd
//     packed) in REGISTERS: lane L holds positions r*32 + L (r = 0..7), radix-2 DIT stages 1..5 pair lanes - every lane multiplies
//     its own value by (W | 1), exchanges the product with one shuffle-xor and finishes with (-1 | +1) * own + partner, no selects -
//     stages 6..8 pair registers, the split post-pass fetches Z[256-k] with one more shuffle; the power spectrum goes to
//     power[frame][bin] in shared memory (pitch 257);
//  C. mel filterbank + log with lane = FRAME (the taps of a filter are the same for the whole warp: broadcast weight reads, no
//     divergence, conflict-free power reads), then thread = (frame, filter) writes every LFR slot that references the frame, CMVN
//     applied on the way out, as coalesced 320-byte row segments.
// Shared memory holds 17 consecutive samples per lane behind a one-pad-word-per-32 layout (conflict-free 64-byte-strided reads).
// HBM traffic = read PCM once + write features once.  Round 2: 105 -> 75 us per 32 x 10 s under ncu (70.0 M -> 46 M warp
// instructions; the FFT phase is 46 % of them).
#include "frontend.cuh"

#include <mutex>

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

// Dynamic shared memory layout (floats unless noted); ~66 KB per CTA, three CTAs per SM; the log-mel staging of phase 5 / 6 reuses the
// sample buffer (dead once the FFT loop is through)
constexpr int kSxWords = kChunkSamples + kChunkSamples / 32 + 2;
constexpr int kWinWords = kFrameLen + kFrameLen / 32 + 2;
constexpr int kPwPitch = 257;                                               // power[frame][bin]: bank = (frame + bin) % 32 either way
constexpr int kOutPitch = kMel + 1;                                         // log-mel[frame][filter]
constexpr int kOffWin = kSxWords;
constexpr int kOffTw = (kOffWin + kWinWords + 1) & ~1;                      // float2 table: 8-byte aligned
constexpr int kOffPw = kOffTw + 512;
constexpr int kOffMelW = kOffPw + kChunkFrames * kPwPitch;
constexpr int kOffShift = kOffMelW + kMelNnzSmem;
constexpr int kOffScale = kOffShift + kMaxCmvnDim;
constexpr int kOffMelIdx = kOffScale + kMaxCmvnDim;                         // 3 x 80 shorts
static_assert(kChunkFrames * kOutPitch <= kSxWords, "the log-mel staging reuses the sample buffer");
constexpr int kSmemWords = kOffMelIdx + (3 * kMel * 2 + 3) / 4;
constexpr int kFrontendSmem = kSmemWords * 4;

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
    extern __shared__ __align__(16) float s_all[];
    float* s_x = s_all;
    float* s_win = s_all + kOffWin;
    float2* s_tw = reinterpret_cast<float2*>(s_all + kOffTw);
    float* s_pw = s_all + kOffPw;
    float* s_melw = s_all + kOffMelW;
    float* s_shift = s_all + kOffShift;
    float* s_scale = s_all + kOffScale;
    float* s_out = s_all;                                  // (after the barrier that ends the FFT phase)
    short* s_mst = reinterpret_cast<short*>(s_all + kOffMelIdx);
    short* s_mln = s_mst + kMel;
    short* s_mof = s_mln + kMel;

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

    // stage the chunk: scale by 32768 in float32 like WavFrontend.cs:34; Kaldi mirrors indices outside the utterance.
    // Interior chunks (no mirrored or missing sample, 16-byte aligned start) take 128-bit loads, all in flight before the first use.
    if (start0 >= 0 && start0 + ns <= n && ((reinterpret_cast<uintptr_t>(x + start0) & 15) == 0)) {
        const float4* x4 = reinterpret_cast<const float4*>(x + start0);
        constexpr int kVec = (kChunkSamples / 4 + kWarps * 32 - 1) / (kWarps * 32);      // 6 float4 per thread
        const int nv = ns >> 2;                                                          // ns is a multiple of 4 (160 k + 400)
        float4 v[kVec];
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
            const int q = threadIdx.x + k * (kWarps * 32);
            v[k] = q < nv ? __ldg(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
            const int q = threadIdx.x + k * (kWarps * 32);
            if (q < nv) {
                float* d = s_x + 4 * q + (q >> 3);                                       // padi(4q + e) = 4q + e + (q >> 3): 4 | 32
                d[0] = v[k].x * 32768.0f; d[1] = v[k].y * 32768.0f; d[2] = v[k].z * 32768.0f; d[3] = v[k].w * 32768.0f;
            }
        }
    } else {
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            int idx = start0 + i;
            if (!snip_edges) idx = reflect_index(idx, n);
            s_x[padi(i)] = (idx >= 0 && idx < n) ? x[idx] * 32768.0f : 0.0f;
        }
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

    // per-lane constants: the 17 consecutive samples behind the lane's 8 bit-reversed positions, and its twiddles.
    // Lane-crossing stage st pairs lanes L and L ^ 2^(st-1): X[lower] = a0 + W a1, X[upper] = a0 - W a1 with a0 held by the lower
    // lane.  Every lane multiplies its OWN value by wsel (W on the upper lane, exactly 1 on the lower one), exchanges the product
    // and finishes with sgn * own + partner (sgn = -1 upper, +1 lower): no selects, and the same roundings as the textbook form.
    constexpr unsigned kFull = 0xffffffffu;
    const int brl = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);       // rev5(lane)
    const int jbase = 16 * brl;                                                          // first sample index of the lane
    float2 wsel[5];
    float sgn[5];
#pragma unroll
    for (int st = 1; st <= 5; ++st) {
        const bool upper = (lane & (1 << (st - 1))) != 0;
        const float2 w = s_tw[(lane & ((1 << (st - 1)) - 1)) * (512 >> st)];
        wsel[st - 1] = upper ? w : make_float2(1.0f, 0.0f);
        sgn[st - 1] = upper ? -1.0f : 1.0f;
    }
    const float2 tw6 = s_tw[lane * 8];
    const float2 tw7[2] = {s_tw[lane * 4], s_tw[(32 + lane) * 4]};
    const float2 tw8[4] = {s_tw[lane * 2], s_tw[(32 + lane) * 2], s_tw[(64 + lane) * 2], s_tw[(96 + lane) * 2]};
    float win[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) win[t] = jbase + t < kFrameLen ? s_win[padi(jbase + t)] : 0.0f;
    const int src_lane = (32 - lane) & 31;
    const int lane_base = jbase + (jbase >> 5);                                          // padi(jbase)
    const int prev_off = brl == 0 ? 0 : padi(jbase - 1);
    const bool lane_real = jbase < kFrameLen;                                            // 16 | 400: a lane's 16 samples are all real or all past the frame

#pragma unroll 1
    for (int fl = warp; fl < nfr; fl += kWarps) {
        const int foff = fl * kFrameShift;
        // 1. the lane's samples j = jbase - 1 .. jbase + 15 (pre-emphasis needs the previous one; x[-1] := x[0]).  foff is a multiple
        //    of 32, so padi(foff + j) = padi(foff) + padi(j), and padi(jbase + t) = lane_base + t for t = 0..15 (16 | jbase)
        float xs[17];
        const float* px = s_x + (foff + (foff >> 5)) + lane_base;
#pragma unroll
        for (int t = 1; t < 17; ++t) xs[t] = lane_real ? px[t - 1] : 0.0f;                 // samples past 400 read as 0
        xs[0] = brl == 0 ? xs[1] : (brl <= 25 ? s_x[(foff + (foff >> 5)) + prev_off] : 0.0f);
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
            const float2 w = wsel[st - 1];
            const float sg = sgn[st - 1];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 t = cmul_tw(w, z[r]);
                const float ox = __shfl_xor_sync(kFull, t.x, half);
                const float oy = __shfl_xor_sync(kFull, t.y, half);
                z[r] = make_float2(fmaf(sg, t.x, ox), fmaf(sg, t.y, oy));
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
        float* power = s_pw + fl * kPwPitch;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 mine = lane == 0 ? z[(8 - r) & 7] : z[7 - r];
            float2 zc;
            zc.x = __shfl_sync(kFull, mine.x, src_lane);
            zc.y = __shfl_sync(kFull, mine.y, src_lane);
            const float2 zk = z[r];
            const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);      // E = (Z[k] + conj(Z[N-k])) / 2
            const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);      // D = (Z[k] - conj(Z[N-k])) / 2
            const float orr = di, oi = -dr;                                        // O = D / i
            const float2 w = s_tw[r * 32 + lane];
            const float xr = er + (w.x * orr - w.y * oi);
            const float xi = ei + (w.x * oi + w.y * orr);
            power[r * 32 + lane] = xr * xr + xi * xi;
        }
    }
    __syncthreads();

    // 5. mel filterbank + log with lane = FRAME: the taps and their count are the same for the whole warp (weights are broadcast
    //    reads, no divergence), the power rows are read with pitch 257 (conflict-free); warp w takes filters w, w + 8, ...
    if (lane < nfr) {
        const float* prow = s_pw + lane * kPwPitch;
#pragma unroll 1
        for (int mb = warp; mb < kMel; mb += kWarps) {
            const int ln = s_mln[mb];
            const float* w = s_melw + s_mof[mb];
            const float* pp = prow + s_mst[mb];
            float e = 0.0f;
#pragma unroll 4
            for (int k = 0; k < ln; ++k) e += w[k] * pp[k];
            s_out[lane * kOutPitch + mb] = logf(fmaxf(e, 1.1920928955078125e-07f));
        }
    }
    __syncthreads();

    // 6. out: thread = (frame, filter), consecutive threads on consecutive filters (coalesced 320-byte row segments); a frame is
    //    scattered into every LFR slot that references it, CMVN applied on the way out
    const int Tl = nlfr[b];
    const int left = (lfr_m - 1) / 2;
    // LFR frame i reads padded frames [i*n, i*n + m): frame p (left-padded index) belongs to i = ceil((p-m+1)/n) .. floor(p/n).  The two
    // divisions are done once per frame (32 threads) instead of once per element.
    int* s_irange = reinterpret_cast<int*>(s_pw);            // [32][2], the power rows are dead by now
    if (threadIdx.x < nfr) {
        const int p = f0 + threadIdx.x + left;
        int i_hi = p / lfr_n;
        int i_lo = p - lfr_m + lfr_n;
        i_lo = i_lo <= 0 ? 0 : i_lo / lfr_n;
        if (i_hi > Tl - 1) i_hi = Tl - 1;
        s_irange[2 * threadIdx.x] = i_lo;
        s_irange[2 * threadIdx.x + 1] = i_hi;
    }
    __syncthreads();
    if (threadIdx.x < 3 * kMel) {                           // 240 threads: (frame % 3, filter) fixed per thread, frames step by 3
        const int fsub = threadIdx.x / kMel, mb = threadIdx.x - fsub * kMel;
        float* fb = fbank_out ? fbank_out + fbank_off[b] + static_cast<long long>(f0) * kMel + mb : nullptr;
        float* dst = feats_out ? feats_out + feats_off[b] + mb : nullptr;
        for (int fl = fsub; fl < nfr; fl += 3) {
            const float v = s_out[fl * kOutPitch + mb];
            if (fb) fb[fl * kMel] = v;
            if (dst) {
                const int p = f0 + fl + left;               // index in the left-padded frame sequence
                const int i_lo = s_irange[2 * fl], i_hi = s_irange[2 * fl + 1];
                for (int i = i_lo; i <= i_hi; ++i) {
                    const int c0 = (p - i * lfr_n) * kMel;
                    float o = (v + s_shift[c0 + mb]) * s_scale[c0 + mb];
                    if (pad_quirk && o == 0.0f) o = pad_value;
                    dst[static_cast<long long>(i) * dim + c0] = o;
                }
            }
        }
    }
    // The (m-1)/2 left-pad frames are zeros in the reference (Q1): after CMVN they read shift*scale.
    if (feats_out && f0 == 0 && Tl > 0) {
        for (int pp = 0; pp < left; ++pp) {
            // padded frame pp belongs to LFR frames i with i*n <= pp < i*n + m
            for (int i = 0; i * lfr_n <= pp && i < Tl; ++i) {
                const int slot = pp - i * lfr_n;
                if (slot >= lfr_m) continue;
                for (int mb = threadIdx.x; mb < kMel; mb += blockDim.x) {
                    const int col = slot * kMel + mb;
                    float o = (0.0f + s_shift[col]) * s_scale[col];
                    if (pad_quirk && o == 0.0f) o = pad_value;
                    feats_out[feats_off[b] + static_cast<long long>(i) * dim + col] = o;
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
    static std::once_flag attr_once;                              // execution lanes call this from several host threads
    std::call_once(attr_once, [&] {
        int ndev = 0, cur = 0;
        PF_CUDA(cudaGetDeviceCount(&ndev));
        PF_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < ndev; ++d) {
            PF_CUDA(cudaSetDevice(d));
            PF_CUDA(cudaFuncSetAttribute(pf_frontend_fbank_lfr_cmvn, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontendSmem));
        }
        PF_CUDA(cudaSetDevice(cur));
    });
    launch_k(pf_frontend_fbank_lfr_cmvn, grid, dim3(kWarps * 32), static_cast<size_t>(kFrontendSmem), stream,
             static_cast<const FrontendTables*>(a.tables), a.pcm, a.pcm_off, a.nsamp, a.nframes, a.nlfr, a.add_shift, a.rescale,
             a.fbank_out, a.fbank_off, a.feats_out, a.feats_off, a.lfr_m, a.lfr_n, a.snip_edges ? 1 : 0, a.pad_quirk ? 1 : 0,
             a.pad_value);
    if (a.feats_out && a.pad_fill && a.tmax_lfr > 0) {
        dim3 g2(32, a.batch);
        launch_k(pf_frontend_pad_fill, g2, dim3(256), 0, stream, a.feats_out, a.feats_off, a.nlfr, a.tmax_lfr, a.lfr_m * kMel, a.pad_value);
    }
}

}  // namespace pf
