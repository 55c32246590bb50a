// Offline engine: weights, workspace, cached launch plans and the forward passes for one device, plus the
// multi-device handle behind the C-ABI.  See engine.cu / abi.cu.
#pragma once
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pf_abi.h"
#include "attention.cuh"
#include "audio.cuh"
#include "common.cuh"
#ifdef PFASR_EXPERIMENTS
#include "ffn_chain.cuh"
#endif
#include "frontend.cuh"
#include "gemm.cuh"
#include "gemm_ln.cuh"
#include "online.cuh"
#include "ops.cuh"
#include "timestamp.cuh"

namespace pf {

struct StatusError {
    pf_status code;
    std::string what;
};

// ------------------------------------------------------------------ PFW1 weight blob (host view)
struct BlobEntry {
    std::string name;
    int ndim = 0;
    int64_t dims[4] = {1, 1, 1, 1};
    const float* data = nullptr;
    size_t count = 0;
};

class Blob {
public:
    void parse(const void* data, size_t bytes);
    const BlobEntry& get(const std::string& name) const;
    bool has(const std::string& name) const { return entries_.count(name) != 0; }

private:
    std::map<std::string, BlobEntry> entries_;
};

// ------------------------------------------------------------------ device weights
struct LnW {
    float* g = nullptr;
    float* b = nullptr;
};
struct EncLayerW {
    int in_size = 0;
    LnW ln1, ln2;
    __half* w_qkv = nullptr; float* b_qkv = nullptr;
    float* fsmn = nullptr;
    __half* w_out = nullptr; float* b_out = nullptr;
    __half* w_ffn1 = nullptr; float* b_ffn1 = nullptr;
    __half* w_ffn2 = nullptr; float* b_ffn2 = nullptr;
};
struct DecFfnW {
    LnW ln_in;                                   // norm1
    __half* w1 = nullptr; float* b1 = nullptr;
    LnW ln_mid;                                  // feed_forward.norm (width dec_ffn)
    __half* w2 = nullptr;                        // no bias
};
struct DecLayerW {
    DecFfnW ffn;
    LnW ln2, ln3;
    float* fsmn = nullptr;
    __half* wq = nullptr; float* bq = nullptr;
    __half* wo = nullptr; float* bo = nullptr;
};

struct EncLayerPlan {
    GemmOp qkv, out, ffn1, ffn2;
    LnGemmOp qkv_ln, ffn1_ln;        // valid: norm1 + QKV / norm2 + FFN1 run as one row-tile-stationary kernel (csrc/gemm_ln.cu)
#ifdef PFASR_EXPERIMENTS
    FfnChainOp chain;                // valid: ffn1 + ffn2 run as one persistent kernel (csrc/ffn_chain.cu)
#endif
    bool ln2_fused = false;          // norm2 is computed by the out-projection's epilogue
    bool next_ln1_fused = false;     // the next layer's norm1 is computed by this layer's FFN2 epilogue
};
struct DecLayerPlan {
    GemmOp w1, w2, q, out;
    bool next_ln1_fused = false;     // norm1 of the next layer / decoders3 is computed by this layer's out-projection epilogue
};
struct EncoderPlan {
    std::vector<EncLayerPlan> layers;            // encoders0 + encoders + tp_encoders
    GemmOp pred_conv, kv_all, ctc_head, ctc_head_pick;
};
struct DecoderPlan {
    std::vector<DecLayerPlan> layers;
    GemmOp d3_w1, d3_w2, head, head_pick;
    std::vector<DecLayerPlan> slayers;           // SeACo bias decoder
    GemmOp s_d3_w1, s_d3_w2, hw_head;
};
struct LstmW {
    __half* w_ih = nullptr; __half* w_hh = nullptr;
    float* b = nullptr;                          // b_ih + b_hh
};

class Barrier {                                   // reusable host barrier for the per-device worker threads
public:
    explicit Barrier(int n) : n_(n) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const int gen = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen != gen_; });
    }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

struct SharedRun {                                // cross-device exchange for one run (Lmax is a batch-wide max)
    explicit SharedRun(int n) : barrier(n), lmax(n, 0), failed(false) {}
    Barrier barrier;
    std::vector<int> lmax;
    bool failed;
};

class DeviceCtx {
public:
    DeviceCtx(int dev, const pf_config& cfg);
    ~DeviceCtx();
    // share == another context on the same device that already loaded this blob: reuse its device weights (read-only)
    void load_weights(const Blob& blob, const DeviceCtx* share = nullptr);
    bool throughput_mode = false;    // several lanes share this device: GEMM tiles are picked for least SM time, not shortest kernel
    void set_cmvn(const float* shift, const float* scale, int dim);
    void set_hotwords(const int32_t* ids, int n);      // SeACo: [n, 10] padded ids -> bias rows + their K|V (n = 0 clears)

    // inputs
    // audio != null: raw interleaved samples (pf_audio) travel instead of float PCM and are converted on the device;
    // nsamp then holds the converted lengths (audio_num_samples)
    void stage_pcm(const float* const* pcm, const int32_t* nsamp, int B, int tmax_lfr, const pf_audio* audio = nullptr);
    void stage_feats(const float* speech, int B, int T);
    // front-end only (single utterance); returns frames written
    int extract(const float* samples, int nsamp, float* out, int capacity_frames, bool raw_fbank);

    // full forward on the staged batch.  shared/idx: cross-device Lmax exchange (may be null for 1 device).
    void run(uint32_t flags, SharedRun* shared, int idx);
    int staged_batch() const { return staged_B_; }

    // results of the last run (host, pinned)
    int B_ = 0, T_ = 0, Lmax_ = 0, Lpad_ = 0;
    int32_t* h_tokens = nullptr;      // [B, Lmax]
    int32_t* h_token_num = nullptr;   // [B]
    float* h_logits = nullptr;        // [B, Lmax, V] (on request)
    float* h_peaks = nullptr;         // [B, T+1] (on request)
    float* h_us = nullptr;            // [2][B, 3T] us_alphas | us_cif_peak (PF_RUN_WANT_TIMESTAMPS, models with the V3 predictor)
    int us_frames = 0;
    bool has_timestamps() const { return w_up16_ != nullptr; }
    float timings_ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [0..5] device (events), [6..9] host clock since the run began
    int64_t launches = 0;
    double gemm_flops = 0.0;

    // per-launch CUDA-event profiling of the GEMM kernel (bench.py roofline leg); adds two events per launch
    void set_profile(int level) { profile_ = level; }   // 0 off, 1 GEMM launches, 2 every kernel (labelled)
    // Re-launch the GEMMs of the last profiled run back to back (PDL-chained, as inside the step) `iters` times between two
    // CUDA events: ms per pass.  The per-launch events of set_profile(1) break the programmatic launch chain and add an
    // event round trip to every launch; this measures the kernel's launch duration the way the step experiences it.
    double replay_gemms(int iters);
    struct ReplayOp { bool fused = false; GemmOp g; LnGemmOp l; };
    std::vector<ReplayOp> replay_;
    double gemm_ms = 0.0;              // sum of GEMM launch durations of the last profiled run
    std::string profile_json;          // per-shape breakdown of the last profiled run

    // ---- streaming (online) path: per-stream state lives in HBM, indexed by slot (OnlineStream.cs:24-37 state,
    // OnlineRecognizer.Forward OnlineRecognizer.cs:341-401).  The sample FIFO / chunk trigger (Q13) is host logic in abi.cu.
    int online_open();
    void online_close(int slot);
    void online_push_chunk(int slot, const float* samples, int nsamp);    // one 160 * chunk_len sample chunk (InputSpeech)
    bool online_ready(int slot) const;                                    // GetDecodeChunk would return a window
    // one Forward over `slots` (caller order); streams without a full chunk are skipped like the reference does.
    // Results: online_working (indices into slots), h_tokens [n_working, Lmax_], h_online_counts [n_working].
    void online_step(const std::vector<int>& slots, uint32_t flags, SharedRun* shared, int idx);
    std::vector<int> online_working;
    int32_t* h_online_counts = nullptr;
    const OnlineDims& online_dims() const { return od_; }
    void online_get_state(int slot, const std::string& name, float* dst, size_t capacity);   // test hook

    void get_tensor(const std::string& name, float* dst, size_t capacity, int32_t* dims4, int32_t* ndim);
    cudaStream_t stream() const { return stream_; }
    void sync_staging();               // staged inputs are resident: the caller may release its host buffers
    int device() const { return dev_; }
    int ldv() const { return (cfg_.vocab + 3) & ~3; }   // fp32 logits row pitch (16-byte aligned rows)

private:
    template <typename T> T* dalloc(size_t n, std::vector<void*>& pool);
    float* up_f32(const BlobEntry& e);
    float* up_f32(const float* host, size_t n);
    __half* up_f16(const float* host, size_t n);
    LnW up_ln(const Blob& b, const std::string& prefix);
    void load_enc_layer(const Blob& b, const std::string& prefix, int in_size, EncLayerW& w);
    void load_dec_ffn(const Blob& b, const std::string& prefix, DecFfnW& w, int ffn_width);
    void load_dec_stack(const Blob& b, const std::string& prefix, int nlayers, int ffn_width, int kernel, std::vector<DecLayerW>& layers,
                        DecFfnW& d3, LnW& after, __half*& w_kv_all, float*& b_kv_all);
    void dec_stack(const std::vector<DecLayerW>& layers, const DecFfnW& d3, const std::vector<DecLayerPlan>& lps, const GemmOp& d3_w1,
                   const GemmOp& d3_w2, int ffn_width, int kernel, const __half* kv16, int ldkv, bool kv_shared, int Tk, int B, int L,
                   bool online);
    void seaco_forward(int B, int L, DecoderPlan& plan);
    void timestamp_forward(int B, int T);

    void ensure_workspace(int B, int T);
    void ensure_host(size_t tokens, size_t logits, size_t peaks);
    EncoderPlan& encoder_plan(int B, int T);
    DecoderPlan& decoder_plan(int B, int T, int L);
    void gemm(const GemmOp& op);
    void ln_gemm(const LnGemmOp& op);
    void encoder_forward(int B, int T, bool online = false);
    void predictor_forward(int B, int T, bool online = false, bool kv_later = false);
    void decoder_forward(int B, int T, int L, bool online = false);
    void online_step_impl(const std::vector<int>& slots, uint32_t flags, SharedRun* shared, int idx);
    void online_grow(int min_slots);
    void free_pool(std::vector<void*>& pool);
    void run_impl(uint32_t flags, SharedRun* shared, int idx);
    bool arrived_ = false, ev0_armed_ = false;
    void finish_profile();
    struct ProfRec { int M, N, K, bn; cudaEvent_t a, b; const char* label; };
    int profile_ = 0;
    template <typename F> void timed(const char* label, F&& f);   // level-2 profiling of a non-GEMM launch
    std::vector<ProfRec> prof_;
    std::vector<cudaEvent_t> prof_pool_;

    int dev_;
    pf_config cfg_;
    static constexpr int kCopyGroups = 4;
    cudaStream_t stream_ = nullptr;
    cudaStream_t copy_stream_ = nullptr;      // H2D of the PCM, overlapped with the front-end of earlier utterance groups
    cudaEvent_t ev_[7] = {};
    cudaEvent_t ev_grp_[kCopyGroups] = {};
    cudaEvent_t ev_compute_ = nullptr;
    int staged_groups_ = 0;
    std::vector<void*> wpool_, apool_;      // wpool_: every weight upload, in load order
    const DeviceCtx* share_from_ = nullptr; size_t share_idx_ = 0, wload_begin_ = 0;
    std::vector<void*> tmp_;           // upload staging, freed after load

    // weights
    std::vector<EncLayerW> enc_, tp_;
    LnW after_norm_, tp_norm_;
    __half* w_conv_ = nullptr; float* b_conv_ = nullptr;
    float* w_alpha_ = nullptr; float* b_alpha_ = nullptr;
    std::vector<DecLayerW> dec_;
    __half* w_kv_all_ = nullptr; float* b_kv_all_ = nullptr;
    DecFfnW dec3_;
    LnW dec_after_;
    __half* w_head_ = nullptr; float* b_head_ = nullptr;
    // SeACo: bias decoder, hot-word head, hot-word encoder
    std::vector<DecLayerW> sdec_;
    DecFfnW sdec3_;
    LnW sdec_after_;
    __half* w_skv_all_ = nullptr; float* b_skv_all_ = nullptr;
    __half* w_hw_head_ = nullptr; float* b_hw_head_ = nullptr;
    float* bias_table_ = nullptr;      // bias_embed.weight [V, d]
    LstmW lstm_[2];
    int nbias_ = 0;                    // rows of bias_embed (10 per hot word, Q8); 0 = no hot words set
    __half* bias16_ = nullptr; __half* skv16_ = nullptr;
    std::vector<void*> hpool_;
    float* emb32_ = nullptr; float* hid32_ = nullptr; float* satt32_ = nullptr; float* logits2_ = nullptr; int* tokens2_ = nullptr;
    // CifPredictorV3 timestamp branch (optional: present when the blob has predictor.upsample_cnn.*)
    __half* w_up16_ = nullptr; float* b_up_ = nullptr;          // ConvTranspose1d as a [3*d, d] GEMM
    __half* w_ih_bi16_ = nullptr; float* b_bi_ = nullptr;       // BiLSTM input projections, forward | reverse
    __half* w_hh_bi16_ = nullptr;                               // [2][4d, d]
    float* w_out2_ = nullptr; float* b_out2_ = nullptr;         // cif_output2
    float* gin32_ = nullptr; float* y32_ = nullptr; float* hbuf_ = nullptr; unsigned int* lstm_bar_ = nullptr;
    float* us_alphas_ = nullptr; float* us_peaks_ = nullptr;
    std::vector<void*> tpool_;
    int ts_capB_ = 0, ts_capT_ = 0;
    size_t h_us_cap_ = 0;
    float* embed_table_ = nullptr;     // SenseVoice prompt table [16, input_size]
    float* inv_ts_ = nullptr;          // PE inverse timescales [input_size/2]
    float* xatt_ws_ = nullptr;         // split workspace of the SeACo bias decoder's cross-attention (attention_split_workspace_bytes)
    size_t xatt_ws_bytes_ = 0;
    float* pe_table_ = nullptr;        // [pe_rows_, input_size] sin | cos((t + 1) * inv_ts): built when a longer batch arrives
    int pe_rows_ = 0;
    float* cmvn_shift_ = nullptr; float* cmvn_scale_ = nullptr;
    void* fe_tables_ = nullptr;

    // workspace (capacity in rows)
    int capB_ = 0, capT_ = 0;
    float* feats_ = nullptr; float* feats_raw_ = nullptr;
    __half* a16_ = nullptr; __half* qkv16_ = nullptr; __half* ctx16_ = nullptr; __half* h16_ = nullptr;
    float* mem32_ = nullptr; float* x32_ = nullptr; float* enc32_ = nullptr; __half* enc16_ = nullptr;
    __half* kv16_ = nullptr;
    float* alphas_ = nullptr; float* wcur_ = nullptr; float* wrem_ = nullptr; float* peaks_ = nullptr;
    int* fire_idx_ = nullptr; int* token_num_ = nullptr; int* fires_ = nullptr; int* meta_ = nullptr;
    float* xd32_ = nullptr; __half* ad16_ = nullptr; float* hd32_ = nullptr; __half* hd16_ = nullptr;
    float* t32_ = nullptr; float* tn32_ = nullptr; __half* q16_ = nullptr; __half* ctxd16_ = nullptr;
    float* logits_ = nullptr; int* tokens_ = nullptr;
    float* pick_ = nullptr; int pick_ld_ = 0;    // partials of the fused greedy pick (head GEMM epilogue)
    bool want_logits_ = false;                   // this run returns log-probs: the head GEMM stores them and the pick is a kernel
    int* prompt_ids_ = nullptr;
#ifdef PFASR_EXPERIMENTS
    FfnChainScratch chain_;                  // dependency flags of the fused feed-forward kernel (this device's compute stream)
    void ffn_chain(const FfnChainOp& op);
#endif
    // staged PCM
    float* pcm_ = nullptr; size_t pcm_cap_ = 0;
    long long* d_off_ = nullptr; int* d_meta_ = nullptr; int meta_capB_ = 0;   // per-utterance tables
    void* h_stage_ = nullptr; size_t h_stage_bytes_ = 0;
    int staged_B_ = 0, staged_T_ = 0, staged_maxframes_ = 0;
    bool staged_is_pcm_ = false;
    // staged raw audio (row f4): bytes as read from the file + one AudioItem per utterance
    unsigned char* raw_ = nullptr; size_t raw_cap_ = 0;
    AudioItem* d_audio_ = nullptr; int audio_capB_ = 0;
    void* h_audio_ = nullptr; size_t h_audio_bytes_ = 0;
    bool staged_audio_ = false;
    int staged_max_nsamp_ = 0;

    // host results
    int* h_meta_ = nullptr;
    size_t h_tokens_cap_ = 0, h_logits_cap_ = 0, h_peaks_cap_ = 0, h_tn_cap_ = 0;

    // streaming state (capacity ocap_ slots)
    struct OnlineSlot { bool open = false; long long pushed = 0, fbanked = 0, dec = 0; };
    OnlineDims od_;
    std::vector<OnlineSlot> oslots_;
    int ocap_ = 0;
    float* ofifo_ = nullptr; float* opcm_ = nullptr; float* osplice_ = nullptr; float* ocache_ = nullptr;
    float* ocif_a_ = nullptr; float* ocif_h_ = nullptr; float* ofsmn_ = nullptr;
    float* inv_ts_online_ = nullptr;
    float* ofresh_ = nullptr; float* ocache_new_ = nullptr; int* otab_ = nullptr; int owork_cap_ = 0;
    long long* ofe_off_ = nullptr; int* ofe_meta_ = nullptr; int ofe_cap_ = 0;
    void* h_ostage_ = nullptr; size_t h_ostage_bytes_ = 0;
    float* h_push_ring_ = nullptr; int push_ring_pos_ = 0; bool push_pending_ = false;   // pinned bounce ring of pushed chunks
    size_t h_ocounts_cap_ = 0;
    std::vector<void*> opool_;
    size_t fsmn_state_stride() const { return static_cast<size_t>(cfg_.dec_layers) * (cfg_.dec_kernel - 1) * cfg_.d_model; }

    std::map<std::pair<int, int>, EncoderPlan> enc_plans_;
    std::map<std::pair<int, int>, DecoderPlan> dec_plans_;   // key (B*capT-independent: B, L) for current T
    int dec_plan_T_ = -1;
};

struct OfflineHandle {
    pf_config cfg;
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    std::recursive_mutex mu;          // recursive: a thread that leased the lane (pf_offline_lane_acquire) keeps calling into it
    // last run layout
    std::vector<int> shard_begin, shard_count;
    int B = 0, Lmax = 0, T = 0;
    std::vector<int32_t> tokens, token_num;
    std::vector<float> logits, peaks, us;
    bool staged = false;
    bool staged_pcm = false;
};

struct OnlineStreamHost {             // host half of OnlineStream: the sample cache of AddSamples (OnlineStream.cs:84-112)
    bool open = false;
    int dev = 0, slot = -1;
    std::vector<float> cache_samples;
};

struct OnlineHandle {
    pf_config cfg;
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    std::recursive_mutex mu;
    std::vector<OnlineStreamHost> streams;
    // last step
    std::vector<int32_t> appended, new_tokens, embeds_len;
    std::vector<float> logits;
};

}  // namespace pf
