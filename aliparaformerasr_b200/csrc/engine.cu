// Per-device offline engine: PFW1 blob -> device weights, workspace, cached TMA launch plans, forward passes.
//
// Replaces the InferenceSession the reference builds in OfflineModel.initModel
// (/root/reference/AliParaformerAsr/OfflineModel.cs:35-70) and runs in IOfflineProj.ModelProj
// (OfflineProjOfParaformer.cs:39-87, OfflineProjOfSenseVoiceSmall.cs:53-175).
#include "engine.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

namespace pf {

// PFASR_DBG_SKIP (tuning aid, results are garbage): bit 0 encoder LayerNorms, 1 encoder attention, 2 decoder LayerNorms,
// 3 decoder FSMN, 4 decoder cross attention - skipped launches show what each family really costs inside the pipelined step
static int dbg_skip() {
    static const int v = [] { const char* e = getenv("PFASR_DBG_SKIP"); return e ? atoi(e) : 0; }();
    return v;
}

// The fused-LayerNorm GEMM epilogue (gemm.cu: cluster statistics exchange + second pass over TMEM) is parity-green but
// measured SLOWER than GEMM + LayerNorm kernel on this path (out-projection 12.7 us fused vs 8.6 + 3.2 us; cfg2 step 5.32
// vs 5.17 ms): the second pass runs on 8 warps per SM where the stand-alone LayerNorm spreads 5312 rows over every warp
// slot of the chip.  It therefore stays opt-in (PFASR_LN_FUSE=1).
static bool ln_fuse_enabled() {
#ifdef PFASR_EXPERIMENTS
    static const bool on = [] { const char* e = getenv("PFASR_LN_FUSE"); return e && *e && *e != '0'; }();
    return on;
#else
    return false;                 // the product build carries only the selected path (build.py: PFASR_BUILD_EXPERIMENTS)
#endif
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("PFASR_NO_PDL");
        return !(e && *e && *e != '0');
    }();
    return on;
}

// ------------------------------------------------------------------ blob
// Layout (little endian): "PFW1" | u32 version | u32 count | u32 reserved |
//   count x { char name[96]; u32 dtype(0=f32); u32 ndim; u64 dims[4]; u64 offset; u64 nbytes } | payload
namespace {
struct RawEntry {
    char name[96];
    uint32_t dtype;
    uint32_t ndim;
    uint64_t dims[4];
    uint64_t offset;
    uint64_t nbytes;
};
static_assert(sizeof(RawEntry) == 96 + 8 + 32 + 16, "PFW1 entry layout");
}  // namespace

void Blob::parse(const void* data, size_t bytes) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    if (bytes < 16 || memcmp(p, "PFW1", 4) != 0) throw StatusError{PF_ERR_WEIGHTS, "weights blob: bad magic (expected PFW1)"};
    uint32_t version, count;
    memcpy(&version, p + 4, 4);
    memcpy(&count, p + 8, 4);
    if (version != 1) throw StatusError{PF_ERR_WEIGHTS, "weights blob: unsupported version"};
    if (16 + static_cast<size_t>(count) * sizeof(RawEntry) > bytes) throw StatusError{PF_ERR_WEIGHTS, "weights blob: truncated table"};
    for (uint32_t i = 0; i < count; ++i) {
        RawEntry e;
        memcpy(&e, p + 16 + static_cast<size_t>(i) * sizeof(RawEntry), sizeof(RawEntry));
        e.name[95] = 0;
        if (e.dtype != 0 || e.ndim > 4) throw StatusError{PF_ERR_WEIGHTS, std::string("weights blob: bad entry ") + e.name};
        if (e.offset > bytes || e.nbytes > bytes - e.offset || (e.offset & 3))      // (no offset + nbytes: it can wrap)
            throw StatusError{PF_ERR_WEIGHTS, std::string("weights blob: bad extent ") + e.name};
        BlobEntry be;
        be.name = e.name;
        be.ndim = static_cast<int>(e.ndim);
        size_t cnt = 1;
        for (int d = 0; d < 4; ++d) {
            if (d < be.ndim && e.dims[d] > (1ull << 40)) throw StatusError{PF_ERR_WEIGHTS, std::string("weights blob: bad shape ") + e.name};
            be.dims[d] = d < be.ndim ? static_cast<int64_t>(e.dims[d]) : 1;
            if (be.dims[d] != 0 && cnt > (~static_cast<size_t>(0) / 8) / static_cast<size_t>(be.dims[d]))
                throw StatusError{PF_ERR_WEIGHTS, std::string("weights blob: bad shape ") + e.name};
            cnt *= static_cast<size_t>(be.dims[d]);
        }
        if (cnt * 4 != e.nbytes) throw StatusError{PF_ERR_WEIGHTS, std::string("weights blob: size mismatch ") + e.name};
        be.count = cnt;
        be.data = reinterpret_cast<const float*>(p + e.offset);
        entries_[be.name] = be;
    }
}

const BlobEntry& Blob::get(const std::string& name) const {
    auto it = entries_.find(name);
    if (it == entries_.end()) throw StatusError{PF_ERR_WEIGHTS, "weights blob: missing tensor '" + name + "'"};
    return it->second;
}

static void expect_shape(const BlobEntry& e, std::initializer_list<int64_t> dims) {
    size_t n = 1;
    for (auto d : dims) n *= static_cast<size_t>(d);
    if (n != e.count) {
        std::string want;
        for (auto d : dims) want += std::to_string(d) + " ";
        throw StatusError{PF_ERR_WEIGHTS, "weights blob: tensor '" + e.name + "' has " + std::to_string(e.count) +
                                              " elements, expected [" + want + "]"};
    }
}

// ------------------------------------------------------------------ DeviceCtx basics
DeviceCtx::DeviceCtx(int dev, const pf_config& cfg) : dev_(dev), cfg_(cfg) {
    PF_CUDA(cudaSetDevice(dev_));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, dev_));
    if (prop.major != 10)
        throw StatusError{PF_ERR_CUDA, "device " + std::to_string(dev_) + " (" + prop.name + ", sm_" + std::to_string(prop.major) +
                                           std::to_string(prop.minor) + ") is not sm_100: libpfasr has no fallback path"};
    PF_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    PF_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (auto& e : ev_) PF_CUDA(cudaEventCreate(&e));
    for (auto& e : ev_grp_) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreateWithFlags(&ev_compute_, cudaEventDisableTiming));
    fe_tables_ = frontend_tables_create();
#ifdef PFASR_EXPERIMENTS
    ffn_chain_scratch_create(chain_);
#endif
    // FunASR SinusoidalPositionEncoder: inv_timescale_i = exp(-i * ln(1e4) / (depth/2 - 1)), float32 arithmetic
    const int half = cfg_.input_size / 2;
    std::vector<float> inv(half);
    const float inc = static_cast<float>(log(10000.0) / (half - 1));
    for (int i = 0; i < half; ++i) inv[i] = expf(static_cast<float>(i) * -inc);
    inv_ts_ = up_f32(inv.data(), inv.size());
    PF_CUDA(cudaMallocHost(&h_meta_, 16 * sizeof(int)));
    const int dim = cfg_.lfr_m * cfg_.n_mels;
    std::vector<float> zeros(dim, 0.0f), ones(dim, 1.0f);
    cmvn_shift_ = up_f32(zeros.data(), dim);
    cmvn_scale_ = up_f32(ones.data(), dim);
}

void DeviceCtx::free_pool(std::vector<void*>& pool) {
    for (void* p : pool) cudaFree(p);
    pool.clear();
}

DeviceCtx::~DeviceCtx() {
    cudaSetDevice(dev_);
    if (stream_) cudaStreamSynchronize(stream_);
    free_pool(wpool_);
    free_pool(apool_);
    free_pool(hpool_);
    free_pool(tpool_);
    if (h_us) cudaFreeHost(h_us);
    if (pe_table_) cudaFree(pe_table_);
    if (xatt_ws_) cudaFree(xatt_ws_);
    free_pool(tmp_);
    frontend_tables_destroy(fe_tables_);
#ifdef PFASR_EXPERIMENTS
    ffn_chain_scratch_destroy(chain_);
#endif
    if (pcm_) cudaFree(pcm_);
    if (raw_) cudaFree(raw_);
    if (d_audio_) cudaFree(d_audio_);
    if (h_audio_) cudaFreeHost(h_audio_);
    if (d_off_) cudaFree(d_off_);
    if (d_meta_) cudaFree(d_meta_);
    if (h_stage_) cudaFreeHost(h_stage_);
    if (h_meta_) cudaFreeHost(h_meta_);
    if (h_tokens) cudaFreeHost(h_tokens);
    if (h_token_num) cudaFreeHost(h_token_num);
    if (h_logits) cudaFreeHost(h_logits);
    if (h_peaks) cudaFreeHost(h_peaks);
    if (h_online_counts) cudaFreeHost(h_online_counts);
    if (h_ostage_) cudaFreeHost(h_ostage_);
    if (h_push_ring_) cudaFreeHost(h_push_ring_);
    for (float* p : {ofifo_, opcm_, osplice_, ocache_, ocif_a_, ocif_h_, ofsmn_}) if (p) cudaFree(p);
    if (ofe_off_) cudaFree(ofe_off_);
    if (ofe_meta_) cudaFree(ofe_meta_);
    free_pool(opool_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
    for (auto& e : ev_grp_) if (e) cudaEventDestroy(e);
    if (ev_compute_) cudaEventDestroy(ev_compute_);
    if (copy_stream_) { cudaStreamSynchronize(copy_stream_); cudaStreamDestroy(copy_stream_); }
    for (auto& e : prof_pool_) cudaEventDestroy(e);
    for (auto& r : prof_) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (stream_) cudaStreamDestroy(stream_);
}

template <typename T>
T* DeviceCtx::dalloc(size_t n, std::vector<void*>& pool) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw StatusError{PF_ERR_OOM, "cudaMalloc of " + std::to_string(n * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e)};
    }
    pool.push_back(p);
    return static_cast<T*>(p);
}

float* DeviceCtx::up_f32(const float* host, size_t n) {
    // a lane that shares another lane's weights walks that lane's upload log instead of uploading again
    if (share_from_) return static_cast<float*>(share_from_->wpool_.at(share_idx_++));
    float* d = dalloc<float>(n, wpool_);
    PF_CUDA(cudaMemcpyAsync(d, host, n * sizeof(float), cudaMemcpyHostToDevice, stream_));
    PF_CUDA(cudaStreamSynchronize(stream_));   // host buffer may be a temporary
    return d;
}
float* DeviceCtx::up_f32(const BlobEntry& e) { return up_f32(e.data, e.count); }

__half* DeviceCtx::up_f16(const float* host, size_t n) {
    if (share_from_) return static_cast<__half*>(share_from_->wpool_.at(share_idx_++));
    float* t = dalloc<float>(n, tmp_);
    PF_CUDA(cudaMemcpyAsync(t, host, n * sizeof(float), cudaMemcpyHostToDevice, stream_));
    __half* d = dalloc<__half>(n, wpool_);
    f32_to_f16_launch(t, d, n, stream_);
    PF_CUDA(cudaStreamSynchronize(stream_));
    free_pool(tmp_);
    return d;
}

LnW DeviceCtx::up_ln(const Blob& b, const std::string& prefix) {
    LnW w;
    w.g = up_f32(b.get(prefix + ".weight"));
    w.b = up_f32(b.get(prefix + ".bias"));
    return w;
}

void DeviceCtx::load_enc_layer(const Blob& b, const std::string& p, int in_size, EncLayerW& w) {
    const int d = cfg_.d_model, f = cfg_.ffn, k = cfg_.enc_kernel;
    w.in_size = in_size;
    expect_shape(b.get(p + ".norm1.weight"), {in_size});
    w.ln1 = up_ln(b, p + ".norm1");
    const BlobEntry& qkv = b.get(p + ".self_attn.linear_q_k_v.weight");
    expect_shape(qkv, {3 * d, in_size});
    w.w_qkv = up_f16(qkv.data, qkv.count);
    expect_shape(b.get(p + ".self_attn.linear_q_k_v.bias"), {3 * d});
    w.b_qkv = up_f32(b.get(p + ".self_attn.linear_q_k_v.bias"));
    expect_shape(b.get(p + ".self_attn.fsmn_block.weight"), {d, 1, k});
    w.fsmn = up_f32(b.get(p + ".self_attn.fsmn_block.weight"));
    const BlobEntry& wo = b.get(p + ".self_attn.linear_out.weight");
    expect_shape(wo, {d, d});
    w.w_out = up_f16(wo.data, wo.count);
    w.b_out = up_f32(b.get(p + ".self_attn.linear_out.bias"));
    expect_shape(b.get(p + ".norm2.weight"), {d});
    w.ln2 = up_ln(b, p + ".norm2");
    const BlobEntry& w1 = b.get(p + ".feed_forward.w_1.weight");
    expect_shape(w1, {f, d});
    w.w_ffn1 = up_f16(w1.data, w1.count);
    w.b_ffn1 = up_f32(b.get(p + ".feed_forward.w_1.bias"));
    const BlobEntry& w2 = b.get(p + ".feed_forward.w_2.weight");
    expect_shape(w2, {d, f});
    w.w_ffn2 = up_f16(w2.data, w2.count);
    w.b_ffn2 = up_f32(b.get(p + ".feed_forward.w_2.bias"));
}

void DeviceCtx::load_dec_ffn(const Blob& b, const std::string& p, DecFfnW& w, int f) {
    const int d = cfg_.d_model;
    w.ln_in = up_ln(b, p + ".norm1");
    const BlobEntry& w1 = b.get(p + ".feed_forward.w_1.weight");
    expect_shape(w1, {f, d});
    w.w1 = up_f16(w1.data, w1.count);
    w.b1 = up_f32(b.get(p + ".feed_forward.w_1.bias"));
    expect_shape(b.get(p + ".feed_forward.norm.weight"), {f});
    w.ln_mid = up_ln(b, p + ".feed_forward.norm");
    const BlobEntry& w2 = b.get(p + ".feed_forward.w_2.weight");
    expect_shape(w2, {d, f});
    w.w2 = up_f16(w2.data, w2.count);
}

void DeviceCtx::load_weights(const Blob& b, const DeviceCtx* share) {
    PF_CUDA(cudaSetDevice(dev_));
    if (share && share->dev_ != dev_) throw StatusError{PF_ERR_BAD_ARG, "weights can only be shared with an owning context on the same device"};
    wload_begin_ = wpool_.size();                 // the constructor's own uploads (PE table, default CMVN) come first
    share_from_ = share;
    share_idx_ = share ? share->wload_begin_ : 0;
    struct Done { DeviceCtx* c; ~Done() { c->share_from_ = nullptr; } } done{this};   // later uploads (CMVN, hot words) are this lane's own
    const int d = cfg_.d_model;
    enc_.resize(cfg_.enc_layers);
    load_enc_layer(b, "encoder.encoders0.0", cfg_.input_size, enc_[0]);
    for (int i = 1; i < cfg_.enc_layers; ++i) load_enc_layer(b, "encoder.encoders." + std::to_string(i - 1), d, enc_[i]);
    after_norm_ = up_ln(b, "encoder.after_norm");
    tp_.resize(cfg_.tp_layers);
    for (int i = 0; i < cfg_.tp_layers; ++i) load_enc_layer(b, "encoder.tp_encoders." + std::to_string(i), d, tp_[i]);
    if (cfg_.tp_layers) tp_norm_ = up_ln(b, "encoder.tp_norm");

    if (cfg_.model_kind == PF_MODEL_SENSEVOICE_SMALL) {
        const BlobEntry& w = b.get("ctc.ctc_lo.weight");
        expect_shape(w, {cfg_.vocab, d});
        w_head_ = up_f16(w.data, w.count);
        b_head_ = up_f32(b.get("ctc.ctc_lo.bias"));
        const BlobEntry& tab = b.get("embed.weight");
        expect_shape(tab, {16, cfg_.input_size});
        embed_table_ = up_f32(tab);
        return;
    }
    // CifPredictorV2: conv1d weight [out, in, 3] -> GEMM weight [out, 3*in] with k = tap*in + cin (im2col order)
    {
        const BlobEntry& cw = b.get("predictor.cif_conv1d.weight");
        expect_shape(cw, {d, d, 3});
        std::vector<float> re(static_cast<size_t>(d) * 3 * d);
        for (int o = 0; o < d; ++o)
            for (int c = 0; c < d; ++c)
                for (int j = 0; j < 3; ++j) re[(static_cast<size_t>(o) * 3 + j) * d + c] = cw.data[(static_cast<size_t>(o) * d + c) * 3 + j];
        w_conv_ = up_f16(re.data(), re.size());
        b_conv_ = up_f32(b.get("predictor.cif_conv1d.bias"));
        expect_shape(b.get("predictor.cif_output.weight"), {1, d});
        w_alpha_ = up_f32(b.get("predictor.cif_output.weight"));
        b_alpha_ = up_f32(b.get("predictor.cif_output.bias"));
    }
    if (b.has("predictor.upsample_cnn.weight")) {
        // CifPredictorV3 timestamp branch [EXT]: ConvTranspose1d(d, d, 3, stride 3) weight [in, out, 3] -> GEMM weight
        // [(j, out), in]; BiLSTM(d) forward | reverse stacked; cif_output2 Linear(2d, 1)
        const BlobEntry& uw = b.get("predictor.upsample_cnn.weight");
        expect_shape(uw, {d, d, 3});
        std::vector<float> re(static_cast<size_t>(3) * d * d), rb(3 * d);
        const BlobEntry& ub = b.get("predictor.upsample_cnn.bias");
        expect_shape(ub, {d});
        for (int j = 0; j < 3; ++j)
            for (int o = 0; o < d; ++o) {
                rb[j * d + o] = ub.data[o];
                for (int c = 0; c < d; ++c) re[(static_cast<size_t>(j) * d + o) * d + c] = uw.data[(static_cast<size_t>(c) * d + o) * 3 + j];
            }
        w_up16_ = up_f16(re.data(), re.size());
        b_up_ = up_f32(rb.data(), rb.size());
        std::vector<float> wih(static_cast<size_t>(8) * d * d), whh(static_cast<size_t>(8) * d * d), bb(8 * d);
        for (int dir = 0; dir < 2; ++dir) {
            const std::string sfx = dir ? "_reverse" : "";
            const BlobEntry& a = b.get("predictor.blstm.weight_ih_l0" + sfx);
            const BlobEntry& h = b.get("predictor.blstm.weight_hh_l0" + sfx);
            const BlobEntry& bi = b.get("predictor.blstm.bias_ih_l0" + sfx);
            const BlobEntry& bh = b.get("predictor.blstm.bias_hh_l0" + sfx);
            expect_shape(a, {4 * d, d}); expect_shape(h, {4 * d, d}); expect_shape(bi, {4 * d}); expect_shape(bh, {4 * d});
            memcpy(wih.data() + static_cast<size_t>(dir) * 4 * d * d, a.data, a.count * sizeof(float));
            memcpy(whh.data() + static_cast<size_t>(dir) * 4 * d * d, h.data, h.count * sizeof(float));
            for (int i = 0; i < 4 * d; ++i) bb[dir * 4 * d + i] = bi.data[i] + bh.data[i];
        }
        w_ih_bi16_ = up_f16(wih.data(), wih.size());
        w_hh_bi16_ = up_f16(whh.data(), whh.size());
        b_bi_ = up_f32(bb.data(), bb.size());
        expect_shape(b.get("predictor.cif_output2.weight"), {1, 2 * d});
        w_out2_ = up_f32(b.get("predictor.cif_output2.weight"));
        b_out2_ = up_f32(b.get("predictor.cif_output2.bias"));
    }
    load_dec_stack(b, "decoder", cfg_.dec_layers, cfg_.dec_ffn, cfg_.dec_kernel, dec_, dec3_, dec_after_, w_kv_all_, b_kv_all_);
    const BlobEntry& wo = b.get("decoder.output_layer.weight");
    expect_shape(wo, {cfg_.vocab, d});
    w_head_ = up_f16(wo.data, wo.count);
    expect_shape(b.get("decoder.output_layer.bias"), {cfg_.vocab});
    b_head_ = up_f32(b.get("decoder.output_layer.bias"));
    if (cfg_.model_kind != PF_MODEL_SEACO_PARAFORMER) return;
    // SeACo: bias decoder, hot-word head, hot-word encoder (model_eb.onnx: Embedding + 2-layer LSTM)
    load_dec_stack(b, "seaco_decoder", cfg_.seaco_layers, cfg_.seaco_ffn, cfg_.seaco_kernel, sdec_, sdec3_, sdec_after_, w_skv_all_, b_skv_all_);
    const BlobEntry& hw = b.get("hotword_output_layer.weight");
    expect_shape(hw, {cfg_.vocab, d});
    w_hw_head_ = up_f16(hw.data, hw.count);
    expect_shape(b.get("hotword_output_layer.bias"), {cfg_.vocab});
    b_hw_head_ = up_f32(b.get("hotword_output_layer.bias"));
    expect_shape(b.get("bias_embed.weight"), {cfg_.vocab, d});
    bias_table_ = up_f32(b.get("bias_embed.weight"));
    for (int layer = 0; layer < 2; ++layer) {
        const std::string p = "bias_encoder.", sfx = "_l" + std::to_string(layer);
        const BlobEntry& wih = b.get(p + "weight_ih" + sfx);
        const BlobEntry& whh = b.get(p + "weight_hh" + sfx);
        expect_shape(wih, {4 * d, d});
        expect_shape(whh, {4 * d, d});
        lstm_[layer].w_ih = up_f16(wih.data, wih.count);
        lstm_[layer].w_hh = up_f16(whh.data, whh.count);
        const BlobEntry& bih = b.get(p + "bias_ih" + sfx);
        const BlobEntry& bhh = b.get(p + "bias_hh" + sfx);
        expect_shape(bih, {4 * d});
        expect_shape(bhh, {4 * d});
        std::vector<float> bsum(4 * d);
        for (int i = 0; i < 4 * d; ++i) bsum[i] = bih.data[i] + bhh.data[i];
        lstm_[layer].b = up_f32(bsum.data(), bsum.size());
    }
}

void DeviceCtx::load_dec_stack(const Blob& b, const std::string& prefix, int nlayers, int ffn_width, int kernel,
                               std::vector<DecLayerW>& layers, DecFfnW& d3, LnW& after, __half*& w_kv_all, float*& b_kv_all) {
    const int d = cfg_.d_model;
    layers.resize(nlayers);
    std::vector<float> kvw(static_cast<size_t>(nlayers) * 2 * d * d), kvb(static_cast<size_t>(nlayers) * 2 * d);
    for (int i = 0; i < nlayers; ++i) {
        const std::string p = prefix + ".decoders." + std::to_string(i);
        DecLayerW& w = layers[i];
        load_dec_ffn(b, p, w.ffn, ffn_width);
        w.ln2 = up_ln(b, p + ".norm2");
        expect_shape(b.get(p + ".self_attn.fsmn_block.weight"), {d, 1, kernel});
        w.fsmn = up_f32(b.get(p + ".self_attn.fsmn_block.weight"));
        w.ln3 = up_ln(b, p + ".norm3");
        const BlobEntry& wq = b.get(p + ".src_attn.linear_q.weight");
        expect_shape(wq, {d, d});
        w.wq = up_f16(wq.data, wq.count);
        w.bq = up_f32(b.get(p + ".src_attn.linear_q.bias"));
        const BlobEntry& wkv = b.get(p + ".src_attn.linear_k_v.weight");
        expect_shape(wkv, {2 * d, d});
        memcpy(kvw.data() + static_cast<size_t>(i) * 2 * d * d, wkv.data, wkv.count * sizeof(float));
        const BlobEntry& bkv = b.get(p + ".src_attn.linear_k_v.bias");
        expect_shape(bkv, {2 * d});
        memcpy(kvb.data() + static_cast<size_t>(i) * 2 * d, bkv.data, bkv.count * sizeof(float));
        const BlobEntry& wo = b.get(p + ".src_attn.linear_out.weight");
        expect_shape(wo, {d, d});
        w.wo = up_f16(wo.data, wo.count);
        w.bo = up_f32(b.get(p + ".src_attn.linear_out.bias"));
    }
    // the K/V projections of all layers of a stack read the same memory: one [layers*2d, d] GEMM
    w_kv_all = up_f16(kvw.data(), kvw.size());
    b_kv_all = up_f32(kvb.data(), kvb.size());
    load_dec_ffn(b, prefix + ".decoders3.0", d3, ffn_width);
    after = up_ln(b, prefix + ".after_norm");
}

void DeviceCtx::set_cmvn(const float* shift, const float* scale, int dim) {
    PF_CUDA(cudaSetDevice(dev_));
    if (dim != cfg_.lfr_m * cfg_.n_mels) throw StatusError{PF_ERR_SHAPE, "cmvn dim must be lfr_m * n_mels"};
    PF_CUDA(cudaMemcpyAsync(cmvn_shift_, shift, dim * sizeof(float), cudaMemcpyHostToDevice, stream_));
    PF_CUDA(cudaMemcpyAsync(cmvn_scale_, scale, dim * sizeof(float), cudaMemcpyHostToDevice, stream_));
    PF_CUDA(cudaStreamSynchronize(stream_));
}

// ------------------------------------------------------------------ workspace
void DeviceCtx::ensure_workspace(int B, int T) {
    if (B <= capB_ && T <= capT_) return;
    PF_CUDA(cudaStreamSynchronize(stream_));
    free_pool(apool_);
    enc_plans_.clear();
    dec_plans_.clear();
    capB_ = std::max(capB_, B);
    capT_ = std::max(capT_, T);
    const size_t M = static_cast<size_t>(capB_) * capT_;
    const size_t Md = static_cast<size_t>(capB_) * (capT_ + 1);
    const int d = cfg_.d_model, din = cfg_.input_size;
    const bool sv = cfg_.model_kind == PF_MODEL_SENSEVOICE_SMALL;
    feats_ = dalloc<float>(M * din, apool_);
    feats_raw_ = sv ? dalloc<float>(M * din, apool_) : nullptr;
    a16_ = dalloc<__half>(M * std::max(din, d), apool_);
    qkv16_ = dalloc<__half>(M * 3 * d, apool_);
    ctx16_ = dalloc<__half>(M * d, apool_);
    h16_ = dalloc<__half>(M * cfg_.ffn, apool_);
    mem32_ = dalloc<float>(M * d, apool_);
    x32_ = dalloc<float>(M * d, apool_);
    enc32_ = dalloc<float>(M * d, apool_);
    enc16_ = dalloc<__half>(M * d, apool_);
    prompt_ids_ = dalloc<int>(4, apool_);
    pick_ld_ = gemm_pick_slots(cfg_.vocab);
    pick_ = dalloc<float>((sv ? M : Md) * pick_ld_ * 3, apool_);
    if (sv) {
        logits_ = dalloc<float>(M * ldv(), apool_);
        tokens_ = dalloc<int>(M, apool_);
        token_num_ = dalloc<int>(capB_, apool_);
        return;
    }
    kv16_ = dalloc<__half>(M * cfg_.dec_layers * 2 * d, apool_);
    const size_t A = static_cast<size_t>(capB_) * (capT_ + 1);
    alphas_ = dalloc<float>(A, apool_);
    wcur_ = dalloc<float>(A, apool_);
    wrem_ = dalloc<float>(A, apool_);
    peaks_ = dalloc<float>(A, apool_);
    fire_idx_ = dalloc<int>(A, apool_);
    token_num_ = dalloc<int>(capB_, apool_);
    fires_ = dalloc<int>(capB_, apool_);
    meta_ = dalloc<int>(4, apool_);
    xd32_ = dalloc<float>(Md * d, apool_);
    ad16_ = dalloc<__half>(Md * d, apool_);
    hd32_ = dalloc<float>(Md * cfg_.dec_ffn, apool_);
    hd16_ = dalloc<__half>(Md * cfg_.dec_ffn, apool_);
    t32_ = dalloc<float>(Md * d, apool_);
    tn32_ = dalloc<float>(Md * d, apool_);
    q16_ = dalloc<__half>(Md * d, apool_);
    ctxd16_ = dalloc<__half>(Md * d, apool_);
    logits_ = dalloc<float>(Md * ldv(), apool_);
    tokens_ = dalloc<int>(Md, apool_);
    if (cfg_.model_kind == PF_MODEL_SEACO_PARAFORMER) {
        emb32_ = dalloc<float>(Md * d, apool_);
        hid32_ = dalloc<float>(Md * d, apool_);
        satt32_ = dalloc<float>(Md * d, apool_);
        logits2_ = dalloc<float>(Md * ldv(), apool_);
        tokens2_ = dalloc<int>(Md, apool_);
    }
}

static void ensure_pinned(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *cap = 0;
        throw StatusError{PF_ERR_OOM, "cudaMallocHost of " + std::to_string(bytes) + " bytes failed"};
    }
    *cap = bytes;
}

void DeviceCtx::ensure_host(size_t tokens, size_t logits, size_t peaks) {
    ensure_pinned(reinterpret_cast<void**>(&h_tokens), &h_tokens_cap_, std::max<size_t>(tokens, 1) * sizeof(int32_t));
    if (logits) ensure_pinned(reinterpret_cast<void**>(&h_logits), &h_logits_cap_, logits * sizeof(float));
    if (peaks) ensure_pinned(reinterpret_cast<void**>(&h_peaks), &h_peaks_cap_, peaks * sizeof(float));
}

// ------------------------------------------------------------------ plans (TMA descriptors are baked per shape)
EncoderPlan& DeviceCtx::encoder_plan(int B, int T) {
    gemm_set_policy(throughput_mode);
    auto key = std::make_pair(B, T);
    auto it = enc_plans_.find(key);
    if (it != enc_plans_.end()) return it->second;
    EncoderPlan plan;
    const int M = B * T, d = cfg_.d_model, f = cfg_.ffn;
    // LayerNorms whose input comes straight out of a residual GEMM are fused into that GEMM's epilogue: norm2 into the
    // out-projection, the NEXT layer's norm1 into FFN2 (next == nullptr: the stack ends, after_norm stays a kernel)
    const bool fuse = ln_fuse_enabled() && gemm_ln_fusable(M, d);
    auto build_layer = [&](const EncLayerW& w, const EncLayerW* next) {
        EncLayerPlan lp;
        GemmEpi e;
        e = GemmEpi{}; e.bias = w.b_qkv; e.out_f16 = qkv16_; e.ld_out = 3 * d;
        gemm_prepare(lp.qkv, a16_, w.in_size, w.w_qkv, w.in_size, M, 3 * d, w.in_size, e);
        // several lanes: LayerNorm + GEMM as one row-tile-stationary kernel (a third less SM time); layer 0 of a
        // stack keeps the two-kernel form (its norm1 is part of embed_pe_ln / runs on 560 columns)
        const bool ln_gemm_on = ln_gemm_mode() == 2 || (ln_gemm_mode() == 1 && throughput_mode);
        if (ln_gemm_on && !fuse && w.in_size == d && ln_gemm_supported(M, 3 * d, d, x32_, d, qkv16_, 3 * d, w.b_qkv))
            ln_gemm_prepare(lp.qkv_ln, x32_, d, w.ln1.g, w.ln1.b, cfg_.ln_eps, w.w_qkv, d, w.b_qkv, qkv16_, 3 * d, 0, M, 3 * d, d);
        // x32_ already holds residual + FSMN memory (the attention kernel accumulates the memory into it)
        e = GemmEpi{}; e.bias = w.b_out; e.resid = x32_; e.ld_resid = d; e.out_f32 = x32_; e.ld_out = d;
        if (fuse) { e.ln_gamma = w.ln2.g; e.ln_beta = w.ln2.b; e.ln_eps = cfg_.ln_eps; e.ln_out16 = a16_; e.ld_ln16 = d; lp.ln2_fused = true; }
        gemm_prepare(lp.out, ctx16_, d, w.w_out, d, M, d, d, e);
        e = GemmEpi{}; e.bias = w.b_ffn1; e.relu = 1; e.out_f16 = h16_; e.ld_out = f;
        gemm_prepare(lp.ffn1, a16_, d, w.w_ffn1, d, M, f, d, e);
        if (ln_gemm_on && !fuse && ln_gemm_supported(M, f, d, x32_, d, h16_, f, w.b_ffn1))
            ln_gemm_prepare(lp.ffn1_ln, x32_, d, w.ln2.g, w.ln2.b, cfg_.ln_eps, w.w_ffn1, d, w.b_ffn1, h16_, f, 1, M, f, d);
        e = GemmEpi{}; e.bias = w.b_ffn2; e.resid = x32_; e.ld_resid = d; e.out_f32 = x32_; e.ld_out = d;
        if (fuse && next != nullptr) {
            e.ln_gamma = next->ln1.g; e.ln_beta = next->ln1.b; e.ln_eps = cfg_.ln_eps; e.ln_out16 = a16_; e.ld_ln16 = d;
            lp.next_ln1_fused = true;
        }
        gemm_prepare(lp.ffn2, h16_, f, w.w_ffn2, f, M, d, f, e);
#ifdef PFASR_EXPERIMENTS
        if (!lp.next_ln1_fused && ffn_chain_enabled() && ffn_chain_supported(M, d, f, chain_))
            ffn_chain_prepare(lp.chain, a16_, d, w.w_ffn1, w.b_ffn1, h16_, f, w.w_ffn2, w.b_ffn2, x32_, d, M, d, f, chain_);
#endif
        plan.layers.push_back(lp);
    };
    for (size_t i = 0; i < enc_.size(); ++i) build_layer(enc_[i], i + 1 < enc_.size() ? &enc_[i + 1] : nullptr);
    for (size_t i = 0; i < tp_.size(); ++i) build_layer(tp_[i], i + 1 < tp_.size() ? &tp_[i + 1] : nullptr);
    if (cfg_.model_kind == PF_MODEL_SENSEVOICE_SMALL) {
        GemmEpi e; e.bias = b_head_; e.out_f32 = logits_; e.ld_out = ldv();
        gemm_prepare(plan.ctc_head, enc16_, d, w_head_, d, M, cfg_.vocab, d, e);
        GemmEpi pk; pk.bias = b_head_; pk.pick_out = pick_; pk.pick_ld = pick_ld_;
        gemm_prepare(plan.ctc_head_pick, enc16_, d, w_head_, d, M, cfg_.vocab, d, pk);
    } else {
        GemmEpi e; e.bias = b_conv_; e.relu = 1; e.out_f32 = mem32_; e.ld_out = d;
        gemm_prepare(plan.pred_conv, qkv16_, 3 * d, w_conv_, 3 * d, M, d, 3 * d, e);   // qkv16_ holds the im2col rows
        GemmEpi k; k.bias = b_kv_all_; k.out_f16 = kv16_; k.ld_out = cfg_.dec_layers * 2 * d;
        gemm_prepare(plan.kv_all, enc16_, d, w_kv_all_, d, M, cfg_.dec_layers * 2 * d, d, k);
    }
    return enc_plans_.emplace(key, std::move(plan)).first->second;
}

DecoderPlan& DeviceCtx::decoder_plan(int B, int T, int L) {
    gemm_set_policy(throughput_mode);
    if (dec_plan_T_ != T) { dec_plans_.clear(); dec_plan_T_ = T; }
    auto key = std::make_pair(B, L);
    auto it = dec_plans_.find(key);
    if (it != dec_plans_.end()) return it->second;
    DecoderPlan plan;
    const int Md = B * L, d = cfg_.d_model, f = cfg_.dec_ffn;
    auto ffn = [&](const DecFfnW& w, GemmOp& g1, GemmOp& g2) {
        GemmEpi e; e.bias = w.b1; e.relu = 1; e.out_f32 = hd32_; e.ld_out = f;
        gemm_prepare(g1, ad16_, d, w.w1, d, Md, f, d, e);
        GemmEpi e2; e2.out_f32 = t32_; e2.ld_out = d;
        gemm_prepare(g2, hd16_, f, w.w2, f, Md, d, f, e2);
    };
    // norm1 of the next layer (or of decoders3) rides on the cross-attention out-projection's epilogue
    const bool fuse = ln_fuse_enabled() && gemm_ln_fusable(Md, d);
    for (size_t i = 0; i < dec_.size(); ++i) {
        const DecLayerW& w = dec_[i];
        DecLayerPlan lp;
        ffn(w.ffn, lp.w1, lp.w2);
        GemmEpi e; e.bias = w.bq; e.out_f16 = q16_; e.ld_out = d;
        gemm_prepare(lp.q, ad16_, d, w.wq, d, Md, d, d, e);
        GemmEpi o; o.bias = w.bo; o.resid = xd32_; o.ld_resid = d; o.out_f32 = xd32_; o.ld_out = d;
        if (fuse) {
            const LnW& nx = i + 1 < dec_.size() ? dec_[i + 1].ffn.ln_in : dec3_.ln_in;
            o.ln_gamma = nx.g; o.ln_beta = nx.b; o.ln_eps = cfg_.ln_eps; o.ln_out16 = ad16_; o.ld_ln16 = d;
            lp.next_ln1_fused = true;
        }
        gemm_prepare(lp.out, ctxd16_, d, w.wo, d, Md, d, d, o);
        plan.layers.push_back(lp);
    }
    ffn(dec3_, plan.d3_w1, plan.d3_w2);
    GemmEpi h; h.bias = b_head_; h.out_f32 = logits_; h.ld_out = ldv();
    gemm_prepare(plan.head, ad16_, d, w_head_, d, Md, cfg_.vocab, d, h);
    GemmEpi pk; pk.bias = b_head_; pk.pick_out = pick_; pk.pick_ld = pick_ld_;
    gemm_prepare(plan.head_pick, ad16_, d, w_head_, d, Md, cfg_.vocab, d, pk);
    if (cfg_.model_kind == PF_MODEL_SEACO_PARAFORMER) {
        const int sf = cfg_.seaco_ffn;
        auto sffn = [&](const DecFfnW& w, GemmOp& g1, GemmOp& g2) {
            GemmEpi e; e.bias = w.b1; e.relu = 1; e.out_f32 = hd32_; e.ld_out = sf;
            gemm_prepare(g1, ad16_, d, w.w1, d, Md, sf, d, e);
            GemmEpi e2; e2.out_f32 = t32_; e2.ld_out = d;
            gemm_prepare(g2, hd16_, sf, w.w2, sf, Md, d, sf, e2);
        };
        for (const DecLayerW& w : sdec_) {
            DecLayerPlan lp;
            sffn(w.ffn, lp.w1, lp.w2);
            GemmEpi e; e.bias = w.bq; e.out_f16 = q16_; e.ld_out = d;
            gemm_prepare(lp.q, ad16_, d, w.wq, d, Md, d, d, e);
            GemmEpi o; o.bias = w.bo; o.resid = xd32_; o.ld_resid = d; o.out_f32 = xd32_; o.ld_out = d;
            gemm_prepare(lp.out, ctxd16_, d, w.wo, d, Md, d, d, o);
            plan.slayers.push_back(lp);
        }
        sffn(sdec3_, plan.s_d3_w1, plan.s_d3_w2);
        GemmEpi hh; hh.bias = b_hw_head_; hh.out_f32 = logits2_; hh.ld_out = ldv();
        gemm_prepare(plan.hw_head, ad16_, d, w_hw_head_, d, Md, cfg_.vocab, d, hh);
    }
    return dec_plans_.emplace(key, std::move(plan)).first->second;
}

double DeviceCtx::replay_gemms(int iters) {
    PF_CUDA(cudaSetDevice(dev_));
    if (replay_.empty() || iters <= 0) return 0.0;
    cudaEvent_t a, b;
    PF_CUDA(cudaEventCreate(&a));
    PF_CUDA(cudaEventCreate(&b));
    auto pass = [&] {
        for (const ReplayOp& op : replay_) {
            if (op.fused) ln_gemm_launch(op.l, stream_);
            else gemm_launch(op.g, stream_);
        }
    };
    pass();                                                               // warm pass
    PF_CUDA(cudaEventRecord(a, stream_));
    for (int i = 0; i < iters; ++i) pass();
    PF_CUDA(cudaEventRecord(b, stream_));
    PF_CUDA(cudaEventSynchronize(b));
    float ms = 0.0f;
    PF_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms / iters;
}

void DeviceCtx::ln_gemm(const LnGemmOp& op) {
    if (profile_ == 1) { ReplayOp r; r.fused = true; r.l = op; replay_.push_back(r); }
    if (profile_) {
        ProfRec r{op.M, op.N, op.K, 256 + 1000 * 11, nullptr, nullptr, nullptr};
        for (cudaEvent_t* e : {&r.a, &r.b}) {
            if (!prof_pool_.empty()) { *e = prof_pool_.back(); prof_pool_.pop_back(); }
            else PF_CUDA(cudaEventCreate(e));
        }
        PF_CUDA(cudaEventRecord(r.a, stream_));
        ln_gemm_launch(op, stream_);
        PF_CUDA(cudaEventRecord(r.b, stream_));
        prof_.push_back(r);
    } else {
        ln_gemm_launch(op, stream_);
    }
    ++launches;
    gemm_flops += ln_gemm_flops(op);
}

void DeviceCtx::gemm(const GemmOp& op) {
    if (profile_ == 1) { ReplayOp r; r.g = op; replay_.push_back(r); }
    if (profile_) {
        ProfRec r{op.M, op.N, op.K, op.bn + 1000 * (op.cm * 10 + op.cn), nullptr, nullptr, nullptr};
        for (cudaEvent_t* e : {&r.a, &r.b}) {
            if (!prof_pool_.empty()) { *e = prof_pool_.back(); prof_pool_.pop_back(); }
            else PF_CUDA(cudaEventCreate(e));
        }
        PF_CUDA(cudaEventRecord(r.a, stream_));
        gemm_launch(op, stream_);
        PF_CUDA(cudaEventRecord(r.b, stream_));
        prof_.push_back(r);
    } else {
        gemm_launch(op, stream_);
    }
    ++launches;
    gemm_flops += pf::gemm_flops(op);
}

#ifdef PFASR_EXPERIMENTS
void DeviceCtx::ffn_chain(const FfnChainOp& op) {
    if (profile_) {
        // one row for both GEMMs: 2 M (2F) D flops = the sum of the two
        ProfRec r{op.M, 2 * op.F, op.D, 256 + 1000 * 11, nullptr, nullptr, nullptr};
        for (cudaEvent_t* e : {&r.a, &r.b}) {
            if (!prof_pool_.empty()) { *e = prof_pool_.back(); prof_pool_.pop_back(); }
            else PF_CUDA(cudaEventCreate(e));
        }
        PF_CUDA(cudaEventRecord(r.a, stream_));
        ffn_chain_launch(op, stream_);
        PF_CUDA(cudaEventRecord(r.b, stream_));
        prof_.push_back(r);
    } else {
        ffn_chain_launch(op, stream_);
    }
    ++launches;
    gemm_flops += ffn_chain_flops(op);
}

#endif

template <typename F>
void DeviceCtx::timed(const char* label, F&& f) {
    if (profile_ < 2) { f(); return; }
    ProfRec r{0, 0, 0, 0, nullptr, nullptr, label};
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!prof_pool_.empty()) { *e = prof_pool_.back(); prof_pool_.pop_back(); }
        else PF_CUDA(cudaEventCreate(e));
    }
    PF_CUDA(cudaEventRecord(r.a, stream_));
    f();
    PF_CUDA(cudaEventRecord(r.b, stream_));
    prof_.push_back(r);
}

void DeviceCtx::finish_profile() {
    gemm_ms = 0.0;
    profile_json.clear();
    if (prof_.empty()) return;
    struct Agg { int count = 0; double ms = 0.0; };
    std::map<std::vector<int>, Agg> agg;
    std::map<std::string, Agg> other;
    for (ProfRec& r : prof_) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (r.label) {
            Agg& a = other[r.label];
            a.count++;
            a.ms += ms;
        } else {
            gemm_ms += ms;
            Agg& a = agg[{r.M, r.N, r.K, r.bn}];
            a.count++;
            a.ms += ms;
        }
        prof_pool_.push_back(r.a);
        prof_pool_.push_back(r.b);
    }
    prof_.clear();
    std::string js = "[";
    for (auto& kv : agg) {
        if (js.size() > 1) js += ",";
        const double fl = 2.0 * kv.first[0] * static_cast<double>(kv.first[1]) * kv.first[2] * kv.second.count;
        js += "{\"name\":\"gemm\",\"M\":" + std::to_string(kv.first[0]) + ",\"N\":" + std::to_string(kv.first[1]) + ",\"K\":" + std::to_string(kv.first[2]) +
              ",\"tile_n\":" + std::to_string(kv.first[3] % 1000) + ",\"cluster\":\"" + std::to_string(kv.first[3] / 10000) + "x" + std::to_string(kv.first[3] / 1000 % 10) + "\"" + ",\"launches\":" + std::to_string(kv.second.count) + ",\"ms\":" +
              std::to_string(kv.second.ms) + ",\"tflops\":" + std::to_string(kv.second.ms > 0 ? fl / (kv.second.ms * 1e9) : 0.0) + "}";
    }
    for (auto& kv : other) {
        if (js.size() > 1) js += ",";
        js += "{\"name\":\"" + kv.first + "\",\"launches\":" + std::to_string(kv.second.count) + ",\"ms\":" + std::to_string(kv.second.ms) + "}";
    }
    js += "]";
    profile_json = js;
}

// ------------------------------------------------------------------ staging
void DeviceCtx::stage_pcm(const float* const* pcm, const int32_t* nsamp, int B, int tmax_lfr, const pf_audio* audio) {
    PF_CUDA(cudaSetDevice(dev_));
    const bool sv = cfg_.model_kind == PF_MODEL_SENSEVOICE_SMALL;
    const int Tenc = tmax_lfr + (sv ? 4 : 0);
    ensure_workspace(B, Tenc);
    // per-utterance tables: offsets (pcm, feats) as int64, then nsamp / nframes / nlfr as int32
    const size_t tab_bytes = static_cast<size_t>(B) * (2 * sizeof(long long) + 3 * sizeof(int));
    ensure_pinned(&h_stage_, &h_stage_bytes_, tab_bytes);
    if (B > meta_capB_) {
        if (d_off_) cudaFree(d_off_);
        if (d_meta_) cudaFree(d_meta_);
        PF_CUDA(cudaMalloc(&d_off_, static_cast<size_t>(B) * 2 * sizeof(long long)));
        PF_CUDA(cudaMalloc(&d_meta_, static_cast<size_t>(B) * 3 * sizeof(int)));
        meta_capB_ = B;
    }
    long long* h_off = static_cast<long long*>(h_stage_);
    int* h_m = reinterpret_cast<int*>(h_off + 2 * B);
    size_t total = 0;
    int maxframes = 0;
    int max_nsamp = 0;
    const int dim = cfg_.lfr_m * cfg_.n_mels;
    AudioItem* h_items = nullptr;
    if (audio) {
        ensure_pinned(&h_audio_, &h_audio_bytes_, static_cast<size_t>(B) * sizeof(AudioItem));
        h_items = static_cast<AudioItem*>(h_audio_);
        if (B > audio_capB_) {
            if (d_audio_) cudaFree(d_audio_);
            d_audio_ = nullptr;
            PF_CUDA(cudaMalloc(&d_audio_, static_cast<size_t>(B) * sizeof(AudioItem)));
            audio_capB_ = B;
        }
        size_t raw_total = 0;
        for (int b = 0; b < B; ++b) {
            h_items[b] = AudioItem{static_cast<long long>(raw_total), audio[b].n_values, audio[b].format, audio[b].channels, audio[b].sample_rate, 0};
            raw_total += static_cast<size_t>(audio[b].n_values) * audio_bytes_per_value(audio[b].format);
            raw_total = (raw_total + 15) & ~static_cast<size_t>(15);
        }
        if (raw_total > raw_cap_) {
            if (raw_) cudaFree(raw_);
            raw_ = nullptr;
            PF_CUDA(cudaMalloc(&raw_, std::max<size_t>(raw_total, 16)));
            raw_cap_ = raw_total;
        }
    }
    for (int b = 0; b < B; ++b) {
        max_nsamp = std::max(max_nsamp, nsamp[b]);
        h_off[b] = static_cast<long long>(total);
        h_off[B + b] = static_cast<long long>(b) * tmax_lfr * dim;
        const int nf = frontend_num_frames(nsamp[b], cfg_.snip_edges != 0);
        h_m[b] = nsamp[b];
        h_m[B + b] = nf;
        h_m[2 * B + b] = nf / cfg_.lfr_n;
        maxframes = std::max(maxframes, nf);
        total += static_cast<size_t>(nsamp[b]);
        total = (total + 3) & ~static_cast<size_t>(3);
    }
    if (total > pcm_cap_) {
        if (pcm_) cudaFree(pcm_);
        pcm_ = nullptr;
        PF_CUDA(cudaMalloc(&pcm_, std::max<size_t>(total, 1) * sizeof(float)));
        pcm_cap_ = total;
    }
    PF_CUDA(cudaEventRecord(ev_[0], stream_));
    ev0_armed_ = true;
    PF_CUDA(cudaMemcpyAsync(d_off_, h_off, static_cast<size_t>(B) * 2 * sizeof(long long), cudaMemcpyHostToDevice, stream_));
    PF_CUDA(cudaMemcpyAsync(d_meta_, h_m, static_cast<size_t>(B) * 3 * sizeof(int), cudaMemcpyHostToDevice, stream_));
    if (audio) PF_CUDA(cudaMemcpyAsync(d_audio_, h_items, static_cast<size_t>(B) * sizeof(AudioItem), cudaMemcpyHostToDevice, stream_));
    // PCM travels on a second stream in up to kCopyGroups groups of utterances; the compute stream waits per group, so the
    // front-end of group g runs while the copy engine still moves group g+1 (the encoder needs them all)
    PF_CUDA(cudaEventRecord(ev_compute_, stream_));                       // pcm_ may still be read by the previous run
    PF_CUDA(cudaStreamWaitEvent(copy_stream_, ev_compute_, 0));
    staged_groups_ = std::min(kCopyGroups, B);
    for (int g = 0; g < staged_groups_; ++g) {
        const int b0 = static_cast<int>(static_cast<long long>(B) * g / staged_groups_);
        const int b1 = static_cast<int>(static_cast<long long>(B) * (g + 1) / staged_groups_);
        for (int b = b0; b < b1; ++b) {
            if (audio) {
                const size_t nb = static_cast<size_t>(audio[b].n_values) * audio_bytes_per_value(audio[b].format);
                if (nb > 0) PF_CUDA(cudaMemcpyAsync(raw_ + h_items[b].raw_off, audio[b].data, nb, cudaMemcpyHostToDevice, copy_stream_));
            } else if (nsamp[b] > 0) {
                PF_CUDA(cudaMemcpyAsync(pcm_ + h_off[b], pcm[b], static_cast<size_t>(nsamp[b]) * sizeof(float), cudaMemcpyHostToDevice, copy_stream_));
            }
        }
        PF_CUDA(cudaEventRecord(ev_grp_[g], copy_stream_));
    }
    staged_audio_ = audio != nullptr;
    staged_max_nsamp_ = max_nsamp;
    staged_B_ = B;
    staged_T_ = tmax_lfr;
    staged_maxframes_ = maxframes;
    staged_is_pcm_ = true;
}

void DeviceCtx::sync_staging() {
    PF_CUDA(cudaSetDevice(dev_));
    PF_CUDA(cudaStreamSynchronize(copy_stream_));
    PF_CUDA(cudaStreamSynchronize(stream_));
    if (staged_groups_ > 1) staged_groups_ = 1;      // everything is resident: one front-end launch for the whole batch
}

void DeviceCtx::stage_feats(const float* speech, int B, int T) {
    PF_CUDA(cudaSetDevice(dev_));
    ensure_workspace(B, T);
    PF_CUDA(cudaEventRecord(ev_[0], stream_));
    ev0_armed_ = true;
    PF_CUDA(cudaMemcpyAsync(feats_, speech, static_cast<size_t>(B) * T * cfg_.input_size * sizeof(float), cudaMemcpyHostToDevice, stream_));
    staged_B_ = B;
    staged_T_ = T;
    staged_is_pcm_ = false;
}

int DeviceCtx::extract(const float* samples, int nsamp, float* out, int capacity_frames, bool raw_fbank) {
    PF_CUDA(cudaSetDevice(dev_));
    const int nf = frontend_num_frames(nsamp, cfg_.snip_edges != 0);
    const int nl = nf / cfg_.lfr_n;
    const int rows = raw_fbank ? nf : nl;
    const int width = raw_fbank ? cfg_.n_mels : cfg_.lfr_m * cfg_.n_mels;
    if (rows > capacity_frames) throw StatusError{PF_ERR_SHAPE, "feature buffer too small: need " + std::to_string(rows) + " frames"};
    if (rows == 0) return 0;
    std::vector<void*> pool;
    float* d_pcm = dalloc<float>(nsamp, pool);
    float* d_out = dalloc<float>(static_cast<size_t>(rows) * width, pool);
    long long* d_off = dalloc<long long>(2, pool);
    int* d_m = dalloc<int>(3, pool);
    try {
        const long long off[2] = {0, 0};
        const int m[3] = {nsamp, nf, nl};
        PF_CUDA(cudaMemcpyAsync(d_pcm, samples, static_cast<size_t>(nsamp) * sizeof(float), cudaMemcpyHostToDevice, stream_));
        PF_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, stream_));
        PF_CUDA(cudaMemcpyAsync(d_m, m, sizeof(m), cudaMemcpyHostToDevice, stream_));
        FrontendLaunch a;
        a.tables = fe_tables_; a.pcm = d_pcm; a.pcm_off = d_off; a.nsamp = d_m; a.nframes = d_m + 1; a.nlfr = d_m + 2;
        a.add_shift = cmvn_shift_; a.rescale = cmvn_scale_;
        if (raw_fbank) { a.fbank_out = d_out; a.fbank_off = d_off + 1; }
        else { a.feats_out = d_out; a.feats_off = d_off + 1; }
        a.batch = 1; a.max_frames = nf; a.tmax_lfr = nl; a.lfr_m = cfg_.lfr_m; a.lfr_n = cfg_.lfr_n;
        a.snip_edges = cfg_.snip_edges != 0;
        frontend_launch(a, stream_);
        PF_CUDA(cudaMemcpyAsync(out, d_out, static_cast<size_t>(rows) * width * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaStreamSynchronize(stream_));
    } catch (...) {
        cudaStreamSynchronize(stream_);
        free_pool(pool);
        throw;
    }
    free_pool(pool);
    return rows;
}

// ------------------------------------------------------------------ forward passes
void DeviceCtx::encoder_forward(int B, int T, bool online) {
    const int M = B * T, d = cfg_.d_model, H = cfg_.heads;
    EncoderPlan& plan = encoder_plan(B, T);
    if (!online && T > pe_rows_) {                          // position-encoding table: rows are independent of T, so it only ever grows
        PF_CUDA(cudaStreamSynchronize(stream_));
        if (pe_table_) cudaFree(pe_table_);
        pe_rows_ = std::max(T, capT_);
        PF_CUDA(cudaMalloc(&pe_table_, static_cast<size_t>(pe_rows_) * cfg_.input_size * sizeof(float)));
        pe_table_launch(pe_table_, pe_rows_, cfg_.input_size, inv_ts_, stream_);
    }
    timed("embed_pe_ln", [&] {
        // streaming windows arrive scaled and position-encoded (OnlineStream.cs:203-206): LayerNorm only
        embed_pe_ln_launch(feats_, M, T, cfg_.input_size, sqrtf(static_cast<float>(d)), online ? nullptr : pe_table_, enc_[0].ln1.g,
                           enc_[0].ln1.b, cfg_.ln_eps, a16_, stream_);
    });
    ++launches;
    auto layer = [&](const EncLayerW& w, const EncLayerPlan& lp, bool have_ln1) {
        if (!have_ln1 && lp.qkv_ln.valid) {
            ln_gemm(lp.qkv_ln);                            // norm1 + QKV in one kernel
        } else {
            if (!have_ln1) {
                if (!(dbg_skip() & 1)) timed("enc_layernorm", [&] { layernorm_f32_launch(x32_, d, M, d, w.ln1.g, w.ln1.b, cfg_.ln_eps, a16_, d, nullptr, 0, stream_); });
                ++launches;
            }
            gemm(lp.qkv);
        }
        if (!(dbg_skip() & 2)) timed("enc_attention_fsmn", [&] {
            // x += mem (layer with a residual) or x = mem (encoders0: in_size != d_model, no residual); the
            // out-projection then adds ctx W_o + b on top of x32_
            launches += attention_fsmn_launch(qkv16_, qkv16_ + d, qkv16_ + 2 * d, ctx16_, B, H, T, 3 * d, d, w.fsmn, cfg_.enc_kernel,
                                              x32_, d, w.in_size == d, stream_);
        });
        gemm(lp.out);
        if (lp.ffn1_ln.valid) {
            ln_gemm(lp.ffn1_ln);                           // norm2 + FFN1 in one kernel
            gemm(lp.ffn2);
            return;
        }
        if (!lp.ln2_fused) {
            if (!(dbg_skip() & 1)) timed("enc_layernorm", [&] { layernorm_f32_launch(x32_, d, M, d, w.ln2.g, w.ln2.b, cfg_.ln_eps, a16_, d, nullptr, 0, stream_); });
            ++launches;
        }
#ifdef PFASR_EXPERIMENTS
        if (lp.chain.valid) {
            ffn_chain(lp.chain);
        } else
#endif
        {
            gemm(lp.ffn1);
            gemm(lp.ffn2);
        }
    };
    size_t li = 0;
    // layer 0's norm1 is part of embed_pe_ln; later layers get it from the previous layer's FFN2 epilogue when fused
    for (size_t i = 0; i < enc_.size(); ++i, ++li) layer(enc_[i], plan.layers[li], i == 0 || plan.layers[li - 1].next_ln1_fused);
    if (tp_.empty()) {
        layernorm_f32_launch(x32_, d, M, d, after_norm_.g, after_norm_.b, cfg_.ln_eps, enc16_, d, enc32_, d, stream_);
        ++launches;
    } else {
        layernorm_f32_launch(x32_, d, M, d, after_norm_.g, after_norm_.b, cfg_.ln_eps, nullptr, 0, x32_, d, stream_);
        ++launches;
        for (size_t i = 0; i < tp_.size(); ++i, ++li) layer(tp_[i], plan.layers[li], i > 0 && plan.layers[li - 1].next_ln1_fused);
        layernorm_f32_launch(x32_, d, M, d, tp_norm_.g, tp_norm_.b, cfg_.ln_eps, enc16_, d, enc32_, d, stream_);
        ++launches;
    }
}

void DeviceCtx::predictor_forward(int B, int T, bool online, bool kv_later) {
    const int d = cfg_.d_model;
    EncoderPlan& plan = encoder_plan(B, T);
    // decoder K/V of all layers: independent of the CIF result.  Offline it is enqueued AFTER the token counts are on their way to the
    // host (kv_later), so that its ~76 us cover the host's wake-up and the first decoder launches instead of preceding an idle gap
    if (!kv_later) gemm(plan.kv_all);
    im2col3_launch(enc16_, B, T, d, qkv16_, stream_);
    gemm(plan.pred_conv);
    alpha_head_launch(mem32_, B, T, d, w_alpha_, b_alpha_, cfg_.smooth_factor, cfg_.noise_threshold, online ? 0.0f : cfg_.cif_tail,
                      alphas_, stream_);
    PF_CUDA(cudaMemsetAsync(meta_, 0, 4 * sizeof(int), stream_));
    launches += 2;
    if (online) return;                                 // the streaming CIF (with carry, no tail) is online_cif_launch
    cif_scan_launch(alphas_, B, T + 1, cfg_.cif_threshold, wcur_, wrem_, fire_idx_, peaks_, token_num_, fires_, meta_, stream_);
    ++launches;
}

// One SANM decoder stack in place on xd32_ [B*L, d]: per layer  t = FFN(LN(x)); x += FSMN(LN(t)); x += CrossAtt(LN(x), memory)
// followed by the FFN-only decoders3 layer; the result (before after_norm) is left in t32_.  kv16: K|V projections of
// the memory for every layer of the stack ([rows, ldkv]); kv_shared: the memory is the same for every batch item (SeACo
// hot-word rows), so all B*L queries attend one [Tk] memory.
void DeviceCtx::dec_stack(const std::vector<DecLayerW>& layers, const DecFfnW& d3, const std::vector<DecLayerPlan>& lps,
                          const GemmOp& d3_w1, const GemmOp& d3_w2, int ffn_width, int kernel, const __half* kv16, int ldkv,
                          bool kv_shared, int Tk, int B, int L, bool online) {
    const int Md = B * L, d = cfg_.d_model, f = ffn_width, H = cfg_.heads;
    const float eps = cfg_.ln_eps;
    // Q11 (OnlineModel.cs:222): stack_states hands every layer the stream's LAYER-0 cache; online_flags bit 0 opts out
    const bool per_layer_cache = (cfg_.online_flags & 1) != 0;
    const size_t cache_layer = static_cast<size_t>(cfg_.dec_kernel - 1) * d;
    auto ffn = [&](const DecFfnW& w, const GemmOp& g1, const GemmOp& g2, bool have_ln) {
        if (!have_ln) {                                    // else: ad16_ = LN(x) came out of the previous out-projection
            if (!(dbg_skip() & 4)) timed("dec_layernorm", [&] { layernorm_f32_launch(xd32_, d, Md, d, w.ln_in.g, w.ln_in.b, eps, ad16_, d, nullptr, 0, stream_); });
            ++launches;
        }
        gemm(g1);
        if (!(dbg_skip() & 4)) timed("dec_layernorm_ffn", [&] { layernorm_f32_launch(hd32_, f, Md, f, w.ln_mid.g, w.ln_mid.b, eps, hd16_, f, nullptr, 0, stream_); });
        gemm(g2);
        ++launches;
    };
    for (size_t i = 0; i < layers.size(); ++i) {
        const DecLayerW& w = layers[i];
        const DecLayerPlan& lp = lps[i];
        ffn(w.ffn, lp.w1, lp.w2, i > 0 && lps[i - 1].next_ln1_fused);
        if (!online) {
            // norm2 -> FSMN memory -> residual -> norm3 in one kernel (three launches otherwise)
            if (!(dbg_skip() & 12)) timed("dec_ln_fsmn_ln", [&] {
                dec_ln_fsmn_ln_launch(t32_, xd32_, w.ln2.g, w.ln2.b, w.fsmn, kernel, w.ln3.g, w.ln3.b, token_num_, B, L, d, eps, ad16_, stream_);
            });
            launches -= 2;
        } else {
            timed("dec_layernorm", [&] { layernorm_f32_launch(t32_, d, Md, d, w.ln2.g, w.ln2.b, eps, nullptr, 0, tn32_, d, stream_); });
            timed("dec_fsmn", [&] {
                online_fsmn_launch(otab_, B, tn32_, fires_, L, d, w.fsmn, kernel, ofsmn_, fsmn_state_stride(),
                                   per_layer_cache ? i * cache_layer : 0, xd32_, ocache_new_, fsmn_state_stride(), i * cache_layer, stream_);
            });
            timed("dec_layernorm", [&] { layernorm_f32_launch(xd32_, d, Md, d, w.ln3.g, w.ln3.b, eps, ad16_, d, nullptr, 0, stream_); });
        }
        gemm(lp.q);
        if (!(dbg_skip() & 16)) timed("dec_cross_attention", [&] {
            if (kv_shared) attention_launch(q16_, kv16 + i * 2 * d, kv16 + i * 2 * d + d, ctxd16_, 1, H, Md, Tk, d, ldkv, ldkv, d, d / H, stream_,
                                            xatt_ws_, xatt_ws_bytes_);
            else attention_launch(q16_, kv16 + i * 2 * d, kv16 + i * 2 * d + d, ctxd16_, B, H, L, Tk, d, ldkv, ldkv, d, d / H, stream_);
        });
        gemm(lp.out);
        launches += 4;
    }
    ffn(d3, d3_w1, d3_w2, !lps.empty() && lps.back().next_ln1_fused);
}

void DeviceCtx::decoder_forward(int B, int T, int L, bool online) {
    const int Md = B * L, d = cfg_.d_model;
    DecoderPlan& plan = decoder_plan(B, T, L);
    const float eps = cfg_.ln_eps;
    const bool seaco = cfg_.model_kind == PF_MODEL_SEACO_PARAFORMER && nbias_ > 0;
    if (!online) {                                      // streaming: xd32_ already holds the compacted CIF frames
        PF_CUDA(cudaMemsetAsync(xd32_, 0, static_cast<size_t>(Md) * d * sizeof(float), stream_));
        cif_gather_launch(enc32_, B, T, d, wcur_, wrem_, fire_idx_, T + 1, xd32_, L, stream_);
        ++launches;
    }
    if (seaco) PF_CUDA(cudaMemcpyAsync(emb32_, xd32_, static_cast<size_t>(Md) * d * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
    dec_stack(dec_, dec3_, plan.layers, plan.d3_w1, plan.d3_w2, cfg_.dec_ffn, cfg_.dec_kernel, kv16_, cfg_.dec_layers * 2 * d, false, T, B, L, online);
    // after_norm -> fp16 operand of the output layer (+ fp32 copy: the decoder hidden the SeACo branch queries with)
    layernorm_f32_launch(t32_, d, Md, d, dec_after_.g, dec_after_.b, eps, ad16_, d, seaco ? hid32_ : nullptr, d, stream_);
    if (want_logits_ || seaco) {
        gemm(plan.head);
        // offline model_out is log-softmax; the streaming decoder graph returns raw logits (the pick is the same)
        logsoftmax_argmax_launch(logits_, Md, cfg_.vocab, ldv(), tokens_, online ? 0 : 1, stream_);
    } else {
        // ids only (the default): the greedy pick rides on the head GEMM's epilogue, the [B*L, V] log-probs are never written
        gemm(plan.head_pick);
        pick_combine_launch(pick_, Md, pick_ld_, ceil_div(cfg_.vocab, plan.head_pick.bn) * 2, tokens_, stream_);
    }
    launches += 2;
    if (seaco) seaco_forward(B, L, plan);
}

// CifPredictorV3.get_upsample_timestmap (row f1): us_alphas / us_cif_peak [B, 3T] from the encoder output and token_num.
void DeviceCtx::timestamp_forward(int B, int T) {
    const int d = cfg_.d_model, M = B * T, T3 = 3 * T;
    if (B > ts_capB_ || T > ts_capT_) {
        PF_CUDA(cudaStreamSynchronize(stream_));
        free_pool(tpool_);
        ts_capB_ = std::max(ts_capB_, B);
        ts_capT_ = std::max(ts_capT_, T);
        const size_t rows = static_cast<size_t>(ts_capB_) * 3 * ts_capT_;
        gin32_ = dalloc<float>(rows * 8 * d, tpool_);
        y32_ = dalloc<float>(rows * 2 * d, tpool_);
        hbuf_ = dalloc<float>(static_cast<size_t>(2) * 2 * kLstmMaxBatch * d, tpool_);
        lstm_bar_ = dalloc<unsigned int>(2, tpool_);
        us_alphas_ = dalloc<float>(rows, tpool_);
        us_peaks_ = dalloc<float>(rows, tpool_);
    }
    // ConvTranspose1d(k = stride = 3) == one GEMM: row (b, t) -> its three output frames [3t, 3t+1, 3t+2]; qkv16_ is free by now
    GemmOp up, gi;
    GemmEpi e; e.bias = b_up_; e.out_f16 = qkv16_; e.ld_out = 3 * d;
    gemm_prepare(up, enc16_, d, w_up16_, d, M, 3 * d, d, e);
    gemm(up);
    GemmEpi g; g.bias = b_bi_; g.out_f32 = gin32_; g.ld_out = 8 * d;
    gemm_prepare(gi, qkv16_, d, w_ih_bi16_, d, 3 * M, 8 * d, d, g);          // qkv16_ viewed as [3M, d]
    gemm(gi);
    timed("ts_bilstm", [&] {
        for (int b0 = 0; b0 < B; b0 += kLstmMaxBatch) {
            const int nb = std::min(kLstmMaxBatch, B - b0);
            bilstm_launch(gin32_ + static_cast<size_t>(b0) * T3 * 8 * d, w_hh_bi16_, nb, T3, d, y32_ + static_cast<size_t>(b0) * T3 * 2 * d, hbuf_,
                          lstm_bar_, stream_);
            ++launches;
        }
    });
    timed("ts_alphas_peaks", [&] {
        us_alphas_peaks_launch(y32_, B, T3, 2 * d, w_out2_, b_out2_, cfg_.smooth_factor2, cfg_.noise_threshold2, token_num_,
                               cfg_.cif_threshold - 1e-4f, us_alphas_, us_peaks_, stream_);
    });
    launches += 2;
}

// SeACo bias branch (FunASR SeacoParaformer export [EXT]): the bias decoder attends the hot-word rows
// twice (queries: CIF embeds, decoder hidden); sum -> hotword_output_layer -> log-softmax dha; rows whose dha pick is
// NO_BIAS keep the ASR posterior, the others take dha.
void DeviceCtx::seaco_forward(int B, int L, DecoderPlan& plan) {
    const int Md = B * L, d = cfg_.d_model;
    // all B * L queries attend the one hot-word memory: 13 query tiles x 4 heads would walk 2010 keys each; the workspace lets the
    // streaming attention kernel cut the memory into runs (attention.cu)
    const size_t ws_need = attention_split_workspace_bytes(1, cfg_.heads, Md);
    if (ws_need > xatt_ws_bytes_) {
        PF_CUDA(cudaStreamSynchronize(stream_));
        if (xatt_ws_) cudaFree(xatt_ws_);
        PF_CUDA(cudaMalloc(&xatt_ws_, ws_need));
        xatt_ws_bytes_ = ws_need;
    }
    const size_t bytes = static_cast<size_t>(Md) * d * sizeof(float);
    const int ldkv = cfg_.seaco_layers * 2 * d;
    const float eps = cfg_.ln_eps;
    for (int pass = 0; pass < 2; ++pass) {
        PF_CUDA(cudaMemcpyAsync(xd32_, pass == 0 ? emb32_ : hid32_, bytes, cudaMemcpyDeviceToDevice, stream_));
        dec_stack(sdec_, sdec3_, plan.slayers, plan.s_d3_w1, plan.s_d3_w2, cfg_.seaco_ffn, cfg_.seaco_kernel, skv16_, ldkv, true, nbias_, B, L, false);
        layernorm_f32_launch(t32_, d, Md, d, sdec_after_.g, sdec_after_.b, eps, nullptr, 0, pass == 0 ? satt32_ : tn32_, d, stream_);
        ++launches;
    }
    add_to_f16_launch(satt32_, tn32_, ad16_, static_cast<size_t>(Md) * d, stream_);
    gemm(plan.hw_head);
    logsoftmax_argmax_launch(logits2_, Md, cfg_.vocab, ldv(), tokens2_, 1, stream_);
    seaco_merge_launch(tokens2_, logits2_, cfg_.seaco_nobias_id, Md, cfg_.vocab, ldv(), tokens_, logits_, stream_);
    launches += 3;
}

// EmbedSeacoModel.Forward (EmbedSeacoModel.cs:70-108): Embedding -> 2-layer LSTM over 10 steps (time-major), then the
// Q8 bias_embed layout (all steps of every hot word, hot-word major) and the K|V projections of the bias decoder.
void DeviceCtx::set_hotwords(const int32_t* ids, int n) {
    PF_CUDA(cudaSetDevice(dev_));
    if (cfg_.model_kind != PF_MODEL_SEACO_PARAFORMER) throw StatusError{PF_ERR_UNSUPPORTED, "hot words need a seacoparaformer model"};
    PF_CUDA(cudaStreamSynchronize(stream_));
    free_pool(hpool_);
    nbias_ = 0;
    bias16_ = nullptr; skv16_ = nullptr;
    if (n <= 0) return;
    const int d = cfg_.d_model, steps = 10, rows = steps * n;
    for (int i = 0; i < n * steps; ++i)
        if (ids[i] < 0 || ids[i] >= cfg_.vocab) throw StatusError{PF_ERR_BAD_ARG, "hot-word token id out of range"};
    std::vector<void*> tmp;
    try {
        int* d_ids = dalloc<int>(static_cast<size_t>(rows), tmp);
        PF_CUDA(cudaMemcpyAsync(d_ids, ids, static_cast<size_t>(rows) * sizeof(int), cudaMemcpyHostToDevice, stream_));
        __half* x16 = dalloc<__half>(static_cast<size_t>(rows) * d, tmp);          // layer input, time-major [t*n + i]
        __half* y16 = dalloc<__half>(static_cast<size_t>(rows) * d, tmp);          // layer output, time-major
        float* gin = dalloc<float>(static_cast<size_t>(rows) * 4 * d, tmp);        // W_ih x + b for every step
        float* gates = dalloc<float>(static_cast<size_t>(n) * 4 * d, tmp);
        float* cstate = dalloc<float>(static_cast<size_t>(n) * d, tmp);
        bias16_ = dalloc<__half>(static_cast<size_t>(rows) * d, hpool_);           // [n*10, d] hot-word major (Q8)
        skv16_ = dalloc<__half>(static_cast<size_t>(rows) * cfg_.seaco_layers * 2 * d, hpool_);
        embed_rows_tmajor_launch(bias_table_, d_ids, n, steps, d, x16, stream_);
        for (int layer = 0; layer < 2; ++layer) {
            const LstmW& w = lstm_[layer];
            GemmOp gi;
            GemmEpi e; e.bias = w.b; e.out_f32 = gin; e.ld_out = 4 * d;
            gemm_prepare(gi, x16, d, w.w_ih, d, rows, 4 * d, d, e);
            gemm_launch(gi, stream_);
            PF_CUDA(cudaMemsetAsync(cstate, 0, static_cast<size_t>(n) * d * sizeof(float), stream_));
            for (int t = 0; t < steps; ++t) {
                const float* g_t = gin + static_cast<size_t>(t) * n * 4 * d;
                if (t > 0) {
                    GemmOp gh;
                    GemmEpi eh; eh.resid = g_t; eh.ld_resid = 4 * d; eh.out_f32 = gates; eh.ld_out = 4 * d;
                    gemm_prepare(gh, y16 + static_cast<size_t>(t - 1) * n * d, d, w.w_hh, d, n, 4 * d, d, eh);
                    gemm_launch(gh, stream_);
                    g_t = gates;
                }
                // last layer also writes the Q8 layout: row (i * 10 + t) of bias_embed
                lstm_cell_launch(g_t, cstate, n, d, y16 + static_cast<size_t>(t) * n * d, layer == 1 ? bias16_ : nullptr, t, steps, stream_);
            }
            std::swap(x16, y16);
        }
        GemmOp gk;
        GemmEpi ek; ek.bias = b_skv_all_; ek.out_f16 = skv16_; ek.ld_out = cfg_.seaco_layers * 2 * d;
        gemm_prepare(gk, bias16_, d, w_skv_all_, d, rows, cfg_.seaco_layers * 2 * d, d, ek);
        gemm_launch(gk, stream_);
        PF_CUDA(cudaStreamSynchronize(stream_));
    } catch (...) {
        cudaStreamSynchronize(stream_);
        free_pool(tmp);
        free_pool(hpool_);
        throw;
    }
    free_pool(tmp);
    nbias_ = rows;
    dec_plans_.clear();
}

void DeviceCtx::run(uint32_t flags, SharedRun* shared, int idx) {
    arrived_ = false;
    if (profile_ == 1) replay_.clear();
    try {
        run_impl(flags, shared, idx);
    } catch (...) {
        // never leave the sibling device threads parked on the Lmax exchange
        if (shared && !arrived_) { shared->failed = true; shared->lmax[idx] = 0; arrived_ = true; shared->barrier.arrive_and_wait(); }
        ev0_armed_ = false;
        throw;
    }
    ev0_armed_ = false;
}

static double host_ms_since(const std::chrono::steady_clock::time_point& t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void DeviceCtx::run_impl(uint32_t flags, SharedRun* shared, int idx) {
    PF_CUDA(cudaSetDevice(dev_));
    const auto host_t0 = std::chrono::steady_clock::now();     // host-side view: enqueue vs blocked time (timings_ms[6..9])
    if (!ev0_armed_) PF_CUDA(cudaEventRecord(ev_[0], stream_));   // run_staged: time from here, not from staging
    const bool sv = cfg_.model_kind == PF_MODEL_SENSEVOICE_SMALL;
    const int B = staged_B_;
    int T = staged_T_;
    launches = 0;
    gemm_flops = 0.0;
    want_logits_ = (flags & PF_RUN_WANT_LOGITS) != 0;
    if (B <= 0) throw StatusError{PF_ERR_BAD_ARG, "nothing staged"};
    const int din = cfg_.input_size;
    // ---- front-end
    if (staged_is_pcm_) {
        float* dst = sv ? feats_raw_ : feats_;
        for (int g = 0; g < staged_groups_; ++g) {
            const int b0 = static_cast<int>(static_cast<long long>(B) * g / staged_groups_);
            const int b1 = static_cast<int>(static_cast<long long>(B) * (g + 1) / staged_groups_);
            PF_CUDA(cudaStreamWaitEvent(stream_, ev_grp_[g], 0));          // this group's PCM has landed
            if (T <= 0 || b1 <= b0) continue;
            if (staged_audio_) {       // raw file samples -> float mono 16 kHz PCM (AudioHelper.GetFileSample), csrc/audio.cu
                audio_convert_launch(raw_, d_audio_ + b0, pcm_, d_off_ + b0, d_meta_ + b0, b1 - b0, staged_max_nsamp_, stream_);
                ++launches;
            }
            FrontendLaunch a;
            a.tables = fe_tables_; a.pcm = pcm_; a.pcm_off = d_off_ + b0; a.nsamp = d_meta_ + b0; a.nframes = d_meta_ + B + b0;
            a.nlfr = d_meta_ + 2 * B + b0;
            a.add_shift = cmvn_shift_; a.rescale = cmvn_scale_;
            a.feats_out = dst; a.feats_off = d_off_ + B + b0;
            a.batch = b1 - b0; a.max_frames = staged_maxframes_; a.tmax_lfr = T; a.lfr_m = cfg_.lfr_m; a.lfr_n = cfg_.lfr_n;
            a.snip_edges = cfg_.snip_edges != 0;
            a.pad_quirk = true; a.pad_fill = true;
            a.pad_value = -23.025850929940457f * 32768.0f;      // Utils/PadHelper.cs:63
            frontend_launch(a, stream_);
            launches += 2;
        }
        if (sv) {
            // Q6: language slot = textnorm id (14 with use_itn else 15), textnorm slot = 15; Q7: [lang, 1, 2, textnorm]
            const int ids[4] = {cfg_.use_itn ? 14 : 15, 1, 2, 15};
            memcpy(h_meta_ + 8, ids, sizeof(ids));
            PF_CUDA(cudaMemcpyAsync(prompt_ids_, h_meta_ + 8, sizeof(ids), cudaMemcpyHostToDevice, stream_));
            prepend_rows_launch(feats_raw_, embed_table_, prompt_ids_, 4, feats_, B, T, din, stream_);
            ++launches;
            T += 4;
        }
    }
    B_ = B;
    T_ = T;
    PF_CUDA(cudaEventRecord(ev_[1], stream_));
    if (T <= 0) {   // no frames at all (utterances shorter than one LFR frame): empty result, like a [B,0,V] tensor
        Lmax_ = 0; Lpad_ = 0;
        ensure_pinned(reinterpret_cast<void**>(&h_token_num), &h_tn_cap_, static_cast<size_t>(B) * sizeof(int32_t));
        for (int b = 0; b < B; ++b) h_token_num[b] = 0;
        if (shared) { shared->lmax[idx] = 0; arrived_ = true; shared->barrier.arrive_and_wait(); }
        PF_CUDA(cudaStreamSynchronize(stream_));
        return;
    }
    // ---- encoder
    encoder_forward(B, T);
    PF_CUDA(cudaEventRecord(ev_[2], stream_));
    ensure_pinned(reinterpret_cast<void**>(&h_token_num), &h_tn_cap_, static_cast<size_t>(B) * sizeof(int32_t));
    if (sv) {
        const int M = B * T;
        if (flags & PF_RUN_WANT_LOGITS) {
            gemm(encoder_plan(B, T).ctc_head);
            PF_CUDA(cudaEventRecord(ev_[3], stream_));
            PF_CUDA(cudaEventRecord(ev_[4], stream_));
            logsoftmax_argmax_launch(logits_, M, cfg_.vocab, ldv(), tokens_, 1, stream_);
        } else {
            // per-frame ids only: fused pick in the CTC head's epilogue, the [B, T, 25055] tensor (879 MB at cfg3) is never written
            EncoderPlan& ep = encoder_plan(B, T);
            gemm(ep.ctc_head_pick);
            PF_CUDA(cudaEventRecord(ev_[3], stream_));
            PF_CUDA(cudaEventRecord(ev_[4], stream_));
            pick_combine_launch(pick_, M, pick_ld_, ceil_div(cfg_.vocab, ep.ctc_head_pick.bn) * 2, tokens_, stream_);
        }
        ++launches;
        PF_CUDA(cudaEventRecord(ev_[5], stream_));
        Lmax_ = T; Lpad_ = T;
        if (shared) { shared->lmax[idx] = T; arrived_ = true; shared->barrier.arrive_and_wait(); }
        ensure_host(static_cast<size_t>(M), (flags & PF_RUN_WANT_LOGITS) ? static_cast<size_t>(M) * cfg_.vocab : 0, 0);
        PF_CUDA(cudaMemcpyAsync(h_tokens, tokens_, static_cast<size_t>(M) * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        if (flags & PF_RUN_WANT_LOGITS)
            PF_CUDA(cudaMemcpy2DAsync(h_logits, static_cast<size_t>(cfg_.vocab) * sizeof(float), logits_, static_cast<size_t>(ldv()) * sizeof(float),
                                      static_cast<size_t>(cfg_.vocab) * sizeof(float), M, cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaEventRecord(ev_[6], stream_));
        PF_CUDA(cudaStreamSynchronize(stream_));
        for (int b = 0; b < B; ++b) h_token_num[b] = T;
    } else {
        // ---- predictor + CIF scan; the token count decides the decoder's shape -> one small D2H + sync
        predictor_forward(B, T, false, true);
        PF_CUDA(cudaMemcpyAsync(h_meta_, meta_, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaMemcpyAsync(h_token_num, token_num_, static_cast<size_t>(B) * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaEventRecord(ev_[3], stream_));
        gemm(encoder_plan(B, T).kv_all);                                       // runs while the host reads the counts and enqueues the decoder
        timings_ms[6] = static_cast<float>(host_ms_since(host_t0));           // front-end + encoder + predictor enqueued
        PF_CUDA(cudaEventSynchronize(ev_[3]));
        timings_ms[7] = static_cast<float>(host_ms_since(host_t0));           // ... and finished (token counts on the host)
        int lmax = h_meta_[0];
        if (shared) {
            shared->lmax[idx] = lmax;
            arrived_ = true; shared->barrier.arrive_and_wait();
            for (int v : shared->lmax) lmax = std::max(lmax, v);
        }
        Lmax_ = lmax;
        Lpad_ = lmax;
        if (lmax > 0) {
            decoder_forward(B, T, lmax);
            PF_CUDA(cudaEventRecord(ev_[4], stream_));
            PF_CUDA(cudaEventRecord(ev_[5], stream_));
            const size_t ntok = static_cast<size_t>(B) * lmax;
            const bool wl = (flags & PF_RUN_WANT_LOGITS) != 0, wp = (flags & PF_RUN_WANT_CIF_PEAK) != 0;
            ensure_host(ntok, wl ? ntok * cfg_.vocab : 0, wp ? static_cast<size_t>(B) * (T + 1) : 0);
            PF_CUDA(cudaMemcpyAsync(h_tokens, tokens_, ntok * sizeof(int), cudaMemcpyDeviceToHost, stream_));
            if (wl) PF_CUDA(cudaMemcpy2DAsync(h_logits, static_cast<size_t>(cfg_.vocab) * sizeof(float), logits_, static_cast<size_t>(ldv()) * sizeof(float),
                                              static_cast<size_t>(cfg_.vocab) * sizeof(float), ntok, cudaMemcpyDeviceToHost, stream_));
            if (wp) PF_CUDA(cudaMemcpyAsync(h_peaks, peaks_, static_cast<size_t>(B) * (T + 1) * sizeof(float), cudaMemcpyDeviceToHost, stream_));
            us_frames = 0;
            if ((flags & PF_RUN_WANT_TIMESTAMPS) != 0) {
                if (!has_timestamps()) throw StatusError{PF_ERR_UNSUPPORTED, "this model has no CifPredictorV3 timestamp branch (predictor.upsample_cnn.* missing)"};
                timestamp_forward(B, T);
                const size_t n = static_cast<size_t>(B) * 3 * T;
                ensure_pinned(reinterpret_cast<void**>(&h_us), &h_us_cap_, 2 * n * sizeof(float));
                PF_CUDA(cudaMemcpyAsync(h_us, us_alphas_, n * sizeof(float), cudaMemcpyDeviceToHost, stream_));
                PF_CUDA(cudaMemcpyAsync(h_us + n, us_peaks_, n * sizeof(float), cudaMemcpyDeviceToHost, stream_));
                us_frames = 3 * T;
            }
        } else {
            PF_CUDA(cudaEventRecord(ev_[4], stream_));
            PF_CUDA(cudaEventRecord(ev_[5], stream_));
        }
        PF_CUDA(cudaEventRecord(ev_[6], stream_));
        timings_ms[8] = static_cast<float>(host_ms_since(host_t0));           // decoder enqueued
        PF_CUDA(cudaStreamSynchronize(stream_));
        timings_ms[9] = static_cast<float>(host_ms_since(host_t0));           // results on the host
    }
    finish_profile();
    float ms = 0.0f;
    for (int i = 0; i < 5; ++i) {
        cudaEventElapsedTime(&ms, ev_[i], ev_[i + 1]);
        timings_ms[i] = ms;
    }
    cudaEventElapsedTime(&ms, ev_[0], ev_[6]);
    timings_ms[5] = ms;
}

// ------------------------------------------------------------------ streaming (online) path
void DeviceCtx::online_grow(int min_slots) {
    if (min_slots <= ocap_) return;
    PF_CUDA(cudaSetDevice(dev_));
    if (ocap_ == 0) {
        // geometry: OnlineModel.cs:15-16,30 (chunk 5, lfr 10 -> 60 frames / 9600 samples per chunk); streaming LFR rule
        od_.mel = cfg_.n_mels; od_.lfr_m = cfg_.lfr_m; od_.lfr_n = cfg_.lfr_n; od_.d_model = cfg_.d_model;
        od_.chunk_len = 60;
        od_.nf = frontend_num_frames(160 * od_.chunk_len, cfg_.snip_edges != 0);
        const int t = od_.chunk_len + 1;
        od_.t_new = (t % od_.lfr_n < od_.lfr_m - od_.lfr_n) ? t / od_.lfr_n - 1 : t / od_.lfr_n;
        od_.cache_rows = 10;
        if (od_.t_new != od_.cache_rows || od_.nf < 1)
            throw StatusError{PF_ERR_UNSUPPORTED, "streaming needs lfr_m/lfr_n that turn 61 frames into 10 LFR rows (7/6)"};
        if (cfg_.model_kind != PF_MODEL_PARAFORMER) throw StatusError{PF_ERR_UNSUPPORTED, "streaming is a paraformer path"};
        // Q12 (OnlineWavFrontend.cs:163-170): inv_timescale_i = exp(-(i+1) * ln(1e4) / (dim/2 - 1)), float arithmetic
        const int half = cfg_.input_size / 2;
        std::vector<float> inv(half);
        const float inc = static_cast<float>(log(10000.0)) / static_cast<float>(half - 1);
        for (int i = 0; i < half; ++i) inv[i] = static_cast<float>(exp(static_cast<double>(static_cast<float>(i + 1) * -inc)));
        inv_ts_online_ = up_f32(inv.data(), inv.size());
    }
    int cap = std::max(16, ocap_);
    while (cap < min_slots) cap *= 2;
    // pushed chunks travel on copy_stream_ into the OLD opcm_: they have to land before the buffers are copied and freed
    PF_CUDA(cudaStreamSynchronize(copy_stream_));
    push_pending_ = false;
    PF_CUDA(cudaStreamSynchronize(stream_));
    const size_t dim = static_cast<size_t>(cfg_.lfr_m) * cfg_.n_mels;
    const size_t per[7] = {static_cast<size_t>(od_.nslot) * (od_.nf + 1) * od_.mel, static_cast<size_t>(od_.nslot) * 160 * od_.chunk_len,
                           static_cast<size_t>(od_.mel), od_.cache_rows * dim, 1, static_cast<size_t>(cfg_.d_model), fsmn_state_stride()};
    float** ptrs[7] = {&ofifo_, &opcm_, &osplice_, &ocache_, &ocif_a_, &ocif_h_, &ofsmn_};
    for (int i = 0; i < 7; ++i) {
        float* fresh = nullptr;
        cudaError_t e = cudaMalloc(&fresh, per[i] * cap * sizeof(float));
        if (e != cudaSuccess) { cudaGetLastError(); throw StatusError{PF_ERR_OOM, "cudaMalloc of streaming state failed"}; }
        PF_CUDA(cudaMemsetAsync(fresh, 0, per[i] * cap * sizeof(float), stream_));
        if (*ptrs[i]) {
            PF_CUDA(cudaMemcpyAsync(fresh, *ptrs[i], per[i] * ocap_ * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
            PF_CUDA(cudaStreamSynchronize(stream_));
            cudaFree(*ptrs[i]);
        }
        *ptrs[i] = fresh;
    }
    PF_CUDA(cudaStreamSynchronize(stream_));
    oslots_.resize(cap);
    ocap_ = cap;
}

int DeviceCtx::online_open() {
    PF_CUDA(cudaSetDevice(dev_));
    int slot = -1;
    for (int i = 0; i < ocap_; ++i) if (!oslots_[i].open) { slot = i; break; }
    if (slot < 0) { slot = ocap_; online_grow(ocap_ + 1); }
    // fresh OnlineStream state (OnlineStream.cs:53-61): zero feature cache, CIF carry, FSMN caches, no splice
    const size_t dim = static_cast<size_t>(cfg_.lfr_m) * cfg_.n_mels;
    PF_CUDA(cudaMemsetAsync(osplice_ + static_cast<size_t>(slot) * od_.mel, 0, od_.mel * sizeof(float), stream_));
    PF_CUDA(cudaMemsetAsync(ocache_ + static_cast<size_t>(slot) * od_.cache_rows * dim, 0, od_.cache_rows * dim * sizeof(float), stream_));
    PF_CUDA(cudaMemsetAsync(ocif_a_ + slot, 0, sizeof(float), stream_));
    PF_CUDA(cudaMemsetAsync(ocif_h_ + static_cast<size_t>(slot) * cfg_.d_model, 0, cfg_.d_model * sizeof(float), stream_));
    PF_CUDA(cudaMemsetAsync(ofsmn_ + static_cast<size_t>(slot) * fsmn_state_stride(), 0, fsmn_state_stride() * sizeof(float), stream_));
    oslots_[slot] = OnlineSlot{};
    oslots_[slot].open = true;
    return slot;
}

void DeviceCtx::online_close(int slot) {
    if (slot < 0 || slot >= ocap_ || !oslots_[slot].open) throw StatusError{PF_ERR_BAD_ARG, "stream slot is not open"};
    oslots_[slot].open = false;
}

void DeviceCtx::online_push_chunk(int slot, const float* samples, int nsamp) {
    PF_CUDA(cudaSetDevice(dev_));
    if (slot < 0 || slot >= ocap_ || !oslots_[slot].open) throw StatusError{PF_ERR_BAD_ARG, "stream slot is not open"};
    const int chunk = 160 * od_.chunk_len;
    if (nsamp != chunk) throw StatusError{PF_ERR_SHAPE, "a streaming chunk is 160 * 60 samples"};
    OnlineSlot& s = oslots_[slot];
    // the oldest chunk a future window still reads must not be overwritten
    const long long base = static_cast<long long>(od_.chunk_len) * s.dec;
    const long long c_min = base < od_.nf + 1 ? 0 : 1 + (base - od_.nf - 1) / od_.nf;
    if (s.pushed - c_min >= od_.nslot)
        throw StatusError{PF_ERR_SHAPE, "stream backlog exceeds " + std::to_string(od_.nslot) + " chunks: call the recognizer's GetResults"};
    float* dst = opcm_ + (static_cast<size_t>(slot) * od_.nslot + static_cast<size_t>(s.pushed % od_.nslot)) * chunk;
    // bounce through a pinned ring: the caller's buffer is pageable (a managed float[] in the reference) and a pageable
    // cudaMemcpyAsync stalls the host for the whole staged copy; from pinned memory the push returns after a 38 KB memcpy
    constexpr int kRing = 256;
    if (!h_push_ring_) PF_CUDA(cudaMallocHost(&h_push_ring_, static_cast<size_t>(kRing) * chunk * sizeof(float)));
    if (push_ring_pos_ == kRing) {                       // wrapped: the copies issued from the ring have to be done
        PF_CUDA(cudaStreamSynchronize(copy_stream_));
        push_ring_pos_ = 0;
    }
    float* bounce = h_push_ring_ + static_cast<size_t>(push_ring_pos_++) * chunk;
    memcpy(bounce, samples, static_cast<size_t>(chunk) * sizeof(float));
    PF_CUDA(cudaMemcpyAsync(dst, bounce, static_cast<size_t>(chunk) * sizeof(float), cudaMemcpyHostToDevice, copy_stream_));
    push_pending_ = true;
    ++s.pushed;
}

bool DeviceCtx::online_ready(int slot) const {
    if (slot < 0 || slot >= ocap_ || !oslots_[slot].open) return false;
    const OnlineSlot& s = oslots_[slot];
    const long long avail = s.pushed > 0 ? s.pushed * od_.nf + 1 : 0;      // first chunk repeats its first frame
    return static_cast<long long>(od_.chunk_len) * (s.dec + 1) <= avail;
}

void DeviceCtx::online_step(const std::vector<int>& slots, uint32_t flags, SharedRun* shared, int idx) {
    arrived_ = false;
    try {
        online_step_impl(slots, flags, shared, idx);
    } catch (...) {
        if (shared && !arrived_) { shared->failed = true; shared->lmax[idx] = 0; arrived_ = true; shared->barrier.arrive_and_wait(); }
        throw;
    }
}

void DeviceCtx::online_step_impl(const std::vector<int>& slots, uint32_t flags, SharedRun* shared, int idx) {
    PF_CUDA(cudaSetDevice(dev_));
    launches = 0;
    gemm_flops = 0.0;
    want_logits_ = (flags & PF_RUN_WANT_LOGITS) != 0;
    online_working.clear();
    Lmax_ = 0; Lpad_ = 0; B_ = 0; T_ = 0;
    PF_CUDA(cudaEventRecord(ev_[0], stream_));
    if (push_pending_) {                                  // pushed chunks travel on the copy stream
        PF_CUDA(cudaEventRecord(ev_grp_[0], copy_stream_));
        PF_CUDA(cudaStreamWaitEvent(stream_, ev_grp_[0], 0));
        push_pending_ = false;
    }
    const int chunk = 160 * od_.chunk_len;
    // ---- per-chunk fbank (InputSpeech, OnlineStream.cs:114-167) for everything pushed since the last step, one launch
    int npend = 0;
    for (int slot : slots) {
        if (slot < 0 || slot >= ocap_ || !oslots_[slot].open) throw StatusError{PF_ERR_BAD_ARG, "stream slot is not open"};
        npend += static_cast<int>(oslots_[slot].pushed - oslots_[slot].fbanked);
    }
    const int nlist = static_cast<int>(slots.size());
    const size_t stage_bytes = static_cast<size_t>(std::max(npend, 1)) * (2 * sizeof(long long) + 3 * sizeof(int)) + static_cast<size_t>(std::max(nlist, 1)) * 4 * sizeof(int);
    ensure_pinned(&h_ostage_, &h_ostage_bytes_, stage_bytes);
    long long* h_off = static_cast<long long*>(h_ostage_);
    int* h_m = reinterpret_cast<int*>(h_off + 2 * std::max(npend, 1));
    int* h_tab = h_m + 3 * std::max(npend, 1);
    if (npend > 0) {
        if (npend > ofe_cap_) {
            if (ofe_off_) cudaFree(ofe_off_);
            if (ofe_meta_) cudaFree(ofe_meta_);
            ofe_cap_ = std::max(npend, 2 * ofe_cap_);
            PF_CUDA(cudaMalloc(&ofe_off_, static_cast<size_t>(ofe_cap_) * 2 * sizeof(long long)));
            PF_CUDA(cudaMalloc(&ofe_meta_, static_cast<size_t>(ofe_cap_) * 3 * sizeof(int)));
        }
        int k = 0;
        for (int slot : slots) {
            OnlineSlot& s = oslots_[slot];
            for (long long c = s.fbanked; c < s.pushed; ++c, ++k) {
                const size_t cs = static_cast<size_t>(slot) * od_.nslot + static_cast<size_t>(c % od_.nslot);
                h_off[k] = static_cast<long long>(cs * chunk);
                // chunk 0 keeps row 0 free: it aliases row 1 (first-frame repeat, OnlineStream.cs:141-153)
                h_off[npend + k] = static_cast<long long>((cs * (od_.nf + 1) + (c == 0 ? 1 : 0)) * od_.mel);
                h_m[k] = chunk;
                h_m[npend + k] = od_.nf;
                h_m[2 * npend + k] = 0;
            }
            s.fbanked = s.pushed;
        }
        PF_CUDA(cudaMemcpyAsync(ofe_off_, h_off, static_cast<size_t>(npend) * 2 * sizeof(long long), cudaMemcpyHostToDevice, stream_));
        PF_CUDA(cudaMemcpyAsync(ofe_meta_, h_m, static_cast<size_t>(npend) * 3 * sizeof(int), cudaMemcpyHostToDevice, stream_));
        FrontendLaunch a;
        a.tables = fe_tables_; a.pcm = opcm_; a.pcm_off = ofe_off_; a.nsamp = ofe_meta_; a.nframes = ofe_meta_ + npend; a.nlfr = ofe_meta_ + 2 * npend;
        a.add_shift = cmvn_shift_; a.rescale = cmvn_scale_;
        a.fbank_out = ofifo_; a.fbank_off = ofe_off_ + npend;
        a.batch = npend; a.max_frames = od_.nf; a.tmax_lfr = 0; a.lfr_m = cfg_.lfr_m; a.lfr_n = cfg_.lfr_n;
        a.snip_edges = cfg_.snip_edges != 0;
        frontend_launch(a, stream_);
        ++launches;
    }
    // ---- working set (GetDecodeChunk returns null without a full chunk: the stream is skipped, OnlineRecognizer.cs:357-361)
    for (int i = 0; i < nlist; ++i) if (online_ready(slots[i])) online_working.push_back(i);
    const int B = static_cast<int>(online_working.size());
    PF_CUDA(cudaEventRecord(ev_[1], stream_));
    auto finish = [&] {
        PF_CUDA(cudaEventRecord(ev_[6], stream_));
        PF_CUDA(cudaStreamSynchronize(stream_));
        finish_profile();
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev_[0], ev_[6]);
        timings_ms[5] = ms;
    };
    if (B == 0) {
        if (shared) { shared->lmax[idx] = 0; arrived_ = true; shared->barrier.arrive_and_wait(); }
        finish();
        return;
    }
    const int T = od_.cache_rows + od_.t_new, d = cfg_.d_model;
    ensure_workspace(B, T);
    if (B > owork_cap_) {
        for (void* p : opool_) cudaFree(p);
        opool_.clear();
        owork_cap_ = std::max(B, 2 * owork_cap_);
        ofresh_ = dalloc<float>(static_cast<size_t>(owork_cap_) * od_.t_new * cfg_.input_size, opool_);
        ocache_new_ = dalloc<float>(static_cast<size_t>(owork_cap_) * fsmn_state_stride(), opool_);
        otab_ = dalloc<int>(static_cast<size_t>(owork_cap_) * 4, opool_);
    }
    for (int b = 0; b < B; ++b) {
        const int slot = slots[online_working[b]];
        const OnlineSlot& s = oslots_[slot];
        h_tab[4 * b + 0] = slot;
        h_tab[4 * b + 1] = static_cast<int>(static_cast<long long>(od_.chunk_len) * s.dec);
        h_tab[4 * b + 2] = s.dec > 0 ? 1 : 0;                            // _cachelfrSplice is empty before the first window
        h_tab[4 * b + 3] = static_cast<int>(od_.t_new * s.dec);          // _startIdx
    }
    PF_CUDA(cudaMemcpyAsync(otab_, h_tab, static_cast<size_t>(B) * 4 * sizeof(int), cudaMemcpyHostToDevice, stream_));
    online_assemble_launch(od_, otab_, B, ofifo_, osplice_, ocache_, cmvn_shift_, cmvn_scale_, inv_ts_online_, feats_, ofresh_, stream_);
    online_commit_launch(od_, otab_, B, ofifo_, ofresh_, osplice_, ocache_, stream_);
    launches += 2;
    for (int b = 0; b < B; ++b) ++oslots_[slots[online_working[b]]].dec;
    B_ = B; T_ = T;
    // ---- encoder.onnx: SAN-M encoder + alpha head (OnlineRecognizer.EncoderProj)
    encoder_forward(B, T, true);
    PF_CUDA(cudaEventRecord(ev_[2], stream_));
    predictor_forward(B, T, true);
    // ---- DynamicMask + host CIF with carry (OnlineRecognizer.PredictorProj); frames land in tn32_ as [B, T+1, d]
    const int lcap = T + 1;
    online_cif_launch(otab_, B, enc32_, alphas_, T + 1, T, d, 5, 15, cfg_.cif_threshold, ocif_a_, ocif_h_, tn32_, lcap, fires_, meta_, stream_);
    ++launches;
    ensure_pinned(reinterpret_cast<void**>(&h_online_counts), &h_ocounts_cap_, static_cast<size_t>(B) * sizeof(int32_t));
    PF_CUDA(cudaMemcpyAsync(h_meta_, meta_, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    PF_CUDA(cudaMemcpyAsync(h_online_counts, fires_, static_cast<size_t>(B) * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    PF_CUDA(cudaEventRecord(ev_[3], stream_));
    PF_CUDA(cudaStreamSynchronize(stream_));
    int lmax = h_meta_[0];
    if (shared) {
        shared->lmax[idx] = lmax;
        arrived_ = true; shared->barrier.arrive_and_wait();
        for (int v : shared->lmax) lmax = std::max(lmax, v);
    }
    if (lmax > lcap) throw StatusError{PF_ERR_SHAPE, "more CIF fires than frames in one chunk"};
    Lmax_ = lmax; Lpad_ = lmax;
    if (lmax > 0) {                                                       // Acoustic_embeds.Length > 0 (OnlineRecognizer.cs:381)
        online_compact_launch(tn32_, lcap, fires_, B, lmax, d, xd32_, stream_);
        ++launches;
        decoder_forward(B, T, lmax, true);
        online_scatter_launch(otab_, B, ocache_new_, ofsmn_, fsmn_state_stride(), stream_);
        ++launches;
        const size_t ntok = static_cast<size_t>(B) * lmax;
        const bool wl = (flags & PF_RUN_WANT_LOGITS) != 0;
        ensure_host(ntok, wl ? ntok * cfg_.vocab : 0, 0);
        PF_CUDA(cudaMemcpyAsync(h_tokens, tokens_, ntok * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        if (wl) PF_CUDA(cudaMemcpy2DAsync(h_logits, static_cast<size_t>(cfg_.vocab) * sizeof(float), logits_, static_cast<size_t>(ldv()) * sizeof(float),
                                          static_cast<size_t>(cfg_.vocab) * sizeof(float), ntok, cudaMemcpyDeviceToHost, stream_));
    }
    PF_CUDA(cudaEventRecord(ev_[4], stream_));
    PF_CUDA(cudaEventRecord(ev_[5], stream_));
    finish();
    float ms = 0.0f;
    for (int i = 0; i < 5; ++i) {
        cudaEventElapsedTime(&ms, ev_[i], ev_[i + 1]);
        timings_ms[i] = ms;
    }
}

void DeviceCtx::online_get_state(int slot, const std::string& name, float* dst, size_t capacity) {
    PF_CUDA(cudaSetDevice(dev_));
    if (slot < 0 || slot >= ocap_ || !oslots_[slot].open) throw StatusError{PF_ERR_BAD_ARG, "stream slot is not open"};
    const float* src = nullptr;
    size_t n = 0;
    const size_t dim = static_cast<size_t>(cfg_.lfr_m) * cfg_.n_mels;
    if (name == "cache_feats") { src = ocache_ + static_cast<size_t>(slot) * od_.cache_rows * dim; n = od_.cache_rows * dim; }
    else if (name == "cif_alpha") { src = ocif_a_ + slot; n = 1; }
    else if (name == "cif_hidden") { src = ocif_h_ + static_cast<size_t>(slot) * cfg_.d_model; n = cfg_.d_model; }
    else if (name == "fsmn") { src = ofsmn_ + static_cast<size_t>(slot) * fsmn_state_stride(); n = fsmn_state_stride(); }   // [layers, k-1, d]
    else if (name == "splice") { src = osplice_ + static_cast<size_t>(slot) * od_.mel; n = od_.mel; }
    else throw StatusError{PF_ERR_BAD_ARG, "unknown stream state '" + name + "'"};
    const size_t cnt = std::min(n, capacity);
    if (dst && cnt) {
        PF_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaStreamSynchronize(stream_));
    }
}

void DeviceCtx::get_tensor(const std::string& name, float* dst, size_t capacity, int32_t* dims4, int32_t* ndim) {
    PF_CUDA(cudaSetDevice(dev_));
    const float* src = nullptr;
    int dims[4] = {1, 1, 1, 1};
    int nd = 0;
    const int d = cfg_.d_model;
    if (name == "feats") { src = feats_; dims[0] = B_; dims[1] = T_; dims[2] = cfg_.input_size; nd = 3; }
    else if (name == "enc") { src = enc32_; dims[0] = B_; dims[1] = T_; dims[2] = d; nd = 3; }
    else if (name == "alphas") { src = alphas_; dims[0] = B_; dims[1] = T_ + 1; nd = 2; }
    else if (name == "cif_peak") { src = peaks_; dims[0] = B_; dims[1] = T_ + 1; nd = 2; }
    else if (name == "logits") { src = logits_; dims[0] = B_; dims[1] = Lpad_; dims[2] = ldv(); nd = 3; }   // padded row pitch
    else if (name == "x") { src = x32_; dims[0] = B_; dims[1] = T_; dims[2] = d; nd = 3; }
    else if (name == "dec_x") { src = xd32_; dims[0] = B_; dims[1] = Lpad_; dims[2] = d; nd = 3; }
    else if (name == "acoustic_embeds") {
        // test hook: the decoder overwrites its input in place, so the CIF frames are gathered again from the tensors the
        // run left behind (encoder output, integrate-and-fire weights) into a scratch buffer
        if (!wcur_ || !tn32_ || Lpad_ <= 0 || dst == nullptr) {
            src = wcur_ && Lpad_ > 0 ? tn32_ : nullptr;
        } else {
            PF_CUDA(cudaMemsetAsync(tn32_, 0, static_cast<size_t>(B_) * Lpad_ * d * sizeof(float), stream_));
            cif_gather_launch(enc32_, B_, T_, d, wcur_, wrem_, fire_idx_, T_ + 1, tn32_, Lpad_, stream_);
            src = tn32_;
        }
        dims[0] = B_; dims[1] = Lpad_; dims[2] = d; nd = 3;
    }
    else throw StatusError{PF_ERR_BAD_ARG, "unknown tensor '" + name + "'"};
    if (!src) throw StatusError{PF_ERR_BAD_ARG, "tensor '" + name + "' not available for this model / before a run"};
    size_t n = 1;
    for (int i = 0; i < nd; ++i) n *= static_cast<size_t>(dims[i]);
    if (dims4) for (int i = 0; i < 4; ++i) dims4[i] = dims[i];
    if (ndim) *ndim = nd;
    const size_t cnt = std::min(n, capacity);
    if (dst && cnt) {
        PF_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        PF_CUDA(cudaStreamSynchronize(stream_));
    }
}

}  // namespace pf
