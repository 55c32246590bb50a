// extern "C" surface of libpfasr (include/pf_abi.h): handle lifecycle, batch sharding over devices, error mapping,
// and the op-level test hooks.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <fstream>
#include <map>
#include <thread>
#include <unordered_map>

#include "engine.cuh"

namespace pf {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

template <typename F>
static pf_status guarded(F&& f) {
    try {
        f();
        return PF_OK;
    } catch (const StatusError& e) {
        set_last_error(e.what);
        return e.code;
    } catch (const CudaError& e) {
        set_last_error(e.what);
        cudaGetLastError();
        return PF_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return PF_ERR_OOM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return PF_ERR_BAD_ARG;
    }
}

static void validate_config(const pf_config& c) {
    if (c.struct_bytes != static_cast<int32_t>(sizeof(pf_config))) throw StatusError{PF_ERR_BAD_ARG, "pf_config.struct_bytes mismatch (ABI version skew)"};
    if (c.model_kind != PF_MODEL_PARAFORMER && c.model_kind != PF_MODEL_SENSEVOICE_SMALL && c.model_kind != PF_MODEL_SEACO_PARAFORMER)
        throw StatusError{PF_ERR_UNSUPPORTED, "unsupported model_kind"};
    if (c.n_mels != 80 || c.fs != 16000) throw StatusError{PF_ERR_UNSUPPORTED, "front-end supports fs=16000, n_mels=80 (the reference hard-codes 80, WavFrontend.cs:75)"};
    if (c.lfr_m < 1 || c.lfr_n < 1 || 2 * c.lfr_n < c.lfr_m + 1) throw StatusError{PF_ERR_UNSUPPORTED, "LFR setting would reach the reference's tail-replicate branch; unsupported"};
    if (c.input_size != c.lfr_m * c.n_mels) throw StatusError{PF_ERR_SHAPE, "input_size must equal lfr_m * n_mels"};
    if (c.input_size % 8 || c.input_size > 640) throw StatusError{PF_ERR_SHAPE, "input_size must be a multiple of 8 and <= 640"};
    if (c.d_model != 512 || c.heads != 4) throw StatusError{PF_ERR_UNSUPPORTED, "kernels are specialised for d_model=512, 4 heads (head dim 128)"};
    if (c.ffn != 2048 && c.ffn != 1024 && c.ffn != 512) throw StatusError{PF_ERR_UNSUPPORTED, "ffn width must be 512/1024/2048"};
    if (c.enc_layers < 1 || c.tp_layers < 0) throw StatusError{PF_ERR_BAD_ARG, "bad layer counts"};
    if (c.enc_kernel != 11 && c.enc_kernel != 21) throw StatusError{PF_ERR_UNSUPPORTED, "FSMN kernel must be 11 or 21"};
    if (c.model_kind == PF_MODEL_SEACO_PARAFORMER) {
        if (c.seaco_layers < 1 || (c.seaco_ffn != 512 && c.seaco_ffn != 1024 && c.seaco_ffn != 2048) || (c.seaco_kernel != 11 && c.seaco_kernel != 21))
            throw StatusError{PF_ERR_UNSUPPORTED, "SeACo bias decoder: need >= 1 layer, ffn 512/1024/2048, kernel 11/21"};
        if (c.seaco_nobias_id < 0 || c.seaco_nobias_id >= c.vocab) throw StatusError{PF_ERR_BAD_ARG, "seaco_nobias_id outside the vocabulary"};
    }
    if (c.model_kind == PF_MODEL_PARAFORMER || c.model_kind == PF_MODEL_SEACO_PARAFORMER) {
        if (c.dec_layers < 1) throw StatusError{PF_ERR_BAD_ARG, "paraformer needs decoder layers"};
        if (c.dec_ffn != 2048 && c.dec_ffn != 1024 && c.dec_ffn != 512) throw StatusError{PF_ERR_UNSUPPORTED, "decoder ffn width must be 512/1024/2048"};
        if (c.dec_kernel != 11 && c.dec_kernel != 21) throw StatusError{PF_ERR_UNSUPPORTED, "decoder FSMN kernel must be 11 or 21"};
    }
    if (c.vocab < 8) throw StatusError{PF_ERR_SHAPE, "vocab too small"};
}

template <typename Handle>
static Handle* create_handle_t(const pf_config* cfg, const void* blob, size_t bytes, const int32_t* devices, int32_t ndev,
                               const Handle* share = nullptr) {
    if (!cfg || !blob) throw StatusError{PF_ERR_BAD_ARG, "null config or weights"};
    validate_config(*cfg);
    Blob b;
    b.parse(blob, bytes);            // a bad weights file is reported as such, with or without a device
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw StatusError{PF_ERR_CUDA, std::string("no CUDA device available (libpfasr has no CPU fallback): ") + cudaGetErrorString(e)};
    }
    std::vector<int> devs;
    if (devices && ndev > 0) devs.assign(devices, devices + ndev);
    else {
        int cur = 0;
        PF_CUDA(cudaGetDevice(&cur));
        devs.push_back(cur);
    }
    for (int d : devs) if (d < 0 || d >= count) throw StatusError{PF_ERR_BAD_ARG, "device ordinal out of range"};
    std::unique_ptr<Handle> h(new Handle());
    h->cfg = *cfg;
    for (size_t i = 0; i < devs.size(); ++i) {
        std::unique_ptr<DeviceCtx> ctx(new DeviceCtx(devs[i], *cfg));
        ctx->load_weights(b, share ? share->devs.at(i).get() : nullptr);     // lanes of one handle read the same device weights
        h->devs.push_back(std::move(ctx));
    }
    return h.release();
}

// pf_offline = a pool of independent lanes (each a full OfflineHandle: own streams, staging, activations; the device
// weights are shared).  Calls from different host threads run concurrently on the GPU: the kernels of two batches
// interleave at CTA granularity and one batch's kernel tails, launch gaps and PCIe copies are filled with the other's
// work.  The reference serialises concurrent GetResults calls on the single ORT session's intra-op pool.
//
// Isolation between threads (what pf_abi.h promises):
//   * thread -> lane binding is a pure function of a process-wide thread ordinal (assigned at the thread's first call
//     into the library) modulo the lane count: it never changes, so a batch staged by a thread is run by the same lane,
//     and no table has to be pruned.  More threads than lanes simply share lanes; every call holds the lane's mutex.
//   * pf_result never points into lane-owned buffers: before the lane mutex is released the outputs are copied into
//     storage owned by the CALLING THREAD (per handle), valid until that thread's next run on the handle.
//   * a sequence of calls that must not be interleaved with another thread sharing the lane (per-call hot words:
//     set_hotwords_local -> run -> restore) is wrapped in pf_offline_lane_acquire / pf_offline_lane_release.
struct OfflinePool {
    ~OfflinePool() { while (!lanes.empty()) lanes.pop_back(); }     // lane 0 owns the device weights: it goes last
    std::vector<std::unique_ptr<OfflineHandle>> lanes;
    pf_config cfg;
    uint64_t serial = 0;               // distinguishes a new pool that reuses a destroyed pool's address
};

static std::atomic<int> g_thread_ordinals{0};
static std::atomic<uint64_t> g_pool_serials{1};
static int thread_ordinal() {
    thread_local int ord = g_thread_ordinals.fetch_add(1);
    return ord;
}

// outputs of the calling thread's last run on a handle
struct ThreadResults {
    uint64_t serial = 0;
    std::vector<int32_t> tokens, token_num;
    std::vector<float> logits, peaks, us;
};
static ThreadResults& thread_results(const OfflinePool* pool) {
    thread_local std::unordered_map<const OfflinePool*, ThreadResults> store;
    ThreadResults& r = store[pool];
    if (r.serial != pool->serial) { r = ThreadResults{}; r.serial = pool->serial; }
    return r;
}

// pf_result as filled by run_all points into lane-owned (pinned) buffers: re-home it into the calling thread's storage
static void own_results(const OfflinePool* pool, pf_result* out) {
    ThreadResults& r = thread_results(pool);
    const size_t B = static_cast<size_t>(std::max(out->batch, 0)), L = static_cast<size_t>(std::max(out->max_len, 0));
    auto take = [](auto& vec, auto*& ptr, size_t n) {
        if (!ptr) { return; }
        vec.assign(ptr, ptr + n);
        ptr = vec.data();
    };
    take(r.tokens, out->tokens, B * L);
    take(r.token_num, out->token_num, B);
    take(r.logits, out->logits, B * L * static_cast<size_t>(out->vocab));
    take(r.peaks, out->cif_peak, B * static_cast<size_t>(out->feat_frames + 1));
    if (out->us_alphas && out->us_cif_peak && out->us_frames > 0) {
        const size_t n = B * static_cast<size_t>(out->us_frames);
        r.us.resize(2 * n);
        memcpy(r.us.data(), out->us_alphas, n * sizeof(float));
        memcpy(r.us.data() + n, out->us_cif_peak, n * sizeof(float));
        out->us_alphas = r.us.data();
        out->us_cif_peak = r.us.data() + n;
    }
}

static int default_lanes() {
    const char* e = getenv("PFASR_LANES");
    const int n = e ? atoi(e) : 1;
    return std::min(std::max(n, 1), 8);
}

static OfflinePool* create_pool(const pf_config* cfg, const void* blob, size_t bytes, const int32_t* devices, int32_t ndev, int lanes) {
    if (lanes < 1 || lanes > 8) throw StatusError{PF_ERR_BAD_ARG, "lanes must be in [1, 8]"};
    std::unique_ptr<OfflinePool> pool(new OfflinePool());
    for (int l = 0; l < lanes; ++l)
        pool->lanes.emplace_back(create_handle_t<OfflineHandle>(cfg, blob, bytes, devices, ndev, l ? pool->lanes[0].get() : nullptr));
    pool->cfg = *cfg;
    pool->serial = g_pool_serials.fetch_add(1);
    if (lanes > 1)
        for (auto& lane : pool->lanes)
            for (auto& d : lane->devs) d->throughput_mode = true;
    return pool.release();
}

// the calling thread's lane
static OfflineHandle* lane_of(pf_offline* hh) {
    OfflinePool* pool = reinterpret_cast<OfflinePool*>(hh);
    if (pool->lanes.size() == 1) return pool->lanes[0].get();
    return pool->lanes[static_cast<size_t>(thread_ordinal()) % pool->lanes.size()].get();
}

static std::vector<char> read_file(const char* path) {
    if (!path || !*path) throw StatusError{PF_ERR_WEIGHTS, "weights path is empty"};
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw StatusError{PF_ERR_WEIGHTS, std::string("cannot open weights file ") + path};
    const std::streamsize sz = f.tellg();
    f.seekg(0);
    std::vector<char> buf(static_cast<size_t>(sz));
    if (!f.read(buf.data(), sz)) throw StatusError{PF_ERR_WEIGHTS, "short read on weights file"};
    return buf;
}

// ------------------------------------------------------------------ streaming (online) handle
static constexpr int kChunkSamples = 160 * 60;        // OnlineStream.cs:61,101 (160 * ChunkLength)

static OnlineStreamHost& stream_of(OnlineHandle* h, int32_t id) {
    if (id < 0 || id >= static_cast<int>(h->streams.size()) || !h->streams[id].open)
        throw StatusError{PF_ERR_BAD_ARG, "unknown or closed stream id " + std::to_string(id)};
    return h->streams[id];
}

static void online_step_all(OnlineHandle* h, const int32_t* ids, int32_t n, uint32_t flags, pf_online_result* out) {
    const int nd = static_cast<int>(h->devs.size());
    std::vector<std::vector<int>> slots(nd), pos(nd);            // per device: slots in caller order, and their index in ids
    for (int i = 0; i < n; ++i) {
        OnlineStreamHost& st = stream_of(h, ids[i]);
        for (int j = 0; j < i; ++j) if (ids[j] == ids[i]) throw StatusError{PF_ERR_BAD_ARG, "duplicate stream id in one step"};
        slots[st.dev].push_back(st.slot);
        pos[st.dev].push_back(i);
    }
    std::vector<int> active;
    for (int d = 0; d < nd; ++d) if (!slots[d].empty()) active.push_back(d);
    const int na = static_cast<int>(active.size());
    if (na == 1) {
        h->devs[active[0]]->online_step(slots[active[0]], flags, nullptr, 0);
    } else if (na > 1) {
        SharedRun shared(na);
        std::vector<std::thread> threads;
        std::vector<std::pair<pf_status, std::string>> errs(na, {PF_OK, ""});
        for (int k = 0; k < na; ++k) {
            threads.emplace_back([&, k] {
                try {
                    h->devs[active[k]]->online_step(slots[active[k]], flags, &shared, k);
                } catch (const StatusError& e) {
                    errs[k] = {e.code, e.what};
                } catch (const CudaError& e) {
                    errs[k] = {PF_ERR_CUDA, e.what};
                } catch (const std::exception& e) {
                    errs[k] = {PF_ERR_BAD_ARG, e.what()};
                }
            });
        }
        for (auto& t : threads) t.join();
        for (auto& er : errs) if (er.first != PF_OK) throw StatusError{er.first, er.second};
        if (shared.failed) throw StatusError{PF_ERR_CUDA, "a device shard failed"};
    }
    int lmax = 0, nwork = 0;
    for (int d : active) { lmax = std::max(lmax, h->devs[d]->Lmax_); nwork += static_cast<int>(h->devs[d]->online_working.size()); }
    const int V = h->cfg.vocab;
    const bool wl = (flags & PF_RUN_WANT_LOGITS) != 0 && lmax > 0;
    h->appended.assign(n, 0);
    h->embeds_len.assign(n, 0);
    h->new_tokens.assign(static_cast<size_t>(n) * lmax, 0);
    if (wl) h->logits.assign(static_cast<size_t>(n) * lmax * V, 0.0f);
    for (int d : active) {
        DeviceCtx* ctx = h->devs[d].get();
        for (size_t b = 0; b < ctx->online_working.size(); ++b) {
            const int i = pos[d][ctx->online_working[b]];
            h->embeds_len[i] = ctx->h_online_counts[b];
            h->appended[i] = lmax;
            if (lmax > 0) memcpy(h->new_tokens.data() + static_cast<size_t>(i) * lmax, ctx->h_tokens + b * lmax, static_cast<size_t>(lmax) * sizeof(int32_t));
            if (wl) memcpy(h->logits.data() + static_cast<size_t>(i) * lmax * V, ctx->h_logits + b * static_cast<size_t>(lmax) * V, static_cast<size_t>(lmax) * V * sizeof(float));
        }
    }
    out->n_streams = n;
    out->max_new = lmax;
    out->vocab = V;
    out->n_working = nwork;
    out->appended = h->appended.data();
    out->new_tokens = h->new_tokens.data();
    out->embeds_len = h->embeds_len.data();
    out->logits = wl ? h->logits.data() : nullptr;
}

// contiguous split of B items over n shards
static void split_batch(int B, int n, std::vector<int>& begin, std::vector<int>& count) {
    begin.assign(n, 0);
    count.assign(n, 0);
    const int base = B / n, rem = B % n;
    int pos = 0;
    for (int i = 0; i < n; ++i) {
        begin[i] = pos;
        count[i] = base + (i < rem ? 1 : 0);
        pos += count[i];
    }
}

static void run_all(OfflineHandle* h, uint32_t flags, pf_result* out) {
    if (!h->staged) throw StatusError{PF_ERR_BAD_ARG, "no batch staged"};
    const int n = static_cast<int>(h->devs.size());
    std::vector<int> active;
    for (int i = 0; i < n; ++i) if (h->shard_count[i] > 0) active.push_back(i);
    const int na = static_cast<int>(active.size());
    if (na == 1) {
        h->devs[active[0]]->run(flags, nullptr, 0);
    } else {
        SharedRun shared(na);
        std::vector<std::thread> threads;
        std::vector<std::pair<pf_status, std::string>> errs(na, {PF_OK, ""});
        for (int k = 0; k < na; ++k) {
            threads.emplace_back([&, k] {
                try {
                    h->devs[active[k]]->run(flags, &shared, k);
                } catch (const StatusError& e) {
                    errs[k] = {e.code, e.what};
                } catch (const CudaError& e) {
                    errs[k] = {PF_ERR_CUDA, e.what};
                } catch (const std::exception& e) {
                    errs[k] = {PF_ERR_BAD_ARG, e.what()};
                }
            });
        }
        for (auto& t : threads) t.join();
        for (auto& er : errs) if (er.first != PF_OK) throw StatusError{er.first, er.second};
        if (shared.failed) throw StatusError{PF_ERR_CUDA, "a device shard failed"};
    }
    // assemble [B, Lmax] across shards
    int lmax = 0, T = 0;
    for (int i : active) { lmax = std::max(lmax, h->devs[i]->Lmax_); T = std::max(T, h->devs[i]->T_); }
    const int B = h->B, V = h->cfg.vocab;
    h->Lmax = lmax;
    h->T = T;
    const bool wl = (flags & PF_RUN_WANT_LOGITS) != 0 && lmax > 0;
    const bool wp = (flags & PF_RUN_WANT_CIF_PEAK) != 0 && lmax > 0 && h->cfg.model_kind != PF_MODEL_SENSEVOICE_SMALL;
    const bool wt = (flags & PF_RUN_WANT_TIMESTAMPS) != 0 && lmax > 0 && h->devs[active[0]]->us_frames > 0;
    const int usf = wt ? h->devs[active[0]]->us_frames : 0;
    out->us_frames = usf;
    out->batch = B;
    out->max_len = lmax;
    out->vocab = V;
    out->feat_frames = T;
    if (na == 1) {
        DeviceCtx* d = h->devs[active[0]].get();
        out->tokens = d->h_tokens;
        out->token_num = d->h_token_num;
        out->logits = wl ? d->h_logits : nullptr;
        out->cif_peak = wp ? d->h_peaks : nullptr;
        out->us_alphas = wt ? d->h_us : nullptr;
        out->us_cif_peak = wt ? d->h_us + static_cast<size_t>(B) * usf : nullptr;
        return;
    }
    h->tokens.assign(static_cast<size_t>(B) * lmax, 0);
    h->token_num.assign(B, 0);
    if (wl) h->logits.resize(static_cast<size_t>(B) * lmax * V);
    if (wp) h->peaks.resize(static_cast<size_t>(B) * (T + 1));
    if (wt) h->us.resize(static_cast<size_t>(2) * B * usf);
    for (int i : active) {
        DeviceCtx* d = h->devs[i].get();
        const int b0 = h->shard_begin[i], nb = h->shard_count[i];
        if (lmax > 0) memcpy(h->tokens.data() + static_cast<size_t>(b0) * lmax, d->h_tokens, static_cast<size_t>(nb) * lmax * sizeof(int32_t));
        memcpy(h->token_num.data() + b0, d->h_token_num, static_cast<size_t>(nb) * sizeof(int32_t));
        if (wl) memcpy(h->logits.data() + static_cast<size_t>(b0) * lmax * V, d->h_logits, static_cast<size_t>(nb) * lmax * V * sizeof(float));
        if (wp) memcpy(h->peaks.data() + static_cast<size_t>(b0) * (T + 1), d->h_peaks, static_cast<size_t>(nb) * (T + 1) * sizeof(float));
        if (wt) {
            memcpy(h->us.data() + static_cast<size_t>(b0) * usf, d->h_us, static_cast<size_t>(nb) * usf * sizeof(float));
            memcpy(h->us.data() + static_cast<size_t>(B + b0) * usf, d->h_us + static_cast<size_t>(nb) * usf, static_cast<size_t>(nb) * usf * sizeof(float));
        }
    }
    out->us_alphas = wt ? h->us.data() : nullptr;
    out->us_cif_peak = wt ? h->us.data() + static_cast<size_t>(B) * usf : nullptr;
    out->tokens = h->tokens.data();
    out->token_num = h->token_num.data();
    out->logits = wl ? h->logits.data() : nullptr;
    out->cif_peak = wp ? h->peaks.data() : nullptr;
}

static void stage_pcm_all(OfflineHandle* h, const float* const* pcm, const int32_t* nsamp, int32_t B) {
    if (!pcm || !nsamp || B <= 0) throw StatusError{PF_ERR_BAD_ARG, "pcm/nsamp null or empty batch"};
    int tmax = 0;
    for (int b = 0; b < B; ++b) {
        if (nsamp[b] < 0 || (nsamp[b] > 0 && !pcm[b])) throw StatusError{PF_ERR_BAD_ARG, "null samples (ArgumentNullException 'source' in the reference, WavFrontend.cs:34)"};
        tmax = std::max(tmax, frontend_num_frames(nsamp[b], h->cfg.snip_edges != 0) / h->cfg.lfr_n);
    }
    const int n = static_cast<int>(h->devs.size());
    split_batch(B, n, h->shard_begin, h->shard_count);
    for (int i = 0; i < n; ++i)
        if (h->shard_count[i] > 0) h->devs[i]->stage_pcm(pcm + h->shard_begin[i], nsamp + h->shard_begin[i], h->shard_count[i], tmax);
    h->B = B;
    h->staged = true;
}

static void stage_audio_all(OfflineHandle* h, const pf_audio* utts, int32_t B) {
    if (!utts || B <= 0) throw StatusError{PF_ERR_BAD_ARG, "utterance table null or empty batch"};
    std::vector<int32_t> nsamp(B);
    int tmax = 0;
    for (int b = 0; b < B; ++b) {
        nsamp[b] = static_cast<int32_t>(audio_num_samples(utts[b]));
        tmax = std::max(tmax, frontend_num_frames(nsamp[b], h->cfg.snip_edges != 0) / h->cfg.lfr_n);
    }
    const int n = static_cast<int>(h->devs.size());
    split_batch(B, n, h->shard_begin, h->shard_count);
    for (int i = 0; i < n; ++i)
        if (h->shard_count[i] > 0)
            h->devs[i]->stage_pcm(nullptr, nsamp.data() + h->shard_begin[i], h->shard_count[i], tmax, utts + h->shard_begin[i]);
    h->B = B;
    h->staged = true;
}

}  // namespace pf

using namespace pf;

extern "C" {

const char* pf_last_error(void) { return g_last_error.c_str(); }
int32_t pf_abi_version(void) { return PF_ABI_VERSION; }

pf_status pf_offline_create_from_memory(const pf_config* cfg, const void* blob, size_t blob_bytes, const int32_t* devices,
                                        int32_t ndev, pf_offline** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        *out = reinterpret_cast<pf_offline*>(create_pool(cfg, blob, blob_bytes, devices, ndev, default_lanes()));
    });
}

pf_status pf_offline_create(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev, pf_offline** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        const std::vector<char> buf = read_file(weights_path);
        *out = reinterpret_cast<pf_offline*>(create_pool(cfg, buf.data(), buf.size(), devices, ndev, default_lanes()));
    });
}

pf_status pf_offline_create_from_memory_mt(const pf_config* cfg, const void* blob, size_t blob_bytes, const int32_t* devices,
                                           int32_t ndev, int32_t lanes, pf_offline** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        *out = reinterpret_cast<pf_offline*>(create_pool(cfg, blob, blob_bytes, devices, ndev, lanes));
    });
}

pf_status pf_offline_create_mt(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev, int32_t lanes,
                               pf_offline** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        const std::vector<char> buf = read_file(weights_path);
        *out = reinterpret_cast<pf_offline*>(create_pool(cfg, buf.data(), buf.size(), devices, ndev, lanes));
    });
}

int32_t pf_offline_lanes(const pf_offline* hh) {
    return hh ? static_cast<int32_t>(reinterpret_cast<const OfflinePool*>(hh)->lanes.size()) : -1;
}

pf_status pf_offline_destroy(pf_offline* hh) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null (ObjectDisposedException in the reference)"};
        delete reinterpret_cast<OfflinePool*>(hh);
    });
}

int32_t pf_offline_lane_acquire(pf_offline* hh) {
    if (!hh) return -1;
    OfflinePool* pool = reinterpret_cast<OfflinePool*>(hh);
    OfflineHandle* h = lane_of(hh);
    h->mu.lock();
    for (size_t i = 0; i < pool->lanes.size(); ++i) if (pool->lanes[i].get() == h) return static_cast<int32_t>(i);
    return 0;
}

pf_status pf_offline_lane_release(pf_offline* hh) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        lane_of(hh)->mu.unlock();
    });
}

pf_status pf_offline_set_cmvn(pf_offline* hh, const float* add_shift, const float* rescale, int32_t dim) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!add_shift || !rescale) throw StatusError{PF_ERR_BAD_ARG, "null cmvn vectors"};
        for (auto& lane : reinterpret_cast<OfflinePool*>(hh)->lanes) {
            OfflineHandle* h = lane.get();
            std::lock_guard<std::recursive_mutex> g(h->mu);
            for (auto& d : h->devs) d->set_cmvn(add_shift, rescale, dim);
        }
    });
}

pf_status pf_offline_set_hotwords(pf_offline* hh, const int32_t* ids, int32_t n) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (n < 0 || (n > 0 && !ids)) throw StatusError{PF_ERR_BAD_ARG, "ids is null"};
        for (auto& lane : reinterpret_cast<OfflinePool*>(hh)->lanes) {
            OfflineHandle* h = lane.get();
            std::lock_guard<std::recursive_mutex> g(h->mu);
            for (auto& d : h->devs) d->set_hotwords(ids, n);
        }
    });
}

pf_status pf_offline_set_hotwords_local(pf_offline* hh, const int32_t* ids, int32_t n) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (n < 0 || (n > 0 && !ids)) throw StatusError{PF_ERR_BAD_ARG, "ids is null"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        for (auto& d : h->devs) d->set_hotwords(ids, n);
    });
}

static pf_status extract_common(pf_offline* hh, const float* samples, int32_t nsamp, float* out, int32_t cap, int32_t* out_frames, bool raw) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!samples) throw StatusError{PF_ERR_BAD_ARG, "samples is null (ArgumentNullException 'source' in the reference, WavFrontend.cs:34)"};
        if (nsamp < 0 || !out_frames || (!out && cap > 0)) throw StatusError{PF_ERR_BAD_ARG, "bad arguments"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        *out_frames = h->devs[0]->extract(samples, nsamp, out, cap, raw);
    });
}

pf_status pf_frontend_extract(pf_offline* hh, const float* samples, int32_t nsamp, float* feats, int32_t capacity_frames, int32_t* out_frames) {
    return extract_common(hh, samples, nsamp, feats, capacity_frames, out_frames, false);
}
pf_status pf_frontend_fbank(pf_offline* hh, const float* samples, int32_t nsamp, float* fbank, int32_t capacity_frames, int32_t* out_frames) {
    return extract_common(hh, samples, nsamp, fbank, capacity_frames, out_frames, true);
}
int32_t pf_frontend_num_frames(const pf_offline* hh, int32_t nsamp) {
    if (!hh || nsamp < 0) return -1;
    const OfflinePool* h = reinterpret_cast<const OfflinePool*>(hh);
    return frontend_num_frames(nsamp, h->cfg.snip_edges != 0) / h->cfg.lfr_n;
}

pf_status pf_offline_stage_pcm(pf_offline* hh, const float* const* pcm, const int32_t* nsamp, int32_t batch) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        stage_pcm_all(h, pcm, nsamp, batch);
        // standalone staging returns only once the PCM is resident in HBM (the caller may free its buffers)
        for (auto& d : h->devs) d->sync_staging();
    });
}

pf_status pf_offline_run_staged(pf_offline* hh, uint32_t flags, pf_result* out) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        memset(out, 0, sizeof(*out));
        run_all(h, flags, out);
        own_results(reinterpret_cast<OfflinePool*>(hh), out);
    });
}

pf_status pf_offline_run_pcm(pf_offline* hh, const float* const* pcm, const int32_t* nsamp, int32_t batch, uint32_t flags, pf_result* out) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        memset(out, 0, sizeof(*out));
        stage_pcm_all(h, pcm, nsamp, batch);
        run_all(h, flags, out);
        own_results(reinterpret_cast<OfflinePool*>(hh), out);
    });
}

pf_status pf_offline_run_audio(pf_offline* hh, const pf_audio* utts, int32_t batch, uint32_t flags, pf_result* out) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        memset(out, 0, sizeof(*out));
        stage_audio_all(h, utts, batch);
        run_all(h, flags, out);
        own_results(reinterpret_cast<OfflinePool*>(hh), out);
    });
}

pf_status pf_offline_run_feats(pf_offline* hh, const float* speech, int32_t batch, int32_t frames, uint32_t flags, pf_result* out) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!out || !speech) throw StatusError{PF_ERR_BAD_ARG, "null argument"};
        if (batch <= 0 || frames <= 0) throw StatusError{PF_ERR_SHAPE, "speech must be [B>0, T>0, input_size]"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        memset(out, 0, sizeof(*out));
        const int n = static_cast<int>(h->devs.size());
        split_batch(batch, n, h->shard_begin, h->shard_count);
        const size_t row = static_cast<size_t>(frames) * h->cfg.input_size;
        for (int i = 0; i < n; ++i)
            if (h->shard_count[i] > 0) h->devs[i]->stage_feats(speech + h->shard_begin[i] * row, h->shard_count[i], frames);
        h->B = batch;
        h->staged = true;
        run_all(h, flags, out);
        own_results(reinterpret_cast<OfflinePool*>(hh), out);
    });
}

pf_status pf_offline_get_tensor(pf_offline* hh, int32_t dev_index, const char* name, float* dst, size_t capacity, int32_t* dims4, int32_t* ndim) {
    return guarded([&] {
        if (!hh || !name) throw StatusError{PF_ERR_BAD_ARG, "null argument"};
        OfflineHandle* h = lane_of(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        if (dev_index < 0 || dev_index >= static_cast<int>(h->devs.size())) throw StatusError{PF_ERR_BAD_ARG, "dev_index out of range"};
        h->devs[dev_index]->get_tensor(name, dst, capacity, dims4, ndim);
    });
}

int32_t pf_offline_get_timings(pf_offline* hh, float* ms, int32_t capacity) {
    if (!hh || !ms) return 0;
    OfflineHandle* h = lane_of(hh);
    const int n = std::min(capacity, 10);
    for (int i = 0; i < n; ++i) ms[i] = h->devs[0]->timings_ms[i];
    return n;
}

int64_t pf_offline_get_launch_count(pf_offline* hh) {
    if (!hh) return 0;
    OfflineHandle* h = lane_of(hh);
    int64_t n = 0;
    for (auto& d : h->devs) n += d->launches;
    return n;
}

double pf_offline_get_gemm_flops(pf_offline* hh) {
    if (!hh) return 0;
    OfflineHandle* h = lane_of(hh);
    double n = 0;
    for (auto& d : h->devs) n += d->gemm_flops;
    return n;
}

pf_status pf_offline_set_profile(pf_offline* hh, int32_t on) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        for (auto& lane : reinterpret_cast<OfflinePool*>(hh)->lanes) {
            OfflineHandle* h = lane.get();
            std::lock_guard<std::recursive_mutex> g(h->mu);
            for (auto& d : h->devs) d->set_profile(on);
        }
    });
}

double pf_offline_get_gemm_ms(pf_offline* hh) {
    if (!hh) return 0;
    return lane_of(hh)->devs[0]->gemm_ms;
}

double pf_offline_replay_gemms(pf_offline* hh, int32_t iters) {
    if (!hh) return 0;
    OfflineHandle* h = lane_of(hh);
    std::lock_guard<std::recursive_mutex> g(h->mu);
    try {
        return h->devs[0]->replay_gemms(iters);
    } catch (...) {
        return 0;
    }
}

int32_t pf_offline_get_profile_json(pf_offline* hh, char* buf, int32_t capacity) {
    if (!hh || !buf || capacity <= 0) return 0;
    const std::string& js = lane_of(hh)->devs[0]->profile_json;
    const int n = std::min<int>(capacity - 1, static_cast<int>(js.size()));
    memcpy(buf, js.data(), n);
    buf[n] = 0;
    return n;
}

void* pf_offline_get_stream(pf_offline* hh, int32_t dev_index) {
    if (!hh) return nullptr;
    OfflineHandle* h = lane_of(hh);
    if (dev_index < 0 || dev_index >= static_cast<int>(h->devs.size())) return nullptr;
    return h->devs[dev_index]->stream();
}

// ------------------------------------------------------------------ streaming (online) exports
static OnlineHandle* create_online(const pf_config* cfg, const void* blob, size_t bytes, const int32_t* devices, int32_t ndev) {
    if (cfg && cfg->model_kind != PF_MODEL_PARAFORMER) throw StatusError{PF_ERR_UNSUPPORTED, "streaming is a paraformer path"};
    return create_handle_t<OnlineHandle>(cfg, blob, bytes, devices, ndev);
}

pf_status pf_online_create_from_memory(const pf_config* cfg, const void* blob, size_t blob_bytes, const int32_t* devices,
                                       int32_t ndev, pf_online** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        *out = reinterpret_cast<pf_online*>(create_online(cfg, blob, blob_bytes, devices, ndev));
    });
}

pf_status pf_online_create(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev, pf_online** out) {
    return guarded([&] {
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        *out = nullptr;
        const std::vector<char> buf = read_file(weights_path);
        *out = reinterpret_cast<pf_online*>(create_online(cfg, buf.data(), buf.size(), devices, ndev));
    });
}

pf_status pf_online_destroy(pf_online* hh) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        delete reinterpret_cast<OnlineHandle*>(hh);
    });
}

pf_status pf_online_set_cmvn(pf_online* hh, const float* add_shift, const float* rescale, int32_t dim) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!add_shift || !rescale) throw StatusError{PF_ERR_BAD_ARG, "null cmvn vectors"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        for (auto& d : h->devs) d->set_cmvn(add_shift, rescale, dim);
    });
}

pf_status pf_online_stream_open(pf_online* hh, int32_t* stream_id) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!stream_id) throw StatusError{PF_ERR_BAD_ARG, "stream_id is null"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        int id = -1;
        for (size_t i = 0; i < h->streams.size(); ++i) if (!h->streams[i].open) { id = static_cast<int>(i); break; }
        if (id < 0) { id = static_cast<int>(h->streams.size()); h->streams.emplace_back(); }
        OnlineStreamHost& st = h->streams[id];
        st.dev = id % static_cast<int>(h->devs.size());               // sticky placement, state never migrates
        st.slot = h->devs[st.dev]->online_open();
        st.cache_samples.assign(kChunkSamples, 0.0f);                  // _cacheSamples = new float[160 * ChunkLength] (Q13)
        st.open = true;
        *stream_id = id;
    });
}

pf_status pf_online_stream_close(pf_online* hh, int32_t stream_id) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        OnlineStreamHost& st = stream_of(h, stream_id);
        h->devs[st.dev]->online_close(st.slot);
        st.open = false;
        st.cache_samples.clear();
        st.cache_samples.shrink_to_fit();
    });
}

pf_status pf_online_stream_push(pf_online* hh, int32_t stream_id, const float* samples, int32_t nsamp) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!samples && nsamp != 0) throw StatusError{PF_ERR_BAD_ARG, "samples is null (NullReferenceException in OnlineStream.AddSamples)"};
        if (nsamp < 0) throw StatusError{PF_ERR_BAD_ARG, "negative sample count"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        OnlineStreamHost& st = stream_of(h, stream_id);
        st.cache_samples.insert(st.cache_samples.end(), samples, samples + nsamp);
        if (static_cast<int>(st.cache_samples.size()) > kChunkSamples) {          // strictly more than one chunk (OnlineStream.cs:102)
            h->devs[st.dev]->online_push_chunk(st.slot, st.cache_samples.data(), kChunkSamples);
            st.cache_samples.erase(st.cache_samples.begin(), st.cache_samples.begin() + kChunkSamples);
        }
    });
}

int32_t pf_online_stream_ready(pf_online* hh, int32_t stream_id) {
    if (!hh) return -1;
    OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
    std::lock_guard<std::recursive_mutex> g(h->mu);
    if (stream_id < 0 || stream_id >= static_cast<int>(h->streams.size()) || !h->streams[stream_id].open) return -1;
    const OnlineStreamHost& st = h->streams[stream_id];
    return h->devs[st.dev]->online_ready(st.slot) ? 1 : 0;
}

pf_status pf_online_step(pf_online* hh, const int32_t* stream_ids, int32_t n, uint32_t flags, pf_online_result* out) {
    return guarded([&] {
        if (!hh) throw StatusError{PF_ERR_DISPOSED, "handle is null"};
        if (!out) throw StatusError{PF_ERR_BAD_ARG, "out is null"};
        memset(out, 0, sizeof(*out));
        if (n == 0) return;                                            // streams.Count == 0 (OnlineRecognizer.cs:343-346)
        if (!stream_ids || n < 0) throw StatusError{PF_ERR_BAD_ARG, "stream_ids is null"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        online_step_all(h, stream_ids, n, flags, out);
    });
}

pf_status pf_online_get_state(pf_online* hh, int32_t stream_id, const char* name, float* dst, size_t capacity) {
    return guarded([&] {
        if (!hh || !name) throw StatusError{PF_ERR_BAD_ARG, "null argument"};
        OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
        std::lock_guard<std::recursive_mutex> g(h->mu);
        OnlineStreamHost& st = stream_of(h, stream_id);
        h->devs[st.dev]->online_get_state(st.slot, name, dst, capacity);
    });
}

int32_t pf_online_get_timings(pf_online* hh, float* ms, int32_t capacity) {
    if (!hh || !ms) return 0;
    OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
    const int n = std::min(capacity, 6);
    for (int i = 0; i < n; ++i) ms[i] = h->devs[0]->timings_ms[i];
    return n;
}

int64_t pf_online_get_launch_count(pf_online* hh) {
    if (!hh) return 0;
    OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
    int64_t n = 0;
    for (auto& d : h->devs) n += d->launches;
    return n;
}

double pf_online_get_gemm_flops(pf_online* hh) {
    if (!hh) return 0;
    OnlineHandle* h = reinterpret_cast<OnlineHandle*>(hh);
    double n = 0;
    for (auto& d : h->devs) n += d->gemm_flops;
    return n;
}

}  // extern "C"
