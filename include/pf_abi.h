/*
 * libpfasr C-ABI: the drop-in boundary for the AliParaformerAsr offline hot path on B200 (sm_100a).
 *
 * Every entry point is what the reference's managed code would bind with [DllImport("pfasr")] in place of the two
 * native crossings it has today:
 *   - Microsoft.ML.OnnxRuntime  InferenceSession.Run      OfflineProjOfParaformer.cs:68,
 *                                                          OfflineProjOfSenseVoiceSmall.cs:156
 *   - ManySpeech.SpeechFeatures OnlineFbank.GetFbank       WavFrontend.cs:21-37
 * Plain pointers and sizes only; no C++/torch types.  All tensors are row-major contiguous.
 *
 * Error model (replaces the reference's exceptions, OfflineProjOfParaformer.cs:82-85): every call returns a
 * pf_status; the message is available from pf_last_error() on the calling thread.  There is NO CPU fallback: a
 * missing GPU / non-sm_100 device is PF_ERR_CUDA.
 *
 * Ownership: inputs stay owned by the caller and may be pageable.  Outputs referenced from pf_result are owned by
 * the handle (pinned host memory) and stay valid until the next run on that handle or pf_offline_destroy.
 * Threading: calls on one handle are serialised internally; distinct handles are independent.
 */
#ifndef PF_ABI_H_
#define PF_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 5

typedef int32_t pf_status;
enum {
    PF_OK = 0,
    PF_ERR_BAD_ARG = -1,
    PF_ERR_CUDA = -2,
    PF_ERR_OOM = -3,
    PF_ERR_SHAPE = -4,
    PF_ERR_DISPOSED = -5,
    PF_ERR_WEIGHTS = -6,
    PF_ERR_UNSUPPORTED = -7
};

/* asr.yaml `model:` dispatch of OfflineRecognizer.cs:39-53 */
enum { PF_MODEL_PARAFORMER = 0, PF_MODEL_SENSEVOICE_SMALL = 1, PF_MODEL_SEACO_PARAFORMER = 2 };

/* Flat POD mirror of the ConfEntity fields the path consumes (Model/ConfEntity.cs:5-43, EncoderConfEntity.cs:13-25,
 * DecoderConfEntity.cs:7-16, PredictorConfEntity.cs:13-17, FrontendConfEntity.cs:7-15).  The C# side keeps parsing
 * asr.yaml / asr.json (Utils/PreloadHelper.cs:39-118) and fills this struct. */
typedef struct pf_config {
    int32_t struct_bytes;      /* = sizeof(pf_config) */
    int32_t model_kind;        /* PF_MODEL_* */
    int32_t input_size;        /* 560 = lfr_m * n_mels */
    int32_t d_model;           /* encoder_conf.output_size 512 */
    int32_t heads;             /* attention_heads 4 */
    int32_t ffn;               /* linear_units 2048 */
    int32_t enc_layers;        /* num_blocks 50 (encoders0 + encoders) */
    int32_t tp_layers;         /* SenseVoice tp_blocks 20, else 0 */
    int32_t enc_kernel;        /* kernel_size 11 */
    int32_t dec_layers;        /* decoder_conf.num_blocks 16 */
    int32_t dec_ffn;           /* decoder_conf.linear_units 2048 */
    int32_t dec_kernel;        /* decoder_conf.kernel_size 11 */
    int32_t vocab;             /* 8404 paraformer-large zh-en / 25055 SenseVoiceSmall */
    float ln_eps;              /* 1e-12 (ESPnet LayerNorm) / 1e-5 */
    float cif_threshold;       /* predictor_conf.threshold 1.0 */
    float cif_tail;            /* predictor_conf.tail_threshold 0.45 */
    float smooth_factor;       /* 1.0 */
    float noise_threshold;     /* 0.0 */
    int32_t fs;                /* frontend_conf.fs 16000 */
    int32_t n_mels;            /* 80 (the reference hard-codes 80, WavFrontend.cs:75) */
    int32_t lfr_m;             /* 7 */
    int32_t lfr_n;             /* 6 */
    int32_t snip_edges;        /* frontend_conf.snip_edges */
    int32_t use_itn;           /* ConfEntity.use_itn (SenseVoice prompt, quirk Q6) */
    int32_t online_flags;      /* streaming, bit 0: 1 = per-layer FSMN caches instead of the reference's stack_states
                                * behaviour that feeds every layer the stream's layer-0 cache (OnlineModel.cs:222) */
    int32_t seaco_layers;      /* seaco_decoder_conf.num_blocks 4 (SeACo only) */
    int32_t seaco_ffn;         /* seaco_decoder_conf.linear_units 1024 */
    int32_t seaco_kernel;      /* seaco_decoder_conf.kernel_size 21 */
    int32_t seaco_nobias_id;   /* class id of "no bias" in the hot-word head (8377 for the 8404-token vocabulary) */
    float smooth_factor2;      /* CifPredictorV3 timestamp branch: predictor_conf.smooth_factor2 0.25 */
    float noise_threshold2;    /* predictor_conf.noise_threshold2 0.01 */
} pf_config;

/* ModelOutputEntity (Model/ModelOutputEntity.cs:10-19) plus the greedy ids that OfflineRecognizer.Forward derives
 * from it (OfflineRecognizer.cs:139-152). */
typedef struct pf_result {
    int32_t batch;             /* B */
    int32_t max_len;           /* L = model_out.Dimensions[1] */
    int32_t vocab;             /* V */
    int32_t feat_frames;       /* T fed to the encoder (incl. SenseVoice prompt rows) */
    const int32_t* tokens;     /* [B, L] greedy ids, ties -> largest index (Q5) */
    const int32_t* token_num;  /* [B] model_out_lens */
    const float* logits;       /* [B, L, V] log-softmax, only with PF_RUN_WANT_LOGITS, else NULL */
    const float* cif_peak;     /* [B, T+1] integrate-and-fire trace, only with PF_RUN_WANT_CIF_PEAK, else NULL */
    /* graph outputs #3 / #4 of the -timestamp- and SeACo models (ModelOutputEntity.cif_peak_tensor = us_cif_peak,
     * OfflineProjOfParaformer.cs:75-79), only with PF_RUN_WANT_TIMESTAMPS on a model that has the V3 predictor */
    int32_t us_frames;         /* 3 * T (the predictor upsamples x3), 0 when absent */
    const float* us_alphas;    /* [B, us_frames] */
    const float* us_cif_peak;  /* [B, us_frames] consumed by OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302) */
} pf_result;

enum { PF_RUN_WANT_LOGITS = 1, PF_RUN_WANT_CIF_PEAK = 2, PF_RUN_WANT_TIMESTAMPS = 4 };

typedef struct pf_offline pf_offline;

/* -------- lifecycle: replaces OfflineModel.initModel (OfflineModel.cs:35-70) + the IOfflineProj constructors.
 * weights_path: PFW1 blob (see aliparaformerasr_b200/weights.py).  devices/ndev: CUDA ordinals to shard batches over
 * (NULL/0 = current device).  The batch is split contiguously across devices, weights are replicated. */
pf_status pf_offline_create(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev,
                            pf_offline** out);
pf_status pf_offline_create_from_memory(const pf_config* cfg, const void* blob, size_t blob_bytes,
                                        const int32_t* devices, int32_t ndev, pf_offline** out);
/* replaces IOfflineProj.Dispose / InferenceSession.Dispose (OfflineProjOfParaformer.cs:88-101) */
/* Concurrent callers: the same handle with `lanes` (1..8) independent execution lanes (own streams, staging buffers and
 * activations per lane; the device weights are shared).  A host thread is bound to lane (thread ordinal mod lanes), the
 * ordinal being assigned at the thread's first call into the library, so the binding never changes: calls from
 * different threads overlap on the GPU (one batch's kernel tails, launch gaps and PCIe copies are filled with another
 * batch's work) and more threads than lanes share lanes, every call holding its lane's lock.
 *   - pf_result points into storage owned by the CALLING THREAD: it stays valid until that thread's next run on the
 *     handle, whatever other threads do (also with one lane).
 *   - a batch staged with pf_offline_stage_pcm must be run (pf_offline_run_staged) by the thread that staged it; wrap
 *     the pair - and any other sequence that another thread sharing the lane must not interleave with, such as
 *     set_hotwords_local / run / restore - in pf_offline_lane_acquire ... pf_offline_lane_release.
 * The plain create functions use one lane (or $PFASR_LANES).  Replaces nothing in the reference: there
 * OfflineRecognizer.GetResults may be called from several threads too, and they queue on the one ORT session. */
pf_status pf_offline_create_mt(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev,
                               int32_t lanes, pf_offline** out);
pf_status pf_offline_create_from_memory_mt(const pf_config* cfg, const void* blob, size_t blob_bytes, const int32_t* devices,
                                           int32_t ndev, int32_t lanes, pf_offline** out);
int32_t pf_offline_lanes(const pf_offline* h);
/* lease the calling thread's lane (re-entrant lock; returns the lane index, -1 on a null handle) / give it back */
int32_t pf_offline_lane_acquire(pf_offline* h);
pf_status pf_offline_lane_release(pf_offline* h);
pf_status pf_offline_destroy(pf_offline* h);

/* SeACo hot words: replaces EmbedSeacoModel.Forward (EmbedSeacoModel.cs:70-108, model_eb.onnx = Embedding + 2-layer
 * LSTM) and the bias_embed assembly of OfflineProjOfSeacoParaformer.cs:85-108 (all 10 LSTM steps of every hot word,
 * Q8).  ids: [n, 10] int32, already truncated / zero padded as EmbedSeacoModel.PadList does (:110-123).  n = 0 clears
 * the hot words (the bias branch is skipped).  Valid for PF_MODEL_SEACO_PARAFORMER handles only. */
pf_status pf_offline_set_hotwords(pf_offline* h, const int32_t* ids, int32_t n);
/* the same for the calling thread's execution lane only: per-call hot words of the streams in one GetResults
 * (OfflineProjOfSeacoParaformer.cs:51-60) when several threads share the handle */
pf_status pf_offline_set_hotwords_local(pf_offline* h, const int32_t* hotword_ids, int32_t n_hotwords);

/* am.mvn vectors parsed by WavFrontend.LoadCmvn (WavFrontend.cs:112-153): <AddShift> and <Rescale>, dim = 560 */
pf_status pf_offline_set_cmvn(pf_offline* h, const float* add_shift, const float* rescale, int32_t dim);

/* -------- front-end only: replaces WavFrontend.GetFbank + LfrCmvn as called by OfflineStream.AddSamples
 * (OfflineStream.cs:40-41).  samples: float PCM in [-1, 1].  feats: [capacity_frames, 560] out; *out_frames = T_lfr.
 * No Q4 substitution here (the reference applies it later, in PadHelper).
 * Fixed fbank settings: Hamming window, 25 ms frames every 10 ms, 80 mel bins, 16 kHz, and NO dither (the reference
 * forwards frontend_conf.dither, C# default 1.0, to OnlineFbank, WavFrontend.cs:21-26: that makes its features random;
 * every published model sets dither 0 or is evaluated without it).  A caller whose asr.yaml asks for another window or
 * frame size must not use this library (the Python host raises; the C# shim should do the same). */
pf_status pf_frontend_extract(pf_offline* h, const float* samples, int32_t nsamp, float* feats,
                              int32_t capacity_frames, int32_t* out_frames);
/* raw Kaldi fbank [T, 80] of one utterance (x32768 scaling included), for parity checks of OnlineFbank.GetFbank */
pf_status pf_frontend_fbank(pf_offline* h, const float* samples, int32_t nsamp, float* fbank, int32_t capacity_frames,
                            int32_t* out_frames);
/* number of LFR frames AddSamples would produce for nsamp samples */
int32_t pf_frontend_num_frames(const pf_offline* h, int32_t nsamp);

/* -------- the hot path.
 * pf_offline_run_pcm: one AddSamples per stream then GetResults(streams) (OfflineRecognizer.cs:110-198): fused
 *   fbank+LFR+CMVN+PadSequence (Q1,Q2,Q4) -> encoder -> CIF -> decoder -> log-softmax -> greedy ids.
 * pf_offline_run_feats: IOfflineProj.ModelProj's InferenceSession.Run on an already padded `speech [B,T,560]`
 *   (speech_lengths = T for every item, Q3), keeping the C# front-end / PadHelper for A/B checks. */
pf_status pf_offline_run_pcm(pf_offline* h, const float* const* pcm, const int32_t* nsamp, int32_t batch,
                             uint32_t flags, pf_result* out);
pf_status pf_offline_run_feats(pf_offline* h, const float* speech, int32_t batch, int32_t frames, uint32_t flags,
                               pf_result* out);
/* Split form used to time the device path alone: stage uploads PCM to HBM, run_staged consumes it. */
pf_status pf_offline_stage_pcm(pf_offline* h, const float* const* pcm, const int32_t* nsamp, int32_t batch);
pf_status pf_offline_run_staged(pf_offline* h, uint32_t flags, pf_result* out);

/* -------- introspection */
/* intermediate of the last run on device `dev_index` (0-based within the handle): "feats", "enc", "alphas",
 * "acoustic_embeds", "logits".  Copies min(capacity, size) floats, reports dims. */
pf_status pf_offline_get_tensor(pf_offline* h, int32_t dev_index, const char* name, float* dst, size_t capacity,
                                int32_t* dims4, int32_t* ndim);
/* per-stage device times of the last run (ms, CUDA events, device 0 of the handle):
 * [0] h2d+frontend [1] encoder [2] predictor+cif [3] decoder [4] head+pick [5] total; then host clock since the run
 * began (ms): [6] first half enqueued [7] token counts on the host [8] decoder enqueued [9] results on the host.
 * Returns count written (at most 10). */
int32_t pf_offline_get_timings(pf_offline* h, float* ms, int32_t capacity);
/* kernels launched by the last run (all devices), algorithmic GEMM flops of the last run */
int64_t pf_offline_get_launch_count(pf_offline* h);
double pf_offline_get_gemm_flops(pf_offline* h);
/* per-launch CUDA-event profiling (two extra events per launch; off by default): on = 1 times the GEMM launches,
 * on = 2 also the other kernels of the encoder / decoder layers (rows carry "name").  After a profiled
 * run: pf_offline_get_gemm_ms = summed GEMM launch durations on device 0 of the handle, pf_offline_get_profile_json =
 * per-shape breakdown [{"M","N","K","tile_n","launches","ms","tflops"}, ...] (returns bytes written). */
pf_status pf_offline_set_profile(pf_offline* h, int32_t on);
double pf_offline_get_gemm_ms(pf_offline* h);
int32_t pf_offline_get_profile_json(pf_offline* h, char* buf, int32_t capacity);
/* after a run with pf_offline_set_profile(h, 1): re-launch that run's GEMMs back to back (programmatic-launch chained,
 * as inside the step) `iters` times between two CUDA events on device 0's stream; returns ms per pass (0 on error) */
double pf_offline_replay_gemms(pf_offline* h, int32_t iters);
/* CUDA stream of device dev_index (as a cudaStream_t), so callers can bracket runs with their own events */
void* pf_offline_get_stream(pf_offline* h, int32_t dev_index);

/* ================= streaming (online) path: OnlineRecognizer / OnlineStream / OnlineModel =================
 * Replaces the two InferenceSessions of OnlineModel (OnlineModel.cs:24-32: encoder.onnx, decoder.onnx), the managed
 * per-stream state of OnlineStream (OnlineStream.cs:24-37: sample cache, fbank FIFO, splice frame, feature cache, CIF
 * carry, 16 FSMN caches) and OnlineRecognizer.Forward (OnlineRecognizer.cs:341-401).  All per-stream state stays in
 * HBM between steps; only PCM goes in and token ids come out.  Weights: the same PFW1 blob layout as paraformer. */
typedef struct pf_online pf_online;

typedef struct pf_online_result {
    int32_t n_streams;          /* streams passed to the step */
    int32_t max_new;            /* L = logits.Dimensions[1] of this step; 0 when no decoder pass ran */
    int32_t vocab;
    int32_t n_working;          /* streams that had a full decode chunk (GetDecodeChunk != null) */
    const int32_t* appended;    /* [n] ids appended to stream i's Tokens: max_new for working streams (the reference
                                 * appends the padded rows too, OnlineRecognizer.cs:390), 0 otherwise */
    const int32_t* new_tokens;  /* [n, max_new] greedy ids (Q5 tie rule); rows of non-working streams are 0 */
    const int32_t* embeds_len;  /* [n] acoustic_embeds_len = CIF fires of this step */
    const float* logits;        /* [n, max_new, V] raw decoder output, only with PF_RUN_WANT_LOGITS */
} pf_online_result;

pf_status pf_online_create(const pf_config* cfg, const char* weights_path, const int32_t* devices, int32_t ndev,
                           pf_online** out);
pf_status pf_online_create_from_memory(const pf_config* cfg, const void* blob, size_t blob_bytes, const int32_t* devices,
                                       int32_t ndev, pf_online** out);
pf_status pf_online_destroy(pf_online* h);
pf_status pf_online_set_cmvn(pf_online* h, const float* add_shift, const float* rescale, int32_t dim);
/* OnlineRecognizer.CreateOnlineStream (OnlineRecognizer.cs:27-31); streams are pinned to device stream_id % ndev */
pf_status pf_online_stream_open(pf_online* h, int32_t* stream_id);
pf_status pf_online_stream_close(pf_online* h, int32_t stream_id);
/* OnlineStream.AddSamples (OnlineStream.cs:84-112): appends to the sample cache (which starts as 9600 zeros) and
 * consumes AT MOST ONE 9600-sample chunk per call, only when the cache holds more than one chunk (Q13) */
pf_status pf_online_stream_push(pf_online* h, int32_t stream_id, const float* samples, int32_t nsamp);
/* 1 when GetDecodeChunk would return a window for this stream, 0 when not, < 0 on error */
int32_t pf_online_stream_ready(pf_online* h, int32_t stream_id);
/* OnlineRecognizer.GetResults -> Forward over the listed streams (OnlineRecognizer.cs:41-48, 341-401) */
pf_status pf_online_step(pf_online* h, const int32_t* stream_ids, int32_t n, uint32_t flags, pf_online_result* out);
/* test hook: per-stream device state "cache_feats" [10,560], "cif_alpha" [1], "cif_hidden" [512], "splice" [80],
 * "fsmn" [layers, kernel-1, 512] (the reference keeps [512, kernel-1] per layer) */
pf_status pf_online_get_state(pf_online* h, int32_t stream_id, const char* name, float* dst, size_t capacity);
int32_t pf_online_get_timings(pf_online* h, float* ms, int32_t capacity);
int64_t pf_online_get_launch_count(pf_online* h);
double pf_online_get_gemm_flops(pf_online* h);

/* ================= audio ingestion =================
 * Replaces AliParaformerAsr.Examples/Utils/AudioHelper.cs:12-32 (GetFileSample: NAudio AudioFileReader -> float) and
 * :223-279 (Resample: stereo -> mono average, linear interpolation to 16 kHz).  The caller passes the file's raw sample
 * bytes; conversion, down-mix and resampling run on the device in front of the fbank kernel, bit-identical to the C#.
 * As in the reference, a 16 kHz file is passed through without down-mixing (AudioHelper.cs:27-30). */
enum { PF_AUDIO_U8 = 0, PF_AUDIO_S16 = 1, PF_AUDIO_S24 = 2, PF_AUDIO_S32 = 3, PF_AUDIO_F32 = 4 };

typedef struct pf_audio {
    const void* data;        /* interleaved little-endian samples */
    int64_t n_values;        /* frames * channels */
    int32_t format;          /* PF_AUDIO_* */
    int32_t channels;
    int32_t sample_rate;
    int32_t reserved;
} pf_audio;

/* locate the fmt / data chunks of a RIFF/WAVE image (PCM 8/16/24/32, IEEE float 32, WAVE_FORMAT_EXTENSIBLE);
 * out->data points into `file`.  PF_ERR_UNSUPPORTED for other containers / codecs. */
pf_status pf_wav_parse(const void* file, size_t bytes, pf_audio* out);
/* samples GetFileSample would return for this audio (after the optional resample), -1 on invalid input */
int64_t pf_audio_num_samples(const pf_audio* a);
/* OfflineRecognizer.GetResults on a batch of files: GetFileSample + AddSamples + Forward in one call */
pf_status pf_offline_run_audio(pf_offline* h, const pf_audio* utts, int32_t batch, uint32_t flags, pf_result* out);

/* ================= host post-processing (consumer half of f1) =================
 * Native equivalents of the string / timestamp loops that follow the graph in the reference, for callers that want the
 * whole GetResults in one library.  Pure host code; UTF-8 in and out. */
typedef struct pf_tokens pf_tokens;

typedef struct pf_text_result {
    char* text;                 /* caller buffer (may be NULL to query sizes): UTF-8, NUL-terminated on return */
    size_t text_capacity;       /* bytes available in text */
    size_t text_bytes;          /* out: bytes of the text, excluding the NUL */
    int32_t text_len;           /* out: OfflineRecognizerResultEntity.TextLen = string.Length (UTF-16 code units) */
    int32_t n_tokens;           /* out: entries of OfflineRecognizerResultEntity.Tokens */
    int32_t n_timestamps;       /* out: entries of .Timestamps (equals n_tokens unless the Remove-by-value quirk fired) */
    int32_t reserved;
    char* tokens;               /* caller buffer or NULL: the merged tokens, each NUL-terminated, back to back */
    size_t tokens_capacity;
    size_t tokens_bytes;        /* out: bytes needed for tokens (terminators included) */
    int32_t* ts;                /* caller buffer or NULL: all timestamp ints back to back (a merged token owns 2, 4, ... ) */
    size_t ts_capacity;         /* int32 elements available in ts */
    size_t ts_count;            /* out: int32 elements needed */
    int32_t* ts_offsets;        /* caller buffer, required with ts: [n_timestamps + 1], entry i spans ts[off[i] .. off[i+1]) */
    size_t ts_offsets_capacity;
} pf_text_result;

/* Utils/PreloadHelper.ReadTokens (PreloadHelper.cs:120-141): File.ReadAllLines of tokens.txt; an empty table is the
 * reference's "invalid tokens file" (OfflineRecognizer.cs:30) -> PF_ERR_BAD_ARG */
pf_status pf_tokens_create(const char* path, pf_tokens** out);
pf_status pf_tokens_create_from_memory(const char* utf8, size_t bytes, pf_tokens** out);
pf_status pf_tokens_destroy(pf_tokens* t);
int32_t pf_tokens_count(const pf_tokens* t);
/* copies line `id` (NUL-terminated, truncated to capacity), returns its byte length or -1 */
int32_t pf_tokens_get(const pf_tokens* t, int32_t id, char* buf, size_t capacity);
/* OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302): one us_cif_peak row + the picked ids ->
 * [start_ms, end_ms] pairs.  out_pairs may be NULL to query n_pairs.  PF_ERR_SHAPE where the C# would throw. */
pf_status pf_timestamps_lfr6(const float* us_cif_peak, int32_t num_frames, const int32_t* token_ids, int32_t n_ids,
                             float begin_time, float total_offset, int32_t* out_pairs, int32_t capacity_pairs,
                             int32_t* n_pairs);
/* OfflineRecognizer.DecodeMulti for one stream (OfflineRecognizer.cs:304-418): ids + [n_timestamps][2] stamps (NULL =
 * the {0,0} per id of the 3-output models) -> Text, Tokens, Timestamps.  PF_ERR_BAD_ARG with the required sizes filled
 * in when a buffer is too small; PF_ERR_SHAPE for an id outside the table (IndexOutOfRangeException in the C#). */
pf_status pf_decode_offline(const pf_tokens* t, const int32_t* token_ids, int32_t n_ids, const int32_t* timestamps,
                            int32_t n_timestamps, pf_text_result* out);
/* the same straight from a run: row `utt` of res->tokens, timestamps from res->us_cif_peak when the run produced it */
pf_status pf_decode_offline_result(const pf_tokens* t, const pf_result* res, int32_t utt, pf_text_result* out);
/* OnlineRecognizer.DecodeMulti for one stream (OnlineRecognizer.cs:403-436), lower-cased */
pf_status pf_decode_online(const pf_tokens* t, const int32_t* token_ids, int32_t n_ids, char* text, size_t capacity,
                           size_t* text_bytes);

const char* pf_last_error(void);
int32_t pf_abi_version(void);
/* 1 when the library was built with PFASR_BUILD_EXPERIMENTS=1 (slower A/B variants compiled in: fused FFN kernel, CTA-pair
 * and sixteen-epilogue-warp GEMMs, fused-LayerNorm GEMM epilogue); the product build returns 0 */
int32_t pf_build_experiments(void);

/* -------- op-level test hooks (device 0 of the process; fp32 host in/out, the library converts to its compute
 * types).  Not used by the product path; they let tests/ check each kernel against the oracle through the C-ABI. */
pf_status pf_dbg_gemm(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias,
                      const float* resid, const float* addend, int32_t relu, int32_t out_half, int32_t tile_n,
                      float* out, float* elapsed_ms, int32_t iters);
/* relu?(LayerNorm(x [M, 512]) W^T + bias) as fp16 through the row-tile-stationary LayerNorm + GEMM kernel (csrc/gemm_ln.cu) */
pf_status pf_dbg_ln_gemm(int32_t M, int32_t N, const float* x, const float* gamma, const float* beta, float eps, const float* W,
                         const float* bias, int32_t relu, float* out, float* elapsed_ms, int32_t iters);
/* greedy ids of the rows of A W^T + bias with the pick fused into the GEMM epilogue (no [M, N] tensor is written) */
pf_status pf_dbg_gemm_pick(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias, int32_t tile_n,
                           int32_t* tokens);
/* x = A W^T + bias + resid (fp32) and LN(x) * gamma + beta (fp16) from the fused-LayerNorm GEMM epilogue */
/* x + relu(a W1^T + b1) W2^T + b2 through the fused feed-forward kernel (csrc/ffn_chain.cu); D, F multiples of 256 */
pf_status pf_dbg_ffn_chain(int32_t M, int32_t D, int32_t F, const float* a, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* x, float* out, float* elapsed_ms, int32_t iters);
pf_status pf_dbg_gemm_ln(int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias,
                         const float* resid, const float* gamma, const float* beta, float eps, float* out, float* out_ln);
/* one utterance through the device audio converter: out receives min(capacity, n) samples, *n the converted length */
pf_status pf_dbg_audio_convert(const pf_audio* a, float* out, int64_t capacity, int64_t* n);
pf_status pf_dbg_layernorm(int32_t M, int32_t D, const float* x, const float* gamma, const float* beta, float eps,
                           float* out);
pf_status pf_dbg_embed_pe_ln(int32_t B, int32_t T, int32_t D, const float* feats, float scale, const float* gamma,
                             const float* beta, float eps, float* out);
pf_status pf_dbg_attention(int32_t B, int32_t H, int32_t Tq, int32_t Tk, const float* q, const float* k, const float* v,
                           float* out);
/* encoder self-attention + FSMN memory of one layer from a packed qkv [B*T, 3*H*128] (fused tcgen05 kernel when T <= 192) */
pf_status pf_dbg_attention_fsmn(int32_t B, int32_t H, int32_t T, int32_t taps, const float* qkv, const float* w, float* ctx,
                                float* mem);
pf_status pf_dbg_fsmn(int32_t B, int32_t T, int32_t D, int32_t K, const float* x, const float* w, const float* resid,
                      const int32_t* lens, int32_t half_input, float* out);
pf_status pf_dbg_cif(int32_t B, int32_t T, int32_t D, const float* hidden, const float* alphas_with_tail,
                     float threshold, int32_t lcap, float* embeds, int32_t* token_num, int32_t* fires, float* peaks);
pf_status pf_dbg_logsoftmax_argmax(int32_t M, int32_t V, float* logits_inout, int32_t* tokens);

#ifdef __cplusplus
}
#endif
#endif /* PF_ABI_H_ */
