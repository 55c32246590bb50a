// P/Invoke surface of libpfasr (include/pf_abi.h, ABI version 5) for the reference's C# host.
//
// Drop into AliParaformerAsr/Native/.  Replaces the two native dependencies of the offline / online hot path:
// Microsoft.ML.OnnxRuntime (InferenceSession.Run at OfflineProjOfParaformer.cs:68, OfflineProjOfSenseVoiceSmall.cs:156,
// OfflineProjOfSeacoParaformer.cs:116, EmbedSeacoModel.cs:97, OnlineRecognizer.cs:80,299) and ManySpeech.SpeechFeatures
// (OnlineFbank at WavFrontend.cs:21-37, OnlineWavFrontend.cs:21-36).  Every struct mirrors the C struct of the same
// name field for field (tests/test_csharp_shim.py checks the field lists and the exported symbols against the header).
using System;
using System.Runtime.InteropServices;

namespace AliParaformerAsr.Native
{
    [StructLayout(LayoutKind.Sequential)]
    internal struct PfConfig          // pf_config (filled from ConfEntity, Model/ConfEntity.cs:5-43)
    {
        public int struct_bytes, model_kind, input_size, d_model, heads, ffn, enc_layers, tp_layers, enc_kernel,
                   dec_layers, dec_ffn, dec_kernel, vocab;
        public float ln_eps, cif_threshold, cif_tail, smooth_factor, noise_threshold;
        public int fs, n_mels, lfr_m, lfr_n, snip_edges, use_itn;
        public int online_flags;                                             // bit 0: per-layer FSMN caches (0 = Q11-compatible)
        public int seaco_layers, seaco_ffn, seaco_kernel, seaco_nobias_id;   // seaco_decoder_conf + NO_BIAS id (8377)
        public float smooth_factor2, noise_threshold2;                       // CifPredictorV3 timestamp branch (0.25, 0.01)
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct PfResult          // pf_result
    {
        public int batch, max_len, vocab, feat_frames;
        public IntPtr tokens;         // int32 [B, L]   greedy ids (OfflineRecognizer.cs:139-152 done on the device, Q5)
        public IntPtr token_num;      // int32 [B]      model_out_lens
        public IntPtr logits;         // float [B,L,V]  only with PF_RUN_WANT_LOGITS
        public IntPtr cif_peak;       // float [B,T+1]  only with PF_RUN_WANT_CIF_PEAK
        public int us_frames;         // 3*T, 0 when absent
        public IntPtr us_alphas;      // float [B,3T]   only with PF_RUN_WANT_TIMESTAMPS (models with the V3 predictor)
        public IntPtr us_cif_peak;    // float [B,3T]   = ModelOutputEntity.cif_peak_tensor (OfflineProjOfParaformer.cs:75-79)
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct PfOnlineResult    // pf_online_result
    {
        public int n_streams, max_new, vocab, n_working;
        public IntPtr appended;       // int32 [n]          ids appended to stream i (max_new for working streams)
        public IntPtr new_tokens;     // int32 [n, max_new] greedy ids of this step
        public IntPtr embeds_len;     // int32 [n]          acoustic_embeds_len
        public IntPtr logits;         // float [n, max_new, V] only with PF_RUN_WANT_LOGITS
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct PfAudio           // pf_audio
    {
        public IntPtr data;
        public long n_values;
        public int format, channels, sample_rate, reserved;
    }

    [StructLayout(LayoutKind.Sequential)]
    internal struct PfTextResult      // pf_text_result
    {
        public IntPtr text; public UIntPtr text_capacity, text_bytes;
        public int text_len, n_tokens, n_timestamps, reserved;
        public IntPtr tokens; public UIntPtr tokens_capacity, tokens_bytes;
        public IntPtr ts; public UIntPtr ts_capacity, ts_count;
        public IntPtr ts_offsets; public UIntPtr ts_offsets_capacity;
    }

    internal static class PfAsr
    {
        const string Lib = "pfasr";   // libpfasr.so / pfasr.dll on the probing path

        internal const int PF_MODEL_PARAFORMER = 0, PF_MODEL_SENSEVOICE_SMALL = 1, PF_MODEL_SEACO_PARAFORMER = 2;
        internal const uint PF_RUN_WANT_LOGITS = 1, PF_RUN_WANT_CIF_PEAK = 2, PF_RUN_WANT_TIMESTAMPS = 4;
        internal const int PF_ERR_DISPOSED = -5;

        // lifecycle (OfflineModel.initModel, OfflineModel.cs:35-70)
        [DllImport(Lib)] internal static extern int pf_abi_version();
        [DllImport(Lib)] internal static extern int pf_offline_create(ref PfConfig cfg, string weightsPath, int[]? devices, int ndev, out IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_create_mt(ref PfConfig cfg, string weightsPath, int[]? devices, int ndev, int lanes, out IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_lanes(IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_lane_acquire(IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_lane_release(IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_destroy(IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_offline_set_cmvn(IntPtr handle, float[] addShift, float[] rescale, int dim);
        // SeACo hot words (EmbedSeacoModel.Forward, EmbedSeacoModel.cs:70-108)
        [DllImport(Lib)] internal static extern int pf_offline_set_hotwords(IntPtr handle, int[]? idsN10, int n);
        [DllImport(Lib)] internal static extern int pf_offline_set_hotwords_local(IntPtr handle, int[]? idsN10, int n);
        // front-end only (WavFrontend.GetFbank + LfrCmvn, OfflineStream.cs:40-41)
        [DllImport(Lib)] internal static extern int pf_frontend_extract(IntPtr handle, float[] samples, int nsamp, float[] feats, int capacityFrames, out int outFrames);
        [DllImport(Lib)] internal static extern int pf_frontend_num_frames(IntPtr handle, int nsamp);
        // the hot path (IOfflineProj.ModelProj + the argmax loop of OfflineRecognizer.Forward)
        [DllImport(Lib)] internal static extern int pf_offline_run_feats(IntPtr handle, float[] speech, int batch, int frames, uint flags, out PfResult result);
        [DllImport(Lib)] internal static extern int pf_offline_run_pcm(IntPtr handle, IntPtr[] pcm, int[] nsamp, int batch, uint flags, out PfResult result);
        // optional: audio files (AudioHelper.GetFileSample) and token -> text / timestamps (DecodeMulti) inside the library
        [DllImport(Lib)] internal static extern int pf_wav_parse(IntPtr file, UIntPtr bytes, out PfAudio audio);
        [DllImport(Lib)] internal static extern int pf_offline_run_audio(IntPtr handle, PfAudio[] utts, int batch, uint flags, out PfResult result);
        [DllImport(Lib)] internal static extern int pf_tokens_create(string path, out IntPtr tokens);
        [DllImport(Lib)] internal static extern int pf_tokens_destroy(IntPtr tokens);
        [DllImport(Lib)] internal static extern int pf_decode_offline_result(IntPtr tokens, in PfResult res, int utt, ref PfTextResult text);
        [DllImport(Lib)] internal static extern int pf_decode_online(IntPtr tokens, int[] ids, int n, byte[]? text, UIntPtr capacity, out UIntPtr bytes);
        // streaming (OnlineStream / OnlineModel / OnlineRecognizer.Forward)
        [DllImport(Lib)] internal static extern int pf_online_create(ref PfConfig cfg, string weightsPath, int[]? devices, int ndev, out IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_online_destroy(IntPtr handle);
        [DllImport(Lib)] internal static extern int pf_online_set_cmvn(IntPtr handle, float[] addShift, float[] rescale, int dim);
        [DllImport(Lib)] internal static extern int pf_online_stream_open(IntPtr handle, out int streamId);
        [DllImport(Lib)] internal static extern int pf_online_stream_close(IntPtr handle, int streamId);
        [DllImport(Lib)] internal static extern int pf_online_stream_push(IntPtr handle, int streamId, float[] samples, int nsamp);
        [DllImport(Lib)] internal static extern int pf_online_step(IntPtr handle, int[] streamIds, int n, uint flags, out PfOnlineResult result);
        [DllImport(Lib)] internal static extern IntPtr pf_last_error();

        /// <summary>status -> the exception types / messages the reference throws at the same place.</summary>
        internal static void Check(int status, string what, string objectName = "OfflineRecognizer")
        {
            if (status == 0) return;
            string msg = Marshal.PtrToStringAnsi(pf_last_error()) ?? "";
            if (status == PF_ERR_DISPOSED) throw new ObjectDisposedException(objectName);       // OfflineRecognizer.cs:94-97
            throw new Exception(what, new Exception($"libpfasr status {status}: {msg}"));        // same outer message as today
        }

        /// <summary>Flat copy of the asr.yaml / asr.json fields the path consumes (Model/ConfEntity.cs:5-43).</summary>
        internal static PfConfig ConfigFrom(Model.ConfEntity conf, int vocab, float lnEps = 1e-12f, int seacoNoBiasId = 8377)
        {
            string model = (conf.model ?? "paraformer").ToLower();               // dispatch of OfflineRecognizer.cs:39-53
            var c = new PfConfig
            {
                model_kind = model == "sensevoicesmall" ? PF_MODEL_SENSEVOICE_SMALL : model == "seacoparaformer" ? PF_MODEL_SEACO_PARAFORMER : PF_MODEL_PARAFORMER,
                d_model = conf.encoder_conf.output_size, heads = conf.encoder_conf.attention_heads, ffn = conf.encoder_conf.linear_units,
                enc_layers = conf.encoder_conf.num_blocks, tp_layers = model == "sensevoicesmall" ? conf.encoder_conf.tp_blocks : 0,
                enc_kernel = conf.encoder_conf.kernel_size,
                dec_layers = model == "sensevoicesmall" ? 0 : conf.decoder_conf.num_blocks, dec_ffn = conf.decoder_conf.linear_units,
                dec_kernel = conf.decoder_conf.kernel_size, vocab = vocab, ln_eps = model == "sensevoicesmall" ? 1e-5f : lnEps,
                cif_threshold = conf.predictor_conf.threshold, cif_tail = conf.predictor_conf.tail_threshold,
                smooth_factor = conf.predictor_conf.smooth_factor, noise_threshold = conf.predictor_conf.noise_threshold,
                fs = conf.frontend_conf.fs, n_mels = conf.frontend_conf.n_mels, lfr_m = conf.frontend_conf.lfr_m, lfr_n = conf.frontend_conf.lfr_n,
                snip_edges = conf.frontend_conf.snip_edges ? 1 : 0, use_itn = conf.use_itn ? 1 : 0, online_flags = 0,
                seaco_layers = 4, seaco_ffn = 1024, seaco_kernel = 21, seaco_nobias_id = seacoNoBiasId,
                smooth_factor2 = 0.25f, noise_threshold2 = 0.01f,
            };
            c.input_size = c.lfr_m * c.n_mels;
            c.struct_bytes = Marshal.SizeOf<PfConfig>();
            // the device front-end is Hamming / 25 ms / 10 ms and never dithers (pf_abi.h): refuse anything else loudly
            if (!string.Equals(conf.frontend_conf.window, "hamming", StringComparison.OrdinalIgnoreCase) ||
                conf.frontend_conf.frame_length != 25 || conf.frontend_conf.frame_shift != 10)
                throw new NotSupportedException("libpfasr: frontend_conf must be window=hamming, frame_length=25, frame_shift=10");
            return c;
        }
    }
}
