// IOfflineProj on libpfasr: the drop-in for OfflineProjOfParaformer / OfflineProjOfSenseVoiceSmall /
// OfflineProjOfSeacoParaformer (IOfflineProj.cs:6-40).  Drop into AliParaformerAsr/ and add one case to the
// switch (conf.model.ToLower()) of OfflineRecognizer.cs:39-53 (or an environment switch such as MANYSPEECH_BACKEND=pfasr).
//
// Two modes:
//  * compatibility (ModelProj): same signature and return type as today, so OfflineRecognizer.Forward
//    (OfflineRecognizer.cs:118-198) is untouched - features come from the managed WavFrontend, PadHelper.PadSequence
//    still runs (Q4), log-probs go back as a DenseTensor and the C# argmax loop picks the ids.
//  * fast (RunPcm): PCM in, greedy ids out - front-end, PadSequence, network and argmax all on the device; this is what
//    bench.py reports as e2e.
using System;
using System.Collections.Generic;
using System.Linq;
using System.Runtime.InteropServices;
using AliParaformerAsr.Model;
using AliParaformerAsr.Native;
using AliParaformerAsr.Utils;
using Microsoft.ML.OnnxRuntime;
using Microsoft.ML.OnnxRuntime.Tensors;

namespace AliParaformerAsr
{
    internal class OfflineProjOfCuda : IOfflineProj, IDisposable
    {
        private IntPtr _h;
        private readonly bool _seaco, _timestamps;
        private readonly int[]? _fileHotwords;                     // [n, 10] padded ids of the hot-word file (ctor) or null
        private readonly int _fileHotwordCount;
        private readonly int _inputSize;

        public InferenceSession ModelSession { get => null!; set { } }   // no ORT session any more
        public int Blank_id { get; set; } = 0;
        public int Sos_eos_id { get; set; } = 1;
        public int Unk_id { get; set; } = 2;
        public int SampleRate { get; set; } = 16000;
        public int FeatureDim { get; set; } = 80;

        /// <param name="weightsPath">PFW1 blob converted once from model.onnx (aliparaformerasr_b200/onnx_weights.py)</param>
        /// <param name="hotwords">file hot words as OfflineRecognizer.GetHotwords returns them (OfflineRecognizer.cs:72-90), SeACo only</param>
        public OfflineProjOfCuda(ConfEntity conf, string weightsPath, float[] addShift, float[] rescale, int vocab,
                                 List<int[]>? hotwords = null, bool timestamps = false, int lanes = 1)
        {
            var c = PfAsr.ConfigFrom(conf, vocab);
            _seaco = c.model_kind == PfAsr.PF_MODEL_SEACO_PARAFORMER;
            _timestamps = timestamps || _seaco;                    // 4-output graphs (OfflineProjOfParaformer.cs:75-79)
            _inputSize = c.input_size;
            PfAsr.Check(PfAsr.pf_offline_create_mt(ref c, weightsPath, null, 0, lanes, out _h), "ModelProj failed");
            PfAsr.Check(PfAsr.pf_offline_set_cmvn(_h, addShift, rescale, addShift.Length), "ModelProj failed");
            if (_seaco && hotwords != null && hotwords.Count > 0)  // EmbedSeacoModel.Forward once (OfflineProjOfSeacoParaformer.cs:29-34)
            {
                _fileHotwords = PadList(hotwords);
                _fileHotwordCount = hotwords.Count;
                PfAsr.Check(PfAsr.pf_offline_set_hotwords(_h, _fileHotwords, _fileHotwordCount), "ModelProj failed");
            }
        }

        /// <summary>EmbedSeacoModel.PadList (EmbedSeacoModel.cs:110-123): truncate to 10 ids, right-pad with 0.</summary>
        private static int[] PadList(List<int[]> hotwords, int maxLen = 10)
        {
            var ids = new int[hotwords.Count * maxLen];
            for (int i = 0; i < hotwords.Count; i++)
                Array.Copy(hotwords[i], 0, ids, i * maxLen, Math.Min(hotwords[i].Length, maxLen));
            return ids;
        }

        // ---------------------------------------------------------------- compatibility mode
        public ModelOutputEntity ModelProj(List<OfflineInputEntity> modelInputs)
        {
            if (_h == IntPtr.Zero) throw new ObjectDisposedException("OfflineRecognizer");
            int B = modelInputs.Count;
            float[] speech = PadHelper.PadSequence(modelInputs);                       // unchanged (Utils/PadHelper.cs:23-65, Q4)
            int T = speech.Length / _inputSize / B;                                    // OfflineProjOfParaformer.cs:53 (Q3)
            // per-stream hot words replace the file ones for this call (OfflineProjOfSeacoParaformer.cs:51-60); the lane
            // stays leased from set to restore so that another request thread cannot run with them
            var callHot = _seaco ? modelInputs.Where(x => x.Hotwords != null).SelectMany(x => x.Hotwords!).ToList() : new List<int[]>();
            uint flags = PfAsr.PF_RUN_WANT_LOGITS | (_timestamps ? PfAsr.PF_RUN_WANT_TIMESTAMPS : 0u);
            PfAsr.pf_offline_lane_acquire(_h);
            try
            {
                if (callHot.Count > 0) PfAsr.Check(PfAsr.pf_offline_set_hotwords_local(_h, PadList(callHot), callHot.Count), "ModelProj failed");
                PfAsr.Check(PfAsr.pf_offline_run_feats(_h, speech, B, T, flags, out var r), "ModelProj failed");
                var logits = new float[B * r.max_len * r.vocab];
                if (logits.Length > 0) Marshal.Copy(r.logits, logits, 0, logits.Length);
                var lens = new int[B];
                Marshal.Copy(r.token_num, lens, 0, B);
                var entity = new ModelOutputEntity
                {
                    model_out = new DenseTensor<float>(logits, new[] { B, r.max_len, r.vocab }),
                    model_out_lens = lens,
                };
                if (r.us_frames > 0 && r.us_cif_peak != IntPtr.Zero)                   // results[3] (OfflineProjOfParaformer.cs:75-79)
                {
                    var peak = new float[B * r.us_frames];
                    Marshal.Copy(r.us_cif_peak, peak, 0, peak.Length);
                    entity.cif_peak_tensor = new DenseTensor<float>(peak, new[] { B, r.us_frames });
                }
                return entity;
            }
            finally
            {
                if (callHot.Count > 0) PfAsr.pf_offline_set_hotwords_local(_h, _fileHotwords, _fileHotwordCount);
                PfAsr.pf_offline_lane_release(_h);
            }
        }

        // ---------------------------------------------------------------- fast mode
        /// <summary>One AddSamples per stream, then GetResults: returns the greedy ids [B][L] the loop at
        /// OfflineRecognizer.cs:139-152 would have produced (Q5) and, for 4-output models, us_cif_peak rows.</summary>
        public (long[][] ids, float[][]? usCifPeak) RunPcm(IReadOnlyList<float[]> pcm)
        {
            if (_h == IntPtr.Zero) throw new ObjectDisposedException("OfflineRecognizer");
            int B = pcm.Count;
            var pins = new GCHandle[B];
            var ptrs = new IntPtr[B];
            var ns = new int[B];
            try
            {
                for (int i = 0; i < B; i++)
                {
                    if (pcm[i] == null) throw new ArgumentNullException("source");      // what WavFrontend.cs:34 throws today
                    pins[i] = GCHandle.Alloc(pcm[i], GCHandleType.Pinned);
                    ptrs[i] = pins[i].AddrOfPinnedObject();
                    ns[i] = pcm[i].Length;
                }
                PfAsr.Check(PfAsr.pf_offline_run_pcm(_h, ptrs, ns, B, _timestamps ? PfAsr.PF_RUN_WANT_TIMESTAMPS : 0u, out var r), "Offline recognition failed");
                var flat = new int[B * r.max_len];
                if (flat.Length > 0) Marshal.Copy(r.tokens, flat, 0, flat.Length);
                var ids = new long[B][];
                for (int i = 0; i < B; i++) ids[i] = flat.Skip(i * r.max_len).Take(r.max_len).Select(x => (long)x).ToArray();
                float[][]? us = null;
                if (r.us_frames > 0 && r.us_cif_peak != IntPtr.Zero)
                {
                    us = new float[B][];
                    for (int i = 0; i < B; i++)
                    {
                        us[i] = new float[r.us_frames];
                        Marshal.Copy(r.us_cif_peak + 4 * i * r.us_frames, us[i], 0, r.us_frames);
                    }
                }
                return (ids, us);
            }
            finally
            {
                foreach (var p in pins) if (p.IsAllocated) p.Free();
            }
        }

        public void Dispose()
        {
            if (_h != IntPtr.Zero) { PfAsr.pf_offline_destroy(_h); _h = IntPtr.Zero; }
        }
    }
}
