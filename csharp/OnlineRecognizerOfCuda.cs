// Streaming on libpfasr: what OnlineStream.AddSamples (OnlineStream.cs:84-112) and OnlineRecognizer.Forward
// (OnlineRecognizer.cs:341-401: EncoderProj + PredictorProj + DecoderProj + stack/unstack_states) become.  All per-stream
// model state (fbank FIFO, splice frame, feature cache, CIF carry, 16 FSMN caches) lives in HBM inside the handle; the C#
// classes keep their signatures and only the token list stays managed.
using System;
using System.Collections.Generic;
using System.Linq;
using System.Runtime.InteropServices;
using AliParaformerAsr.Model;
using AliParaformerAsr.Native;

namespace AliParaformerAsr
{
    internal sealed class OnlineSessionOfCuda : IDisposable
    {
        private IntPtr _h;

        public OnlineSessionOfCuda(ConfEntity conf, string weightsPath, float[] addShift, float[] rescale, int vocab, bool perLayerCaches = false)
        {
            var c = PfAsr.ConfigFrom(conf, vocab);
            c.online_flags = perLayerCaches ? 1 : 0;               // 0 = the reference's stack_states behaviour (OnlineModel.cs:222, Q11)
            PfAsr.Check(PfAsr.pf_online_create(ref c, weightsPath, null, 0, out _h), "Online recognition failed", "OnlineRecognizer");
            PfAsr.Check(PfAsr.pf_online_set_cmvn(_h, addShift, rescale, addShift.Length), "Online recognition failed", "OnlineRecognizer");
        }

        /// <summary>OnlineRecognizer.CreateOnlineStream (OnlineRecognizer.cs:27): a fresh state slot on the device.</summary>
        public int OpenStream()
        {
            PfAsr.Check(PfAsr.pf_online_stream_open(_h, out int id), "CreateOnlineStream", "OnlineRecognizer");
            return id;
        }

        public void CloseStream(int id) => PfAsr.pf_online_stream_close(_h, id);

        /// <summary>OnlineStream.AddSamples: the 9600-zero sample cache and the one-chunk-per-call rule (Q13) live in the library.</summary>
        public void AddSamples(int id, float[] samples)
        {
            if (samples == null) throw new NullReferenceException();           // what samples.Length throws today (OnlineStream.cs:88)
            PfAsr.Check(PfAsr.pf_online_stream_push(_h, id, samples, samples.Length), "AddSamples", "OnlineRecognizer");
        }

        /// <summary>OnlineRecognizer.Forward over the listed streams: appends the new ids to every stream's Tokens (:384-393).</summary>
        public void Forward(IReadOnlyList<(int id, List<long> tokens)> streams)
        {
            if (streams.Count == 0) return;                                     // OnlineRecognizer.cs:343-346
            int[] ids = streams.Select(s => s.id).ToArray();
            PfAsr.Check(PfAsr.pf_online_step(_h, ids, ids.Length, 0, out var r), "Online recognition failed", "OnlineRecognizer");
            for (int i = 0; i < ids.Length; i++)
            {
                int n = Marshal.ReadInt32(r.appended, 4 * i);                   // max_new for working streams (padded rows included, :390)
                for (int j = 0; j < n; j++) streams[i].tokens.Add(Marshal.ReadInt32(r.new_tokens, 4 * (i * r.max_new + j)));
            }
        }

        public void Dispose()
        {
            if (_h != IntPtr.Zero) { PfAsr.pf_online_destroy(_h); _h = IntPtr.Zero; }
        }
    }
}
