"""CPU oracle for the AliParaformerAsr offline hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``aliparaformerasr_b200/`` (the product)
imports this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker (or as the CPU arm that is timed *beside* the GPU
path, never instead of it).

Parity status
-------------
* ``oracle.frontend`` (Kaldi fbank -> LFR -> CMVN -> pad quirk) follows
  ``/root/reference/AliParaformerAsr/WavFrontend.cs:31-153`` and
  ``Utils/PadHelper.cs:23-65`` line by line.  The fbank arithmetic itself lives in
  the un-vendored NuGet package ``ManySpeech.SpeechFeatures 1.1.7`` (Kaldi
  ``compute-fbank-feats`` semantics); it is PINNED against
  ``torchaudio.compliance.kaldi.fbank`` golden vectors committed under
  ``tests/golden/`` (generator: ``scripts/make_golden.py``).
* ``oracle.sanm`` (SAN-M encoder, CIF predictor, SANM decoder, log-softmax,
  greedy pick) restates the FunASR export graph that the reference executes
  through ``Microsoft.ML.OnnxRuntime 1.22.*`` (``OfflineProjOfParaformer.cs:68``).
  Neither onnxruntime nor any ``model.onnx`` exists in the build container and the
  reference's tests hold no numeric vectors for it: **parity unpinned** for the
  network graph.  The only numeric artefact of the
  reference, ``data/embed.onnx`` (16x560 SenseVoice prompt table), is pinned by
  sha256 and committed as ``tests/golden/sensevoice_embed.npy``.
"""
