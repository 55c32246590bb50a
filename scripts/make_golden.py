"""Generates tests/golden/*: independent known-answer vectors for the oracle and the CUDA front-end.

* kaldi_fbank.npz      torchaudio.compliance.kaldi.fbank (an independent implementation of Kaldi
                       compute-fbank-feats, the semantics of SpeechFeatures.OnlineFbank) on a seeded waveform.
* sensevoice_embed.npy the 16x560 float32 prompt table stored in the reference's data/embed.onnx
                       (single Gather initialiser; raw payload at byte offset 98), sha256-pinned.
Run in the build container (needs torchaudio and /root/reference):  python scripts/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aliparaformerasr_b200 import synth  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

pcm = synth.make_pcm(4242, 1.3)
kw = dict(num_mel_bins=80, frame_length=25.0, frame_shift=10.0, dither=0.0, window_type="hamming", energy_floor=0.0,
          sample_frequency=16000.0)
wave = torch.from_numpy(pcm)[None] * 32768.0
np.savez_compressed(os.path.join(out, "kaldi_fbank.npz"), pcm=pcm,
                    fbank_snip0=torchaudio.compliance.kaldi.fbank(wave, snip_edges=False, **kw).numpy(),
                    fbank_snip1=torchaudio.compliance.kaldi.fbank(wave, snip_edges=True, **kw).numpy())

ref = "/root/reference/AliParaformerAsr/data/embed.onnx"
if os.path.exists(ref):
    raw = open(ref, "rb").read()
    assert hashlib.sha256(raw).hexdigest().startswith("5c69dceb"), "embed.onnx changed"
    table = np.frombuffer(raw[98:98 + 16 * 560 * 4], dtype="<f4").reshape(16, 560).copy()
    np.save(os.path.join(out, "sensevoice_embed.npy"), table)
    print("embed table row norms", np.linalg.norm(table, axis=1).round(2))
print("golden vectors written to", out)
