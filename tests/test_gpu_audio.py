"""Audio ingestion on the device: csrc/audio.cu against oracle/audio.py, bit for bit, and the
end-to-end equivalence run_audio(raw file samples) == run_pcm(GetFileSample(...))."""
import numpy as np
import pytest

from aliparaformerasr_b200 import _lib, audio, synth
from aliparaformerasr_b200.engine import Engine
from oracle import audio as oaudio

pytestmark = pytest.mark.gpu


def _payload(rng, fmt, n):
    if fmt == _lib.PF_AUDIO_U8:
        return rng.integers(0, 256, n, dtype=np.uint8)
    if fmt == _lib.PF_AUDIO_S16:
        return rng.integers(-32768, 32768, n).astype(np.int16)
    if fmt == _lib.PF_AUDIO_S24:
        return rng.integers(0, 256, 3 * n, dtype=np.uint8)
    if fmt == _lib.PF_AUDIO_S32:
        return rng.integers(-2**31, 2**31, n).astype(np.int32)
    return (rng.standard_normal(n) * 0.3).astype(np.float32)


@pytest.mark.parametrize("fmt", [_lib.PF_AUDIO_U8, _lib.PF_AUDIO_S16, _lib.PF_AUDIO_S24, _lib.PF_AUDIO_S32, _lib.PF_AUDIO_F32])
def test_convert_bit_exact(fmt):
    rng = np.random.default_rng(fmt)
    for rate in (16000, 8000, 11025, 22050, 32000, 44100, 48000, 96000):
        for ch in (1, 2):
            n = int(rng.integers(1, 40000))
            data = _payload(rng, fmt, n)
            clip = audio.Audio(data, fmt, ch, rate)
            got = clip.to_pcm()
            ref = oaudio.get_file_sample(data, fmt, ch, rate)
            assert got.shape == ref.shape, (rate, ch, n)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (rate, ch, n)


def test_convert_edges():
    one = audio.Audio(np.asarray([1234], np.int16), _lib.PF_AUDIO_S16, 1, 8000)
    assert one.to_pcm().tolist() == oaudio.get_file_sample(one.data, oaudio.S16, 1, 8000).tolist() == [1234 / 32768, 1234 / 32768]
    empty = audio.Audio(np.zeros(0, np.int16), _lib.PF_AUDIO_S16, 2, 44100)
    assert empty.to_pcm().size == 0
    # int32 -> float rounds to nearest even before the division
    x = np.asarray([2**31 - 1, 2**24 + 1, -(2**24) - 3], np.int32)
    clip = audio.Audio(x, _lib.PF_AUDIO_S32, 1, 16000)
    assert np.array_equal(clip.to_pcm(), oaudio.to_float(x, oaudio.S32))


def test_run_audio_equals_run_pcm_of_converted_samples():
    cfg = synth.tiny()
    eng = Engine(cfg, synth.make_weights(cfg))
    rng = np.random.default_rng(5)
    clips = []
    for fmt, rate, ch, secs in ((_lib.PF_AUDIO_S16, 16000, 1, 1.3), (_lib.PF_AUDIO_S16, 44100, 2, 0.9), (_lib.PF_AUDIO_F32, 8000, 1, 1.1),
                                (_lib.PF_AUDIO_S24, 48000, 2, 0.7), (_lib.PF_AUDIO_U8, 22050, 1, 1.0), (_lib.PF_AUDIO_S32, 16000, 2, 0.4)):
        n = int(secs * rate) * ch
        t = np.arange(n) / (rate * ch)
        wave = 0.4 * np.sin(2 * np.pi * 220.0 * t) + 0.05 * rng.standard_normal(n)
        if fmt == _lib.PF_AUDIO_S16:
            data = np.clip(wave * 32767, -32768, 32767).astype(np.int16)
        elif fmt == _lib.PF_AUDIO_F32:
            data = wave.astype(np.float32)
        elif fmt == _lib.PF_AUDIO_U8:
            data = np.clip(wave * 127 + 128, 0, 255).astype(np.uint8)
        elif fmt == _lib.PF_AUDIO_S32:
            data = np.clip(wave * 2**31, -2**31, 2**31 - 1).astype(np.int32)
        else:
            v = np.clip(wave * 2**23, -2**23, 2**23 - 1).astype(np.int32)
            data = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=1).astype(np.uint8).reshape(-1)
        clips.append(audio.Audio(data, fmt, ch, rate))
    pcm = [oaudio.get_file_sample(c.data, c.format, c.channels, c.sample_rate) for c in clips]
    a = eng.run_audio(clips, want_logits=True)
    feats_a = eng.tensor("feats")
    b = eng.run_pcm(pcm, want_logits=True)
    feats_b = eng.tensor("feats")
    assert np.array_equal(feats_a, feats_b)                      # identical PCM -> identical features
    assert np.array_equal(a.tokens, b.tokens) and np.array_equal(a.token_num, b.token_num)
    assert np.array_equal(a.logits, b.logits)
    # a second audio batch after a PCM batch reuses the staging buffers
    c = eng.run_audio(clips[:2])
    assert np.array_equal(c.tokens[:, : c.tokens.shape[1]], eng.run_pcm(pcm[:2]).tokens)
    eng.close()


def test_plain_c_consumer_end_to_end(tmp_path):
    """The C example (WAV files -> text through the C-ABI alone) prints what the Python host mirror decodes."""
    import shutil
    import struct
    import subprocess
    import os
    from aliparaformerasr_b200 import weights as W
    from aliparaformerasr_b200.text import TokenTable
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "aliparaformerasr_b200")
    exe = tmp_path / "offline_wav"
    r = subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "offline_wav.c"), "-L", libdir,
                        "-lpfasr", f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cfg = synth.tiny()
    w = synth.make_weights(cfg)
    blob = tmp_path / "model.pfw"
    W.pack(w).tofile(str(blob))
    vocab = ["<blank>", "<s>", "</s>"] + [chr(0x4E00 + i) if i % 3 else f"w{i}@@" for i in range(cfg.vocab - 3)]
    tok = tmp_path / "tokens.txt"
    tok.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    wavs, clips = [], []
    for i, (rate, ch) in enumerate(((16000, 1), (44100, 2))):
        pcm = synth.make_pcm(i, 1.5)
        n = int(1.5 * rate)
        x = np.interp(np.arange(n) * 16000.0 / rate, np.arange(pcm.size), pcm)
        s16 = np.clip(x * 32767, -32768, 32767).astype(np.int16)
        data = np.repeat(s16, ch) if ch == 2 else s16
        hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + data.nbytes, b"WAVE", b"fmt ", 16, 1, ch, rate, rate * ch * 2, ch * 2, 16,
                          b"data", data.nbytes)
        path = tmp_path / f"u{i}.wav"
        path.write_bytes(hdr + data.tobytes())
        wavs.append(str(path))
        clips.append(audio.Audio(data, _lib.PF_AUDIO_S16, ch, rate))
    env = dict(os.environ, PFASR_EXAMPLE_LAYERS=f"{cfg.enc_layers},{cfg.dec_layers}")
    r = subprocess.run([str(exe), str(blob), str(tok)] + wavs, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    got = [line.split("\t", 1)[1] for line in r.stdout.splitlines()]
    eng = Engine(cfg, w)                                  # default CMVN (shift 0, scale 1), as the example uses
    out = eng.run_audio(clips)
    table = TokenTable(lines=vocab)
    want = [table.decode_offline(out.tokens[i])[0] for i in range(2)]
    assert got == want and any(len(t) > 0 for t in want)
    eng.close()
