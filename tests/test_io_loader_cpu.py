"""on-disk formats.  WAV I/O (the torchaudio.load / save contract of the reference's example.py)
and FlowHighSR.from_local on a checkpoint directory laid out like the ResembleAI/FlowHigh hub repo
(flowhighsr.py:109-137, init_vocoder.py:8-23): BigVGAN JSON + ['generator'] in weight-norm form,
FLowHigh_basic_400k.pt ['model'] in folded form.  No GPU: loading stops before the engine is built."""
import json
import struct
import wave

import numpy as np
import pytest
import torch

from flowhigh_b200 import FlowHighSR, VocoderConfig
from flowhigh_b200.config import BackboneConfig
from flowhigh_b200.io import load_wav, save_wav
from flowhigh_b200.weights import fold_weight_norm, random_state_dict
from util import write_hub_dir

VOC = "flowhigh.audio_enc_dec.vocoder."


def test_wav_roundtrip_pcm16_and_float32(tmp_path):
    rng = np.random.default_rng(0)
    x = np.clip(rng.standard_normal((2, 4801)) * 0.3, -0.99, 0.99).astype(np.float32)
    save_wav(tmp_path / "a.wav", x, 22050)
    y, sr = load_wav(tmp_path / "a.wav")
    assert sr == 22050 and y.shape == (2, 4801) and y.dtype == torch.float32
    assert np.abs(y.numpy() - x).max() <= 0.5 / 32768 + 1e-7  # PCM16 quantisation only
    save_wav(tmp_path / "b.wav", torch.from_numpy(x[0]), 48000, bits_per_sample=32)
    z, sr = load_wav(tmp_path / "b.wav")
    assert sr == 48000 and z.shape == (1, 4801) and np.array_equal(z.numpy()[0], x[0])  # float32: bit exact


def test_wav_reader_matches_stdlib_writer_and_24bit(tmp_path):
    pcm = (np.arange(-500, 500, dtype=np.int16) * 60)
    with wave.open(str(tmp_path / "c.wav"), "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(pcm.astype("<i2").tobytes())
    y, sr = load_wav(tmp_path / "c.wav")
    assert sr == 16000 and np.array_equal(y.numpy()[0], pcm.astype(np.float32) / 32768.0)
    # 24-bit PCM with an odd-sized LIST chunk before the data chunk (chunk word alignment)
    v = np.array([0, 1, -1, 8388607, -8388608, 123456], dtype=np.int64)
    body = b"".join(struct.pack("<i", int(s))[:3] for s in v)
    fmt = struct.pack("<HHIIHH", 1, 1, 44100, 44100 * 3, 3, 24)
    junk = b"LIST" + struct.pack("<I", 3) + b"abc\0"
    blob = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + junk + b"data" + struct.pack("<I", len(body)) + body
    (tmp_path / "d.wav").write_bytes(b"RIFF" + struct.pack("<I", len(blob)) + blob)
    y, sr = load_wav(tmp_path / "d.wav")
    assert sr == 44100 and np.allclose(y.numpy()[0], v.astype(np.float64) / 8388608.0, atol=1e-7)
    with pytest.raises(ValueError):
        (tmp_path / "e.wav").write_bytes(b"not a wav file at all")
        load_wav(tmp_path / "e.wav")


def test_from_local_reads_hub_layout(tmp_path):
    vcfg = VocoderConfig.tiny()
    sd = random_state_dict(BackboneConfig(), vcfg, seed=3, vocoder_gain=0.7)
    gen = write_hub_dir(tmp_path, vcfg, sd)
    folded = fold_weight_norm(gen)
    for k, t in sd.items():
        if k.startswith(VOC):
            assert torch.allclose(folded[k[len(VOC):]], t, rtol=1e-6, atol=1e-7), k

    model = FlowHighSR.from_local(tmp_path, device="cpu")
    got = model.state_dict()
    assert list(got.keys()) == list(sd.keys())  # the reference's key layout, order included
    for k in sd:
        assert torch.equal(got[k], sd[k]), k
    # flowhighsr.py:124-129: from_local builds the wrapper with its defaults (F4)
    assert model.cfm_method == "basic_cfm" and model.odeint_kwargs["method"] == "midpoint" and float(model.sigma) == 0.0
    # a strict load must reject a missing key, like nn.Module.load_state_dict
    bad = dict(sd)
    bad.pop(next(iter(bad)))
    with pytest.raises((RuntimeError, KeyError)):
        model.load_state_dict(bad)
    # no CPU fallback: using the model without a CUDA device is an explicit error
    with pytest.raises(RuntimeError):
        model.generate(np.zeros(16000, np.float32) + 0.1, 16000, 48000)
