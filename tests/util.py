import ast
import os

import numpy as np
import torch

from flowhigh_b200.config import BackboneConfig, VocoderConfig
from flowhigh_b200.weights import random_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def vcfg_from_golden(g) -> VocoderConfig:
    d = ast.literal_eval(str(g["vcfg"]))
    return VocoderConfig(resblock=d["resblock"], upsample_rates=tuple(d["upsample_rates"]),
                         upsample_kernel_sizes=tuple(d["upsample_kernel_sizes"]),
                         upsample_initial_channel=d["upsample_initial_channel"],
                         resblock_kernel_sizes=tuple(d["resblock_kernel_sizes"]),
                         resblock_dilation_sizes=tuple(tuple(x) for x in d["resblock_dilation_sizes"]),
                         activation=d["activation"], snake_logscale=d["snake_logscale"], num_mels=d["num_mels"])


def golden_weights(g):
    vcfg = vcfg_from_golden(g)
    sd = random_state_dict(BackboneConfig(), vcfg, seed=int(g["seed"]), vocoder_gain=float(g["gain"]))
    cs = float(sum(v.double().abs().sum().item() for k, v in sd.items() if k.endswith("weight")))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * abs(cs), "weight RNG drifted from the golden fixtures"
    return sd, vcfg


def snr_db(ref, x):
    ref, x = ref.double().flatten(), x.double().flatten()
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum().clamp_min(1e-300)))


def lsd_db(ref, x):
    """log-spectral distance in dB: mean over frames of the RMS over frequency of the difference of
    10*log10 power spectra (n_fft 2048, hop 480, Hann)."""
    w = torch.hann_window(2048, dtype=torch.float64)
    A = torch.stft(ref.double().flatten(), 2048, 480, window=w, return_complex=True).abs() ** 2
    B = torch.stft(x.double().flatten(), 2048, 480, window=w, return_complex=True).abs() ** 2
    d = 10 * torch.log10(A.clamp_min(1e-10)) - 10 * torch.log10(B.clamp_min(1e-10))
    return float(d.pow(2).mean(0).sqrt().mean())


def _unfold(w):
    """weight -> (weight_g, weight_v) of torch.nn.utils.weight_norm (dim 0), with v deliberately not unit-norm."""
    scale = torch.linspace(0.5, 2.0, w.shape[0]).reshape(-1, *([1] * (w.dim() - 1)))
    v = w * scale
    g = w.reshape(w.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (w.dim() - 1)))
    return g, v


def write_hub_dir(path, vcfg, sd):
    """Lays out `path` like the ResembleAI/FlowHigh hub repo (flowhighsr.py:109-149): BigVGAN JSON (the AttrDict fields
    bigvgan/models.py reads), bigvgan .pt = {'generator': weight_g / weight_v pairs} as upstream BigVGAN ships it,
    FLowHigh_basic_400k.pt = {'model': full state dict, vocoder folded}, and the unused FLowHigh json.
    Returns the raw generator dict."""
    import json
    VOC = "flowhigh.audio_enc_dec.vocoder."
    (path / "bigvgan_48khz_256band.json").write_text(json.dumps(vcfg.to_attr_json()))
    gen = {}
    for k, t in sd.items():
        if not k.startswith(VOC):
            continue
        name = k[len(VOC):]
        if name.endswith(".weight") and t.dim() == 3 and ("conv" in name or name.startswith("ups.")):
            g, v = _unfold(t)
            gen[name + "_g"], gen[name + "_v"] = g, v
        else:
            gen[name] = t.clone()
    torch.save({"generator": gen}, path / "bigvgan_48khz_256band.pt")
    torch.save({"model": sd}, path / "FLowHigh_basic_400k.pt")
    (path / "FLowHigh_basic_400k.json").write_text("{}")
    return gen
