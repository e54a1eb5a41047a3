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
