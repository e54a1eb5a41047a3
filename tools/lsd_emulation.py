"""CPU experiment: which 16-bit rounding site of the tensor-core vocoder costs how much
log-spectral distance?  Runs the oracle vocoder (fp64 reference run vs fp32 runs with fp16 rounding injected at
selected sites) on a small golden fixture.  Test / analysis tooling only (imports oracle/).

    python tools/lsd_emulation.py [fixture]

Sites: A = conv activation operand, W = conv weights, X = snake input, S = snake 2x-rate samples, T = snake taps.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model  # noqa: E402
from util import golden_weights, load_golden, lsd_db, snr_db  # noqa: E402

h = lambda t: t.half().to(t.dtype)
SITES = set()
SPLIT = set()   # sites kept as hi + lo pairs (22 bits)


def rnd(t, site):
    if site not in SITES:
        return t
    if site in SPLIT:
        hi = h(t)
        return hi + h(t - hi)
    return h(t)


_conv1d, _convT = F.conv1d, F.conv_transpose1d


def conv1d(x, w, b=None, **kw):
    if kw.get("groups", 1) != 1:  # the depthwise anti-alias filters: taps are site T
        return _conv1d(rnd(x, "S"), rnd(w, "T"), b, **kw)
    return _conv1d(rnd(x, "A"), rnd(w, "W"), b, **kw)


def convT(x, w, b=None, **kw):
    if kw.get("groups", 1) != 1:
        return _convT(rnd(x, "X"), rnd(w, "T"), b, **kw)
    return _convT(rnd(x, "A"), rnd(w, "W"), b, **kw)


def run(sd, vcfg, mel, sites, split=()):
    SITES.clear(); SITES.update(sites)
    SPLIT.clear(); SPLIT.update(split)
    F.conv1d, F.conv_transpose1d = conv1d, convT
    try:
        return model.vocoder_forward(sd, vcfg, mel).squeeze(1)
    finally:
        F.conv1d, F.conv_transpose1d = _conv1d, _convT


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "gen_c1_adaptive_euler"
    g = load_golden(name)
    sd, vcfg = golden_weights(g)
    mel = torch.from_numpy(g["ref_mel"])
    ref = model.vocoder_forward({k: v.double() for k, v in sd.items()}, vcfg, mel.double()).squeeze(1).float()
    for sites, split in [("", ""), ("A", ""), ("W", ""), ("AW", ""), ("X", ""), ("S", ""), ("T", ""), ("AWXST", ""),
                         ("AWXST", "A"), ("AWXST", "AX"), ("AWXST", "AXS"), ("AWXST", "W"), ("AWXST", "AW"), ("AWXST", "AWXS")]:
        out = run(sd, vcfg, mel, set(sites), set(split))
        print(f"sites {sites or '-':6s} split {split or '-':5s}: SNR {snr_db(ref, out):6.1f} dB  LSD {lsd_db(ref, out):.4f} dB")
