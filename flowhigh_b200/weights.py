"""state_dict layout of the reference `FlowHighSR` module and deterministic random init.

Key layout follows what `FlowHighSR(...).state_dict()` produces in the reference
(flow.py:92-142, transformer.py:141-165, attend.py:165-171,
bigvgan/models.py:126-170 after `remove_weight_norm()`), so a checkpoint's
`['model']` dict (flowhighsr.py:131-135) loads without renaming.

`random_state_dict` is the "random-init weights of the named architecture" used by
parity tests and the bench: no checkpoints are available offline.  Identity-initialised
parameters of the reference (adaptive-norm projections, Snake alpha/beta, q/k-norm
gamma, null_cond) are perturbed so that those sub-paths are exercised.
The generator is numpy PCG64 keyed by (seed, key index) so every machine with the same
numpy builds identical tensors; nothing here depends on torch's RNG.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch

from .config import BackboneConfig, VocoderConfig

FH = "flowhigh."
VOC = "flowhigh.audio_enc_dec.vocoder."


def kaiser_sinc_filter12() -> np.ndarray:
    """12-tap Kaiser-windowed sinc low-pass, cutoff 0.25, half-width 0.3.

    Restates alias_free_torch/filter.py:28-57 for (cutoff=0.5/2, half_width=0.6/2,
    kernel_size=12): beta from A = 2.285*(6-1)*pi*(4*0.3)+7.95, window
    kaiser(12, beta) (symmetric), taps 2*cutoff*sinc(2*cutoff*t), normalised to sum 1.
    """
    ksz, cutoff, half_width = 12, 0.25, 0.3
    half = ksz // 2
    delta_f = 4 * half_width
    A = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    # torch.kaiser_window(periodic=False) in float32
    window = torch.kaiser_window(ksz, periodic=False, beta=beta, dtype=torch.float32)
    time = torch.arange(-half, half, dtype=torch.float32) + 0.5
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.numpy().astype(np.float32)


def state_dict_spec(bcfg: BackboneConfig, vcfg: VocoderConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (key, shape, init-kind) list in the reference's registration order."""
    D, Din, H, Dh = bcfg.dim, bcfg.dim_in, bcfg.heads, bcfg.dim_head
    inner = bcfg.ff_inner
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    add = spec.append
    # FLowHigh registers audio_enc_dec first (flow.py:88), then sinu_pos_emb, to_embed, ...
    # but nn.Module.state_dict lists parameters of a module before its children, so
    # `null_cond` comes first, then children in registration order.
    add((FH + "null_cond", (Din,), "small"))
    # --- vocoder (audio_enc_dec is registered before the other children, flow.py:88)
    C0 = vcfg.upsample_initial_channel
    add((VOC + "conv_pre.bias", (C0,), "bias"))
    add((VOC + "conv_pre.weight", (C0, vcfg.num_mels, 7), "conv"))
    for s, (u, k) in enumerate(zip(vcfg.upsample_rates, vcfg.upsample_kernel_sizes)):
        cin, cout = C0 // (2 ** s), C0 // (2 ** (s + 1))
        add((VOC + f"ups.{s}.0.bias", (cout,), "bias"))
        add((VOC + f"ups.{s}.0.weight", (cin, cout, k), "convT"))
    beta = vcfg.activation == "snakebeta"

    def act_keys(prefix, ch):
        add((prefix + "act.alpha", (ch,), "snake"))
        if beta:
            add((prefix + "act.beta", (ch,), "snake"))
        add((prefix + "upsample.filter", (1, 1, 12), "aafilter"))
        add((prefix + "downsample.lowpass.filter", (1, 1, 12), "aafilter"))

    ch = C0
    for s in range(vcfg.num_stages):
        ch = vcfg.stage_channels(s)
        for j, (k, dil) in enumerate(zip(vcfg.resblock_kernel_sizes, vcfg.resblock_dilation_sizes)):
            r = s * vcfg.num_kernels + j
            p = VOC + f"resblocks.{r}."
            if vcfg.resblock == "1":
                for i in range(len(dil)):
                    add((p + f"convs1.{i}.bias", (ch,), "bias"))
                    add((p + f"convs1.{i}.weight", (ch, ch, k), "resconv"))
                for i in range(len(dil)):
                    add((p + f"convs2.{i}.bias", (ch,), "bias"))
                    add((p + f"convs2.{i}.weight", (ch, ch, k), "resconv"))
                nact = 2 * len(dil)
            else:
                for i in range(len(dil)):
                    add((p + f"convs.{i}.bias", (ch,), "bias"))
                    add((p + f"convs.{i}.weight", (ch, ch, k), "resconv"))
                nact = len(dil)
            for a in range(nact):
                act_keys(p + f"activations.{a}.", ch)
    act_keys(VOC + "activation_post.", ch)
    add((VOC + "conv_post.bias", (1,), "bias"))
    add((VOC + "conv_post.weight", (1, ch, 7), "convpost"))
    # --- backbone
    add((FH + "sinu_pos_emb.0.weights", (D // 2,), "normal1"))
    add((FH + "sinu_pos_emb.1.weight", (D, D), "linear"))
    add((FH + "sinu_pos_emb.1.bias", (D,), "bias"))
    add((FH + "to_embed.weight", (D, 2 * Din), "linear"))
    add((FH + "to_embed.bias", (D,), "bias"))
    add((FH + "conv_embed.dw_conv1d.0.weight", (D, 1, bcfg.conv_pos_kernel), "dwconv"))
    add((FH + "conv_embed.dw_conv1d.0.bias", (D,), "bias"))
    if bcfg.architecture == "convnext":  # flow.py:124-139, convnext.py:9-93; a module's own parameter (gamma) comes first
        I = D * bcfg.convnext_mult
        for i in range(bcfg.convnext_layers):
            p = FH + f"convnext.{i}."
            add((p + "gamma", (D,), "gamma1"))
            add((p + "dwconv.weight", (D, 1, 7), "dwconv"))
            add((p + "dwconv.bias", (D,), "bias"))
            add((p + "norm.scale.weight", (D, D), "adaw"))   # zero / one initialised in the reference (convnext.py:79-82):
            add((p + "norm.scale.bias", (D,), "gamma1"))     # perturbed here so that the time conditioning is exercised
            add((p + "norm.shift.weight", (D, D), "adaw"))
            add((p + "norm.shift.bias", (D,), "small"))
            add((p + "pwconv1.weight", (I, D), "linear"))
            add((p + "pwconv1.bias", (I,), "bias"))
            add((p + "pwconv2.weight", (D, I), "linear"))
            add((p + "pwconv2.bias", (D,), "bias"))
        add((FH + "final_layer_norm.weight", (D,), "gamma1"))
        add((FH + "final_layer_norm.bias", (D,), "small"))
        add((FH + "to_pred.weight", (Din, D), "linear"))
        return spec
    for l in range(bcfg.depth):
        p = FH + f"transformer.layers.{l}."
        if bcfg.use_unet_skip_connection and l + 1 > bcfg.depth // 2:  # transformer.py:148-151: registered first
            add((p + "0.weight", (D, 2 * D), "linear"))
            add((p + "0.bias", (D,), "bias"))
        for idx in (2, 4):
            if idx == 4:
                # attention (index 3) is registered between the two norms
                add((p + "3.q_norm.gamma", (H, 1, Dh), "gamma1"))
                add((p + "3.k_norm.gamma", (H, 1, Dh), "gamma1"))
                add((p + "3.to_qkv.weight", (3 * H * Dh, D), "linear"))
                add((p + "3.to_out.weight", (D, H * Dh), "linear"))
            add((p + f"{idx}.to_gamma.weight", (D, D), "adaw"))
            add((p + f"{idx}.to_gamma.bias", (D,), "gamma1"))
            add((p + f"{idx}.to_beta.weight", (D, D), "adaw"))
            add((p + f"{idx}.to_beta.bias", (D,), "small"))
        add((p + "5.0.weight", (2 * inner, D), "linear"))
        add((p + "5.0.bias", (2 * inner,), "bias"))
        add((p + "5.3.weight", (D, inner), "linear"))
        add((p + "5.3.bias", (D,), "bias"))
    add((FH + "transformer.rotary_emb.inv_freq", (Dh // 2,), "inv_freq"))
    add((FH + "transformer.final_norm.gamma", (D,), "gamma1"))
    add((FH + "to_pred.weight", (Din, D), "linear"))
    return spec


def _fan_in(shape, kind):
    if kind == "convT":
        # ConvTranspose1d weight is (C_in, C_out, k); each output sees C_in * ceil(k/stride) taps
        return shape[0] * shape[2]
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def random_state_dict(bcfg: BackboneConfig = BackboneConfig(), vcfg: VocoderConfig = VocoderConfig(),
                      seed: int = 0, vocoder_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 state_dict with the reference key layout."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    filt = kaiser_sinc_filter12()
    for idx, (key, shape, kind) in enumerate(state_dict_spec(bcfg, vcfg)):
        rng = np.random.Generator(np.random.PCG64([seed, idx]))
        if kind == "aafilter":
            a = filt.reshape(1, 1, 12).copy()
        elif kind == "inv_freq":
            dh = shape[0] * 2
            a = (1.0 / (bcfg.rotary_theta ** (torch.arange(0, dh, 2).float() / dh))).numpy()
        elif kind in ("linear", "conv", "dwconv"):
            a = rng.standard_normal(shape, dtype=np.float32) * (1.0 / math.sqrt(_fan_in(shape, kind)))
        elif kind == "resconv":
            a = rng.standard_normal(shape, dtype=np.float32) * (0.6 * vocoder_gain / math.sqrt(_fan_in(shape, kind)))
        elif kind == "convT":
            # every output sample sees ~k/stride taps of each input channel
            a = rng.standard_normal(shape, dtype=np.float32) * (1.6 * vocoder_gain / math.sqrt(_fan_in(shape, kind)))
        elif kind == "convpost":
            a = rng.standard_normal(shape, dtype=np.float32) * (1.0 / math.sqrt(_fan_in(shape, kind)))
        elif kind == "adaw":
            a = rng.standard_normal(shape, dtype=np.float32) * (0.3 / math.sqrt(shape[1]))
        elif kind == "bias":
            a = rng.standard_normal(shape, dtype=np.float32) * 0.02
        elif kind == "small":
            a = rng.standard_normal(shape, dtype=np.float32) * 0.1
        elif kind == "gamma1":
            a = 1.0 + rng.standard_normal(shape, dtype=np.float32) * 0.1
        elif kind == "normal1":
            a = rng.standard_normal(shape, dtype=np.float32)
        elif kind == "snake":
            if vcfg.snake_logscale:
                a = rng.standard_normal(shape, dtype=np.float32) * 0.3
            else:
                a = np.abs(1.0 + rng.standard_normal(shape, dtype=np.float32) * 0.25) + 0.05
        else:
            raise AssertionError(kind)
        out[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def fold_weight_norm(generator_sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Folds legacy `weight_g`/`weight_v` pairs of a raw BigVGAN `['generator']` dict.

    Same arithmetic as torch's remove_weight_norm with the default dim=0 used at
    bigvgan/models.py:27-43,134,143,165: w = g * v / ||v|| with the norm over every
    dim but 0 (for ConvTranspose1d dim 0 is C_in, and that is what the reference does).
    """
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in generator_sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            vv = generator_sd[base + "weight_v"].float()
            g = v.float()
            norm = vv.reshape(vv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (vv.dim() - 1)))
            out[base + "weight"] = vv * (g / norm)
        elif k.endswith("weight_v"):
            continue
        else:
            out[k] = v
    return out
