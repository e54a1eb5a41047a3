"""ORACLE (test infrastructure, not product code) -- networks of the FLowHigh hot path.

Functional, state_dict-driven restatement in plain torch-CPU fp32 of
  * the vector-field network        FLowHigh.forward            models/flow.py:180-274
  * the fixed-grid ODE solvers      torchdiffeq.odeint          call site cfm_superresolution.py:243
  * the CFM sampler                 ConditionalFlowMatcherWrapper.sample   cfm_superresolution.py:162-284
  * the BigVGAN generator           BigVGAN.forward             models/bigvgan/models.py:172-194
Pinned against the reference itself: tests/golden/make_golden.py loads the very same
state_dict into the unmodified reference modules (imported from /root/reference with stub
modules for the four packages missing offline) and stores the reference outputs as fixtures;
tests/test_oracle_cpu.py replays them through this file.

torchdiffeq (pyproject.toml:14 `torchdiffeq>=0.2.3`, not installed offline) is restated from
its published fixed-grid solvers: the grid is exactly `t`; euler: y1 = y0 + dt f(t0,y0);
midpoint: y1 = y0 + dt f(t0+dt/2, y0 + dt/2 f(t0,y0)).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

FH = "flowhigh."
VOC = "flowhigh.audio_enc_dec.vocoder."


# --------------------------------------------------------------------------------------
# backbone
# --------------------------------------------------------------------------------------
def time_embedding(sd: Dict[str, torch.Tensor], t: torch.Tensor) -> torch.Tensor:
    """pos_emb.py:22-26 + flow.py:92-96.  t [B] -> [B, dim]."""
    w = sd[FH + "sinu_pos_emb.0.weights"]
    freqs = t[:, None] * w[None, :] * 2 * math.pi
    four = torch.cat((freqs.sin(), freqs.cos()), dim=-1)
    return F.silu(F.linear(four, sd[FH + "sinu_pos_emb.1.weight"], sd[FH + "sinu_pos_emb.1.bias"]))


def _ada_rmsnorm(sd, prefix, x, temb):
    # transformer.py:82-88
    dim = x.shape[-1]
    normed = F.normalize(x, dim=-1) * dim ** 0.5
    gamma = F.linear(temb, sd[prefix + "to_gamma.weight"], sd[prefix + "to_gamma.bias"])[:, None, :]
    beta = F.linear(temb, sd[prefix + "to_beta.weight"], sd[prefix + "to_beta.bias"])[:, None, :]
    return normed * gamma + beta


def _rotate_half(x):
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def _attention(sd, prefix, x, heads, rot, scale=10.0):
    # attend.py:173-189 (math path :123-137), qk-norm :144-151, rotary pos_emb.py:45-60
    B, N, _ = x.shape
    qkv = F.linear(x, sd[prefix + "to_qkv.weight"])
    q, k, v = [t.reshape(B, N, heads, -1).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
    dh = q.shape[-1]
    q = F.normalize(q, dim=-1) * sd[prefix + "q_norm.gamma"] * dh ** 0.5
    k = F.normalize(k, dim=-1) * sd[prefix + "k_norm.gamma"] * dh ** 0.5
    q = q * rot.cos() + _rotate_half(q) * rot.sin()
    k = k * rot.cos() + _rotate_half(k) * rot.sin()
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, N, heads * dh)
    return F.linear(out, sd[prefix + "to_out.weight"])


def vector_field(sd: Dict[str, torch.Tensor], x: torch.Tensor, cond: torch.Tensor, t: torch.Tensor,
                 depth: int = 2, heads: int = 16, null_cond: bool = False,
                 skip_connect_scale: float = 2 ** -0.5) -> torch.Tensor:
    """FLowHigh.forward (inference branch).  x, cond [B,N,256]; t 0-dim or [B]."""
    B, N, _ = x.shape
    if t.ndim == 0:
        t = t.repeat(B)
    if null_cond:  # flow.py:224-230 with cond_drop_prob = 1
        cond = sd[FH + "null_cond"].expand_as(cond)
    emb = F.linear(torch.cat((x, cond), dim=-1), sd[FH + "to_embed.weight"], sd[FH + "to_embed.bias"])
    wc = sd[FH + "conv_embed.dw_conv1d.0.weight"]
    pos = F.conv1d(emb.transpose(1, 2), wc, sd[FH + "conv_embed.dw_conv1d.0.bias"],
                   padding=wc.shape[-1] // 2, groups=wc.shape[0])
    h = F.gelu(pos).transpose(1, 2) + emb
    temb = time_embedding(sd, t)
    if FH + "convnext.0.gamma" in sd:  # architecture == 'convnext' (flow.py:247-253, convnext.py:46-63,86-93)
        i = 0
        while FH + f"convnext.{i}.gamma" in sd:
            p = FH + f"convnext.{i}."
            y = F.conv1d(h.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3,
                         groups=h.shape[-1]).transpose(1, 2)
            scale = F.linear(temb, sd[p + "norm.scale.weight"], sd[p + "norm.scale.bias"])[:, None, :]
            shift = F.linear(temb, sd[p + "norm.shift.weight"], sd[p + "norm.shift.bias"])[:, None, :]
            y = F.layer_norm(y, (y.shape[-1],), eps=1e-6) * scale + shift
            y = F.linear(F.gelu(F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"])), sd[p + "pwconv2.weight"],
                         sd[p + "pwconv2.bias"])
            h = h + sd[p + "gamma"] * y
            i += 1
        h = F.layer_norm(h, (h.shape[-1],), sd[FH + "final_layer_norm.weight"], sd[FH + "final_layer_norm.bias"], eps=1e-6)
        return F.linear(h, sd[FH + "to_pred.weight"])
    inv_freq = sd[FH + "transformer.rotary_emb.inv_freq"]
    pos_idx = torch.arange(N, dtype=inv_freq.dtype)
    fr = torch.einsum("i,j->ij", pos_idx, inv_freq)
    rot = torch.cat((fr, fr), dim=-1)
    skips = []
    for l in range(depth):
        p = FH + f"transformer.layers.{l}."
        if p + "0.weight" in sd:  # transformer.py:213-218 (use_unet_skip_connection): second-half layers pop a skip
            h = F.linear(torch.cat((h, skips.pop() * skip_connect_scale), dim=-1), sd[p + "0.weight"], sd[p + "0.bias"])
        else:
            skips.append(h)
        a = _ada_rmsnorm(sd, p + "2.", h, temb)
        h = _attention(sd, p + "3.", a, heads, rot) + h
        f = _ada_rmsnorm(sd, p + "4.", h, temb)
        u = F.linear(f, sd[p + "5.0.weight"], sd[p + "5.0.bias"])
        xg, gate = u.chunk(2, dim=-1)
        h = F.linear(F.gelu(gate) * xg, sd[p + "5.3.weight"], sd[p + "5.3.bias"]) + h
    h = F.normalize(h, dim=-1) * h.shape[-1] ** 0.5 * sd[FH + "transformer.final_norm.gamma"]
    return F.linear(h, sd[FH + "to_pred.weight"])


def vector_field_cfg(sd, x, cond, t, cond_scale=1.0, **kw):
    """flow.py:165-178."""
    v = vector_field(sd, x, cond, t, **kw)
    if cond_scale == 1.0:
        return v
    vn = vector_field(sd, x, cond, t, null_cond=True, **kw)
    return vn + (v - vn) * cond_scale


def odeint_fixed(fn: Callable, y0: torch.Tensor, t: torch.Tensor, method: str) -> torch.Tensor:
    """Final state of torchdiffeq's fixed-grid solver on grid t."""
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        if method == "euler":
            y = y + dt * fn(t0, y)
        elif method == "midpoint":
            half = 0.5 * dt
            ymid = y + fn(t0, y) * half
            y = y + dt * fn(t0 + half, ymid)
        else:
            raise ValueError(method)
    return y


def mel_cutoff_bin(mel_logclip: torch.Tensor, percentile: float = 0.9995) -> int:
    """cfm_superresolution.py:134-144 applied to exp(mel) of one clip [N,256]."""
    mag = torch.abs(torch.exp(mel_logclip))
    energy = torch.cumsum(torch.sum(mag, dim=0), dim=0)
    thr = energy[-1] * percentile
    n = energy.shape[0]
    for i in range(1, n):
        if energy[-i] < thr:
            return n - i
    return 0


def cfm_prior(cond_mel, eps, cfm_method, sigma):
    """cfm_superresolution.py:176-183,219-237 (std_1/std_2 quirk: both reset to 1, sigma)."""
    if cfm_method == "basic_cfm":
        return eps
    low = cond_mel * 1.0 + eps * sigma
    if cfm_method in ("independent_cfm_adaptive", "independent_cfm_constant"):
        return low
    if cfm_method == "independent_cfm_mix":
        y0 = torch.zeros_like(eps)
        for i in range(eps.shape[0]):
            c = mel_cutoff_bin(cond_mel[i])
            y0[i][..., c:] = eps[i][..., c:]
            y0[i][..., :c] = low[i][..., :c]
        return y0
    raise ValueError(cfm_method)


def cfm_sample_mel(sd, cond_mel, eps, *, steps, ode_method, cfm_method, sigma, cond_scale=1.0,
                   mel_pp=False, depth=2, heads=16, adaptive=None):
    y0 = cfm_prior(cond_mel, eps, cfm_method, sigma)
    t = torch.linspace(0, 1, steps + 1)
    fn = lambda tt, yy: vector_field_cfg(sd, yy, cond_mel, tt, cond_scale=cond_scale, depth=depth, heads=heads)
    if adaptive is not None:  # use_torchode (cfm_superresolution.py:259-276): every clip is its own problem instance
        from . import ode_adaptive
        fb = lambda b, tt, yy: vector_field_cfg(sd, yy, cond_mel[b: b + 1], tt.to(yy.dtype), cond_scale=cond_scale,
                                                depth=depth, heads=heads)
        out, _ = ode_adaptive.odeint_adaptive_batch(fb, y0, float(t[0]), float(t[-1]), **adaptive)
    else:
        out = odeint_fixed(fn, y0, t, ode_method)
    if mel_pp:  # cfm_superresolution.py:146-152,278-279
        res = torch.zeros_like(out)
        for i in range(out.shape[0]):
            c = mel_cutoff_bin(cond_mel[i])
            res[i][..., c:] = out[i][..., c:]
            res[i][..., :c] = cond_mel[i][..., :c]
        out = res
    return out


# --------------------------------------------------------------------------------------
# vocoder
# --------------------------------------------------------------------------------------
def aa_activation(x: torch.Tensor, alpha: torch.Tensor, beta: Optional[torch.Tensor], filt_up: torch.Tensor,
                  filt_down: torch.Tensor, logscale: bool) -> torch.Tensor:
    """Activation1d (act.py:23-28): 2x Kaiser-sinc upsample -> Snake/SnakeBeta -> 2x downsample.

    x [B,C,L].  UpSample1d resample.py:25-33 (replicate pad 5, conv_transpose stride 2, x2,
    crop 15/15); activations.py:48-59,107-119; LowPassFilter1d filter.py:86-94 (replicate
    pad 5/6, stride-2 depthwise conv).
    """
    C = x.shape[1]
    xp = F.pad(x, (5, 5), mode="replicate")
    u = 2 * F.conv_transpose1d(xp, filt_up.expand(C, -1, -1), stride=2, groups=C)
    u = u[..., 15:-15]
    a = alpha[None, :, None]
    b = a if beta is None else beta[None, :, None]
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    s = u + (1.0 / (b + 1e-9)) * torch.sin(u * a) ** 2
    sp = F.pad(s, (5, 6), mode="replicate")
    return F.conv1d(sp, filt_down.expand(C, -1, -1), stride=2, groups=C)


def _act(sd, prefix, x, logscale):
    return aa_activation(x, sd[prefix + "act.alpha"], sd.get(prefix + "act.beta"),
                         sd[prefix + "upsample.filter"], sd[prefix + "downsample.lowpass.filter"], logscale)


def vocoder_forward(sd: Dict[str, torch.Tensor], vcfg, mel: torch.Tensor) -> torch.Tensor:
    """mel [B,N,256] -> wave [B,1,480N]  (melvoco.py:114-121 + bigvgan/models.py:172-194)."""
    x = mel.transpose(1, 2)
    x = F.conv1d(x, sd[VOC + "conv_pre.weight"], sd[VOC + "conv_pre.bias"], padding=3)
    nk = len(vcfg.resblock_kernel_sizes)
    ls = vcfg.snake_logscale
    for s, (u, k) in enumerate(zip(vcfg.upsample_rates, vcfg.upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, sd[VOC + f"ups.{s}.0.weight"], sd[VOC + f"ups.{s}.0.bias"],
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j, (kk, dil) in enumerate(zip(vcfg.resblock_kernel_sizes, vcfg.resblock_dilation_sizes)):
            p = VOC + f"resblocks.{s * nk + j}."
            y = x
            if vcfg.resblock == "1":  # AMPBlock1 models.py:63-72
                for i, d in enumerate(dil):
                    xt = _act(sd, p + f"activations.{2 * i}.", y, ls)
                    xt = F.conv1d(xt, sd[p + f"convs1.{i}.weight"], sd[p + f"convs1.{i}.bias"],
                                  dilation=d, padding=(kk * d - d) // 2)
                    xt = _act(sd, p + f"activations.{2 * i + 1}.", xt, ls)
                    xt = F.conv1d(xt, sd[p + f"convs2.{i}.weight"], sd[p + f"convs2.{i}.bias"],
                                  padding=(kk - 1) // 2)
                    y = xt + y
            else:  # AMPBlock2 models.py:111-117
                for i, d in enumerate(dil):
                    xt = _act(sd, p + f"activations.{i}.", y, ls)
                    xt = F.conv1d(xt, sd[p + f"convs.{i}.weight"], sd[p + f"convs.{i}.bias"],
                                  dilation=d, padding=(kk * d - d) // 2)
                    y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = _act(sd, VOC + "activation_post.", x, ls)
    x = F.conv1d(x, sd[VOC + "conv_post.weight"], sd[VOC + "conv_post.bias"], padding=3)
    return torch.tanh(x)
