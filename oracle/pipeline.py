"""ORACLE (test infrastructure, not product code) -- end-to-end `generate` restatement.

Follows FlowHighSR.generate (flowhighsr.py:51-102) -> sample (cfm_superresolution.py:162-284)
-> post_processing (postprocessing.py:18-41), one clip at a time (the reference is batch-1,
SURVEY.md F8).  The CFM noise epsilon is an explicit argument: the reference draws it with
torch.randn_like; the parity tests draw it once and hand the same tensor to both sides.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import dsp, model


@dataclass
class OracleFlowHigh:
    sd: Dict[str, torch.Tensor]
    vcfg: object
    depth: int = 2
    heads: int = 16
    sigma: float = 0.0
    cfm_method: str = "basic_cfm"
    ode_method: str = "midpoint"
    upsampling_method: str = "scipy"   # 'librosa' = the soxr_hq branch (flowhighsr.py:74-80; dsp.resample_soxr_hq)
    use_torchode: bool = False         # adaptive Tsit5 + IntegralController (cfm_superresolution.py:259-276; ode_adaptive.py)
    torchode_method: str = "tsit5"
    ode_atol: float = 1e-5
    ode_rtol: float = 1e-5

    def sample(self, cond_audio: torch.Tensor, eps: torch.Tensor, time_steps: int, cfm_method: Optional[str] = None,
               cond_scale: float = 1.0, mel_pp: bool = False, decode: bool = True):
        cfm_method = cfm_method or self.cfm_method
        cond_mel = dsp.encode_logmel(cond_audio)
        adaptive = dict(method=self.torchode_method, atol=self.ode_atol, rtol=self.ode_rtol) if self.use_torchode else None
        mel = model.cfm_sample_mel(self.sd, cond_mel, eps, steps=time_steps, ode_method=self.ode_method,
                                   cfm_method=cfm_method, sigma=self.sigma, cond_scale=cond_scale,
                                   mel_pp=mel_pp, depth=self.depth, heads=self.heads, adaptive=adaptive)
        if not decode:
            return mel
        return model.vocoder_forward(self.sd, self.vcfg, mel)

    @torch.no_grad()
    def generate(self, audio: np.ndarray, sr: int, eps: torch.Tensor, target_sampling_rate: int = 48000,
                 timestep: int = 1, return_stages: bool = False):
        cond = dsp.preprocess_audio(audio, sr, target_sampling_rate,
                                    method="scipy" if self.upsampling_method == "scipy" else "soxr_hq")
        cond = torch.from_numpy(np.ascontiguousarray(cond)).float().unsqueeze(0)
        hr = self.sample(cond, eps, timestep).squeeze(1)
        out = dsp.postprocess(hr, cond, cond.shape[-1])
        if return_stages:
            return out, {"cond": cond, "vocoder_out": hr}
        return out
