"""Static configuration of the FLowHigh hot path.

Backbone hyper-parameters are the ones hard-coded by the reference loader
(/root/reference/src/flowhigh/flowhighsr.py:112-129 -> models/flow.py:55-75):
dim 1024, depth 2, 16 heads x 64, ff_mult 4 (inner = int(1024*4*2/3) = 2730),
conv-pos-embed kernel 31, qk-norm with scale 10, rotary theta 50000.

The vocoder is config driven (models/bigvgan/models.py:126-170 reads an
AttrDict).  The real `bigvgan_48khz_256band.json` is fetched from the HF hub at
run time and is not in the reference tree, so `VocoderConfig.assumed_48k()` is
an ASSUMPTION that only satisfies the hard constraints:
num_mels = 256, prod(upsample_rates) = 480 = hop.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import List, Sequence


@dataclass(frozen=True)
class MelConfig:
    # models/melvoco.py:17-31
    n_fft: int = 2048
    win_length: int = 2048
    hop_length: int = 480
    n_mels: int = 256
    sampling_rate: int = 48000
    f_min: float = 20.0
    f_max: float = 24000.0

    @property
    def n_freq(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def pad(self) -> int:
        return (self.n_fft - self.hop_length) // 2  # 784, melvoco.py:74


@dataclass(frozen=True)
class BackboneConfig:
    dim_in: int = 256
    dim: int = 1024
    depth: int = 2
    heads: int = 16
    dim_head: int = 64
    ff_mult: int = 4
    conv_pos_kernel: int = 31
    qk_norm_scale: float = 10.0   # attend.py:155
    rotary_theta: float = 50000.0  # pos_emb.py:34
    # transformer.py:123-124,150-153: U-Net style skip connections (OFF in the shipped model): layers of the
    # second half combine [x, skip * scale] with a Linear(2 dim -> dim) before the attention block
    use_unet_skip_connection: bool = False
    # flow.py:124-139: alternative vector-field network, 8 ConvNeXt blocks (dwconv7 -> AdaLayerNorm -> Linear 3x -> GELU ->
    # Linear -> layer scale -> residual) + final LayerNorm instead of the transformer (not used by the shipped checkpoint)
    architecture: str = "transformer"
    convnext_layers: int = 8
    convnext_mult: int = 3
    skip_connect_scale: float = 2 ** -0.5

    @property
    def ff_inner(self) -> int:
        return int(self.dim * self.ff_mult * 2 / 3)  # transformer.py:98 -> 2730


@dataclass(frozen=True)
class VocoderConfig:
    resblock: str = "1"
    upsample_rates: Sequence[int] = (5, 4, 3, 2, 2, 2)
    upsample_kernel_sizes: Sequence[int] = (11, 8, 7, 4, 4, 4)
    upsample_initial_channel: int = 1536
    resblock_kernel_sizes: Sequence[int] = (3, 7, 11)
    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    activation: str = "snakebeta"
    snake_logscale: bool = True
    num_mels: int = 256

    # ------------------------------------------------------------------
    @staticmethod
    def assumed_48k() -> "VocoderConfig":
        """BigVGAN-large-like 48 kHz / 256-band layout (ASSUMPTION)."""
        return VocoderConfig()

    @staticmethod
    def tiny(resblock: str = "1", activation: str = "snakebeta", logscale: bool = True) -> "VocoderConfig":
        """Small config with the same topology; used by parity tests and goldens."""
        if resblock == "1":
            return VocoderConfig(resblock="1", upsample_rates=(5, 4, 3, 2, 2, 2),
                                 upsample_kernel_sizes=(11, 8, 7, 4, 4, 4),
                                 upsample_initial_channel=256,
                                 resblock_kernel_sizes=(3, 7, 11),
                                 resblock_dilation_sizes=((1, 3, 5),) * 3,
                                 activation=activation, snake_logscale=logscale)
        return VocoderConfig(resblock="2", upsample_rates=(10, 6, 2, 2, 2),
                             upsample_kernel_sizes=(20, 12, 4, 4, 4),
                             upsample_initial_channel=128,
                             resblock_kernel_sizes=(3, 5),
                             resblock_dilation_sizes=((1, 2), (2, 6)),
                             activation=activation, snake_logscale=logscale)

    @staticmethod
    def from_json(path) -> "VocoderConfig":
        """Reads the AttrDict-style JSON that init_vocoder.py:10-11 consumes."""
        with open(path) as f:
            h = json.load(f)
        return VocoderConfig(
            resblock=str(h["resblock"]),
            upsample_rates=tuple(h["upsample_rates"]),
            upsample_kernel_sizes=tuple(h["upsample_kernel_sizes"]),
            upsample_initial_channel=int(h["upsample_initial_channel"]),
            resblock_kernel_sizes=tuple(h["resblock_kernel_sizes"]),
            resblock_dilation_sizes=tuple(tuple(d) for d in h["resblock_dilation_sizes"]),
            activation=h.get("activation", "snakebeta"),
            snake_logscale=bool(h.get("snake_logscale", True)),
            num_mels=int(h["num_mels"]),
        )

    def to_attr_json(self) -> dict:
        d = asdict(self)
        d["upsample_rates"] = list(self.upsample_rates)
        d["upsample_kernel_sizes"] = list(self.upsample_kernel_sizes)
        d["resblock_kernel_sizes"] = list(self.resblock_kernel_sizes)
        d["resblock_dilation_sizes"] = [list(x) for x in self.resblock_dilation_sizes]
        return d

    # ------------------------------------------------------------------
    @property
    def num_stages(self) -> int:
        return len(self.upsample_rates)

    @property
    def num_kernels(self) -> int:
        return len(self.resblock_kernel_sizes)

    def stage_channels(self, i: int) -> int:
        """Output channels of upsampler i (models.py:143-144)."""
        return self.upsample_initial_channel // (2 ** (i + 1))

    @property
    def total_upsample(self) -> int:
        p = 1
        for u in self.upsample_rates:
            p *= u
        return p

    def validate(self) -> None:
        if self.total_upsample != 480:
            raise ValueError(f"prod(upsample_rates) must be 480 (= mel hop), got {self.total_upsample}")
        for k, u in zip(self.upsample_kernel_sizes, self.upsample_rates):
            if (k - u) % 2:
                raise ValueError("(kernel - rate) must be even for every upsampler (models.py:143-145)")
        if self.activation not in ("snake", "snakebeta"):
            raise ValueError("activation must be snake|snakebeta (models.py:156-163)")
        if self.resblock not in ("1", "2"):
            raise ValueError("resblock must be '1' or '2'")


CFM_METHODS = ("basic_cfm", "independent_cfm_adaptive", "independent_cfm_constant", "independent_cfm_mix")
