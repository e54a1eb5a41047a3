"""flowhigh_b200 -- B200-native (sm_100a) inference engine for FLowHigh audio super-resolution.

Drop-in for the reference's `FlowHighSR.from_pretrained(...)` / `model.generate(wav, sr_in, target_sr)`
path; every stage runs as a hand-written CUDA kernel behind a C ABI (include/flowhigh_b200.h).
"""
from .config import BackboneConfig, MelConfig, VocoderConfig
from .flowhighsr import FLowHigh, FlowHighSR, MelVoco, PostProcessing

__all__ = ["FlowHighSR", "FLowHigh", "MelVoco", "PostProcessing", "VocoderConfig", "BackboneConfig", "MelConfig"]
