"""CPU suite: the oracle restatement against the committed golden vectors (outputs of the
unmodified reference, tests/golden/make_golden.py), against the installed scipy, and the host
logic / ABI surface of the product.  No GPU work."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from flowhigh_b200 import tables
from flowhigh_b200.config import BackboneConfig, VocoderConfig
from flowhigh_b200.synth import synth_speech
from flowhigh_b200.weights import kaiser_sinc_filter12, random_state_dict, state_dict_spec, fold_weight_norm
from oracle import dsp, model, pipeline
from util import golden_weights, load_golden, vcfg_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ resampler
@pytest.mark.parametrize("sr", [8000, 12000, 16000, 24000, 22050, 44100])
def test_resample_restatement_matches_scipy(sr):
    import scipy.signal
    x = synth_speech(sr // 3 + 11, sr, seed=1)
    for dt, tol in ((np.float32, 1e-6), (np.float64, 1e-12)):
        a = scipy.signal.resample_poly(x.astype(dt), 48000, sr)
        b = dsp.resample_poly(x.astype(dt), 48000, sr)
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.abs(a - b).max() <= tol


@pytest.mark.parametrize("sr", [8000, 12000, 16000, 24000, 22050, 44100])
def test_resample_golden(sr):
    g = load_golden("frontend")
    y = dsp.preprocess_audio(g[f"wav_{sr}"], sr)
    assert np.abs(y - g[f"cond_{sr}"]).max() <= 2e-6


def test_resample_int16_quirk_golden():
    g = load_golden("frontend")
    y = dsp.preprocess_audio(g["wav_int16"], 16000)  # max > 1 -> /32768 in fp64 (flowhighsr.py:62-63)
    assert y.dtype == np.float64
    assert np.abs(y - g["cond_int16"]).max() <= 1e-6


def test_resample_taps_table_matches_oracle():
    for sr in (8000, 12000, 16000, 24000, 22050):
        h, up, down, npp, npr = tables.resample_plan(sr, 48000)
        u2, d2, half, npp2, npr2 = dsp.resample_poly_params(sr, 48000)
        assert (up, down, npp, npr) == (u2, d2, npp2, npr2)
        ho = (dsp.firwin_kaiser_lowpass(2 * half + 1, 1.0 / max(up, down)).astype(np.float32) * np.float32(up))
        assert np.abs(h - ho).max() <= 1e-7


# ------------------------------------------------------------------ mel
def test_mel_filterbank_structure():
    a = dsp.mel_filterbank()
    b = tables.mel_filterbank_dense()
    assert a.shape == b.shape == (256, 1025) and a.dtype == np.float32
    assert np.abs(a - b).max() <= 1e-7
    assert (a >= 0).all() and int((a > 0).sum()) == 2030
    start, length, w, stride = tables.mel_filterbank_sparse()
    dense = np.zeros_like(b)
    for m in range(256):
        dense[m, start[m]: start[m] + length[m]] = w[m, : length[m]]
    assert np.array_equal(dense, b) and length.max() <= 33


def test_mel_filterbank_pinned_against_torchaudio():
    """librosa is not installable offline, so `librosa.filters.mel` (melvoco.py:64-70) is restated; torchaudio's
    independent Slaney implementation pins the restatement and the engine's table (fp32 rounding apart)."""
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(1025, 20.0, 24000.0, 256, 48000, norm="slaney", mel_scale="slaney").T.numpy()
    a = dsp.mel_filterbank()
    b = tables.mel_filterbank_dense()
    assert np.abs(a - fb).max() <= 1e-6 and np.abs(b - fb).max() <= 1e-6  # peak weight 6.2e-2
    # same support: the only disagreement allowed is an edge tap below fp32 resolution of the bin frequencies
    assert int(((a > 0) != (fb > 0)).sum()) <= 2 and a[(a > 0) != (fb > 0)].max(initial=0.0) <= 1e-6


@pytest.mark.parametrize("sr", [8000, 24000])
def test_logmel_golden(sr):
    g = load_golden("frontend")
    a = dsp.encode_logmel(torch.from_numpy(g[f"cond_{sr}"])[None])
    assert a.shape == g[f"logmel_{sr}"].shape
    assert (a - torch.from_numpy(g[f"logmel_{sr}"])).abs().max() <= 1e-5  # same torch build -> identical


def test_aa_filter_constants():
    f = kaiser_sinc_filter12()
    ref = np.array([0.00202896, 0.00938947, -0.02554346, -0.05765738, 0.12857258, 0.44320980, 0.44320980, 0.12857258,
                    -0.05765738, -0.02554346, 0.00938947, 0.00202896], np.float32)  # SURVEY.md A.5 [probe]
    assert np.abs(f - ref).max() <= 1e-7 and abs(f.sum() - 1) < 1e-6


# ------------------------------------------------------------------ networks vs reference goldens
@pytest.mark.parametrize("name", ["gen_c1_adaptive_euler", "gen_basic_midpoint", "gen_basic_euler4"])
def test_generate_golden(name):
    g = load_golden(name)
    sd, vcfg = golden_weights(g)
    eps = torch.from_numpy(g["eps"])
    cond_mel = dsp.encode_logmel(torch.from_numpy(g["ref_cond"]))
    assert (cond_mel - torch.from_numpy(g["ref_cond_mel"])).abs().max() <= 1e-5
    v = model.vector_field(sd, eps, cond_mel, torch.tensor(0.5))
    # two fp32 evaluation orders of the same network: compare with the reference's own distance to fp64
    floor = float(np.abs(g["ref_vfield_t05"] - g["f64_vfield_t05"]).max())
    assert float((v - torch.from_numpy(g["f64_vfield_t05"])).abs().max()) <= 3 * floor + 1e-5
    o = pipeline.OracleFlowHigh(sd, vcfg, sigma=float(g["sigma"]), cfm_method=str(g["cfm_method"]),
                                ode_method=str(g["ode_method"]))
    out, st = o.generate(g["wav"], int(g["sr"]), eps, timestep=int(g["steps"]), return_stages=True)
    assert np.abs(st["cond"].numpy() - g["ref_cond"]).max() <= 2e-6
    floor = float(np.abs(g["ref_final"] - g["f64_final"]).max())
    assert float((out - torch.from_numpy(g["f64_final"])).abs().max()) <= 3 * floor + 1e-5
    # vocoder alone, from the reference's own mel: tight
    voc = model.vocoder_forward(sd, vcfg, torch.from_numpy(g["ref_mel"])).squeeze(1)
    assert (voc - torch.from_numpy(g["ref_vocoder"])).abs().max() <= 2e-5


@pytest.mark.parametrize("name", ["voc_resblock2_snake", "voc_resblock1_snakebeta"])
def test_vocoder_golden(name):
    g = load_golden(name)
    sd, vcfg = golden_weights(g)
    out = model.vocoder_forward(sd, vcfg, torch.from_numpy(g["mel"]))
    assert out.shape == g["ref_vocoder"].shape
    assert (out - torch.from_numpy(g["ref_vocoder"])).abs().max() <= 2e-5


def test_postprocess_properties():
    torch.manual_seed(0)
    T = 480 * 20 + 123  # T % 480 != 0: pred is shorter than src (postprocessing.py:30-34)
    src = torch.from_numpy(synth_speech(T, 48000, 3))[None]
    pred = torch.randn(1, 480 * 20) * 0.1
    out = dsp.postprocess(pred, src, T)
    assert out.shape == (1, T) and abs(float(out.abs().max()) - 0.99) < 1e-6
    spec = dsp._stft_center_zero(src)
    cr = dsp.cutoff_index(spec)
    assert 1 <= cr <= 1024
    # degenerate: all energy in bin 0 -> loop never tests index 0 -> returns 0
    e = torch.zeros(1, 1025, 5, dtype=torch.complex64)
    e[:, 0] = 1
    assert dsp.cutoff_index(e) == 0


# ------------------------------------------------------------------ host logic
def test_state_dict_layout_and_determinism():
    for vcfg in (VocoderConfig.tiny(), VocoderConfig.tiny(resblock="2", activation="snake", logscale=False)):
        a = random_state_dict(BackboneConfig(), vcfg, seed=3)
        b = random_state_dict(BackboneConfig(), vcfg, seed=3)
        assert list(a) == [k for k, _, _ in state_dict_spec(BackboneConfig(), vcfg)]
        assert all(torch.equal(a[k], b[k]) for k in a)
    n = len(state_dict_spec(BackboneConfig(), VocoderConfig.assumed_48k()))
    assert n == 711


def test_fold_weight_norm_matches_torch():
    torch.manual_seed(0)
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 4, 3))
    convt = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(6, 4, 4, 2))
    gen = {"a." + k: v.detach().clone() for k, v in conv.state_dict().items()}
    gen.update({"b." + k: v.detach().clone() for k, v in convt.state_dict().items()})
    folded = fold_weight_norm(gen)
    torch.nn.utils.remove_weight_norm(conv)
    torch.nn.utils.remove_weight_norm(convt)
    assert torch.allclose(folded["a.weight"], conv.weight, atol=1e-6)
    assert torch.allclose(folded["b.weight"], convt.weight, atol=1e-6)


def test_vocoder_config_validation():
    with pytest.raises(ValueError):
        VocoderConfig(upsample_rates=(4, 4, 4, 4), upsample_kernel_sizes=(8, 8, 8, 8)).validate()
    with pytest.raises(ValueError):
        VocoderConfig(upsample_rates=(5, 4, 3, 2, 2, 2), upsample_kernel_sizes=(10, 8, 7, 4, 4, 4)).validate()
    VocoderConfig.assumed_48k().validate()


def test_api_surface_cpu():
    from flowhigh_b200 import FlowHighSR
    m = FlowHighSR.from_random(VocoderConfig.tiny(), device="cpu")
    assert m.cfm_method == "basic_cfm" and m.odeint_kwargs["method"] == "midpoint" and m.sigma == 0.0  # F4
    m.set_cfm_method("independent_cfm_adaptive")
    assert m.cfm_method == "independent_cfm_adaptive"
    sd = m.state_dict()
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):  # no CPU fallback
        m.generate(np.zeros(2000, np.float32), 16000)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly the symbols include/flowhigh_b200.h declares."""
    from flowhigh_b200 import _lib
    from flowhigh_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "flowhigh_b200.h")).read()
    declared = set(re.findall(r"\b(fh_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.fh_version() == 1
    lib.fh_tc_packed_weight_bytes.restype = ctypes.c_int64
    assert lib.fh_tc_packed_weight_bytes(32, 48, 3, 1, 48) == 1 * 1 * 2 * 3 * 48 * 32


def test_unet_skip_variant_pinned_against_reference_transformer():
    """Transformer(use_unet_skip_connection=True).  The golden was produced by the unmodified
    reference Transformer swapped into the reference FLowHigh (oracle/ref_harness.py); the oracle must reproduce it and
    the state_dict must carry the combiner keys at the reference's position (first in the layer)."""
    from flowhigh_b200.config import BackboneConfig
    g = load_golden("vf_unet_skip")
    vcfg = vcfg_from_golden(g)
    bcfg = BackboneConfig(use_unet_skip_connection=True)
    sd = random_state_dict(bcfg, vcfg, seed=int(g["seed"]), vocoder_gain=float(g["gain"]))
    cs = float(sum(v.double().abs().sum().item() for k, v in sd.items() if k.endswith("weight")))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * abs(cs)
    keys = [k for k in sd if ".transformer.layers.1." in k]
    assert keys[0].endswith("layers.1.0.weight") and keys[1].endswith("layers.1.0.bias")
    assert sd[keys[0]].shape == (1024, 2048) and not any(".layers.0.0." in k for k in sd)
    x, cond = torch.from_numpy(g["x"]), torch.from_numpy(g["cond"])
    sd64 = {k: v.double() for k, v in sd.items()}
    v64 = model.vector_field(sd64, x.double(), cond.double(), torch.tensor(0.25, dtype=torch.float64))
    assert float((v64.float() - torch.from_numpy(g["f64_vfield_t025"])).abs().max()) <= 1e-6
    assert float((v64.float() - torch.from_numpy(g["ref_vfield_t025"])).abs().max()) <= 1e-4   # reference fp32 noise
    v32 = model.vector_field(sd, x, cond, torch.tensor(0.25))
    assert float((v32 - torch.from_numpy(g["ref_vfield_t025"])).abs().max()) <= 1e-4
    # without the flag the keys are absent and the default layout is unchanged (711 keys for the assumed config)
    assert not any(".layers.1.0." in k for k in random_state_dict(BackboneConfig(), vcfg, seed=0))


def test_convnext_variant_pinned_against_reference():
    """architecture='convnext' (flow.py:124-139).  Golden from the unmodified reference FLowHigh; the key
    order of the spec was asserted equal to the reference's state_dict() when the fixture was generated."""
    g = load_golden("vf_convnext")
    vcfg = vcfg_from_golden(g)
    bcfg = BackboneConfig(architecture="convnext")
    sd = random_state_dict(bcfg, vcfg, seed=int(g["seed"]), vocoder_gain=float(g["gain"]))
    cs = float(sum(v.double().abs().sum().item() for k, v in sd.items() if k.endswith("weight")))
    assert abs(cs - float(g["weight_checksum"])) <= 1e-6 * abs(cs)
    assert not any(".transformer." in k for k in sd) and sum(".convnext." in k for k in sd) == 8 * 11
    x, cond = torch.from_numpy(g["x"]), torch.from_numpy(g["cond"])
    sd64 = {k: v.double() for k, v in sd.items()}
    v64 = model.vector_field(sd64, x.double(), cond.double(), torch.tensor(0.25, dtype=torch.float64))
    assert float((v64.float() - torch.from_numpy(g["f64_vfield_t025"])).abs().max()) <= 1e-6
    assert float((v64.float() - torch.from_numpy(g["ref_vfield_t025"])).abs().max()) <= 1e-4
    mel = model.cfm_sample_mel(sd, cond, x, steps=2, ode_method="euler", cfm_method="basic_cfm", sigma=0.0)
    assert float((mel - torch.from_numpy(g["ref_mel"])).abs().max()) <= 1e-4


SAMPLE_VARIANTS = {  # tests/golden/make_golden.py:sample_variants_case
    "cfg": dict(cfm_method="basic_cfm", ode_method="midpoint", sigma=0.0, steps=2, cond_scale=1.7),
    "mix": dict(cfm_method="independent_cfm_mix", ode_method="euler", sigma=1e-4, steps=2),
    "mel_pp": dict(cfm_method="independent_cfm_adaptive", ode_method="euler", sigma=1e-4, steps=1, mel_pp=True),
    "cfg_mix_pp": dict(cfm_method="independent_cfm_mix", ode_method="midpoint", sigma=1e-4, steps=1, cond_scale=0.6,
                       mel_pp=True),
}


@pytest.mark.parametrize("variant", sorted(SAMPLE_VARIANTS))
def test_sample_variants_pinned_against_reference(variant):
    """CFG (flow.py:165-178), independent_cfm_mix (cfm_superresolution.py:232-237) and mel_pp (:278-279): the oracle's
    fp64 result against the mel the unmodified reference's sample() produced (its fp32 rounding is the only gap)."""
    g = load_golden("sample_variants")
    sd, vcfg = golden_weights(g)
    cond, eps = torch.from_numpy(g["cond"]), torch.from_numpy(g["eps"])
    assert [model.mel_cutoff_bin(cond[i]) for i in range(cond.shape[0])] == g["cuts"].tolist()
    sd64 = {k: v.double() for k, v in sd.items()}
    out = model.cfm_sample_mel(sd64, cond.double(), eps.double(), **SAMPLE_VARIANTS[variant]).float()
    ref = torch.from_numpy(g["ref_mel_" + variant])
    d = (out - ref).abs()
    # guided sampling from pure noise amplifies the reference's fp32 attention noise (logits reach +-640): mean 3e-4
    tol_max, tol_mean = (2e-2, 1e-3) if variant == "cfg" else (1e-4, 5e-6)
    assert float(d.max()) <= tol_max and float(d.mean()) <= tol_mean, (float(d.max()), float(d.mean()))
