"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported with
the stub modules of oracle/ref_harness.py) on CPU.  Run in the build container only:

    python tests/golden/make_golden.py

Weights are not stored: they are `flowhigh_b200.weights.random_state_dict(seed=...)`, regenerated
bit-identically from numpy's PCG64 wherever the tests run; a checksum guards against drift.
Each fixture also carries the fp64 result of the oracle restatement so that tests can compare an
implementation's error with the reference's own fp32 rounding noise.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from flowhigh_b200.config import BackboneConfig, VocoderConfig  # noqa: E402
from flowhigh_b200.synth import synth_speech  # noqa: E402
from flowhigh_b200.weights import random_state_dict  # noqa: E402
from oracle import dsp, model, pipeline, ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GAIN = 0.7


def checksum(sd):
    return float(sum(v.double().abs().sum().item() for k, v in sd.items() if k.endswith("weight")))


def sd64(sd):
    return {k: v.double() for k, v in sd.items()}


def generate_case(name, vcfg, sr, n_in, steps, cfm_method, ode_method, sigma, seed):
    bcfg = BackboneConfig()
    sd = random_state_dict(bcfg, vcfg, seed=seed, vocoder_gain=GAIN)
    ref = ref_harness.build_reference_model(sd, vcfg, cfm_method=cfm_method, ode_method=ode_method, sigma=sigma)
    wav = synth_speech(n_in, sr, seed=seed + 100)
    T = -(-n_in * 48000 // sr)
    N = T // 480
    eps = torch.from_numpy(np.random.default_rng(seed + 7).standard_normal((1, N, 256)).astype(np.float32))
    # --- reference, stage by stage (same calls generate() makes) and end to end
    import scipy.signal
    cond_np = scipy.signal.resample_poly(wav, 48000, sr)
    cond_np = cond_np / np.max(np.abs(cond_np))
    cond = torch.tensor(cond_np).unsqueeze(0).float()
    cond_mel = ref.flowhigh.audio_enc_dec.encode(cond)
    with ref_harness.patched_randn_like(eps):
        mel = ref.sample(cond=cond, time_steps=steps, cfm_method=cfm_method, decode_to_audio=False)
    voc = ref.flowhigh.audio_enc_dec.decode(mel).squeeze(1)
    with ref_harness.patched_randn_like(eps):
        final = ref.generate(wav, sr, 48000, timestep=steps)
    v0 = ref.flowhigh.forward_with_cond_scale(eps, times=torch.tensor(0.5), cond=cond_mel)
    # --- fp64 oracle (noise-floor yardstick)
    o64 = sd64(sd)
    cond64 = torch.from_numpy(dsp.preprocess_audio(wav.astype(np.float64), sr)).unsqueeze(0)
    cond_mel64 = dsp.encode_logmel(cond64)
    mel64 = model.cfm_sample_mel(o64, cond_mel64, eps.double(), steps=steps, ode_method=ode_method,
                                 cfm_method=cfm_method, sigma=sigma)
    voc64 = model.vocoder_forward(o64, vcfg, mel64).squeeze(1)
    final64 = dsp.postprocess(voc64, cond64, cond64.shape[-1])
    v064 = model.vector_field(o64, eps.double(), cond_mel64, torch.tensor(0.5, dtype=torch.float64))
    # vocoder / postproc in fp64 from the REFERENCE's fp32 intermediates (isolates each stage)
    voc64_from_ref_mel = model.vocoder_forward(o64, vcfg, mel.double()).squeeze(1)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        wav=wav, sr=sr, steps=steps, cfm_method=cfm_method, ode_method=ode_method, sigma=sigma, seed=seed, gain=GAIN,
        vcfg=str(vcfg.to_attr_json()), weight_checksum=checksum(sd), eps=eps.numpy(),
        ref_cond=cond.numpy(), ref_cond_mel=cond_mel.numpy(), ref_mel=mel.numpy(), ref_vocoder=voc.numpy(),
        ref_final=final.numpy(), ref_vfield_t05=v0.numpy(),
        f64_cond_mel=cond_mel64.float().numpy(), f64_mel=mel64.float().numpy(), f64_vocoder=voc64.float().numpy(),
        f64_final=final64.float().numpy(), f64_vfield_t05=v064.float().numpy(),
        f64_vocoder_from_ref_mel=voc64_from_ref_mel.float().numpy())
    print(name, "T", T, "N", N, "final absmax", float(final.abs().max()),
          "ref-vs-f64: mel %.3g voc %.3g final %.3g" % ((mel - mel64).abs().max(), (voc - voc64).abs().max(),
                                                        (final - final64).abs().max()))


def vocoder_case(name, vcfg, n_frames, seed):
    sd = random_state_dict(BackboneConfig(), vcfg, seed=seed, vocoder_gain=GAIN)
    ref = ref_harness.build_reference_model(sd, vcfg)
    rng = np.random.default_rng(seed)
    mel = torch.from_numpy((rng.standard_normal((2, n_frames, 256)) * 2.0 - 5.0).astype(np.float32))
    out = ref.flowhigh.audio_enc_dec.decode(mel)
    out64 = model.vocoder_forward(sd64(sd), vcfg, mel.double())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), mel=mel.numpy(), ref_vocoder=out.numpy(),
                        f64_vocoder=out64.float().numpy(), seed=seed, gain=GAIN, vcfg=str(vcfg.to_attr_json()),
                        weight_checksum=checksum(sd))
    print(name, out.shape, "absmax", float(out.abs().max()), "ref-vs-f64 %.3g" % (out - out64).abs().max())


def skip_case(name, seed):
    """Vector field and CFM mel of the U-Net-skip transformer variant from the unmodified reference
    Transformer(use_unet_skip_connection=True) swapped into FLowHigh (oracle/ref_harness.py)."""
    vcfg = VocoderConfig.tiny()
    bcfg = BackboneConfig(use_unet_skip_connection=True)
    sd = random_state_dict(bcfg, vcfg, seed=seed, vocoder_gain=GAIN)
    ref = ref_harness.build_reference_model(sd, vcfg, cfm_method="basic_cfm", ode_method="midpoint", sigma=0.0,
                                            use_unet_skip_connection=True)
    rng = np.random.default_rng(seed)
    B, N = 2, 37
    x = torch.from_numpy(rng.standard_normal((B, N, 256)).astype(np.float32))
    cond = torch.from_numpy((rng.standard_normal((B, N, 256)) * 2.0 - 5.0).astype(np.float32))
    v = ref.flowhigh.forward_with_cond_scale(x, times=torch.tensor(0.25), cond=cond)
    v64 = model.vector_field(sd64(sd), x.double(), cond.double(), torch.tensor(0.25, dtype=torch.float64))
    mel64 = model.cfm_sample_mel(sd64(sd), cond.double(), x.double(), steps=2, ode_method="midpoint",
                                 cfm_method="basic_cfm", sigma=0.0)
    fn = lambda tt, yy: ref.flowhigh.forward_with_cond_scale(yy, times=tt, cond=cond)
    mel = model.odeint_fixed(fn, x, torch.linspace(0, 1, 3), "midpoint")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), cond=cond.numpy(), ref_vfield_t025=v.detach().numpy(),
                        f64_vfield_t025=v64.float().numpy(), ref_mel=mel.detach().numpy(), f64_mel=mel64.float().numpy(),
                        seed=seed, gain=GAIN, vcfg=str(vcfg.to_attr_json()), weight_checksum=checksum(sd))
    print(name, "vfield absmax", float(v.abs().max()), "ref-vs-f64 vfield %.3g mel %.3g" % ((v - v64).abs().max(),
                                                                                          (mel - mel64).abs().max()))


def convnext_case(name, seed):
    """architecture='convnext' (SURVEY 8f row 4): the reference FLowHigh builds it natively (flow.py:124-139)."""
    vcfg = VocoderConfig.tiny()
    bcfg = BackboneConfig(architecture="convnext")
    sd = random_state_dict(bcfg, vcfg, seed=seed, vocoder_gain=GAIN)
    ref = ref_harness.build_reference_model(sd, vcfg, cfm_method="basic_cfm", ode_method="euler", sigma=0.0,
                                            architecture="convnext")
    assert list(ref.state_dict().keys()) == list(sd.keys()), "convnext key order differs from the reference"
    rng = np.random.default_rng(seed)
    B, N = 2, 41
    x = torch.from_numpy(rng.standard_normal((B, N, 256)).astype(np.float32))
    cond = torch.from_numpy((rng.standard_normal((B, N, 256)) * 2.0 - 5.0).astype(np.float32))
    v = ref.flowhigh.forward_with_cond_scale(x, times=torch.tensor(0.25), cond=cond)
    v64 = model.vector_field(sd64(sd), x.double(), cond.double(), torch.tensor(0.25, dtype=torch.float64))
    mel64 = model.cfm_sample_mel(sd64(sd), cond.double(), x.double(), steps=2, ode_method="euler", cfm_method="basic_cfm",
                                 sigma=0.0)
    fn = lambda tt, yy: ref.flowhigh.forward_with_cond_scale(yy, times=tt, cond=cond)
    mel = model.odeint_fixed(fn, x, torch.linspace(0, 1, 3), "euler")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), cond=cond.numpy(), ref_vfield_t025=v.detach().numpy(),
                        f64_vfield_t025=v64.float().numpy(), ref_mel=mel.detach().numpy(), f64_mel=mel64.float().numpy(),
                        seed=seed, gain=GAIN, vcfg=str(vcfg.to_attr_json()), weight_checksum=checksum(sd))
    print(name, "vfield absmax", float(v.abs().max()), "ref-vs-f64 vfield %.3g mel %.3g" % ((v - v64).abs().max(),
                                                                                          (mel - mel64).abs().max()))


def frontend_case(name):
    import scipy.signal
    d = {}
    for sr in (8000, 12000, 16000, 24000, 22050, 44100):
        wav = synth_speech(sr // 2 + 37, sr, seed=sr)
        y = scipy.signal.resample_poly(wav, 48000, sr)
        y = y / np.max(np.abs(y))
        d[f"wav_{sr}"] = wav
        d[f"cond_{sr}"] = y.astype(np.float32)
    wav16 = (synth_speech(8000, 16000, seed=3) * 20000).astype(np.int16)  # int16-range input: /32768 path, fp64
    y = scipy.signal.resample_poly(wav16 / 32768.0, 48000, 16000)
    d["wav_int16"] = wav16
    d["cond_int16"] = (y / np.max(np.abs(y))).astype(np.float32)
    # log-mel of band-limited audio at 48 k (empty high bands exercise the clamp floor)
    ref = ref_harness.build_reference_model(random_state_dict(BackboneConfig(), VocoderConfig.tiny(), 0),
                                            VocoderConfig.tiny())
    for sr in (8000, 24000):
        a = torch.from_numpy(d[f"cond_{sr}"])[None]
        d[f"logmel_{sr}"] = ref.flowhigh.audio_enc_dec.encode(a).numpy()
        d[f"logmel64_{sr}"] = dsp.encode_logmel(a.double()).float().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "ok")


def big_case(name, sr, n_in, steps, cfm_method, ode_method, sigma, seed):
    """The configuration bench.py times (VocoderConfig.assumed_48k(): C0 = 1536, stages 768 ... 24) at the BASELINE clip
    sizes, from the unmodified reference.  Slim fixture: input, noise seed and the reference's mel / vocoder / final
    outputs only (no fp64 restatement; that pins the oracle on the small fixtures above)."""
    vcfg = VocoderConfig.assumed_48k()
    sd = random_state_dict(BackboneConfig(), vcfg, seed=seed, vocoder_gain=GAIN)
    ref = ref_harness.build_reference_model(sd, vcfg, cfm_method=cfm_method, ode_method=ode_method, sigma=sigma)
    wav = synth_speech(n_in, sr, seed=seed + 100)
    T = -(-n_in * 48000 // sr)
    N = T // 480
    eps = torch.from_numpy(np.random.default_rng(seed + 7).standard_normal((1, N, 256)).astype(np.float32))
    import scipy.signal
    cond_np = scipy.signal.resample_poly(wav, 48000, sr)
    cond_np = cond_np / np.max(np.abs(cond_np))
    cond = torch.tensor(cond_np).unsqueeze(0).float()
    with ref_harness.patched_randn_like(eps):
        mel = ref.sample(cond=cond, time_steps=steps, cfm_method=cfm_method, decode_to_audio=False)
    voc = ref.flowhigh.audio_enc_dec.decode(mel).squeeze(1)
    with ref_harness.patched_randn_like(eps):
        final = ref.generate(wav, sr, 48000, timestep=steps)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        wav=wav, sr=sr, steps=steps, cfm_method=cfm_method, ode_method=ode_method, sigma=sigma, seed=seed, gain=GAIN,
        eps_seed=seed + 7, vcfg=str(vcfg.to_attr_json()), weight_checksum=checksum(sd),
        ref_mel=mel.numpy(), ref_vocoder=voc.numpy(), ref_final=final.numpy())
    print(name, "T", T, "N", N, "vocoder absmax", float(voc.abs().max()), "final absmax", float(final.abs().max()))


def sample_variants_case(name, seed):
    """`sample()` with cond_scale != 1 (classifier-free guidance, flow.py:165-178), cfm_method='independent_cfm_mix'
    (cfm_superresolution.py:232-237) and mel_pp=True (:278-279) from the unmodified reference (mel output)."""
    vcfg = VocoderConfig.tiny()
    sd = random_state_dict(BackboneConfig(), vcfg, seed=seed, vocoder_gain=GAIN)
    rng = np.random.default_rng(seed)
    B, N = 2, 45
    # band-limited log-mel-like conditioning: different cutoff bins per clip
    cond = rng.standard_normal((B, N, 256)).astype(np.float32) * 1.5 - 4.0
    cond[0, :, 90:] = -11.5129 + 0.01 * rng.standard_normal((N, 166)).astype(np.float32)
    cond[1, :, 170:] = -11.5129 + 0.01 * rng.standard_normal((N, 86)).astype(np.float32)
    cond = torch.from_numpy(cond)
    eps = torch.from_numpy(rng.standard_normal((B, N, 256)).astype(np.float32))
    d = dict(cond=cond.numpy(), eps=eps.numpy(), seed=seed, gain=GAIN, vcfg=str(vcfg.to_attr_json()),
             weight_checksum=checksum(sd))
    cases = {"cfg": dict(cfm_method="basic_cfm", ode="midpoint", sigma=0.0, kw=dict(cond_scale=1.7, time_steps=2)),
             "mix": dict(cfm_method="independent_cfm_mix", ode="euler", sigma=1e-4, kw=dict(time_steps=2)),
             "mel_pp": dict(cfm_method="independent_cfm_adaptive", ode="euler", sigma=1e-4, kw=dict(time_steps=1, mel_pp=True)),
             "cfg_mix_pp": dict(cfm_method="independent_cfm_mix", ode="midpoint", sigma=1e-4,
                                kw=dict(cond_scale=0.6, time_steps=1, mel_pp=True))}
    for tag, c in cases.items():
        ref = ref_harness.build_reference_model(sd, vcfg, cfm_method=c["cfm_method"], ode_method=c["ode"], sigma=c["sigma"])
        with ref_harness.patched_randn_like(eps):
            mel = ref.sample(cond=cond, decode_to_audio=False, cfm_method=c["cfm_method"], **c["kw"])
        cuts = ref.mel_cutoff_bins(cond)
        mel64 = model.cfm_sample_mel(sd64(sd), cond.double(), eps.double(), steps=c["kw"]["time_steps"], ode_method=c["ode"],
                                     cfm_method=c["cfm_method"], sigma=c["sigma"], cond_scale=c["kw"].get("cond_scale", 1.0),
                                     mel_pp=c["kw"].get("mel_pp", False))
        d["ref_mel_" + tag] = mel.numpy()
        d["f64_mel_" + tag] = mel64.float().numpy()
        d["cuts"] = np.asarray(cuts, dtype=np.int32)
        print(name, tag, "cut bins", cuts, "ref-vs-f64 %.3g" % (mel - mel64).abs().max())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


if __name__ == "__main__":
    assert ref_harness.available(), "needs /root/reference"
    torch.manual_seed(0)
    only = set(sys.argv[1:])
    want = lambda n: not only or n in only
    if want("frontend"):
        frontend_case("frontend")
    if want("gen_c1_adaptive_euler"):
        generate_case("gen_c1_adaptive_euler", VocoderConfig.tiny(), 16000, 16000, 1, "independent_cfm_adaptive", "euler",
                      1e-4, seed=1)
    if want("gen_basic_midpoint"):
        generate_case("gen_basic_midpoint", VocoderConfig.tiny(), 12000, 9000, 1, "basic_cfm", "midpoint", 0.0, seed=2)
    if want("gen_basic_euler4"):
        generate_case("gen_basic_euler4", VocoderConfig.tiny(), 24000, 12240, 4, "basic_cfm", "euler", 0.0, seed=3)
    if want("voc_resblock2_snake"):
        vocoder_case("voc_resblock2_snake", VocoderConfig.tiny(resblock="2", activation="snake", logscale=False), 24, seed=4)
    if want("voc_resblock1_snakebeta"):
        vocoder_case("voc_resblock1_snakebeta", VocoderConfig.tiny(), 30, seed=5)
    if want("vf_unet_skip"):
        skip_case("vf_unet_skip", seed=6)
    if want("vf_convnext"):
        convnext_case("vf_convnext", seed=7)
    if want("sample_variants"):
        sample_variants_case("sample_variants", seed=8)
    # the configuration bench.py times: BASELINE configs[0] exactly (4 s, 16 kHz, adaptive, euler 1 step, N = 400) and one
    # clip of configs[1] (10 s, 12 kHz, basic_cfm, midpoint, N = 1000)
    if want("big_c0_4s_adaptive_euler"):
        big_case("big_c0_4s_adaptive_euler", 16000, 64000, 1, "independent_cfm_adaptive", "euler", 1e-4, seed=21)
    if want("big_c1_10s_basic_midpoint"):
        big_case("big_c1_10s_basic_midpoint", 12000, 120000, 1, "basic_cfm", "midpoint", 0.0, seed=0)
