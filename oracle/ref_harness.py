"""ORACLE tooling -- runs the UNMODIFIED reference from /root/reference on CPU.

Runs from /root/reference in the build container (tests/golden/make_golden.py pins oracle/ against the
reference's own code there) or from baseline/_ref, the `pip install --target` copy of the unmodified reference
that travels to the GPU box for `bench.py --impl reference` and the `cpu_baseline` leg.

Four packages the reference imports are missing offline: `librosa`,
`torchdiffeq`, `torchode`, `gateloop_transformer`.  They are replaced by stub modules:
`torchode` / `gateloop_transformer` names are never called on the hot path;
`torchdiffeq.odeint` and `librosa.filters.mel` are the oracle restatements (so those two
pieces are NOT independently pinned by this harness -- resample_poly is pinned against the
installed scipy, the mel table is "parity unpinned").  `.cuda()` is neutralised because the
reference hard-codes CUDA.  No reference file is modified or copied.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree of the build container, else the pip --target install that travels to the GPU box
# (baseline/_ref: git-ignored, created by __graft_entry__.build(); unmodified reference package)
_CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))


def ref_src():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "flowhigh")):
            return c
    return None


REF_SRC = ref_src() or _CANDIDATES[0]


def available() -> bool:
    return ref_src() is not None


_loaded = {}


def load_reference():
    """Imports `flowhigh` from the reference tree with stubs; returns the package."""
    if "pkg" in _loaded:
        return _loaded["pkg"]
    if not available():
        raise RuntimeError("reference tree not present")
    from . import dsp, model

    def _stub(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "torchode" not in sys.modules:
        m = _stub("torchode")
        for n in ("Tsit5", "ODETerm", "IntegralController", "AutoDiffAdjoint", "InitialValueProblem"):
            setattr(m, n, type(n, (), {}))
    if "gateloop_transformer" not in sys.modules:
        m = _stub("gateloop_transformer")
        m.SimpleGateLoopLayer = type("SimpleGateLoopLayer", (), {})
    if "torchdiffeq" not in sys.modules:
        m = _stub("torchdiffeq")

        def odeint(fn, y0, t, atol=None, rtol=None, method="midpoint"):
            # returns a 2-entry "trajectory" so that trajectory[-1] is the final state
            return [y0, model.odeint_fixed(fn, y0, t, method)]

        m.odeint = odeint
    if "librosa" not in sys.modules:
        m = _stub("librosa")
        f = _stub("librosa.filters")
        f.mel = lambda sr, n_fft, n_mels, fmin, fmax: dsp.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        m.filters = f
        u = _stub("librosa.util")
        u.normalize = lambda x, **k: x
        m.util = u
        m.load = None
        m.resample = None

    # .cuda() -> no-op on CPU-only hosts (reference pins device 0 everywhere)
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, device=None: self
        torch.Tensor.cuda = lambda self, *a, **k: self

    # the reference configures logging to ./model_debug.log at import: run from a scratch dir
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="fh_ref_")
    os.chdir(scratch)
    try:
        src = ref_src()
        sys.path.insert(0, src)
        import flowhigh  # noqa: F401
        import logging
        logging.getLogger().setLevel(logging.ERROR)  # F9b: avoid tensor repr formatting cost
    finally:
        sys.path.remove(src)
        os.chdir(cwd)
    _loaded["pkg"] = sys.modules["flowhigh"]
    return _loaded["pkg"]


def build_reference_model(sd, vcfg, *, cfm_method="basic_cfm", ode_method="midpoint", sigma=0.0, depth=2,
                          use_unet_skip_connection=False, architecture="transformer"):
    """Reference FlowHighSR with `sd` loaded (strict), bypassing the checkpoint-file loaders.

    use_unet_skip_connection: the reference's FLowHigh never forwards that flag to its Transformer, so the
    unmodified reference `Transformer(..., use_unet_skip_connection=True)` (transformer.py:108-165) is constructed with
    the arguments of flow.py:109-122 and swapped in before the weights are loaded."""
    pkg = load_reference()
    from flowhigh.models import melvoco as ref_melvoco
    from flowhigh.models.bigvgan.models import BigVGAN
    from flowhigh.models.bigvgan.env import AttrDict

    def init_bigvgan(config, checkpoint, vocoder_freeze=False):
        voc = BigVGAN(AttrDict(vcfg.to_attr_json()))
        voc.eval()
        voc.remove_weight_norm()
        for p in voc.parameters():
            p.requires_grad = False
        return voc

    ref_melvoco.init_bigvgan = init_bigvgan
    voc = pkg.models.MelVoco(vocoder_config=None, vocoder_path=None)
    net = pkg.models.FLowHigh(dim_in=voc.n_mels, audio_enc_dec=voc, depth=depth, architecture=architecture).eval()
    if use_unet_skip_connection:
        from flowhigh.models.transformer import Transformer
        net.transformer = Transformer(dim=1024, depth=depth, dim_head=64, heads=16, ff_mult=4, ff_dropout=0.0,
                                      attn_dropout=0.0, attn_flash=False, attn_qk_norm=True, adaptive_rmsnorm=True,
                                      adaptive_rmsnorm_cond_dim_in=1024, use_gateloop_layers=False,
                                      use_unet_skip_connection=True).eval()
    m = pkg.FlowHighSR(flowhigh=net, cfm_method=cfm_method, torchdiffeq_ode_method=ode_method, sigma=sigma)
    m.load_state_dict(sd, strict=True)
    return m.eval()


class patched_randn_like:
    """Makes the reference's `torch.randn_like(cond)` (cfm_superresolution.py:220-234) return a
    caller-provided epsilon so both implementations integrate from the same prior."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self._orig = torch.randn_like
        eps = self.eps
        torch.randn_like = lambda t, *a, **k: eps.to(t.dtype).reshape(t.shape).clone()
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig
        return False
