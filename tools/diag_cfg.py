"""GPU diagnostic: where does the 16-bit classifier-free-guidance path lose accuracy?  (sample_variants golden)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from flowhigh_b200.engine import Engine
from util import golden_weights, load_golden, snr_db
g = load_golden("sample_variants")
sd, vcfg = golden_weights(g)
cond, eps = torch.from_numpy(g["cond"]).cuda(), torch.from_numpy(g["eps"]).cuda()
e32 = Engine(sd, vcfg, device="cuda:0", precision="fp32")
for prec in ("fp16", "bf16"):
    e16 = Engine(sd, vcfg, device="cuda:0", precision=prec)
    null = sd["flowhigh.null_cond"].cuda().expand_as(cond).contiguous()
    zero = torch.zeros_like(eps)
    for name, c in (("cond", cond), ("null", null)):
        for t in (0.0, 0.25, 0.5):
            a, b = torch.empty_like(eps), torch.empty_like(eps)
            e32.vector_field_step(eps, c, t, zero, 1.0, a)
            e16.vector_field_step(eps, c, t, zero, 1.0, b)
            print(prec, name, "t", t, "field SNR %.1f dB" % snr_db(a.cpu(), b.cpu()), "|v| rms %.3f" % float(a.pow(2).mean().sqrt()),
                  "max err %.3g" % float((a - b).abs().max()))
    for cs in (1.0, 1.7):
        kw = dict(steps=2, ode_method="midpoint", cfm_method="basic_cfm", sigma=0.0, cond_scale=cs)
        a = e32.sample_mel(cond, eps, **kw).cpu(); b = e16.sample_mel(cond, eps, **kw).cpu()
        print(prec, "sample cond_scale", cs, "mel SNR %.1f dB" % snr_db(a, b))

# ---- where inside the field does the null-conditioned branch lose accuracy?  compare the engines' named buffers
def bufs(e, names):
    out = {}
    for (name, shape, dt), t in e._bufs.items():
        if name in names:
            out[name] = t.float().cpu().clone()
    return out
e16 = Engine(sd, vcfg, device="cuda:0", precision="fp16")
null = sd["flowhigh.null_cond"].cuda().expand_as(cond).contiguous()
zero = torch.zeros_like(eps)
for name, c in (("cond", cond), ("null", null)):
    a, b = torch.empty_like(eps), torch.empty_like(eps)
    e32.vector_field_step(eps, c, 0.25, zero, 1.0, a)
    e16.vector_field_step(eps, c, 0.25, zero, 1.0, b)
    torch.cuda.synchronize()
    A, Bf = bufs(e32, {"bb_E", "bb_h", "bb_qkv"}), bufs(e16, {"bb_E", "bb_h", "bb_qkv"})
    for k in ("bb_E", "bb_qkv", "bb_h"):
        x, y = A[k], Bf[k]
        print(name, k, tuple(x.shape), "SNR %.1f dB" % snr_db(x, y), "rms %.3g absmax %.3g" % (float(x.pow(2).mean().sqrt()), float(x.abs().max())))
print("null_cond rms", float(sd["flowhigh.null_cond"].pow(2).mean().sqrt()), "cond rms", float(cond.pow(2).mean().sqrt()))
