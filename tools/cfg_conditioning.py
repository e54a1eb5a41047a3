"""CPU experiment: why does classifier-free guidance lose ~40 dB on the 16-bit path with RANDOM-INIT
weights?  The unconditional branch (cond = null_cond for every token) is ill-conditioned: all tokens share the
conditioning, q / k differ only through the noise input, and `softmax(10 q k^T)` with logits of +-640 turns rounding
noise into attention flips.  The oracle's own fp32 run is 33 dB further from fp64 on that branch (93 vs 127 dB), and
fp16 rounding of ANY single GEMM operand (activations or weights, any layer) lands at 35-41 dB -- what the GPU measures
(tools/diag_cfg.py: 32 dB).  Not an implementation defect; `precision='fp32'` gives reference-level CFG parity.

    python tools/cfg_conditioning.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model  # noqa: E402
from util import golden_weights, load_golden, snr_db  # noqa: E402

if __name__ == "__main__":
    g = load_golden("sample_variants")
    sd, vcfg = golden_weights(g)
    cond, eps = torch.from_numpy(g["cond"]), torch.from_numpy(g["eps"])
    null = sd["flowhigh.null_cond"].expand_as(cond).contiguous()
    sd64 = {k: v.double() for k, v in sd.items()}
    t = torch.tensor(0.25)
    ref = {n: model.vector_field(sd64, eps.double(), c.double(), t.double()).float() for n, c in (("cond", cond), ("null", null))}
    _lin = F.linear
    h = lambda x: x.half().to(x.dtype)
    MODE = {}

    def lin(x, w, b=None):
        big = w.shape[0] >= 256 and w.shape[-1] >= 256 and x.dim() == 3
        return _lin(h(x) if big and MODE.get("A") else x, h(w) if big and MODE.get("W") else w, b)
    F.linear = lin
    for name, mode in (("fp32", {}), ("fp16 activations", {"A": 1}), ("fp16 weights", {"W": 1}), ("fp16 both", {"A": 1, "W": 1})):
        MODE.clear()
        MODE.update(mode)
        print(f"{name:18s} " + "  ".join(f"{n}: {snr_db(ref[n], model.vector_field(sd, eps, c, t)):6.1f} dB" for n, c in (("cond", cond), ("null", null))))
