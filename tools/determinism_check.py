import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from flowhigh_b200 import FlowHighSR, packing
from flowhigh_b200.engine import HALO
from util import golden_weights, load_golden
g = load_golden("gen_basic_midpoint")
sd, vcfg = golden_weights(g)
m = FlowHighSR.from_random(vcfg, device="cuda:0", precision="fp16"); m.load_state_dict(sd); m = m.cuda()
eng = m._engine()
# kernel-level determinism
torch.manual_seed(0)
filt = eng.sd["flowhigh.audio_enc_dec.vocoder.activation_post.upsample.filter"].flatten().float().cuda().contiguous()
for (B, C, L) in [(4, 96, 20000), (2, 768, 5000), (1, 24, 48000), (1, 16, 700)]:
    Lp = HALO + packing.round_up(L, 128) + 64
    cs, bs = Lp * 8, (C // 8) * Lp * 8
    xc = torch.zeros(B, C // 8, Lp, 8, device="cuda:0"); xc[:, :, HALO:HALO + L] = torch.randn(B, C // 8, L, 8, device="cuda:0") * 2
    a = torch.rand(C, device="cuda:0") + 0.5; ib = torch.rand(C, device="cuda:0") + 0.5
    outs = []
    for rep in range(4):
        y = torch.zeros(B, C // 8, Lp, 8, dtype=torch.float16, device="cuda:0")
        eng._call("fh_snake_aa_chunked", xc.data_ptr(), y.data_ptr(), a.data_ptr(), ib.data_ptr(), filt.data_ptr(), bs, cs, HALO, B, C, L, 2, eng.stream)
        torch.cuda.synchronize(); outs.append(y.clone())
    d = [int((outs[0] != o).sum()) for o in outs[1:]]
    print("kernel determinism", (B, C, L), "mismatching elements:", d, "finite:", bool(torch.isfinite(outs[0].float()).all()))
    if any(d):
        idx = (outs[0] != outs[1]).nonzero()[:8].tolist(); print("  first mismatches (b, chunk, row, ch):", idx, "L rows start at", HALO)
eps = torch.from_numpy(g["eps"]); wav, sr = g["wav"], int(g["sr"])
m.cuda_graphs = False
r1 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu(); r2 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu()
print("eager vs eager max diff", float((r1 - r2).abs().max()))
eng.branch_streams = False
r3 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu(); r4 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu()
print("eager(no branch streams) self diff", float((r3 - r4).abs().max()), "vs branch-streams", float((r1 - r3).abs().max()))
eng.branch_streams = True
m.cuda_graphs = True
a1 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu(); a2 = m.generate(wav, sr, 48000, timestep=1, eps=eps).cpu()
print("graph vs eager", float((a1 - r1).abs().max()), "graph vs graph", float((a1 - a2).abs().max()))
