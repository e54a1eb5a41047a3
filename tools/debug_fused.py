"""GPU debug: one fused snake-prologue conv (fh_tc_conv with x_f32) against snake + conv launched separately."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from flowhigh_b200 import packing, _lib
from flowhigh_b200.engine import Engine, HALO
from util import golden_weights, load_golden, snr_db

g = load_golden("voc_resblock1_snakebeta")
sd, vcfg = golden_weights(g)
eng = Engine(sd, vcfg, device="cuda:0", precision="fp16")
eng.new_call()
dbg = torch.zeros(1, dtype=torch.int32).pin_memory()
_lib.check(eng.lib.fh_set_debug_word(dbg.data_ptr()), "dbg")
torch.manual_seed(0)
cases = [(int(a) for a in c.split(",")) for c in (sys.argv[1:] or ["2,32,300,3,1", "2,96,1000,11,5", "3,24,700,7,3", "1,8,200,3,1", "2,128,520,11,1"])]
for B, C, L, k, d in cases:
    w = torch.randn(C, C, k) / (C * k) ** 0.5
    b = torch.randn(C) * 0.1
    rec = eng._mk_tc(packing.conv1d_taps(w.cuda(), b.cuda(), d), cin_pad=C, cout_pad=C)
    alpha = torch.exp(torch.randn(C) * 0.3).cuda(); ib = (1.0 / (torch.exp(torch.randn(C) * 0.3) + 1e-9)).cuda()
    filt = eng.voc["post_act"][2]
    X, cs, bs = eng._cbuf("dbg_X", B, C, L, torch.float32)
    A, _, _ = eng._cbuf("dbg_A", B, C, L, eng.h16)
    O1, _, _ = eng._cbuf("dbg_O1", B, C, L, torch.float32)
    O2, _, _ = eng._cbuf("dbg_O2", B, C, L, torch.float32)
    Xh, _, _ = eng._cbuf("dbg_Xh", B, C, L, eng.h16)
    x = torch.randn(B, C, L).cuda() * 1.5
    v = X[: B * bs].view(B, C // 8, cs // 8, 8)
    v[:, :, HALO:HALO + L, :] = x.view(B, C // 8, 8, L).permute(0, 1, 3, 2)
    Xh[: B * bs].view(B, C // 8, cs // 8, 8)[:, :, HALO:HALO + L, :] = x.view(B, C // 8, 8, L).permute(0, 1, 3, 2).half()
    o = HALO * 8
    st = eng.stream
    for x16 in (False, True):
        src = Xh if x16 else X
        O1.zero_(); O2.zero_()
        if x16:
            eng._call("fh_snake_aa_chunked_h", src.data_ptr(), A.data_ptr(), alpha.data_ptr(), ib.data_ptr(), filt.data_ptr(), bs, cs, HALO, B, C, L, st)
        else:
            eng._call("fh_snake_aa_chunked", src.data_ptr(), A.data_ptr(), alpha.data_ptr(), ib.data_ptr(), filt.data_ptr(), bs, cs, HALO, B, C, L, 2, st)
        eng._tc_conv(rec, A, bs, cs, HALO, O1[o:], (bs, cs, 8), 0, B, L, res=X[o:], res_strides=(bs, cs, 8), beta=1.0)
        torch.cuda.synchronize()
        try:
            eng._tc_conv(rec, None, bs, cs, HALO, O2[o:], (bs, cs, 8), 0, B, L, res=X[o:], res_strides=(bs, cs, 8), beta=1.0,
                         xf=src, snake=(alpha, ib, filt), x16=x16)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"B{B} C{C} L{L} k{k} d{d} x16={x16}: FAILED {str(e)[:80]} debug word {int(dbg[0]):#x}")
            sys.exit(1)
        a = O1[: B * bs].view(B, C // 8, cs // 8, 8)[:, :, HALO:HALO + L].cpu()
        c = O2[: B * bs].view(B, C // 8, cs // 8, 8)[:, :, HALO:HALO + L].cpu()
        halo_clean = float(O2[: B * bs].view(B, C // 8, cs // 8, 8)[:, :, :HALO].abs().max()) == 0.0
        print(f"B{B} C{C} L{L} k{k} d{d} x16={x16}: fused vs separate max-abs {float((a - c).abs().max()):.3g} SNR {snr_db(a, c):.1f} dB halo clean {halo_clean}")
