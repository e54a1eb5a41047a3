"""GPU debug: CTA-pair conv kernel (fh_tc_conv with two_cta) against the single-CTA kernel on random shapes."""
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
cases = [tuple(int(a) for a in c.split(",")) for c in (sys.argv[1:] or
         ["2,32,32,300,3,1", "2,96,96,1000,11,5", "3,24,24,700,7,3", "1,384,384,2000,7,1", "2,192,192,5000,3,1", "1,768,768,640,11,1", "4,1024,3072,1,1,1"])]
for B, Ci, Co, L, k, d in cases:
    if k == 1 and L == 1:  # Linear: M tokens
        L = 4000
    w = torch.randn(Co, Ci, k) / (Ci * k) ** 0.5
    b = torch.randn(Co) * 0.1
    tconv = packing.conv1d_taps(w.cuda(), b.cuda(), d)
    r1 = eng._mk_tc(tconv, cin_pad=Ci, cout_pad=Co, two_cta=False)
    r2 = eng._mk_tc(tconv, cin_pad=Ci, cout_pad=Co, two_cta=True)
    A, cs, bs = eng._cbuf("d2_A", B, Ci, L, eng.h16)
    O1, ocs, obs = eng._cbuf("d2_O1", B, Co, L, torch.float32)
    O2, _, _ = eng._cbuf("d2_O2", B, Co, L, torch.float32)
    R, _, _ = eng._cbuf("d2_R", B, Co, L, torch.float32)
    x = torch.randn(B, Ci, L).cuda()
    A.zero_()
    A[: B * bs].view(B, Ci // 8, cs // 8, 8)[:, :, HALO:HALO + L, :] = x.view(B, Ci // 8, 8, L).permute(0, 1, 3, 2).half()
    R.normal_()
    o = HALO * 8
    for res in (False, True):
        O1.zero_(); O2.zero_()
        kw = dict(res=R[o:], res_strides=(obs, ocs, 8), beta=1.0) if res else {}
        eng._tc_conv(r1, A, bs, cs, HALO, O1[o:], (obs, ocs, 8), 0, B, L, **kw)
        torch.cuda.synchronize()
        try:
            eng._tc_conv(r2, A, bs, cs, HALO, O2[o:], (obs, ocs, 8), 0, B, L, **kw)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"B{B} Ci{Ci} Co{Co} L{L} k{k} d{d} res={res}: FAILED {str(e)[:80]} debug word {int(dbg[0]):#x}")
            sys.exit(1)
        a = O1[: B * obs].view(B, Co // 8, ocs // 8, 8)[:, :, HALO:HALO + L].cpu()
        c = O2[: B * obs].view(B, Co // 8, ocs // 8, 8)[:, :, HALO:HALO + L].cpu()
        clean = float(O2[: B * obs].view(B, Co // 8, ocs // 8, 8)[:, :, :HALO].abs().max()) == 0.0
        print(f"B{B} Ci{Ci} Co{Co} L{L} k{k} d{d} res={res}: pair vs single max-abs {float((a - c).abs().max()):.3g} SNR {snr_db(a, c):.1f} dB halo clean {clean}")
