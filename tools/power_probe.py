"""Which kernel classes hold the chip at its power cap?  Records the launches of one B = 64 step, then replays the whole step and
three subsets of it (snake only / HBM-bound convolutions only / tensor-bound convolutions only) for ~2 s each while
nvidia-smi samples SM clock and board power every 20 ms.

  python tools/power_probe.py > gpurun_out/power_probe.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowhigh_b200 import FlowHighSR, VocoderConfig  # noqa: E402
from flowhigh_b200.synth import synth_speech  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
model = FlowHighSR.from_random(VocoderConfig.assumed_48k(), device=dev, precision="fp16")
eng = model._engine()
x = torch.from_numpy(np.stack([synth_speech(120000, 12000, 0)] * B)).to(dev)
eps = torch.randn((B, 1000, 256), device=dev)


def step():
    cond = eng.resample_normalise(x, 12000)
    mel = eng.sample_mel(eng.encode(cond), eps, steps=1, ode_method="midpoint", cfm_method="basic_cfm", sigma=0.0)
    return eng.postprocess(eng.vocoder(mel), cond)


for _ in range(2):
    out = step()
torch.cuda.synchronize()
rec = []
orig_call, orig_conv = eng._call, eng._launch_conv


def rec_call(name, *args, work=None):
    rec.append(("k", name, args, None))
    return orig_call(name, *args, work=work)


def rec_conv(args, work):
    rec.append(("c", work["tag"], args, work))
    return orig_conv(args, work)


eng._call, eng._launch_conv = rec_call, rec_conv
keep = step()  # outputs stay referenced so that no buffer the recording points at is freed
torch.cuda.synchronize()
eng._call, eng._launch_conv = orig_call, orig_conv


def replay(items):
    for kind, name, args, _ in items:
        if kind == "k":
            getattr(eng.lib, name)(*args)
        else:
            eng.lib.fh_tc_conv(C.byref(args), eng.stream)


def hbm_bound(w):
    return w["bytes"] / 6549.4e9 > w["flops"] / 1346.6e12


groups = {
    "full_step": rec,
    "snake_only": [r for r in rec if r[1].startswith("fh_snake_aa_chunked")],
    "conv_hbm_bound": [r for r in rec if r[0] == "c" and hbm_bound(r[3])],
    "conv_tensor_bound": [r for r in rec if r[0] == "c" and not hbm_bound(r[3])],
}
log = "/tmp/smi_probe.csv"
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                        "-lms", "20", "-f", log])
time.sleep(0.5)
marks, res = [], {}
t_origin = time.time()
for name, items in groups.items():
    torch.cuda.synchronize()
    time.sleep(1.0)  # idle gap between phases
    replay(items)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    reps = 0
    e0.record()
    while time.time() - t0 < 2.5:
        replay(items)
        reps += 1
        if reps % 4 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    marks.append((name, t0 - t_origin, t1 - t_origin))
    res[name] = {"launches": len(items), "reps": reps, "ms_per_pass": e0.elapsed_time(e1) / reps}
time.sleep(0.3)
smi.terminate()
smi.wait()
rows = [ln.strip().split(", ") for ln in open(log) if ln.strip()]
n = len(rows)
t_total = time.time() - t_origin + 0.5
for name, a, b in marks:  # samples are evenly spaced over [-0.5, t_total): take the middle 70 % of each phase
    i0 = int((a + 0.5 + 0.15 * (b - a)) / t_total * n)
    i1 = int((a + 0.5 + 0.85 * (b - a)) / t_total * n)
    seg = rows[i0:i1]
    clk = [float(r[0]) for r in seg if r[0].replace(".", "").isdigit()]
    pw = [float(r[1]) for r in seg if r[1].replace(".", "").isdigit()]
    res[name].update({"sm_mhz_median": float(np.median(clk)) if clk else None, "sm_mhz_min": min(clk) if clk else None,
                      "sm_mhz_max": max(clk) if clk else None, "power_w_median": float(np.median(pw)) if pw else None,
                      "reasons": sorted({r[2] for r in seg if len(r) > 2})[:4], "samples": len(seg)})
print(json.dumps(res))
