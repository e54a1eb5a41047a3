"""Is a kernel slow, or is the step power-limited?  Per-launch CUDA-event times of one B = 64 step, twice: back to back
(the chip at its power cap, as in the timed loop) and with the GPU drained + idled FH_PROFILE_COOL_MS before every launch.

  python tools/cool_vs_hot.py hot  > gpurun_out/hot.json ; FH_PROFILE_COOL_MS=15 python tools/cool_vs_hot.py cool > gpurun_out/cool.json
"""
import json
import os
import sys

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


for _ in range(3):
    step()
torch.cuda.synchronize()
eng.start_profile()
step()
agg = eng.stop_profile()
print(json.dumps({"mode": sys.argv[1] if len(sys.argv) > 1 else "", "cool_ms": eng._cool_ms, "per_kernel": agg}))
