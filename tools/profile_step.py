"""One resident step of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --batch 8
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:tc_conv_kernel -s 24 -c 2 -o gpurun_out/prof_tc python tools/profile_step.py --batch 8
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowhigh_b200 import FlowHighSR, VocoderConfig  # noqa: E402
from flowhigh_b200.synth import synth_speech  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
dev = torch.device("cuda:0")
model = FlowHighSR.from_random(VocoderConfig.assumed_48k(), device=dev, precision=args.precision)
eng = model._engine()
n_in = int(args.seconds * 12000)
x = torch.from_numpy(np.stack([synth_speech(n_in, 12000, 0)] * args.batch)).to(dev)
N = int(args.seconds * 48000) // 480
eps = torch.randn((args.batch, N, 256), device=dev)


def step():
    cond = eng.resample_normalise(x, 12000)
    mel = eng.sample_mel(eng.encode(cond), eps, steps=1, ode_method="midpoint", cfm_method="basic_cfm", sigma=0.0)
    return eng.postprocess(eng.vocoder(mel), cond)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
