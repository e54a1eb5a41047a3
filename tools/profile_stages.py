"""One pass of the whole path at the bench shapes (B clips x 10 s) with the CUDA profiler API switched on only around it:

  ncu --set full --clock-control none --import-source on --profile-from-start off \\
      -k regex:'^(?!.*(tc_conv|snake)).*' -o gpurun_out/r2_ncu_stages python tools/profile_stages.py 16

captures every kernel class EXCEPT the tcgen05 conv and the snake (captured separately) once: resampler, log-mel,
backbone helpers (dwconv, rmsnorm, q/k-norm + rotary, attention), conv_post + tanh, post-processing."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowhigh_b200 import FlowHighSR, VocoderConfig
from flowhigh_b200.synth import synth_speech

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model = FlowHighSR.from_random(VocoderConfig.assumed_48k(), device=dev, seed=0, precision="fp16")
eng = model._engine()
host = np.stack([synth_speech(120000, 12000, seed=i % 4) for i in range(B)])
x = torch.from_numpy(host).to(dev)
eps = torch.randn((B, 1000, 256), device=dev)


def step():
    eng.new_call(); eng.status_begin()
    cond = eng.resample_normalise(x, 12000)
    mel = eng.sample_mel(eng.encode(cond), eps, steps=1, ode_method="midpoint", cfm_method="basic_cfm", sigma=0.0)
    out = eng.postprocess(eng.vocoder(mel), cond)
    eng.status_end()
    return out


step(); step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
