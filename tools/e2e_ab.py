"""A/B of the host-buffer path inside ONE process (same box, same clocks): resident step vs generate_batch with per-clip D2H
copies by the caller vs generate_batch(out_host=...) with overlapped sub-batches.  python tools/e2e_ab.py [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowhigh_b200 import FlowHighSR, VocoderConfig  # noqa: E402
from flowhigh_b200.synth import synth_speech  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
B = 64
model = FlowHighSR.from_random(VocoderConfig.assumed_48k(), device=dev, precision="fp16")
eng = model._engine()
host_t = torch.from_numpy(np.stack([synth_speech(120000, 12000, i) for i in range(B)])).pin_memory()
x_dev = host_t.to(dev)
eps = torch.randn((B, 1000, 256), device=dev)
out_host = torch.empty((B, 480000), dtype=torch.float32).pin_memory()


def resident():
    cond = eng.resample_normalise(x_dev, 12000)
    mel = eng.sample_mel(eng.encode(cond), eps, steps=1, ode_method="midpoint", cfm_method="basic_cfm", sigma=0.0)
    return eng.postprocess(eng.vocoder(mel), cond)


def old():
    outs = model.generate_batch(list(host_t), 12000, 48000, timestep=1, eps=list(eps), pinned=True)
    for i, o in enumerate(outs):
        out_host[i].copy_(o[0], non_blocking=True)


def new():
    model.generate_batch(list(host_t), 12000, 48000, timestep=1, eps=list(eps), pinned=True, out_host=out_host)


def timeit(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return 1000 * (time.perf_counter() - t0) / reps


for fn in (resident, old, new):
    fn()
for rnd in range(3):
    print({f.__name__: round(timeit(f), 2) for f in (resident, old, new)}, flush=True)
for mb in (8, 11, 16, 1000):
    model.overlap_min_batch = mb
    new()
    print("overlap_min_batch", mb, round(timeit(new), 2), flush=True)
