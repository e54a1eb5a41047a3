"""Drop-in counterpart of the reference's example.py (reference example.py:1-12) on the B200 engine.

    python example.py LOW-RES-AUDIO.wav OUTPUT.wav [--ckpt-dir DIR] [--steps 1]

With --ckpt-dir the four files of the ResembleAI/FlowHigh hub repo are read from DIR (`FlowHighSR.from_local`);
without it `from_pretrained` downloads them (needs network).  `--random` builds random-init weights of the same
architecture, for smoke runs on a box without checkpoints.
"""
import argparse

from flowhigh_b200 import FlowHighSR
from flowhigh_b200.io import load_wav, save_wav

TARGET_SR = 48000

ap = argparse.ArgumentParser()
ap.add_argument("input")
ap.add_argument("output")
ap.add_argument("--ckpt-dir", default=None)
ap.add_argument("--random", action="store_true")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--long", action="store_true", help="chunked long-form generation (clips longer than ~60 s)")
args = ap.parse_args()

if args.random:
    model = FlowHighSR.from_random(device="cuda")
elif args.ckpt_dir:
    model = FlowHighSR.from_local(args.ckpt_dir, device="cuda")
else:
    model = FlowHighSR.from_pretrained(device="cuda")

wav, sr_in = load_wav(args.input)
if wav.shape[0] > 1:  # the reference takes mono [1, T] (flowhighsr.py:59-60)
    wav = wav.mean(0, keepdim=True)
if args.long:
    wav_hr = model.generate_long(wav, sr_in, TARGET_SR, timestep=args.steps)
else:
    wav_hr = model.generate(wav, sr_in, TARGET_SR, timestep=args.steps)
save_wav(args.output, wav_hr.cpu(), TARGET_SR)
