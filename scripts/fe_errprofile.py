"""Diagnostic: linear-domain error of the GPU dB spectrogram vs the float64 oracle, as a function of the
frame position inside the 64-frame group (anchor / chain structure).  Not a test."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from birdsoundclassif_b200 import frontend, synth
from oracle import frontend_oracle as fo

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
pcm = synth.synth_pcm(secs, 77)
plan = frontend.get_plan()
print("impl:", plan.impl)
plan.run(torch.from_numpy(pcm).cuda())
torch.cuda.synchronize()
db = plan.spectrogram_view(0).cpu().numpy().astype(np.float64)
ref = fo.db_spectrogram(fo.to_float(pcm), fo.derive_params())[0]
mag, rmag = 10 ** (db / 20), 10 ** (ref / 20)
rms = np.sqrt((rmag ** 2).mean(axis=0, keepdims=True))
e = (mag - rmag) / rms                       # error relative to the frame's rms bin magnitude
T = e.shape[1]
print("overall: rms %.3e  max %.3e  mean(bias) %.3e" % (np.sqrt((e ** 2).mean()), np.abs(e).max(), e.mean()))
pos = np.arange(T) % 64
print("pos  rms_err   max_err")
for p in range(0, 64, 4):
    sel = e[:, pos == p]
    print("%3d  %.3e  %.3e" % (p, np.sqrt((sel ** 2).mean()), np.abs(sel).max()))
rb = np.sqrt((e ** 2).mean(axis=1))
print("by bin (every 25):", " ".join("%d:%.2e" % (b, rb[b]) for b in range(0, 375, 25)))
