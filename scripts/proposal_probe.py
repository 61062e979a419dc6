"""One fused ProposalLayer call repeated (run under `ncu --metrics gpu__time_duration.sum` for its kernel list):
    python scripts/proposal_probe.py [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import postproc as pp, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
args = synth.default_args("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
cls = torch.rand((4, 15, 2, 24, 64), device="cuda", generator=g).softmax(2).reshape(4, 30, 24, 64)
reg = torch.randn((4, 60, 24, 64), device="cuda", generator=g) * 0.2
layer = pp.ProposalLayer(args, args.n_layers).eval()
for _ in range(3):
    rois, sc = layer(cls, reg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t = time.perf_counter()
for _ in range(reps):
    rois, sc = layer(cls, reg)
e1.record()
torch.cuda.synchronize()
print("rois", tuple(rois.shape), "per call: device span %.1f us, host %.1f us" % (e0.elapsed_time(e1) * 1e3 / reps, (time.perf_counter() - t) * 1e6 / reps))
