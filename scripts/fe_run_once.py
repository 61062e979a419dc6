"""One front-end run on a batch of 60 s clips (for ncu captures of single kernels): python scripts/fe_run_once.py [clips] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import frontend, synth  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw = dict(freq_accuracy=10.0, dt=0.001) if os.environ.get("NBM_STRESS") else {}
base = [synth.synth_pcm(60.0, 300 + i, calls_per_s=2.0 + i) for i in range(6)]
pcm = torch.from_numpy(np.concatenate([base[i % 6] for i in range(clips)])).cuda()
offs = (np.arange(clips + 1) * len(base[0])).tolist()
plan = frontend.FrontendPlan(**kw)
for _ in range(reps):
    tiles, toff, mm = plan.run_batch(pcm, offs)
torch.cuda.synchronize()
print("frames", int(sum(plan.query_batch([len(base[0])] * clips)[0])), "listed", plan.last_listed())
