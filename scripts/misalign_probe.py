"""Diagnostic: front-end rate when file starts are not multiples of 8 samples (odd-length files in a batch)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from birdsoundclassif_b200 import frontend
plan = frontend.get_plan()
clips = 256
for n in (2646000, 2646004, 2646001):
    pcm = torch.randint(-3000, 3000, (clips * n,), dtype=torch.int16, device='cuda')
    offs = [i * n for i in range(clips + 1)]
    _, tile_off, _ = plan.query_batch([n] * clips)
    tiles = torch.empty((tile_off[-1], 1, 375, 1024), dtype=torch.float32, device='cuda')
    plan.set_profiling(True)
    for _ in range(4):
        plan.run_batch(pcm, offs, out=tiles)
    ms, runs = plan.get_profile_kernels()
    print(n, {k: round(v / runs, 3) for k, v in ms.items()})
    plan.set_profiling(False)
