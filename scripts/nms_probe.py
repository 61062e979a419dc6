"""Kernel times of the greedy NMS for one shape (run under `ncu --metrics gpu__time_duration.sum`):
    python scripts/nms_probe.py B N thresh [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import postproc as pp  # noqa: E402

B, N, th = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rng = np.random.default_rng(0)
x1 = rng.integers(0, 600, (B, N)); y1 = rng.integers(0, 300, (B, N))
boxes = torch.from_numpy(np.stack([x1, y1, x1 + rng.integers(5, 90, (B, N)), y1 + rng.integers(5, 60, (B, N))], -1).astype(np.float32)).cuda()
for _ in range(reps):
    keep_idx, keep_cnt = pp.nms_keep(boxes, th)
torch.cuda.synchronize()
print("kept", keep_cnt.tolist())
