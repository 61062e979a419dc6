"""How deep below its frame's level must a pixel be before float32 misses the 1e-4 tolerance?  Runs the front-end on a few
60 s clips for several flag levels (NBM_REFINE_REL_DB, read at plan creation) and reports the worst normalised-tile error
against the oracle, the pixels above 1e-4, and the time of the refinement stage on a 256-clip batch.

    python scripts/refine_probe.py [n_clips]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import frontend, synth  # noqa: E402
from oracle import frontend_oracle as fo  # noqa: E402

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pcms = [synth.synth_pcm(60.0, 300 + i, calls_per_s=2.0 + i) for i in range(n_clips)]
refs = [np.stack(fo.process(p).tiles) for p in pcms]
flat = torch.from_numpy(np.concatenate(pcms)).cuda()
offs = np.concatenate([[0], np.cumsum([len(p) for p in pcms])]).tolist()
big = flat.repeat(max(1, 256 // n_clips))
big_offs = (np.arange(big.numel() // len(pcms[0]) + 1) * len(pcms[0])).tolist()
for rel in ("-200", "-62"):
    os.environ["NBM_REFINE_REL_DB"] = rel
    plan = frontend.FrontendPlan()
    tiles, toff, mm = plan.run_batch(flat, offs)
    torch.cuda.synchronize()
    worst, over, sq, cnt = 0.0, 0, 0.0, 0
    for i, ref in enumerate(refs):
        err = np.abs(tiles[toff[i]:toff[i + 1], 0].cpu().numpy().astype(np.float64) - ref)
        worst = max(worst, err.max()); over += int((err > 1e-4).sum()); sq += (err ** 2).sum(); cnt += err.size
    del tiles
    out = torch.empty((big_offs[-1] // len(pcms[0]) * 25, 1, 375, 1024), dtype=torch.float32, device="cuda")
    plan.run_batch(big, big_offs, out=out)
    plan.set_profiling(True)
    for _ in range(3):
        plan.run_batch(big, big_offs, out=out)
    k, runs = plan.get_profile_kernels()
    plan.set_profiling(False)
    listed, cap, n_px = plan.last_listed()
    print(f"rel {rel:>5s} dB: worst {worst:.3e}  >1e-4: {over:6d} of {cnt}  rms {np.sqrt(sq / cnt):.2e} | {len(big_offs) - 1} clips: "
          f"anchors {k['anchor'] / runs:.2f} slides {k['stft'] / runs:.2f} refine+minmax {k['minmax'] / runs:.3f} tiles {k['tile'] / runs:.2f} ms, {listed} blocks listed (cap {cap}), {n_px} pixels recomputed")
    del out
    plan.close()
