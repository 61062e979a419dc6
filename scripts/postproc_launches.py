"""Every post-processing kernel of one detector batch (bs 4), a few launches each, for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/postproc_launches.py
RPN NMS 4 x 500 @ 0.7, final NMS 4 x 50 @ 0.3, file-merge NMS 1 x 5000 @ 0.3, decode, fused ProposalLayer, ROIPooling, fused
FastRCNN tail, per-file merge."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import postproc as pp, synth  # noqa: E402

rng = np.random.default_rng(0)
REPS = 5


def boxes(B, N, span=600):
    x1 = rng.integers(0, span, (B, N)); y1 = rng.integers(0, 300, (B, N))
    return torch.from_numpy(np.stack([x1, y1, x1 + rng.integers(5, 90, (B, N)), y1 + rng.integers(5, 60, (B, N))], -1).astype(np.float32)).cuda()


cfg = synth.default_args("cuda")
for B, N, th in ((4, 500, 0.7), (4, 50, 0.3), (1, 5000, 0.3)):
    b = boxes(B, N, 600 if N < 1000 else 900)
    for _ in range(REPS):
        pp.nms_keep(b, th)
g = torch.Generator(device="cuda").manual_seed(0)
cls = torch.rand((4, 15, 2, 24, 64), device="cuda", generator=g).softmax(2).reshape(4, 30, 24, 64)
reg = torch.randn((4, 60, 24, 64), device="cuda", generator=g) * 0.2
anchors = pp.make_anchors(cfg.base_size, cfg.ratios, 2 ** np.arange(cfg.n_layers), 64, 24, cfg.anchor_stride, torch.device("cuda"))
deltas = torch.randn((4, anchors.shape[0], 4), device="cuda", generator=g) * 0.2
layer = pp.ProposalLayer(cfg, cfg.n_layers).eval()
rp = pp.ROIPooling(cfg, want_levels=False)
feats = [torch.from_numpy(f).cuda() for f in synth.fpn_features(78, 4, cfg.out_fpn_chan, cfg.n_layers)]
for _ in range(REPS):
    pp.bbox_reg_to_coord(deltas, anchors)
    rois, sc = layer(cls, reg)
    rp(rois, feats)
    R = rois.shape[1]
    bbox_reg = torch.randn((4 * R, 4 * (cfg.num_classes + 1)), device="cuda", generator=g) * 0.1
    probs = torch.rand((4 * R, cfg.num_classes + 1), device="cuda", generator=g).softmax(1)
    rec = pp.final_detections_flat(bbox_reg, probs, rois, cfg.num_classes, cfg.img_width, cfg.img_height, 0.3, 0.0)
n = 3000
mb = boxes(1, n, 900)[0]
for _ in range(REPS):
    pp.merge_flat(mb, torch.rand(n, device="cuda", generator=g), torch.randint(1, 151, (n,), device="cuda", generator=g, dtype=torch.int32),
                  torch.sort(torch.randint(0, 245, (n,), device="cuda", generator=g, dtype=torch.int32))[0], 245, 1024, 819, 200455)
torch.cuda.synchronize()
print("rois", tuple(rois.shape), "done")
