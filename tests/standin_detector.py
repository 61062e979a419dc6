"""TEST DOUBLE for the detector network (tests only, not product code).

The reference's CNN (nbm_model/nets) is not present on the GPU box and its checkpoint is a
Git-LFS stub, so the pipeline tests drive the host driver with this small seeded network that
honours the same contract: ``model(batch[:, None], min_score=...)`` -> list of per-image dicts.
Its two heads emit tensors with the reference's shapes (RPN cls [B,30,24,64] softmaxed pairs,
reg [B,60,24,64]; RCNN bbox_reg [B*R,604], probs [B*R,151]); the post-processing between and
after them is pluggable: ``backend='nbm'`` = libnbm_b200 (product), ``backend='oracle'`` = the
CPU oracle fed with the SAME head outputs, so the two can be compared bit for bit."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from birdsoundclassif_b200 import synth


class StandInDetector(nn.Module):
    def __init__(self, args=None, backend="nbm", seed=0):
        super().__init__()
        self.args = args or synth.default_args("cuda")
        self.backend = backend
        g = torch.Generator().manual_seed(seed)
        self.rpn = nn.Conv2d(4, 90, 1)
        self.rcnn = nn.Linear(8, 151 * 5)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.8)
        self.eval()
        if backend == "nbm":
            from birdsoundclassif_b200 import postproc
            self.prop = postproc.ProposalLayer(self.args, self.args.n_layers).eval()

    def _features(self, x):
        p = F.adaptive_avg_pool2d(x, (24, 64))
        q = F.adaptive_max_pool2d(x, (24, 64))
        return torch.cat([p, q, p * p, (q - p)], dim=1) * 4

    def first_stage(self, x):
        B = x.shape[0]
        o = self.rpn(self._features(x))
        cls = o[:, :30].reshape(B, 15, 2, 24, 64).softmax(2).reshape(B, 30, 24, 64)
        reg = o[:, 30:] * 0.15
        return cls, reg

    def second_stage(self, x, rois):
        B, R = rois.shape[:2]
        m = x.mean(dim=(1, 2, 3))
        f = torch.cat([rois / 512.0, (rois[..., 2:] - rois[..., :2]) / 256.0, m[:, None, None].expand(B, R, 2)], dim=-1)
        o = self.rcnn(f.reshape(B * R, 8))
        logits = o[:, :151] * 3
        logits[:, 12:] -= 6          # concentrate mass on a dozen classes so scores clear min_score
        return (o[:, 151:] * 0.1).reshape(B * R, 604).contiguous(), logits.softmax(1)

    @torch.no_grad()
    def forward(self, x, nms_thresh=0.3, min_score=0.5):
        a = self.args
        cls, reg = self.first_stage(x)
        if self.backend == "nbm":
            from birdsoundclassif_b200 import postproc
            rois, _ = self.prop(cls, reg)
            bbox_reg, probs = self.second_stage(x, rois)
            return postproc.fastrcnn_inference_tail(bbox_reg, probs, rois, a, nms_thresh, min_score)
        from oracle import postproc_oracle as po
        rois, _ = po.proposal_layer(cls.cpu().numpy(), reg.cpu().numpy())
        rois_t = torch.from_numpy(rois).to(x.device)
        bbox_reg, probs = self.second_stage(x, rois_t)
        dets = po.final_detections(bbox_reg.cpu().numpy(), probs.cpu().numpy(), rois, nms_thresh=nms_thresh,
                                   min_score=min_score)
        return [{k: {kk: torch.from_numpy(np.ascontiguousarray(vv)) for kk, vv in v.items()} for k, v in d.items()}
                for d in dets]
