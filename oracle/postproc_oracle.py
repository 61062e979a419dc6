"""CPU oracle (numpy, float32 arithmetic) for the detector post-processing.

TEST INFRASTRUCTURE: checker only, never on the product path (see
``oracle/__init__.py``).  Restates

  make_anchors        <- generate_anchors_frcnn / get_anchor_shifts_frcnn  nets_utils.py:35-59
                         as combined in ProposalLayer.forward             layers.py:252-258
  decode              <- bbox_reg_to_coord                                nets_utils.py:169-186
  iou_row / greedy    <- batch_self_overlap + the greedy loop of nms      nets_utils.py:189-232
  nms                 <- nms (batch-min truncation, return_idx)           nets_utils.py:210-245
  proposal_layer      <- ProposalLayer.forward (eval)                     layers.py:226-303
  final_detections    <- FastRCNN.forward inference branch                layers.py:688-778
  merge_images        <- merge_images                                     run_detection.py:163-249

Parity status: PINNED -- each function is compared with the reference's own function run
in the build container (tests/test_oracle_postproc.py, container only) and with the
committed vectors in tests/golden/postproc_*.npz produced by oracle/make_golden.py.
All box arithmetic is done in float32 one IEEE operation at a time (no FMA), which is what
eager PyTorch does; sorting is a STABLE descending sort (ties -> lower index first), the
reference's ``argsort(descending=True)`` being unspecified on ties.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def make_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(1, 2, 4, 8, 16),
                 width=64, height=24, stride=16) -> np.ndarray:
    """All anchors [height*width*A, 4] (x1,y1,x2,y2), index = (y*width + x)*A + a."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales)
    side = np.sqrt(float(base_size) * float(base_size))
    wh = np.stack([np.sqrt(ratios), 1 / np.sqrt(ratios)], axis=1) * side          # [R, 2]
    whs = (wh.reshape(-1) * scales[:, None]).reshape(-1, 2)                       # scale-major
    base = (np.concatenate([-whs / 2, whs / 2], axis=1) + int(base_size / 2)).astype(int)
    xs, ys = np.arange(width) * stride, np.arange(height) * stride
    shift = np.stack([np.tile(xs, height), np.repeat(ys, width)], axis=1)         # [K, 2]
    shift = np.tile(shift, 2)                                                     # [K, 4]
    return (base[None, :, :] + shift[:, None, :]).reshape(-1, 4).astype(np.float32)


def _round_half_even(x: np.ndarray) -> np.ndarray:
    return np.rint(x)     # torch.round is half-to-even


def decode(deltas: np.ndarray, anchors: np.ndarray) -> np.ndarray:
    """deltas [..., N, 4] float32, anchors [N, 4] float32 -> boxes [..., N, 4] float32."""
    d = np.asarray(deltas, dtype=np.float32)
    a = np.asarray(anchors, dtype=np.float32)
    wa = (a[:, 2] - a[:, 0]) + f32(1)
    ha = (a[:, 3] - a[:, 1]) + f32(1)
    xa = a[:, 0] + f32(0.5) * wa
    ya = a[:, 1] + f32(0.5) * ha
    x = d[..., 0] * wa + xa
    y = d[..., 1] * ha + ya
    w = np.exp(d[..., 2]) * wa
    h = np.exp(d[..., 3]) * ha
    hw, hh = f32(0.5) * w, f32(0.5) * h
    out = np.stack([_round_half_even(x - hw), _round_half_even(y - hh),
                    _round_half_even(x + hw), _round_half_even(y + hh)], axis=-1)
    return out.astype(np.float32)


def decode_tie_mask(deltas, anchors, ulps=4) -> np.ndarray:
    """True where some pre-round coordinate lies within `ulps` of k+0.5 (or exp() differs
    by an ulp between libms): the only places two correct decoders may disagree."""
    d = np.asarray(deltas, np.float64)
    a = np.asarray(anchors, np.float64)
    wa, ha = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    xa, ya = a[:, 0] + 0.5 * wa, a[:, 1] + 0.5 * ha
    x, y = d[..., 0] * wa + xa, d[..., 1] * ha + ya
    w, h = np.exp(d[..., 2]) * wa, np.exp(d[..., 3]) * ha
    pre = np.stack([x - 0.5 * w, y - 0.5 * h, x + 0.5 * w, y + 0.5 * h], -1)
    mag = np.maximum.reduce([np.abs(x), np.abs(y), w, h])[..., None] + 1.0
    frac = np.abs(pre - np.floor(pre) - 0.5)
    return (frac <= ulps * np.spacing(mag.astype(np.float32)).astype(np.float64)).any(-1)


def iou_row(boxes: np.ndarray, i: int, start: int | None = None) -> np.ndarray:
    """IoU (float32, +1 pixel convention) of box i against boxes[start:], one rounding per op,
    in the reference's operation order (nets_utils.py:193-205)."""
    b = np.asarray(boxes, dtype=np.float32)
    o = b if start is None else b[start:]
    xi = np.maximum((np.minimum(o[:, 2], b[i, 2]) - np.maximum(o[:, 0], b[i, 0])) + f32(1), f32(0))
    yi = np.maximum((np.minimum(o[:, 3], b[i, 3]) - np.maximum(o[:, 1], b[i, 1])) + f32(1), f32(0))
    inter = xi * yi
    area = ((b[:, 2] - b[:, 0]) + f32(1)) * ((b[:, 3] - b[:, 1]) + f32(1))
    ao = area if start is None else area[start:]
    union = (ao + area[i]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / union


def greedy(boxes: np.ndarray, thresh: float) -> list[int]:
    """In-order greedy suppression: box idx (ascending) is kept unless an earlier KEPT box has
    IoU >= float32(thresh) with it.  NaN IoU never suppresses."""
    b = np.asarray(boxes, dtype=np.float32)
    n = len(b)
    t = f32(thresh)
    dead = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        if i + 1 < n:
            dead[i + 1:] |= iou_row(b, i, i + 1) >= t
    return keep


def nms(bbox: np.ndarray, scores: np.ndarray, nms_thresh=0.7, post_nms_topN=300, return_idx=False):
    bbox = np.asarray(bbox, np.float32)
    scores = np.asarray(scores, np.float32)
    keeps = [greedy(bbox[b], nms_thresh) for b in range(len(bbox))]
    m = min(min(len(k) for k in keeps), post_nms_topN)
    out_s = np.stack([scores[b, keeps[b][:m]] for b in range(len(bbox))])
    out_b = np.stack([bbox[b, keeps[b][:m], :] for b in range(len(bbox))])
    return (out_b, out_s, keeps) if return_idx else (out_b, out_s)


def stable_desc_order(scores: np.ndarray) -> np.ndarray:
    return np.argsort(-np.asarray(scores, np.float32), axis=-1, kind="stable")


def proposal_layer(cls: np.ndarray, reg: np.ndarray, *, base_size=16, ratios=(0.5, 1, 2), n_layers=5,
                   anchor_stride=16, img_width=1024, img_height=375, min_threshold=5,
                   nms_thresh=0.7, pre_nms_topN=500, post_nms_topN=50, rcnn_batch_size=16):
    """cls [B, 2A, H, W] (softmaxed pairs), reg [B, 4A, H, W] -> (rois [B,M,4], scores [B,M]) or
    (empty, empty) when fewer than rcnn_batch_size candidates survive (layers.py:288-290)."""
    B, _, H, W = cls.shape
    anchors = make_anchors(base_size, ratios, 2 ** np.arange(n_layers), W, H, anchor_stride)
    N = anchors.shape[0]
    scores = np.transpose(cls, (0, 2, 3, 1)).reshape(B, N, 2)[..., 1].astype(np.float32)
    deltas = np.transpose(reg, (0, 2, 3, 1)).reshape(B, N, 4).astype(np.float32)
    boxes = decode(deltas, anchors)
    boxes[..., [0, 2]] = np.clip(boxes[..., [0, 2]], 0, img_width - 1)
    boxes[..., [1, 3]] = np.clip(boxes[..., [1, 3]], 0, img_height - 1)
    keep = ((boxes[..., 2] - boxes[..., 0] + f32(1)) >= min_threshold) & \
           ((boxes[..., 3] - boxes[..., 1] + f32(1)) >= min_threshold)
    pre = min(pre_nms_topN, int(keep.sum(axis=1).min()))
    if pre < rcnn_batch_size:
        return np.zeros((0,), np.float32), np.zeros((0,), np.float32)
    order = stable_desc_order(scores)
    sel = np.stack([order[b][keep[b, order[b]]][:pre] for b in range(B)])
    s = np.stack([scores[b, sel[b]] for b in range(B)])
    bx = np.stack([boxes[b, sel[b]] for b in range(B)])
    return nms(bx, s, nms_thresh, post_nms_topN)


def final_detections(bbox_reg: np.ndarray, probs: np.ndarray, rois: np.ndarray, *, num_classes=150,
                     img_width=1024, img_height=375, proposal_number=50, nms_thresh=0.3,
                     min_score=0.5) -> list[dict]:
    """bbox_reg [B*R, 4*(C+1)], probs [B*R, C+1], rois [B, R, 4] -> list(B) of
    {str(c): {'bbox_coord': [n,4] f32, 'scores': [1,n] f32}} for c in 1..C (empty -> shape (0,))."""
    B, R = rois.shape[:2]
    probs = np.asarray(probs, np.float32)
    cls = probs.argmax(axis=1)                       # first maximum, like torch.max(dim=1)
    sc = probs[np.arange(len(probs)), cls]
    reg = np.asarray(bbox_reg, np.float32).reshape(B * R, num_classes + 1, 4)[np.arange(B * R), cls]
    reg, cls, sc = reg.reshape(B, R, 4), cls.reshape(B, R), sc.reshape(B, R)
    order = stable_desc_order(sc)
    out = []
    for b in range(B):
        box = decode(reg[b][None], rois[b])[0]
        box[:, [0, 2]] = np.clip(box[:, [0, 2]], 0, img_width - 1)
        box[:, [1, 3]] = np.clip(box[:, [1, 3]], 0, img_height - 1)
        s_s, s_b, s_c = sc[b, order[b]], box[order[b]], cls[b, order[b]]
        nz = np.nonzero(s_c > 0)[0]
        if len(nz) > 0:
            nb, ns, idx = nms(s_b[nz][None], s_s[nz][None], nms_thresh, len(s_b), True)
            s_b, s_s, s_c = nb[0], ns[0], s_c[nz][idx[0]]
        res = {}
        for c in range(1, num_classes + 1):
            w = np.nonzero(s_c == c)[0]
            if len(w) == 0:
                res[str(c)] = dict(bbox_coord=np.zeros((0,), np.float32), scores=np.zeros((0,), np.float32))
                continue
            cb, cs = nms(s_b[w][None], s_s[w][None], nms_thresh, proposal_number)
            ok = np.nonzero(cs[0] > f32(min_score))[0]
            if len(ok) == 0:
                res[str(c)] = dict(bbox_coord=np.zeros((0,), np.float32), scores=np.zeros((0,), np.float32))
            else:
                res[str(c)] = dict(bbox_coord=cb[0][ok], scores=cs[:, ok])
        out.append(res)
    return out


def merge_images(tiles_out: list[dict], *, w_pix=1024, hop_spectro=819, spectrogram_length: int,
                 num_classes=150, nms_thresh=0.3) -> dict:
    """Per-tile dicts (flattened over batches, file order) -> per-file dict
    {str(c): {'bbox_coord': [m,4], 'scores': [m]}}; class-major candidate order, one
    class-agnostic in-order NMS (run_detection.py:180-247)."""
    min_border = 0.9 * (w_pix - hop_spectro)
    n = len(tiles_out)
    boxes, scores, species = [], [], []
    for c in range(1, num_classes + 1):
        for i, t in enumerate(tiles_out):
            bb = np.asarray(t[str(c)]["bbox_coord"], np.float32)
            if len(bb) == 0:
                continue
            ss = np.asarray(t[str(c)]["scores"], np.float32).reshape(-1)
            bb = bb.reshape(-1, 4).copy()
            width = bb[:, 2] - bb[:, 0]
            right, left = bb[:, 2] >= w_pix - 5, bb[:, 0] <= 4
            edge = right if i == 0 else (left if i == n - 1 else (left | right))
            ok = ~(edge & (width < min_border))
            bb, ss = bb[ok], ss[ok]
            if len(bb) == 0:
                continue
            bb[:, 0] += f32(hop_spectro * i)
            bb[:, 2] += f32(hop_spectro * i)
            ok = ~(bb[:, 2] >= spectrogram_length)
            bb, ss = bb[ok], ss[ok]
            if len(bb) == 0:
                continue
            boxes.append(bb); scores.append(ss); species += [c] * len(bb)
    empty = lambda: dict(bbox_coord=np.zeros((0,), np.float32), scores=np.zeros((0,), np.float32))
    if not boxes:
        return {str(c): empty() for c in range(1, num_classes + 1)}
    boxes, scores, species = np.concatenate(boxes), np.concatenate(scores), np.asarray(species)
    keep = greedy(boxes, nms_thresh)
    kb, ks, ksp = boxes[keep], scores[keep], species[keep]
    return {str(c): (dict(bbox_coord=kb[ksp == c], scores=ks[ksp == c]) if (ksp == c).any() else empty())
            for c in range(1, num_classes + 1)}


# ------------------------------------------------------------------------------ RoI pooling ----
def positional_encoding_1d(length: int, cn: int, temp=10000) -> np.ndarray:
    """position_encoding.py:10-15 in float32 numpy (sin/cos may differ from torch by an ulp; tests
    feed the kernel and this oracle the SAME table)."""
    pos = np.arange(1, length + 1, dtype=np.float32)
    dt = (np.float32(temp) ** (2 * (np.arange(cn, dtype=np.float32) // 2) / np.float32(cn))).astype(np.float32)
    posenc = (pos[:, None] / dt[None, :]).astype(np.float32)
    pe = np.stack([np.sin(posenc[:, 0::2]), np.cos(posenc[:, 1::2])], axis=2).reshape(length, -1)
    return pe.astype(np.float32)


def _adaptive_avg_pool(x: np.ndarray, ph: int, pw: int) -> np.ndarray:
    """ATen cpu_adaptive_avg_pool2d on x [C, H, W]: window summed row by row in float32, then / kh / kw."""
    Cn, H, W = x.shape
    out = np.zeros((Cn, ph, pw), dtype=np.float32)
    for oh in range(ph):
        ih0, ih1 = (oh * H) // ph, -((-(oh + 1) * H) // ph)
        for ow in range(pw):
            iw0, iw1 = (ow * W) // pw, -((-(ow + 1) * W) // pw)
            acc = np.zeros(Cn, dtype=np.float32)
            for ih in range(ih0, ih1):
                for iw in range(iw0, iw1):
                    acc = (acc + x[:, ih, iw]).astype(np.float32)
            out[:, oh, ow] = (acc / np.float32(ih1 - ih0)).astype(np.float32) / np.float32(iw1 - iw0)
    return out


def roi_pool(rois: np.ndarray, feats: list, pe_freq: np.ndarray, pe_time: np.ndarray, *, n_layers=5, pool_h=2, pool_w=2):
    """ROIPooling.forward (layers.py:399-497).  rois [B, R, 4] float32, feats: n_layers arrays [B, C, H_l, W_l]."""
    rois = np.asarray(rois, np.float32)
    B, R = rois.shape[:2]
    Cn = feats[0].shape[1]
    pool = np.zeros((B, R, Cn, pool_h, pool_w), np.float32)
    pe = np.zeros_like(pool)
    lvl = np.zeros((B, R), np.int32)
    ln2 = np.float32(np.log(2))
    for b in range(B):
        for i in range(R):
            x1, y1, x2, y2 = [np.float32(v) for v in rois[b, i]]
            side = np.sqrt(np.float32(np.float32(x2 - x1) * np.float32(y2 - y1)))
            with np.errstate(divide="ignore", invalid="ignore"):
                lv = np.float32(np.log(np.float32(side * np.float32(0.1)))) / ln2
            level = int(np.clip(int(lv) if np.isfinite(lv) else 0, 0, n_layers - 1))
            s = 2 ** (level + 1)
            fx1, fy1, fx2, fy2 = [int(np.rint(np.float32(v) / np.float32(s))) for v in (x1, y1, x2, y2)]
            H, W = feats[level].shape[-2:]
            fy2 = min(fy2, H - 1)
            while fy2 - fy1 + 1 < pool_h:
                fy1 = max(0, fy1 - 1); fy2 = min(H - 1, fy2 + 1)
            while fx2 - fx1 + 1 < pool_w:
                fx1 = max(0, fx1 - 1); fx2 = min(W - 1, fx2 + 1)
            lvl[b, i] = level
            pool[b, i] = _adaptive_avg_pool(feats[level][b, :, fy1:fy2 + 1, fx1:fx2 + 1].astype(np.float32), pool_h, pool_w)
            fpe, tpe = pe_freq[s * fy1:s * fy2], pe_time[:s * (fx2 - fx1)]
            roi_pe = np.concatenate([np.broadcast_to(fpe[:, None, :], (len(fpe), len(tpe), fpe.shape[1])),
                                     np.broadcast_to(tpe[None, :, :], (len(fpe), len(tpe), tpe.shape[1]))], axis=-1)
            pe[b, i] = _adaptive_avg_pool(np.ascontiguousarray(np.transpose(roi_pe, (2, 0, 1))), pool_h, pool_w)
    return pool, pe, lvl
