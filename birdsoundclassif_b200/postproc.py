"""Detector post-processing on the GPU behind the reference's Python symbols.

Same names, argument meaning and return conventions as the reference functions they
replace, so they can be patched into ``nbm_model.nets.layers`` / ``nbm_model.run_detection``
(see INTEGRATION.md):

  bbox_reg_to_coord   nets_utils.py:169-186
  nms                 nets_utils.py:210-245   (in-order greedy, batch-min truncation, return_idx)
  ProposalLayer       layers.py:219-303       (parameter-free nn.Module)
  fastrcnn_inference_tail   layers.py:688-778 (the inference branch of FastRCNN.forward)
  merge_images        run_detection.py:163-249
  ROIPooling          layers.py:399-497       (parameter-free nn.Module)

All arithmetic is done by libnbm_b200.so on the tensors' CUDA device; torch only owns the
memory.  There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import gc

import numpy as np
import torch
from torch import nn

from . import _lib


def _stream() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.NbmError("libnbm_b200 post-processing needs CUDA tensors (no CPU fallback)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


# --------------------------------------------------------------------------- host hygiene -----
@contextlib.contextmanager
def gc_paused():
    """Python's cyclic collector off for the duration of a detection loop (restored on exit if it was on).

    The reference's output format is a dictionary of 150 dictionaries per tile, kept until the per-file merge: a
    ten-minute recording holds ~150 000 of them, and every full collection the allocation counters trigger walks all of
    them again -- 12 % of the wall clock of a graph-replayed ten-minute file and 23 % of an eager one went there
    (scripts/detect_probe.py, 248 batches: 13.2 -> 11.5 and 24.6 -> 18.9 ms per batch).  Nothing in these loops builds
    reference cycles; reference counting frees everything as before."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


# ------------------------------------------------------------------------------- anchors ------
_ANCHORS: dict = {}


def make_anchors(base_size, ratios, scales, width, height, stride, device) -> torch.Tensor:
    """[height*width*A, 4] float32 anchors (nets_utils.py:35-59 as combined at layers.py:252-258);
    built once per configuration instead of every forward (layers.py:252-271)."""
    key = (base_size, tuple(float(r) for r in ratios), tuple(int(s) for s in scales), width, height, stride,
           str(device))
    if key not in _ANCHORS:
        r = (C.c_double * len(ratios))(*[float(x) for x in ratios])
        s = (C.c_int64 * len(scales))(*[int(x) for x in scales])
        out = np.empty((height * width * len(ratios) * len(scales), 4), dtype=np.float32)
        _lib.check(_lib.lib().nbm_make_anchors(int(base_size), r, len(ratios), s, len(scales), int(width),
                                               int(height), int(stride), out.ctypes.data), "nbm_make_anchors")
        _ANCHORS[key] = torch.from_numpy(out).to(device)
    return _ANCHORS[key]


# -------------------------------------------------------------------------------- decode ------
def decode_boxes(deltas: torch.Tensor, anchors: torch.Tensor, clip_w=0.0, clip_h=0.0, min_size=0.0,
                 want_valid=False):
    """deltas [B,N,4]; anchors [N,4] (shared) or [B,N,4] (per image)."""
    _need_cuda(deltas, anchors)
    deltas, anchors = _f32c(deltas), _f32c(anchors)
    B, N = deltas.shape[0], deltas.shape[1]
    boxes = torch.empty_like(deltas)
    valid = torch.empty((B, N), dtype=torch.uint8, device=deltas.device) if want_valid else None
    with torch.cuda.device(deltas.device):
        _lib.check(_lib.lib().nbm_decode_boxes(deltas.data_ptr(), anchors.data_ptr(), B, N, int(anchors.dim() == 3),
                                               float(clip_w), float(clip_h), float(min_size), boxes.data_ptr(),
                                               valid.data_ptr() if want_valid else None, _stream()),
                   "nbm_decode_boxes")
    return (boxes, valid) if want_valid else boxes


def bbox_reg_to_coord(bbox_pred: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """Drop-in for nets_utils.bbox_reg_to_coord: bbox_pred [..., N, 4], anchors [N, 4] -> [B, N, 4]."""
    return decode_boxes(bbox_pred.reshape(-1, bbox_pred.shape[-2], 4), anchors)


# ----------------------------------------------------------------------------------- NMS ------
_WS: dict = {}          # (device, stream) -> grow-only scratch tensor for the NMS bit matrix


def _scratch(dev, nbytes: int) -> torch.Tensor:
    """Scratch memory reused across calls on the same stream (same-stream calls are ordered, so one buffer serves
    them all; the reference's nms allocates eight N x N temporaries per call, nets_utils.py:193-205)."""
    key = (dev, _stream())
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        t = _WS[key] = torch.empty((max(int(nbytes * 1.5), 1 << 16),), dtype=torch.uint8, device=dev)
    return t


def nms_keep(boxes: torch.Tensor, thresh: float, n_valid: torch.Tensor | None = None, packed: bool = False):
    """boxes [B,N,4] in priority order -> (keep_idx int32 [B,N] (first keep_cnt[b] valid), keep_cnt int32 [B]);
    with ``packed`` also the int32 buffer [B + B*N] holding both (counts first), for a single device-to-host copy."""
    _need_cuda(boxes)
    boxes = _f32c(boxes)
    B, N = boxes.shape[0], boxes.shape[1]
    Nn = max(N, 1)
    buf = torch.empty((B * (Nn + 1),), dtype=torch.int32, device=boxes.device)
    keep_cnt, keep_idx = buf[:B], buf[B:].view(B, Nn)
    ws_bytes = _lib.lib().nbm_nms_workspace_bytes(B, N)
    ws = _scratch(boxes.device, max(ws_bytes, 8))
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.lib().nbm_nms_greedy(boxes.data_ptr(), n_valid.data_ptr() if n_valid is not None else None,
                                             B, N, float(thresh), keep_idx.data_ptr(), keep_cnt.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream()), "nbm_nms_greedy")
    return (keep_idx, keep_cnt, buf) if packed else (keep_idx, keep_cnt)


def nms(bbox_pred: torch.Tensor, scores: torch.Tensor, nms_thresh=0.7, post_nms_topN=300, return_idx=False):
    """Drop-in for nets_utils.nms.  Greedy IN INPUT ORDER (no sort), IoU >= thresh suppresses,
    every row truncated to min(min_b len(keep_b), post_nms_topN); with return_idx the untruncated
    keep lists come back as list[list[int]] (callers fancy-index with them, layers.py:746).  One device-to-host
    copy per call (the counts, or counts + index table with return_idx)."""
    keep_idx, keep_cnt, buf = nms_keep(bbox_pred, nms_thresh, packed=True)
    B = keep_cnt.shape[0]
    if return_idx:
        host = buf.cpu().numpy()
        cnt = host[:B].tolist()
        rows = host[B:].reshape(B, -1)
    else:
        cnt = keep_cnt.tolist()
    m = min(min(cnt), int(post_nms_topN))
    sel = keep_idx[:, :m].long()
    out_scores = torch.gather(scores, 1, sel)
    out_boxes = torch.gather(bbox_pred, 1, sel[..., None].expand(-1, -1, 4))
    if return_idx:
        return out_boxes, out_scores, [rows[b, :cnt[b]].tolist() for b in range(B)]
    return out_boxes, out_scores


# ------------------------------------------------------------------------------ proposals -----
class ProposalLayer(nn.Module):
    """Drop-in for layers.ProposalLayer (eval branch): decode 15*24*64 anchors per image, clamp,
    min-size filter, stable score sort, batch-coupled top-N, NMS, batch-coupled truncation -- one
    library call, no per-image Python.  Holds no parameters (state_dict keys unaffected)."""

    def __init__(self, config, n_layers):
        super().__init__()
        self.n_layers = n_layers
        self.config = config
        self._ws = None

    def _launch(self, labels_pred, bbox_reg, M_dev=None, ws_holder=None):
        cfg = self.config
        if self.training:
            raise NotImplementedError("training-time proposals are out of scope (inference hot path only)")
        _need_cuda(labels_pred, bbox_reg)
        B = labels_pred.shape[0]
        H, W = labels_pred.shape[-2:]
        scales = 2 ** np.arange(self.n_layers)
        A = len(cfg.ratios) * len(scales)
        dev = labels_pred.device
        anchors = make_anchors(cfg.base_size, cfg.ratios, scales, W, H, cfg.anchor_stride, dev)
        p = _lib.ProposalParams(A=A, H=H, W=W, img_width=float(cfg.img_width), img_height=float(cfg.img_height),
                                min_size=float(cfg.min_threshold), nms_thresh=float(cfg.nms_thresh),
                                pre_nms_topN=int(cfg.pre_nms_topN_eval), post_nms_topN=int(cfg.post_nms_topN_eval),
                                rcnn_batch_size=int(cfg.rcnn_batch_size))
        ws_bytes = _lib.lib().nbm_proposals_workspace_bytes(C.byref(p), B)
        if ws_holder is not None:           # a caller that keeps several calls in flight brings a workspace per lane
            if ws_holder[0] is None or ws_holder[0].numel() < ws_bytes or ws_holder[0].device != dev:
                ws_holder[0] = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            ws = ws_holder[0]
        else:
            if self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
                self._ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            ws = self._ws
        rois = torch.empty((B, p.post_nms_topN, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((B, p.post_nms_topN), dtype=torch.float32, device=dev)
        cls, reg = _f32c(labels_pred), _f32c(bbox_reg)
        with torch.cuda.device(dev):
            if M_dev is None:
                M = C.c_int32(0)
                _lib.check(_lib.lib().nbm_proposals(C.byref(p), cls.data_ptr(), reg.data_ptr(), anchors.data_ptr(), B,
                                                    rois.data_ptr(), scores.data_ptr(), C.byref(M), ws.data_ptr(),
                                                    ws.numel(), _stream()), "nbm_proposals")
                return rois, scores, M.value
            _lib.check(_lib.lib().nbm_proposals_async(C.byref(p), cls.data_ptr(), reg.data_ptr(), anchors.data_ptr(), B,
                                                      rois.data_ptr(), scores.data_ptr(), M_dev.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), _stream()),
                       "nbm_proposals_async")
        return rois, scores, M_dev

    def forward(self, labels_pred: torch.Tensor, bbox_reg: torch.Tensor):
        rois, scores, M = self._launch(labels_pred, bbox_reg)
        if M < 0:
            print("Not enough possible RoIs, RPN failed")                  # layers.py:288-290
            dev = labels_pred.device
            return torch.tensor([]).to(dev), torch.tensor([]).to(dev)
        return rois[:, :M], scores[:, :M]

    def forward_async(self, labels_pred: torch.Tensor, bbox_reg: torch.Tensor, M_dev: torch.Tensor, ws_holder=None):
        """No host read, no synchronisation (CUDA-graph capturable): returns the FULL ``rois [B, post_nms_topN, 4]``
        and ``scores [B, post_nms_topN]``; the number of valid rows (or -1 for the reference's "RPN failed" branch)
        is written to ``M_dev`` (int32 [1] on the device).  ``ws_holder`` (a one-element list) keeps the call's
        workspace apart from the module's own, for callers with several calls in flight on different streams."""
        rois, scores, _ = self._launch(labels_pred, bbox_reg, M_dev, ws_holder)
        return rois, scores


# ----------------------------------------------------------------------------- final tail -----
def final_detections_flat(bbox_reg, bbox_classes, rois, num_classes, img_width, img_height,
                          nms_thresh=0.3, min_score=0.5):
    """Fused inference tail -> flat records (boxes [B,R,4], scores [B,R], classes int32 [B,R],
    counts int32 [B]); the first counts[b] rows of image b are its detections in surviving
    (score-descending) order."""
    _need_cuda(bbox_reg, bbox_classes, rois)
    B, R = rois.shape[0], rois.shape[1]
    dev = rois.device
    bbox_reg, bbox_classes, rois = _f32c(bbox_reg), _f32c(bbox_classes), _f32c(rois)
    boxes = torch.empty((B, R, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((B, R), dtype=torch.float32, device=dev)
    classes = torch.empty((B, R), dtype=torch.int32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nbm_final_detections(bbox_reg.data_ptr(), bbox_classes.data_ptr(), rois.data_ptr(), B, R,
                                                   int(num_classes), float(img_width), float(img_height),
                                                   float(nms_thresh), float(min_score), boxes.data_ptr(),
                                                   scores.data_ptr(), classes.data_ptr(), counts.data_ptr(),
                                                   _stream()), "nbm_final_detections")
    return boxes, scores, classes, counts


_EMPTY = torch.Tensor()


class TileDetections(dict):
    """The reference's per-image dictionary plus the flat record it was built from (boxes [n,4], scores [n],
    classes int32 [n], device tensors in surviving order), so that the per-file merge can take the records as
    they are instead of re-collecting them from 150 dictionary entries per tile.  ``flat`` describes the
    dictionary as it was built; code that edits the entries afterwards must set it to None."""
    __slots__ = ("flat",)


def records_sort_device(boxes, scores, classes, counts, num_classes):
    """Device half of ``records_to_dicts`` (no host read: CUDA-graph capturable): one stable sort by class for the
    whole batch -> (skey int32 [B,R] sorted class keys with num_classes+1 for dead rows, sb [B,R,4], ss [B,R])."""
    B, R = classes.shape
    invalid = int(num_classes) + 1
    key = torch.where(torch.arange(R, device=boxes.device)[None] < counts[:, None], classes, invalid)
    skey, order = torch.sort(key, dim=1, stable=True)               # per-class order = surviving order
    sb = torch.gather(boxes, 1, order[..., None].expand(-1, -1, 4))
    ss = torch.gather(scores, 1, order)
    return skey, sb, ss


def records_build_dicts(skey_h, sb, ss, boxes, scores, classes, num_classes, proposal_number=None):
    """Host half of ``records_to_dicts``: ``skey_h`` is the sorted key table on the host (numpy [B,R])."""
    invalid = int(num_classes) + 1
    out = []
    for b in range(skey_h.shape[0]):
        row = skey_h[b]
        n = int((row < invalid).sum())
        # empty classes: CPU torch.Tensor() like the reference (layers.py:753-755, 765-766) -- ONE shared empty tensor, read-only
        # by convention, instead of 300 new ones per tile
        d = TileDetections((str(c), dict(bbox_coord=_EMPTY, scores=_EMPTY)) for c in range(1, num_classes + 1))
        truncated = False
        if n:
            starts = [0] + (np.flatnonzero(row[1:n] != row[:n - 1]) + 1).tolist() + [n]
            for s0, s1 in zip(starts[:-1], starts[1:]):
                if proposal_number is not None and s1 - s0 > proposal_number:
                    s1 = s0 + proposal_number
                    truncated = True
                d[str(int(row[s0]))] = dict(bbox_coord=sb[b, s0:s1], scores=ss[b, s0:s1][None])
        d.flat = None if truncated else (boxes[b, :n], scores[b, :n], classes[b, :n])
        out.append(d)
    return out


def records_to_dicts(boxes, scores, classes, counts, num_classes, proposal_number=None):
    """Flat records -> the reference's list(B) of {str(c): {'bbox_coord': [n,4], 'scores': [1,n]}}
    (layers.py:750-775); empty classes are CPU ``torch.Tensor()`` like the reference.  One stable sort by class
    for the whole batch and one device-to-host copy; the per-class entries are views of the sorted rows (the
    reference spends 150 iterations with a nonzero + sync each per image, layers.py:757-775)."""
    skey, sb, ss = records_sort_device(boxes, scores, classes, counts, num_classes)
    return records_build_dicts(skey.cpu().numpy(), sb, ss, boxes, scores, classes, num_classes, proposal_number)


def fastrcnn_inference_tail(bbox_reg, bbox_classes, rois, config, nms_thresh=0.3, min_score=0.5):
    """Drop-in for the inference branch of FastRCNN.forward given the head outputs."""
    rec = final_detections_flat(bbox_reg, bbox_classes, rois, config.num_classes, config.img_width,
                                config.img_height, nms_thresh, min_score)
    return records_to_dicts(*rec, config.num_classes, config.proposal_number)


# ---------------------------------------------------------------------------------- merge -----
def merge_flat(boxes, scores, classes, tiles, n_tiles, w_pix, hop_spectro, spectrogram_length, nms_thresh=0.3):
    """Flat per-file candidates in tile-major order -> survivors (boxes [m,4], scores [m], classes [m])
    in the reference's NMS order (class-major candidates, run_detection.py:180-233)."""
    n = boxes.shape[0]
    dev = boxes.device
    if n == 0:
        return boxes.new_zeros((0, 4)), scores.new_zeros((0,)), classes.new_zeros((0,))
    _need_cuda(boxes, scores, classes, tiles)
    boxes, scores = _f32c(boxes), _f32c(scores)
    classes, tiles = classes.to(torch.int32).contiguous(), tiles.to(torch.int32).contiguous()
    ob = torch.empty((n, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((n,), dtype=torch.float32, device=dev)
    oc = torch.empty((n,), dtype=torch.int32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = _scratch(dev, _lib.lib().nbm_merge_workspace_bytes(n))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nbm_merge_detections(boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(),
                                                   tiles.data_ptr(), n, int(n_tiles), int(w_pix), int(hop_spectro),
                                                   int(spectrogram_length), float(nms_thresh), ob.data_ptr(),
                                                   os_.data_ptr(), oc.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), _stream()), "nbm_merge_detections")
    m = int(cnt.item())
    return ob[:m], os_[:m], oc[:m]


def flatten_tile_dicts(out: list, num_classes: int, device):
    """list of per-tile dicts (the model's output format) -> flat (boxes, scores, classes, tiles),
    tile-major, per tile class-major (any within-tile order works: merge sorts by class stably and
    the per-class order inside a tile is preserved)."""
    if out and all(isinstance(d, TileDetections) and d.flat is not None for d in out):
        # the records the dictionaries were built from: three concatenations instead of 150 x n_tiles look-ups
        counts = torch.tensor([len(d.flat[1]) for d in out], dtype=torch.int64)
        tt = torch.repeat_interleave(torch.arange(len(out), dtype=torch.int32), counts).to(device, non_blocking=True)
        return (torch.cat([d.flat[0] for d in out]).to(device), torch.cat([d.flat[1] for d in out]).to(device),
                torch.cat([d.flat[2] for d in out]).to(device), tt)
    bb, ss, cc, tt = [], [], [], []
    for i, d in enumerate(out):
        for c in range(1, num_classes + 1):
            e = d[str(c)]
            n = len(e["bbox_coord"])
            if n == 0:
                continue
            bb.append(e["bbox_coord"].reshape(-1, 4).to(device))
            ss.append(e["scores"].reshape(-1).to(device))
            cc.append(torch.full((n,), c, dtype=torch.int32, device=device))
            tt.append(torch.full((n,), i, dtype=torch.int32, device=device))
    if not bb:
        z = torch.zeros((0,), device=device)
        return z.reshape(0, 4), z, z.to(torch.int32), z.to(torch.int32)
    return torch.cat(bb), torch.cat(ss), torch.cat(cc), torch.cat(tt)


def survivors_to_class_dict(boxes, scores, classes, num_classes):
    """-> {str(j): {'bbox_coord': [m,4], 'scores': [m]}} with ``torch.tensor([])`` for empty classes
    (run_detection.py:238-247)."""
    out = {}
    cls_host = classes.cpu()
    present = set(torch.unique(cls_host).tolist()) if len(cls_host) else set()
    for j in range(1, num_classes + 1):
        if j in present:
            w = torch.nonzero(cls_host == j)[:, 0].to(boxes.device)
            out[str(j)] = dict(bbox_coord=boxes[w], scores=scores[w])
        else:
            out[str(j)] = dict(bbox_coord=torch.tensor([]), scores=torch.tensor([]))
    return out


def _merge_survivors(fp, outputs, num_classes, nms_thresh):
    out = []
    for b in outputs:
        out.extend(b)
    dev = torch.device("cuda", torch.cuda.current_device())
    boxes, scores, classes, tiles = flatten_tile_dicts(out, num_classes, dev)
    return merge_flat(boxes, scores, classes, tiles, len(out), fp.W_PIX, fp.HOP_SPECTRO, int(fp.spectrogram_length), nms_thresh)


def merge_images(fp, outputs, num_classes, nms_thresh=0.3):
    """Drop-in for run_detection.merge_images: `outputs` is the list of per-batch lists of per-tile
    dicts the model returned; `fp` carries W_PIX, HOP_SPECTRO, spectrogram_length."""
    kb, ks, kc = _merge_survivors(fp, outputs, num_classes, nms_thresh)
    return survivors_to_class_dict(kb, ks, kc, num_classes)


def merge_to_output(fp, outputs, num_classes, reverse_dict, nms_thresh=0.3) -> dict:
    """merge_images followed by run_detection.py:69-77 in one step: the per-file output dictionary
    ``{species: {'bbox_coord': [[x1,y1,x2,y2],...], 'scores': [...]}}`` (classes ascending, empty ones left out)
    from ONE device-to-host copy of the survivors instead of two per detected class."""
    kb, ks, kc = _merge_survivors(fp, outputs, num_classes, nms_thresh)
    kb, ks, kc = kb.cpu().numpy(), ks.cpu().numpy(), kc.cpu().numpy()
    output = {}
    for idx in sorted(set(kc.tolist())):
        if 1 <= idx <= num_classes:
            w = kc == idx
            output[reverse_dict[idx]] = {"bbox_coord": kb[w].tolist(), "scores": ks[w].tolist()}
    return output


# ------------------------------------------------------------------------------ RoI pooling ----
def one_dimension_positional_encoding(length, cn, temp=10000):
    """position_encoding.py:10-15, restated (computed on the CPU like the reference, then moved)."""
    pos = torch.arange(1, length + 1, dtype=torch.float32)
    dt = temp ** (2 * torch.div(torch.arange(cn, dtype=torch.float32), 2, rounding_mode='trunc') / cn)
    posenc = pos[:, None] / dt[None, :]
    return torch.stack([posenc[:, 0::2].sin(), posenc[:, 1::2].cos()], dim=2).flatten(start_dim=1)


class ROIPooling(nn.Module):
    """Drop-in for layers.ROIPooling (layers.py:399-497): same forward(rois, conv_out) ->
    (roi_pool_out [B,R,C,ph,pw], roi_pe_out [B,R,C,ph,pw], level assignment numpy [B,R]), one kernel
    launch instead of a Python loop over batch x RoIs with .item() syncs.  Holds no parameters."""

    def __init__(self, config, want_levels=True):
        super().__init__()
        self.config = config
        self.want_levels = want_levels     # False: skip the device-to-host copy of the level table (a stream sync)
        self._pe = {}

    def _tables(self, device):
        cfg = self.config
        key = (int(cfg.img_height), int(cfg.img_width), int(cfg.out_fpn_chan), str(device))
        if key not in self._pe:
            self._pe[key] = (one_dimension_positional_encoding(cfg.img_height, cfg.out_fpn_chan // 2).to(device).contiguous(),
                             one_dimension_positional_encoding(cfg.img_width, cfg.out_fpn_chan // 2).to(device).contiguous())
        return self._pe[key]

    def forward(self, rois: torch.Tensor, conv_out):
        cfg = self.config
        _need_cuda(rois, *conv_out)
        B, R = rois.shape[:2]
        n_layers = int(cfg.n_layers)
        feats = [_f32c(f) for f in conv_out[:n_layers]]
        Cc = feats[0].shape[1]
        if Cc != int(cfg.out_fpn_chan):
            raise _lib.NbmError("feature maps must have out_fpn_chan channels")
        ph, pw = int(cfg.roi_pool_h), int(cfg.roi_pool_w)
        pe_f, pe_t = self._tables(rois.device)
        r = _f32c(rois)
        pool = torch.empty((B, R, Cc, ph, pw), dtype=torch.float32, device=rois.device)
        pe = torch.empty_like(pool)
        lvl = torch.empty((B, R), dtype=torch.int32, device=rois.device)
        ptrs = (C.c_void_p * n_layers)(*[f.data_ptr() for f in feats])
        hs = (C.c_int32 * n_layers)(*[f.shape[-2] for f in feats])
        ws = (C.c_int32 * n_layers)(*[f.shape[-1] for f in feats])
        _lib.check(_lib.lib().nbm_roi_pool(r.data_ptr(), B, R, ptrs, hs, ws, n_layers, Cc, ph, pw, int(cfg.img_height),
                                           int(cfg.img_width), pe_f.data_ptr(), pe_t.data_ptr(), pool.data_ptr(),
                                           pe.data_ptr(), lvl.data_ptr(), _stream()), "nbm_roi_pool")
        return pool, pe, (lvl.cpu().numpy() if self.want_levels else lvl)
