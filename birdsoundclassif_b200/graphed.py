"""CUDA-graph replay of the detector forward ("CUDA-graph the fixed-shape bs=4 forward").

The reference network (``nbm_model.nets`` NbmModel, unchanged, weights from ``model_chkpt.pt``) is a few hundred
small eager launches per batch of four tiles -- ResNet-50 whose FrozenBatchNorm recomputes scale and bias from
five elementwise kernels per layer (backbone.py:55-62), two self-attention blocks, an FPN, an RPN with per-level
heads -- and on a B200 the Python interpreter, not the GPU, sets its pace (~20 ms per batch, GPU mostly idle).  The
launch sequence is fixed for a fixed input shape once the post-processing no longer reads sizes back to the host, so
it is recorded once per batch size and replayed:

  graph 1   samples [B,1,H,W] -> backbone -> attention -> FPN -> RPN (nbm_model.py:39-53, head.py:32-38) ->
            ``nbm_proposals_async`` (ProposalLayer, layers.py:226-303): full ``rois [B,50,4]`` + RoI count M on the device
  host      reads M (4 bytes; the reference's batch-coupled truncation makes the RoI count data-dependent)
  graph 2   one per distinct M: ROIPooling kernel -> RCNN (layers.py:560-586) -> ``nbm_final_detections`` (FastRCNN
            tail, layers.py:688-778) -> class sort of the records
  host      one device-to-host copy of the sorted class keys -> the reference's per-image dictionaries

Same kernels, same order, same inputs as the eager accelerated model, hence bit-identical outputs (tested).  The
position embedding the reference's ``Joiner`` computes for every level (backbone.py:139-147, a host-side
``torch.ones(...).to(device)`` per call) is skipped when ``args.add_posenc`` is False, because ``NbmModel`` then
discards it (nbm_model.py:44-46); with ``add_posenc`` it is computed once per shape outside the graph (it depends
on the shape only).

Two replay LANES (``lanes=2``): each lane has its own stream, graphs, graph memory pool, static buffers, cuBLAS and
ProposalLayer workspaces -- the weights are shared -- and ``detect_stream`` runs consecutive batches on alternating lanes
as a software pipeline, across recording boundaries, so that a batch's small late-stage kernels and the host's read of
M are covered by the other batch's work (13.6 -> 11.4 ms per batch; a third lane is slower).  Each batch is still the
same kernels on the same inputs: bit-identical to one lane.  Not thread-safe: one caller at a time.

Anything the capture cannot digest (a configuration whose forward still syncs, e.g. ``pyramid_top_n_attn`` = all
levels with its per-call ``.to(device)``, self_attention.py:27-31) raises at capture time; ``GraphedDetector`` then
falls back to calling the eager accelerated model and says so once (still the GPU path, no CPU fallback).
"""
from __future__ import annotations

import warnings

import torch

from . import postproc


class _Stage1:
    __slots__ = ("graph", "x", "rois", "M_dev", "M_host", "fpn_out", "stage2")


class _Lane:
    """One replay lane: its own stream, capture stream (cuBLAS keeps one workspace per stream, and the graphs bake its
    address in), graph memory pool and ProposalLayer workspace, so that two lanes can be in flight at the same time."""
    __slots__ = ("stream", "cap_stream", "pool", "s1", "ws")

    def __init__(self):
        self.stream = self.cap_stream = self.pool = None
        self.s1: dict = {}              # (shape, nms_thresh, min_score) -> _Stage1
        self.ws: list = [None]          # ProposalLayer workspace holder (filled at the first launch)


class _Stage2:
    __slots__ = ("graph", "rec", "skey", "sb", "ss", "skey_host")


class GraphedDetector:
    """``GraphedDetector(model)(batch[:, None], min_score=...)`` == ``model(batch[:, None], min_score=...)`` for a
    reference NbmModel that went through ``accelerate_model`` (inference only).  ``lanes``: batches kept in flight by
    ``detect_tiles`` / ``detect_stream`` (a single ``__call__`` uses lane 0)."""

    def __init__(self, model, warmup: int = 2, lanes: int = 2):
        head = getattr(model, "head", None)
        if head is None or not isinstance(head.prop_layer, postproc.ProposalLayer) or \
                not isinstance(head.fast_rcnn.roi_pooling, postproc.ROIPooling):
            raise ValueError("GraphedDetector needs a model prepared by run_detection.accelerate_model()")
        self.model = model
        self.args = model.args
        self.warmup = warmup
        self._lanes = [_Lane() for _ in range(max(1, int(lanes)))]
        self._pos: dict = {}
        self._eager_only = False
        self.training = False

    def eval(self):
        return self

    # ------------------------------------------------------------------ the network, as nbm_model.py runs it
    def _first_stage(self, samples, M_dev, ws):
        m, a = self.model, self.args
        xs = m.backbone[0](samples)                                     # backbone.py:139-142 (Joiner -> Backbone)
        features = [x for _, x in xs.items()]
        if a.add_posenc:                                                # nbm_model.py:45-46
            features = [f + self._pos_for(f) for f in features]
        if a.fpn_first:                                                 # nbm_model.py:47-52
            fpn_out = m.attn(m.fpn(features))
        elif a.sandwich_attn:
            fpn_out = m.attn[1](m.fpn(m.attn[0](features)))
        else:
            fpn_out = m.fpn(m.attn(features))
        cls_scores, bbox_reg = m.head.rpn(fpn_out)                      # head.py:34
        rois, _ = m.head.prop_layer.forward_async(cls_scores, bbox_reg, M_dev, ws)
        return rois, fpn_out

    def _pos_for(self, f):
        key = tuple(f.shape)
        if key not in self._pos:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("position embedding must be cached before capture")
            self._pos[key] = self.model.backbone[1](f).to(f.dtype)
        return self._pos[key]

    def _second_stage(self, fpn_out, rois, nms_thresh, min_score):
        frc, a = self.model.head.fast_rcnn, self.args
        roi_pool_out, roi_pe_out, _ = frc.roi_pooling(rois, fpn_out)    # layers.py:674
        bbox_reg, bbox_classes = frc.rcnn(roi_pool_out, roi_pe_out)     # layers.py:676
        rec = postproc.final_detections_flat(bbox_reg, bbox_classes, rois, a.num_classes, a.img_width, a.img_height,
                                             nms_thresh, min_score)
        return rec, postproc.records_sort_device(*rec, a.num_classes)

    # ------------------------------------------------------------------ capture
    def _capture1(self, samples, lane):
        dev = samples.device
        if lane.stream is None:
            lane.stream, lane.cap_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        s1 = _Stage1()
        s1.x = torch.empty_like(samples)
        s1.x.copy_(samples)
        s1.M_dev = torch.zeros((1,), dtype=torch.int32, device=dev)
        s1.M_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        s1.stage2 = {}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):            # cuDNN plans, workspaces, anchors, position tables: all outside the graph
                self._first_stage(s1.x, s1.M_dev, lane.ws)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        s1.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(s1.graph, pool=lane.pool, stream=lane.cap_stream):
            s1.rois, s1.fpn_out = self._first_stage(s1.x, s1.M_dev, lane.ws)
            s1.M_host.copy_(s1.M_dev, non_blocking=True)
        if lane.pool is None:
            lane.pool = s1.graph.pool()
        return s1

    def _capture2(self, s1, M, nms_thresh, min_score, lane):
        dev = s1.x.device
        s2 = _Stage2()
        rois = s1.rois[:, :M]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):
                self._second_stage(s1.fpn_out, rois, nms_thresh, min_score)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        s2.skey_host = torch.empty((rois.shape[0], M), dtype=torch.int32).pin_memory()    # not inside the capture
        s2.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(s2.graph, pool=lane.pool, stream=lane.cap_stream):
            s2.rec, (s2.skey, s2.sb, s2.ss) = self._second_stage(s1.fpn_out, rois, nms_thresh, min_score)
            s2.skey_host.copy_(s2.skey, non_blocking=True)
        return s2

    # ------------------------------------------------------------------ call
    # A batch goes through three steps; with two lanes, batch i's first graph runs on the GPU beside batch i-1's (the
    # late layers of the network are too small at bs 4 to fill 148 SMs, and the host's read of M no longer leaves
    # the GPU idle):
    #   _stage1   copy the tiles into the lane's static input, replay graph 1 on the lane's stream
    #   _stage2   wait for that lane, read M, replay graph 2 (captured at its first use for this M), copy the records
    #             out of the graph's static buffers on the caller's stream
    #   _finish   wait for the copies, build the reference's dictionaries on the host
    def _stage1(self, samples, nms_thresh, min_score, lane):
        if self._eager_only:
            return self.model(samples, nms_thresh=nms_thresh, min_score=min_score)
        if not samples.is_cuda:
            raise postproc._lib.NbmError("GraphedDetector needs CUDA tensors (no CPU fallback)")
        key = (tuple(samples.shape), float(nms_thresh), float(min_score))
        s1 = lane.s1.get(key)
        dev = samples.device
        if s1 is None:
            try:
                s1 = self._capture1(samples.contiguous(), lane)
            except Exception as e:          # a configuration whose forward still talks to the host: stay eager, say so
                torch.cuda.synchronize(dev)
                warnings.warn(f"GraphedDetector: capture failed ({type(e).__name__}: {e}); running the eager accelerated model")
                self._eager_only = True
                return self._stage1(samples, nms_thresh, min_score, lane)
            lane.s1[key] = s1
        # behind the caller's stream: the tiles come from there, and so do the copies that read this lane's static
        # outputs for its previous batch
        lane.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(lane.stream):
            s1.x.copy_(samples)
            s1.graph.replay()
        samples.record_stream(lane.stream)
        return (s1, lane, nms_thresh, min_score)

    def _stage2(self, started):
        if isinstance(started, list):        # eager: already the dictionaries
            return started
        s1, lane, nms_thresh, min_score = started
        lane.stream.synchronize()
        M = int(s1.M_host[0])
        if M < 0:                            # layers.py:288-290; the reference then fails inside ROIPooling on the empty RoIs
            print("Not enough possible RoIs, RPN failed")
            raise RuntimeError("RPN produced fewer than rcnn_batch_size candidate boxes (the reference crashes here too)")
        s2 = s1.stage2.get(M)
        if s2 is None:
            s2 = s1.stage2[M] = self._capture2(s1, M, nms_thresh, min_score, lane)
        with torch.cuda.stream(lane.stream):
            s2.graph.replay()
        main = torch.cuda.current_stream(s1.x.device)
        main.wait_stream(lane.stream)
        # the records live in the graph's static buffers and are overwritten by the lane's next replay: the dictionaries
        # (which callers keep until the per-file merge) get their own copies, made on the caller's stream
        boxes, scores, classes, _ = (t.clone() for t in s2.rec)
        sb, ss = s2.sb.clone(), s2.ss.clone()
        done = torch.cuda.Event()
        done.record(main)
        return (s2, boxes, scores, classes, sb, ss, done)

    def _finish(self, pending):
        if isinstance(pending, list):        # eager: already the dictionaries
            return pending
        s2, boxes, scores, classes, sb, ss, done = pending
        done.synchronize()                   # graph 2 and its copy of the sorted class keys to the host have finished
        a = self.args
        return postproc.records_build_dicts(s2.skey_host.numpy().copy(), sb, ss, boxes, scores, classes, a.num_classes,
                                            a.proposal_number)

    @torch.no_grad()
    def __call__(self, samples, nms_thresh=0.3, min_score=0.5):
        return self._finish(self._stage2(self._stage1(samples, nms_thresh, min_score, self._lanes[0])))

    @torch.no_grad()
    def detect_tiles(self, tiles, min_score, bs, nms_thresh=0.3):
        """`run_detection.detect_tiles` for this detector: the reference's batching (run_detection.py:47-67) as a
        software pipeline (see `detect_stream`)."""
        with postproc.gc_paused():
            for outputs in self.detect_stream([tiles], min_score, bs, nms_thresh):
                return outputs

    @torch.no_grad()             # on a generator function: grad mode is switched per resume, not left off across a yield
    def detect_stream(self, files, min_score, bs, nms_thresh=0.3):
        """``files``: an iterable of device tile tensors [n, H, W], one per recording.  Yields, in order, each recording's
        list of per-batch outputs -- exactly ``detect_tiles`` of that recording (batches never span recordings:
        run_detection.py:47-67 batches one file) -- while the lanes already run the first batches of the NEXT
        recordings: graph 1 of batch i is enqueued before the host turns to batch i-(L-1)'s RoI count and second graph,
        and the dictionaries of the batches before that are built while both run."""
        lanes = self._lanes
        L = len(lanes)
        started, pending, open_files = [], [], []        # FIFOs of (recording ordinal, state); open_files: [n batches, all enqueued, outputs so far]
        i = 0

        def attributed(fn, item):
            """Run one pipeline step of a batch; an exception leaves with the ordinal of the recording it belongs to
            (``nbm_file_index``): the stream runs ahead, so the caller cannot tell from where it stands."""
            try:
                return (item[0], fn(item[1]))
            except Exception as e:
                e.nbm_file_index = item[0]
                raise

        def finish_oldest():
            f = next(e for e in open_files if e[0] > len(e[2]))
            f[2].append(attributed(self._finish, pending.pop(0))[1])

        def second(item):
            pending.append(attributed(self._stage2, item))

        def ready():
            while open_files and open_files[0][0] == len(open_files[0][2]) and open_files[0][1]:
                yield open_files.pop(0)[2]

        for ordinal, tiles in enumerate(files):
            n_batches = (len(tiles) + bs - 1) // bs
            entry = [n_batches, False, []]
            open_files.append(entry)
            for s in range(0, len(tiles), bs):
                started.append(attributed(lambda t: self._stage1(t, nms_thresh, min_score, lanes[i % L]),
                                          (ordinal, tiles[s:s + bs][:, None])))
                i += 1
                if L == 1:
                    # one lane: the second graph's host copy of the class keys is free only after the previous batch's
                    # dictionaries were built from it
                    while pending:
                        finish_oldest()
                    second(started.pop(0))
                else:
                    if len(started) == L:
                        second(started.pop(0))
                    while len(pending) > 1:
                        finish_oldest()
                if s + bs >= len(tiles):
                    entry[1] = True
                yield from ready()
            if n_batches == 0:
                entry[1] = True
                yield from ready()
        while started:
            second(started.pop(0))
            while len(pending) > 1:
                finish_oldest()
        while pending:
            finish_oldest()
        yield from ready()
