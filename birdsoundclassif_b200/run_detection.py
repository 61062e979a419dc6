"""GPU host driver with the reference's ``run_detection`` / ``merge_images`` call surface.

Mirrors ``nbm_model/run_detection.py``: ``run_detection(model, config, wav_path,
bird_dicts_path, min_score=0.5, bs=10, ...)`` (:28-84) returns the same
``{species: {'bbox_coord': [[x1,y1,x2,y2],...], 'scores': [...]}}`` dictionary and batches the
detector windows identically (bs, order, partial last batch -- the reference's ``nms`` and
``ProposalLayer`` are batch-coupled).  Differences are only WHERE the work
happens: PCM16 goes to the GPU once, the front-end produces the ``[n,1,375,1024]`` float32
tensor on the device (no ``np.stack`` + host cast + H2D per batch, run_detection.py:53), and
the per-file merge runs in one library call instead of a 150 x n_tiles Python loop.

``model`` is the reference's ``NbmModel`` (loaded by its own ``load_model``, unchanged) or any
callable with the same contract: ``model(batch[:, None], min_score=...)`` -> list (one per
image) of ``{'1'..str(num_classes): {'bbox_coord': [n,4], 'scores': [1,n]}}``.
``patch_reference()`` swaps the reference's hot-path symbols for the library-backed ones.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import torch

from . import postproc
from .frontend import File_Processor


_PATCHED: dict = {}     # (module name, symbol) -> the reference's original object


def patch_reference() -> list[str]:
    """Monkey-patch the reference modules (when importable as ``nbm_model.*``) so its unchanged
    model code calls libnbm_b200: ``nms`` and ``bbox_reg_to_coord`` are imported BY NAME into
    ``nbm_model.nets.layers`` (layers.py:4) and ``nbm_model.run_detection`` (run_detection.py:18).
    Returns the list of patched symbols; ``unpatch_reference()`` restores the originals."""
    done = []
    repl = {"nms": postproc.nms, "bbox_reg_to_coord": postproc.bbox_reg_to_coord,
            "merge_images": postproc.merge_images, "File_Processor": File_Processor}
    for modname, names in (("nbm_model.nets.layers", ("nms", "bbox_reg_to_coord")),
                           ("nbm_model.nets.util.nets_utils", ("nms", "bbox_reg_to_coord")),
                           ("nbm_model.run_detection", ("nms", "merge_images", "File_Processor"))):
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for n in names:
            if hasattr(mod, n):
                if getattr(mod, n) is not repl[n]:
                    _PATCHED.setdefault((modname, n), getattr(mod, n))
                    setattr(mod, n, repl[n])
                done.append(f"{modname}.{n}")
    return done


def unpatch_reference() -> None:
    """Put the reference's own symbols back (tests compare the two in one process)."""
    for (modname, n), orig in _PATCHED.items():
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, n, orig)
    _PATCHED.clear()


def _fast_rcnn_forward(self, conv_out, rois, nms_thresh=0.3, min_score=0.5, training=None):
    """FastRCNN.forward (layers.py:668-778) with the inference branch (:688-778: argmax, class delta gather,
    decode, clamp, sort, drop class 0, NMS, per-class NMS + ``> min_score``) done by one kernel launch
    (``nbm_final_detections``) instead of a Python loop over images x 150 classes.  The network part
    (roi_pooling -> rcnn, :674-676) and the training return (:678-681) are the reference's."""
    roi_pool_out, roi_pe_out, _ = self.roi_pooling(rois, conv_out)
    bbox_reg, bbox_classes = self.rcnn(roi_pool_out, roi_pe_out)
    if training is None:
        training = self.training
    if training:
        return bbox_reg, bbox_classes
    return postproc.fastrcnn_inference_tail(bbox_reg, bbox_classes, rois, self.config, nms_thresh, min_score)


def accelerate_model(model, tail: bool = True):
    """Swap the parameter-free ProposalLayer (head.prop_layer, head.py:18) and ROIPooling
    (head.fast_rcnn.roi_pooling, layers.py:661) of a reference NbmModel for the library-backed ones and, with
    ``tail``, bind ``_fast_rcnn_forward`` over ``head.fast_rcnn.forward`` (the fused inference tail).  None of
    them holds weights, so state_dict keys are unaffected."""
    head = getattr(model, "head", None)
    if head is not None and hasattr(head, "prop_layer"):
        old = head.prop_layer
        head.prop_layer = postproc.ProposalLayer(old.config, old.n_layers).train(old.training)
    # second stage: the parameter-free ROIPooling of FastRCNN (layers.py:661) -> one-launch kernel
    frcnn = getattr(head, "fast_rcnn", None) if head is not None else None
    if frcnn is not None and hasattr(frcnn, "roi_pooling"):
        frcnn.roi_pooling = postproc.ROIPooling(frcnn.roi_pooling.config, want_levels=not tail)
        if tail:
            frcnn.forward = types.MethodType(_fast_rcnn_forward, frcnn)
    return model


def build_model(args, device=None):
    """The model-building half of the reference's ``load_model`` (run_detection.py:100-118), restated on the
    reference's own builders: only ``nbm_model.nets`` has to be importable (``nbm_model/run_detection.py``
    itself imports matplotlib and librosa at module level, :7-12, which an inference box need not have)."""
    try:
        from nbm_model.nets.backbone import build_backbone
        from nbm_model.nets.fpn import build_fpn
        from nbm_model.nets.head import build_head
        from nbm_model.nets.nbm_model import NbmModel
        from nbm_model.nets.self_attention import build_sa_layers
    except ImportError as e:
        raise ImportError("nbm_model.nets (the reference checkout: it provides the detector network) must be "
                          f"importable, e.g. PYTHONPATH=<reference checkout> ({e})") from e
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"          # run_detection.py:22-25
    backbone = build_backbone(args)
    if args.fpn_first:
        attn_channels = [args.out_fpn_chan] * len(backbone.num_channels)
    elif args.sandwich_attn:
        attn_channels = (backbone.num_channels, [args.out_fpn_chan] * len(backbone.num_channels))
    else:
        attn_channels = backbone.num_channels
    attn = build_sa_layers(args, attn_channels)
    fpn = build_fpn(args, backbone.num_channels)
    head = build_head(args)
    return NbmModel(args, backbone, attn, fpn, head).to(device)


def load_model(mod_p, device=None):
    """Same contract as the reference's ``load_model`` (run_detection.py:87-122): reads ``<mod_p>/args`` (the JSON
    dump of the training namespace) and ``<mod_p>/model_chkpt.pt`` (``{'checkpoints': state_dict, ...}``,
    train.py:171-187) and returns ``(model.eval(), args)``.  The network classes are the reference's."""
    from nbm_model.nets.nbm_model import initialize_model
    from nbm_model.nets.util.nets_utils import setattr_others

    class Args:
        def __init__(self, **kwargs):
            for k, v in kwargs.items():
                setattr(self, k, v)

    with open(os.path.join(mod_p, "args"), "rb") as f:
        args = Args(**json.load(f))
    setattr_others(args)
    model = build_model(args, device)
    model = initialize_model(model, path=os.path.join(mod_p, "model_chkpt.pt"), train=False)
    return model, args


def detect_tiles(model, tiles: torch.Tensor, min_score: float, bs: int) -> list:
    """Run the detector over device-resident tiles [n, H, W] in the reference's batching
    (run_detection.py:47-67).  Returns the list of per-batch outputs."""
    if hasattr(model, "detect_tiles"):          # graphed.GraphedDetector overlaps its host work with the next batch
        return model.detect_tiles(tiles, min_score, bs)
    outputs = []
    n_img = len(tiles)
    with postproc.gc_paused(), torch.no_grad():
        for s in range(0, n_img, bs):
            batch = tiles[s:s + bs]
            outputs.append(model(batch[:, None], min_score=min_score))
    return outputs


def detect_stream(model, files, min_score: float, bs: int):
    """``detect_tiles`` over an iterable of per-recording tile tensors: yields each recording's outputs in order.  A
    ``GraphedDetector`` keeps its replay lanes busy across recording boundaries (the first batches of the next
    recording run while this one's last batch finishes and its merge runs); any other model is called file by file."""
    if hasattr(model, "detect_stream"):
        yield from model.detect_stream(files, min_score, bs)
        return
    for tiles in files:
        yield detect_tiles(model, tiles, min_score, bs)


def run_detection(model, config, wav_path, bird_dicts_path, min_score=0.5, bs=10, visualise_outputs=False,
                  show_sp_name=True, timings: dict | None = None):
    """Same signature and return value as the reference's run_detection (plotting is not
    offered: visualise_outputs must be False)."""
    if visualise_outputs:
        raise NotImplementedError("matplotlib visualisation is out of scope")
    with postproc.gc_paused():          # the per-tile dictionaries live until the merge below
        return _run_detection(model, config, wav_path, bird_dicts_path, min_score, bs, timings)


def _run_detection(model, config, wav_path, bird_dicts_path, min_score, bs, timings):
    t0 = time.perf_counter()
    fp = File_Processor(wav_path)
    img_db, _ = fp.process_file()
    if img_db is None:
        raise RuntimeError(f"{wav_path}: could not be loaded")      # the reference crashes at len(None), :47
    with open(bird_dicts_path, "r") as f:
        birds_dict = json.load(f)
    birds_dict.update({"Non bird sound": 0})
    reverse_dict = {idx: name for name, idx in birds_dict.items()}
    t1 = time.perf_counter()
    if isinstance(img_db, list):
        # Recording longer than 3401 s: the reference's File_Processor returns one image list per max_l-sample piece
        # (prepare_dataset.py:187-225) and its run_detection cannot consume that.  Extension: every piece is detected
        # and merged as a file of its own, its boxes are shifted to the recording's time axis (frames) and the
        # per-species lists are concatenated in piece order.  Pieces do not overlap, so no NMS runs across them.
        output, n_tiles, t_model = {}, 0, 0.0
        for k, tiles in enumerate(img_db):
            tm0 = time.perf_counter()
            outs = detect_tiles(model, tiles, min_score, bs)
            t_model += time.perf_counter() - tm0
            piece = types.SimpleNamespace(W_PIX=fp.W_PIX, HOP_SPECTRO=fp.HOP_SPECTRO,
                                          spectrogram_length=fp.piece_spectrogram_lengths[k])
            shift = float(round(k * fp.piece_samples / fp.HOP_LENGTH))
            for name, v in postproc.merge_to_output(piece, outs, config.num_classes, reverse_dict).items():
                e = output.setdefault(name, {"bbox_coord": [], "scores": []})
                e["bbox_coord"] += [[b[0] + shift, b[1], b[2] + shift, b[3]] for b in v["bbox_coord"]]
                e["scores"] += v["scores"]
            n_tiles += len(tiles)
        output = {name: output[name] for name in sorted(output, key=lambda nm: birds_dict[nm])}     # classes ascending
        if timings is not None:
            t3 = time.perf_counter()
            timings.update(frontend_s=t1 - t0, model_s=t_model, post_s=t3 - t1 - t_model, tiles=n_tiles,
                           frames=int(sum(fp.piece_spectrogram_lengths)),
                           detections=sum(len(v["scores"]) for v in output.values()))
        return output
    outputs = detect_tiles(model, img_db, min_score, bs)
    t2 = time.perf_counter()

    # == merge_images + the dictionary comprehension of run_detection.py:69-77 (tested equal)
    output = postproc.merge_to_output(fp, outputs, config.num_classes, reverse_dict)
    if timings is not None:
        t3 = time.perf_counter()
        timings.update(frontend_s=t1 - t0, model_s=t2 - t1, post_s=t3 - t2, tiles=len(img_db),
                       frames=int(fp.spectrogram_length),
                       detections=sum(len(v["scores"]) for v in output.values()))
    return output
