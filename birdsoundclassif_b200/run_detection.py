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
import sys
import time
import types

import torch

from . import postproc
from .frontend import File_Processor


def patch_reference() -> list[str]:
    """Monkey-patch the reference modules (when importable as ``nbm_model.*``) so its unchanged
    model code calls libnbm_b200: ``nms`` and ``bbox_reg_to_coord`` are imported BY NAME into
    ``nbm_model.nets.layers`` (layers.py:4) and ``nbm_model.run_detection`` (run_detection.py:18).
    Returns the list of patched symbols."""
    done = []
    for modname, names in (("nbm_model.nets.layers", ("nms", "bbox_reg_to_coord")),
                           ("nbm_model.nets.util.nets_utils", ("nms", "bbox_reg_to_coord")),
                           ("nbm_model.run_detection", ("nms", "merge_images", "File_Processor"))):
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for n in names:
            if hasattr(mod, n):
                setattr(mod, n, {"nms": postproc.nms, "bbox_reg_to_coord": postproc.bbox_reg_to_coord,
                                 "merge_images": postproc.merge_images, "File_Processor": File_Processor}[n])
                done.append(f"{modname}.{n}")
    return done


def accelerate_model(model):
    """Swap the parameter-free ProposalLayer (head.prop_layer, head.py:18) and ROIPooling
    (head.fast_rcnn.roi_pooling, layers.py:661) of a reference NbmModel for the library-backed ones;
    state_dict keys are unaffected."""
    head = getattr(model, "head", None)
    if head is not None and hasattr(head, "prop_layer"):
        old = head.prop_layer
        head.prop_layer = postproc.ProposalLayer(old.config, old.n_layers).train(old.training)
    # second stage: the parameter-free ROIPooling of FastRCNN (layers.py:661) -> one-launch kernel
    frcnn = getattr(head, "fast_rcnn", None) if head is not None else None
    if frcnn is not None and hasattr(frcnn, "roi_pooling"):
        frcnn.roi_pooling = postproc.ROIPooling(frcnn.roi_pooling.config)
    return model


def detect_tiles(model, tiles: torch.Tensor, min_score: float, bs: int) -> list:
    """Run the detector over device-resident tiles [n, H, W] in the reference's batching
    (run_detection.py:47-67).  Returns the list of per-batch outputs."""
    outputs = []
    n_img = len(tiles)
    for s in range(0, n_img, bs):
        batch = tiles[s:s + bs]
        with torch.no_grad():
            outputs.append(model(batch[:, None], min_score=min_score))
    return outputs


def run_detection(model, config, wav_path, bird_dicts_path, min_score=0.5, bs=10, visualise_outputs=False,
                  show_sp_name=True, timings: dict | None = None):
    """Same signature and return value as the reference's run_detection (plotting is not
    offered: visualise_outputs must be False)."""
    if visualise_outputs:
        raise NotImplementedError("matplotlib visualisation is out of scope")
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
