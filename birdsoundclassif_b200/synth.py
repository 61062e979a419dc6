"""Synthetic inputs for tests and benchmarks: seeded night-recording-like PCM16 audio,
wav writing, and the stand-in detector config.

There is no network and the reference's checkpoint/args are Git-LFS stubs
(/root/reference/model_weights/{args,model_chkpt.pt}), so the detector config is the
JSON dump of the reference training parser's defaults (train.py:21-168, serialised at
train.py:286-288), which is what ``load_model`` (run_detection.py:87-99) reads.
"""
from __future__ import annotations

import json
import os
import wave

import numpy as np

SAMPLE_RATE = 44100   # prepare_dataset.py:98

# Defaults of the reference training argparse (train.py:21-168) == the schema of `args`.
DEFAULT_ARGS = {
    "lr": 1e-4, "lr_backbone": 1e-5, "batch_size": 2, "weight_decay": 1e-4, "lr_drop": 383,
    "clip_max_norm": 0.1, "model_name": "new_model", "data_path": "dataset", "save_dir": "models",
    "max_steps": 5e5, "first_neg_step": 0, "neg_step_freq": 10, "save_step": None,
    "img_width": 1024, "img_height": 375, "inpt_channels": 1, "backbone": "resnet50",
    "dilation": False, "position_embedding": "sine", "add_posenc": False, "one_dim_posenc": True,
    "norm_layer_backbone": "frozen_batchnorm",
    "fs_cls_loss_coef": 1, "fs_neg_cls_loss_coef": 1, "fs_reg_loss_coef": 1,
    "sec_cls_loss_coef": 1, "sec_neg_cls_loss_coef": 1, "sec_reg_loss_coef": 1,
    "focal_loss": False, "device": "cuda", "seed": 42, "num_workers": 4,
    "n_ratios": 3, "anchor_stride": 16, "base_size": 16,
    "rpn_neg_label": 0.3, "rpn_pos_label": 0.7, "rpn_batchsize": 16, "rpn_fg_fraction": 0.5,
    "rcnn_batch_size": 16, "rcnn_fg_prop": 0.4, "fg_threshold": 0.5,
    "bg_threshold_lo": 0.1, "bg_threshold_hi": 0.5, "depth_rcnn": 3,
    "pre_nms_topN": 3000, "min_threshold": 5, "nms_thresh": 0.7, "post_nms_topN": 1000,
    "post_nms_topN_eval": 50, "pre_nms_topN_eval": 500, "roi_pool_h": 2, "roi_pool_w": 2,
    "hidden_size_rcnn": 512, "dropout": 0, "proposal_number": 50,
    "fpn": "fpn", "n_bifpn_layers": 5, "fpn_p_chan": 384, "out_fpn_chan": 256,
    "fpn_first": False, "sandwich_attn": False, "tf_rcnn": False, "tf_pe_qk": False,
    "tf_model_dim": 512, "tf_nhead": 8, "tf_num_encoder_layers": 6, "tf_dim_feedforward": 1024,
    "pyramid_top_n_attn": 2, "num_classes": 150, "validation_prop": 0.03,
}


class Args:
    """Bare attribute bag, like the dummy class in run_detection.py:90-94."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def default_args(device: str = "cuda") -> Args:
    """DEFAULT_ARGS re-hydrated and extended the way load_model does
    (run_detection.py:95-99 + nets_utils.setattr_others, nets_utils.py:405-416)."""
    a = Args(**DEFAULT_ARGS)
    a.device = device
    if a.n_ratios == 3:
        a.ratios = [0.5, 1, 2]
    elif a.n_ratios == 5:
        a.ratios = [0.2, 0.5, 1, 2, 5]
    if "vgg" in a.backbone:
        a.n_layers, a.top_size = 4, (23, 64)
    else:
        a.n_layers, a.top_size = 5, (24, 64)
    a.scales = 2 ** np.arange(a.n_layers)
    return a


def write_args(dirpath: str, **overrides) -> str:
    """``<dirpath>/args``: the JSON the training script writes (train.py:286-288) with the parser's defaults;
    ``overrides`` e.g. ``device='cpu'`` (layers.py:271,290,419 read ``config.device``)."""
    os.makedirs(dirpath, exist_ok=True)
    p = os.path.join(dirpath, "args")
    with open(p, "w") as f:
        json.dump({**DEFAULT_ARGS, **overrides}, f)
    return p


def write_standin_checkpoint(dirpath: str, seed: int = 0, sharpen: float = 0.0, **arg_overrides) -> str:
    """Stand-in for the reference's ``model_weights/`` directory (its ``args`` and ``model_chkpt.pt`` are Git-LFS
    stubs): ``args`` from the training parser's defaults and ``model_chkpt.pt`` in the format ``train.save``
    writes (train.py:171-187: ``{'checkpoints': state_dict, 'steps', 'epoch', 'best_val_cls_loss'}``) and
    ``initialize_model`` reads (nbm_model.py:325-334), holding a seeded random initialisation of the reference's
    own network (``nbm_model.nets`` must be importable).

    A random 151-way softmax peaks near 1/151, so nothing clears a useful ``min_score``.  ``sharpen``
    > 0 scales the weights of the second-stage classifier (``head.fast_rcnn.rcnn.bbox_classif_layer``) by that
    factor and lowers the background logit, so that a few classes collect the probability mass and the tail, the
    per-file merge and the ``.txt`` output see real detections; the network is otherwise untouched."""
    import torch
    from .run_detection import build_model

    write_args(dirpath, **arg_overrides)
    with open(os.path.join(dirpath, "args")) as f:
        a = Args(**json.load(f))
    from nbm_model.nets.util.nets_utils import setattr_others
    setattr_others(a)
    rng_state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = build_model(a, device="cpu")
        sd = model.state_dict()
        if sharpen:
            w, b = "head.fast_rcnn.rcnn.bbox_classif_layer.weight", "head.fast_rcnn.rcnn.bbox_classif_layer.bias"
            sd[w] = sd[w] * float(sharpen)
            sd[b] = sd[b].clone()
            sd[b][0] -= 2.0
    finally:
        torch.set_rng_state(rng_state)
    p = os.path.join(dirpath, "model_chkpt.pt")
    torch.save({"checkpoints": sd, "steps": 0, "epoch": 0, "best_val_cls_loss": float("inf")}, p)
    return p


def synth_pcm(seconds: float, seed: int, noise_sigma: float = 0.05, calls_per_s: float = 2.0,
              sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """Seeded mono PCM16: white Gaussian noise (sigma in full-scale units) plus Poisson
    'calls' (Hann-tapered linear chirps 1-10 kHz, 20-200 ms, amplitude U(0.05, 0.4))."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sample_rate))
    x = rng.standard_normal(n).astype(np.float32) * np.float32(noise_sigma)
    for _ in range(rng.poisson(calls_per_s * seconds)):
        dur = rng.uniform(0.02, 0.2)
        m = max(8, int(dur * sample_rate))
        start = int(rng.integers(0, max(1, n - m)))
        m = min(m, n - start)
        f0, f1 = rng.uniform(1000, 10000, size=2)
        amp = rng.uniform(0.05, 0.4)
        t = np.arange(m) / sample_rate
        phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / max(dur, 1e-9) * t * t)
        x[start:start + m] += (amp * np.hanning(m) * np.sin(phase)).astype(np.float32)
    return np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)


def write_wav(path: str, pcm: np.ndarray, sample_rate: int = SAMPLE_RATE) -> str:
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if pcm.ndim == 1 else pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())
    return path


def read_wav_pcm16(path: str) -> tuple[np.ndarray, int]:
    """PCM16 wav -> (int16 [n] mono or [n, ch], sample_rate).  The product path uploads
    int16 and converts on the device (x / 32768, exact in fp32; cf. prepare_dataset.py:162)."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{path}: only uncompressed PCM16 wav is supported")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        raw = w.readframes(n)
    raw = raw[:len(raw) - len(raw) % (2 * ch)]          # a truncated file yields the whole frames it holds (as libsndfile does)
    pcm = np.frombuffer(raw, dtype="<i2")
    return (pcm if ch == 1 else pcm.reshape(-1, ch)), sr


def fpn_features(seed: int, batch: int, channels: int, n_layers: int = 5, img_height: int = 375, img_width: int = 1024):
    """Seeded stand-in FPN outputs for the second-stage pooling tests: level l is [batch, channels,
    ceil(img_height / 2^(l+1)), img_width / 2^(l+1)] float32 (strides 2 .. 2^n_layers, layers.py:415)."""
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((batch, channels, -(-img_height // 2 ** (l + 1)), img_width // 2 ** (l + 1))).astype(np.float32)
            for l in range(n_layers)]
