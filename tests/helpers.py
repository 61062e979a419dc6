"""Shared helpers for the golden-vector tests."""
import os
import zlib

import numpy as np

from birdsoundclassif_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must match oracle/make_golden.py
FRONTEND_CASES = [
    ("fe_2s", 2.0, 11, {}),
    ("fe_half_s", 0.5, 12, {}),
    ("fe_6s", 6.2, 13, {}),
    ("fe_exact", (1024 + 819) * 132 / 44100.0 - 0.001, 14, {}),
    ("fe_stress", 1.0, 15, dict(freq_accuracy=10.0, dt=0.001)),
]
ROW_STRIDE, COL_STRIDE = 5, 7


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def frontend_pcm(case, gold):
    name, secs, seed, _ = case
    pcm = synth.synth_pcm(secs, seed)
    assert np.uint32(zlib.crc32(pcm.tobytes())) == gold[name + "/pcm_crc"], "synthetic PCM not reproducible"
    return pcm


def split_keep(gold, name):
    lens = gold[f"{name}/keep_len"]
    flat = gold[f"{name}/keep_flat"]
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n].tolist())
        o += n
    return out


def dets_from_flat(counts, boxes, scores):
    """counts [B, C] + class-major flat boxes/scores -> list(B) of {class: (boxes[n,4], scores[n])}."""
    out, o = [], 0
    for b in range(counts.shape[0]):
        d = {}
        for c in range(counts.shape[1]):
            n = int(counts[b, c])
            if n:
                d[c + 1] = (boxes[o:o + n], scores[o:o + n])
                o += n
        out.append(d)
    return out


def dets_to_flat(dets, num_classes):
    counts = np.zeros((len(dets), num_classes), dtype=np.int64)
    bb, ss = [], []
    for b, d in enumerate(dets):
        for c in range(1, num_classes + 1):
            e = d[str(c)]
            bc = np.asarray(e["bbox_coord"].cpu() if hasattr(e["bbox_coord"], "cpu") else e["bbox_coord"])
            n = len(bc)
            counts[b, c - 1] = n
            if n:
                sc = np.asarray(e["scores"].cpu() if hasattr(e["scores"], "cpu") else e["scores"])
                bb.append(bc.reshape(-1, 4).astype(np.float32))
                ss.append(sc.reshape(-1).astype(np.float32))
    return counts, (np.concatenate(bb) if bb else np.zeros((0, 4), np.float32)), \
        (np.concatenate(ss) if ss else np.zeros((0,), np.float32))


def requant_piece(p):
    """What a piece of a > 3401 s recording becomes on its way through the reference's temp wav file
    (prepare_dataset.py:199: soundfile.write of float32 data as PCM_16 = lrintf(x * 0x7FFF); read back / 32768)."""
    x = p.astype(np.float32) / np.float32(32768.0)
    if x.ndim == 2:
        x = x.mean(axis=1, dtype=np.float32)
    return np.clip(np.rint(x * np.float32(32767.0)), -32768, 32767).astype(np.int16)
