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


# annotated recordings (tests/golden/labels.npz, oracle/make_golden.py labels_golden)
LABEL_CASES = [            # (name, seconds, seed, end of the last annotation in s)
    ("lab_tail", 13.0, 21, 12.8827),        # 40 empty frames after the last call: padding in steps 40, 80, 160, 320, 175
    ("lab_over", 13.0, 22, 13.2),           # annotation runs past the end: single-column steps first
    ("lab_far", 13.0, 23, 6.0),             # plenty of room: the unannotated single-step padding
]


def label_table(name, last_end):
    """A small annotation table in the reference's format (utils.py:95-173) that exercises every branch of
    merge_and_filter_labels: boxes across tile borders, slivers, a box wider than a tile, a flat box, frequencies outside
    the band, background rows (-1) with and without real calls in the same tile, rows of another recording."""
    import pandas as pd
    rows = [  # t_start, t_end, f_start, f_end, bird_id
        (0.5, 1.2, 2000.0, 4000.0, 12), (2.3, 2.6, 1500.0, 6000.0, 7), (3.00, 3.08, 3000.0, 3500.0, 7),
        (3.04, 3.40, 800.0, 900.0, 31), (2.0, 7.0, 5000.0, 5600.0, 44), (5.3968, 9.0, 700.0, 9000.0, 3),
        (4.0, 4.5, 1000.0, 1010.0, 5), (8.0, 8.4, 100.0, 15000.0, 150), (8.1, 8.3, 2500.0, 2600.0, -1),
        (10.2, 10.6, 2500.0, 2600.0, -1), (11.0, last_end, 4000.0, 4400.0, 9),
    ]
    df = pd.DataFrame(rows, columns=["t_start", "t_end", "f_start", "f_end", "bird_id"])
    df["filename"] = name
    other = df.iloc[:3].copy()
    other["filename"] = "another_recording"
    return pd.concat([other, df], ignore_index=True)


def annotations_from_gold(gold, name):
    """-> [(tile index, [(x1, y1, x2, y2), ...], [bird ids])]"""
    out, o = [], 0
    for idx, n in zip(gold[name + "/index"], gold[name + "/count"]):
        out.append((int(idx), [tuple(int(v) for v in b) for b in gold[name + "/coord"][o:o + n]],
                    [int(v) for v in gold[name + "/bird_id"][o:o + n]]))
        o += n
    return out


def annotations_from_frame(df):
    return [(int(i), [tuple(int(v) for v in b) for b in c], [int(v) for v in bb])
            for i, c, bb in zip(df["index"], df["coord"], df["bird_id"])]
