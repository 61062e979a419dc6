"""Annotated recordings: the labelled branches of the reference's ``File_Processor`` (training-set preparation).

The detector tiles of an annotated recording differ from those of an unannotated one in ONE place -- how the last,
partial tile is padded (prepare_dataset.py:280-292) -- and come with a table of bounding boxes per tile
(``merge_and_filter_labels``, prepare_dataset.py:297-376).  Both are host arithmetic on a few dozen rows; the tiles
themselves come from the same CUDA front-end.  ``labels`` is the reference's table (``create_label_dataset``,
utils.py:95-173): columns ``filename, t_start, t_end`` (s), ``f_start, f_end`` (Hz), ``bird_id``.
"""
from __future__ import annotations

import numpy as np


class NoLabelsForFile(Exception):
    """The label table has no row for this recording (the reference raises pd.errors.IntCastingNaNError at
    prepare_dataset.py:314 and process_file answers (None, None), :150-152)."""


def file_rows(labels, filename):
    return labels.loc[labels["filename"] == filename]


def labelled_pad_map(w: int, w_pix: int, empty_width: int) -> np.ndarray:
    """Source column (< w) of every column of the last tile, real width ``w`` < ``w_pix``.

    prepare_dataset.py:280-292 pads in steps: ``pad_width = max(1, min(empty_width, w_pix - width))`` columns by
    ``np.pad(..., mode='reflect')``, then ``empty_width += pad_width`` -- so that no annotated call is mirrored into
    the padding: ``empty_width`` starts as the number of frames after the last annotation's end (it may be <= 0; the
    steps are then single columns until it has grown).  An unannotated recording starts at ``empty_width = w_pix``
    and is padded in one step.  The steps are replayed here on a row of column indices with numpy's own ``reflect``,
    which is what defines the result (including its iteration when a step is wider than the image)."""
    idx = np.arange(w, dtype=np.int64)[None, :]
    empty_width = int(empty_width)
    while idx.shape[-1] < w_pix:
        pad_width = max(1, min(empty_width, w_pix - idx.shape[-1]))
        idx = np.pad(idx, ((0, 0), (0, pad_width)), mode="reflect")
        empty_width += pad_width
    return idx[0]


def empty_width_of(rows, spectrogram_length: int, dt: float) -> int:
    """Frames after the end of the last annotation (prepare_dataset.py:282-286)."""
    return int(spectrogram_length) - int(rows["t_end"].max() / dt)


def merge_and_filter_labels(labels, filename: str, ext: str, n_img: int, c: dict):
    """One row per tile that holds annotations: ``index`` (tile number), ``coord`` (list of (x1, y1, x2, y2) in tile
    pixels), ``bird_id`` (list) -- prepare_dataset.py:297-376.  ``c``: the constants ``process_file`` derives
    (``DT, FREQ_ACCURACY, LOW_FREQ, HIGH_FREQ`` -- the band edges as overwritten at :137-138 --, ``W_PIX, HOP_SPECTRO,
    H_PIX``).  Boxes are kept for every tile they overlap by at least min(20 px, half their width) and min(45 px, a
    tenth of their width), clipped to the tile; rows with ``bird_id`` -1 (background) survive nowhere: the reference's
    inner join with the per-tile count of real annotations (:366-367) drops the tiles that have none, and its filter
    drops the -1 rows of the others."""
    import pandas as pd
    rows = file_rows(labels, filename)
    if len(rows) == 0:
        raise NoLabelsForFile(filename)
    t0 = rows["t_start"].astype(float).to_numpy().copy()
    t1 = rows["t_end"].astype(float).to_numpy().copy()
    if ext == "mp3":                                   # the offset Audacity adds to mp3 imports (:308-310)
        t0 -= 0.03
        t1 -= 0.03
    f0 = np.clip(rows["f_start"].to_numpy().astype(float), c["LOW_FREQ"], c["HIGH_FREQ"])
    f1 = np.clip(rows["f_end"].to_numpy().astype(float), c["LOW_FREQ"], c["HIGH_FREQ"])
    bird = rows["bird_id"].to_numpy()
    x1 = (t0 / c["DT"]).astype(np.int64)               # .astype(int): truncation (:319)
    x2 = (t1 / c["DT"]).astype(np.int64)
    y1 = ((f0 - c["LOW_FREQ"]) / c["FREQ_ACCURACY"]).astype(np.int64)
    y2 = ((f1 - c["LOW_FREQ"]) / c["FREQ_ACCURACY"]).astype(np.int64)
    w = x2 - x1 + 1
    h = y2 - y1 + 1
    live = (y1 != y2) & (w > 0) & (h > 0)              # :325, :331-332
    x1, x2, y1, y2, w, bird = (a[live] for a in (x1, x2, y1, y2, w, bird))
    start = np.arange(n_img, dtype=np.int64) * c["HOP_SPECTRO"]         # :302
    end = start + c["W_PIX"] - 1
    # every annotation against every tile, annotation-major like the reference's cross join (:339)
    X1, X2, S, E = x1[:, None], x2[:, None], start[None, :], end[None, :]
    touches = ((X1 >= S) & (X1 <= E)) | ((X2 >= S) & (X2 <= E)) | ((X1 < S) & (X2 > E))      # :340-341
    inside = np.minimum(X2, E) - np.maximum(X1, S) + 1                  # :344
    W_ = w[:, None]
    small = ((inside < 0.5 * W_) & (inside < 20)) | ((inside < 0.1 * W_) & (inside < 45))   # :346-349
    keep = touches & ~small
    li, ti = np.nonzero(keep)                          # row-major: annotation order, then tile order
    cx1 = np.maximum(x1[li] - start[ti], 0)            # :357-360
    cx2 = np.minimum(x2[li] - start[ti], c["W_PIX"] - 1)
    cy1 = np.maximum(y1[li], 0)
    cy2 = np.minimum(y2[li], c["H_PIX"] - 1)
    b = bird[li]
    real = b != -1
    out = {"index": [], "coord": [], "bird_id": []}
    for t in np.unique(ti[real]):                      # tiles with at least one real annotation, ascending (:366-372)
        sel = np.nonzero((ti == t) & real)[0]
        out["index"].append(int(t))
        out["coord"].append([(int(cx1[j]), int(cy1[j]), int(cx2[j]), int(cy2[j])) for j in sel])
        out["bird_id"].append([int(b[j]) for j in sel])
    return pd.DataFrame(out, columns=["index", "coord", "bird_id"])


def piece_labels(labels, filename: str, k: int, time_increment: float):
    """Annotations of piece ``k`` of a recording longer than 3401 s, on the piece's own time axis
    (prepare_dataset.py:205-214): those that START inside the piece, their end clipped to it; ``None`` when there are
    none (the piece is then processed as unannotated)."""
    rows = file_rows(labels, filename).copy()
    rows["t_start"] = rows["t_start"] - k * time_increment
    rows["t_end"] = rows["t_end"] - k * time_increment
    rows = rows.loc[rows["t_start"].between(0, time_increment)].copy()
    rows["t_end"] = rows["t_end"].clip(upper=time_increment)
    rows["filename"] = f"temp{k}"
    return rows if len(rows) else None
