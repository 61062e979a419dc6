"""Dataset-preparation consumer of the front-end: ``prepare_dataset``, mirroring
``nbm_model/nbm_datasets/prepare_dataset.py:12-89``.

Same file selection, same output layout and names
(``<out>/{positive,negative}_files/<top>__<file>/<top>__<file>__00042.png`` plus ``annotations.csv`` beside the
positive images, at most 1000 negative images per file, existing output directories skipped), same image content:
``uint8(round(img * 255))`` as an 8-bit grey-scale PNG of 375 x 1024.  The tiles come from the batched GPU
front-end and are quantised on the device (``nbm_tiles_to_u8``), so 1 byte per pixel crosses PCIe instead of the
reference's 8 (float64 images).  With ``annotations=True`` the label table is passed in (``labels=``, the format of
the reference's ``create_label_dataset``, utils.py:95-173 -- that function itself reads Audacity exports against a
species dictionary at a hard-coded path, utils.py:109, and is not restated); annotated tiles go to
``positive_files`` with their box table, the others to ``negative_files`` (labels.py has the joins).

The PNG encoder is the stdlib's zlib (the reference uses imageio, absent here); PNG is lossless, so a
decoder returns the same array whichever encoder wrote it.
"""
from __future__ import annotations

import glob
import json
import os
import struct
import zlib

import numpy as np
import torch

from . import _lib, labels as labels_mod
from .frontend import File_Processor, _stream_ptr


def tiles_to_u8(tiles: torch.Tensor, stream=None) -> torch.Tensor:
    """float32 CUDA tiles in [0, 1] -> uint8 CUDA tensor of the same shape (prepare_dataset.py:85)."""
    if not tiles.is_cuda:
        raise _lib.NbmError("tiles_to_u8 needs a CUDA tensor (no CPU fallback)")
    assert tiles.dtype == torch.float32 and tiles.is_contiguous()
    out = torch.empty(tiles.shape, dtype=torch.uint8, device=tiles.device)
    with torch.cuda.device(tiles.device):
        _lib.check(_lib.lib().nbm_tiles_to_u8(tiles.data_ptr(), tiles.numel(), out.data_ptr(), _stream_ptr(stream)),
                   "nbm_tiles_to_u8")
    return out


def encode_png_gray8(img: np.ndarray, level: int = 6) -> bytes:
    """8-bit grey-scale PNG (colour type 0, filter 0 on every row)."""
    assert img.dtype == np.uint8 and img.ndim == 2
    h, w = img.shape

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    raw = np.zeros((h, w + 1), dtype=np.uint8)
    raw[:, 1:] = img
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + chunk(b"IEND", b"")


def decode_png_gray8(data: bytes) -> np.ndarray:
    """Inverse of encode_png_gray8 (tests; handles only what that encoder writes)."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", 0
    h = 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 0)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, w + 1)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].copy()


def prepare_dataset(directory, out_directory, freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024,
                    annotations=True, audio_format="", keep_files_p=None, labels=None):
    """Reference signature (prepare_dataset.py:12-13) plus ``labels``: the annotation table the reference builds with
    ``create_label_dataset`` (:33).  Returns the number of images written."""
    if annotations and labels is None:
        raise NotImplementedError("annotations=True needs the label table (labels=DataFrame with filename, t_start, t_end, "
                                  "f_start, f_end, bird_id): the reference's create_label_dataset reads it from Audacity "
                                  "exports against a dictionary at a hard-coded path (utils.py:109) and is not restated")
    if not annotations:
        labels = None                                               # :35-36
    top_dir = directory.split("/")[-1]                              # :19
    extra_str_label = ""
    if keep_files_p is not None:
        with open(keep_files_p, "r") as f:
            keep_files = json.load(f)
        dir_keep_files = [e.replace(f"{top_dir}__", "").replace(extra_str_label, "") for e in keep_files if top_dir in e]
    if audio_format != "":
        audio_files = glob.glob(directory + f"/*.{audio_format}")
    else:
        audio_files = glob.glob(directory + "/*.wav") + glob.glob(directory + "/*.mp3")
    written = 0
    for file in audio_files:
        filename = os.path.basename(file).replace(".mp3", "").replace(".wav", "").replace(".WAV", "")
        if keep_files_p is not None and filename not in dir_keep_files:
            print(f"** File {filename} not included, going to next file **")
            continue
        fp = File_Processor(file, extra_str_label, labels)
        stem = top_dir + "__" + fp.filename.replace("#", "__")
        out_pos_dir = os.path.join(out_directory, "positive_files", stem)
        out_neg_dir = os.path.join(out_directory, "negative_files", stem)
        if os.path.exists(out_pos_dir) or os.path.exists(out_neg_dir):
            continue
        print(f"~~~ Processing file {fp.filename} ~~~")
        img_db, ann = fp.process_file(freq_accuracy=freq_accuracy, dt=dt, overlap_spectro=overlap_spectro, w_pix=w_pix)
        if img_db is None:
            continue
        if isinstance(img_db, list):
            # recording longer than 3401 s: one image list per piece (prepare_dataset.py:187-225); the images are numbered
            # consecutively across the pieces as the reference's cumulative `lengths` do (:59-61, :77-80).  (Upstream,
            # the unlabelled case then dies in np.concatenate([]) at :63; the labelled one works.)  The box tables of
            # the annotated pieces get the offset of the piece they belong to -- upstream pairs the j-th TABLE with the
            # j-th piece's offset (zip at :62), which misnumbers them when an earlier piece has no annotation.
            lengths = np.cumsum([0] + [len(t) for t in img_db])
            pos_idx = []
            if ann:
                import pandas as pd
                annotated = [k for k in range(len(img_db)) if labels_mod.piece_labels(labels, fp.filename, k, fp.piece_samples / fp.FREQ) is not None]
                for a, k in zip(ann, annotated):
                    a["index"] = a["index"].to_numpy() + lengths[k]
                ann = pd.concat(ann)
                pos_idx = ann["index"].to_numpy().astype(int).tolist()
            img_db = torch.cat([t for t in img_db if len(t)], dim=0)
        else:
            pos_idx = [] if ann is None else ann["index"].to_numpy().astype(int).tolist()      # :55-57, :66-68
        n_img = len(img_db)
        pos = set(pos_idx)
        if pos:                                                     # :70-72
            os.makedirs(out_pos_dir, exist_ok=True)
            ann.to_csv(os.path.join(out_pos_dir, "annotations.csv"), sep=";", index=False)
        if len(pos) < n_img:                                        # :73-74
            os.makedirs(out_neg_dir, exist_ok=True)
        wanted = [i for i in range(n_img) if i in pos or i <= 999]  # `if i in pos_idx ... elif i <= 999` (:86-89)
        for s0 in range(0, len(wanted), 256):                       # quantised on the device, 256 images per copy
            part = wanted[s0:s0 + 256]
            u8 = tiles_to_u8(img_db[torch.as_tensor(part, device=img_db.device)].contiguous()).cpu().numpy()
            for j, i in enumerate(part):
                file_idx = "__".join([top_dir, fp.filename.replace("#", "__"), format(i, "05d")]) + ".png"
                with open(os.path.join(out_pos_dir if i in pos else out_neg_dir, file_idx), "wb") as f:
                    f.write(encode_png_gray8(u8[j]))
                written += 1
    return written
