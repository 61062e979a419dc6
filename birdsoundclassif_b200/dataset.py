"""Dataset-preparation consumer of the front-end: ``prepare_dataset`` for UNLABELLED
directories, mirroring ``nbm_model/nbm_datasets/prepare_dataset.py:12-89``.

Same file selection, same output layout and names
(``<out>/negative_files/<top>__<file>/<top>__<file>__00042.png``, at most 1000 images per file, existing
output directories skipped), same image content: ``uint8(round(img * 255))`` as an 8-bit grey-scale PNG
of 375 x 1024.  The tiles come from the batched GPU front-end and are quantised on the device
(``nbm_tiles_to_u8``), so 1 byte per pixel crosses PCIe instead of the reference's 8 (float64 images).
Label joins (``annotations=True``: ``create_label_dataset`` and the positive / negative split) are
training-set preparation and out of scope: they raise.

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

from . import _lib
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
                    annotations=True, audio_format="", keep_files_p=None):
    """Reference signature (prepare_dataset.py:12-13).  Returns the number of images written."""
    if annotations:
        raise NotImplementedError("label joins (create_label_dataset) are training-set preparation, out of scope; "
                                  "call with annotations=False")
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
        fp = File_Processor(file, extra_str_label, None)
        stem = top_dir + "__" + fp.filename.replace("#", "__")
        out_pos_dir = os.path.join(out_directory, "positive_files", stem)
        out_neg_dir = os.path.join(out_directory, "negative_files", stem)
        if os.path.exists(out_pos_dir) or os.path.exists(out_neg_dir):
            continue
        print(f"~~~ Processing file {fp.filename} ~~~")
        img_db, _ = fp.process_file(freq_accuracy=freq_accuracy, dt=dt, overlap_spectro=overlap_spectro, w_pix=w_pix)
        if img_db is None:
            continue
        if isinstance(img_db, list):
            # recording longer than 3401 s: one image list per piece (prepare_dataset.py:187-225); the images are numbered
            # consecutively across the pieces as the reference's cumulative `lengths` do (:59-61, :77-80).  (Upstream,
            # the unlabelled case then dies in np.concatenate([]) at :63; the labelled one works.)
            img_db = torch.cat([t for t in img_db if len(t)], dim=0)
        n_img = len(img_db)
        os.makedirs(out_neg_dir, exist_ok=True)                     # no labels: every image is a negative (:66-74)
        keep = min(n_img, 1000)                                     # `elif i <= 999` (:87)
        u8 = tiles_to_u8(img_db[:keep].contiguous()).cpu().numpy()
        for i in range(keep):
            file_idx = "__".join([top_dir, fp.filename.replace("#", "__"), format(i, "05d")]) + ".png"
            with open(os.path.join(out_neg_dir, file_idx), "wb") as f:
                f.write(encode_png_gray8(u8[i]))
            written += 1
    return written
