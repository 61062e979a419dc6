"""Dataset-preparation consumer: PNG writer on the CPU; on the GPU the uint8 quantisation against
np.round(img * 255) of the float64 oracle image (prepare_dataset.py:85) and the reference's output layout."""
import glob
import os

import numpy as np
import pytest
import torch

from birdsoundclassif_b200 import dataset, synth


def test_png_roundtrip_and_structure():
    rng = np.random.default_rng(3)
    for shape in [(375, 1024), (1, 1), (7, 13)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        png = dataset.encode_png_gray8(img)
        assert png[:8] == b"\x89PNG\r\n\x1a\n" and png[12:16] == b"IHDR" and png[-8:-4] == b"IEND"
        assert np.array_equal(dataset.decode_png_gray8(png), img)


def test_annotations_need_the_label_table(tmp_path):
    with pytest.raises(NotImplementedError, match="labels="):
        dataset.prepare_dataset(str(tmp_path), str(tmp_path / "out"))


def test_no_cpu_fallback():
    with pytest.raises(Exception, match="CUDA"):
        dataset.tiles_to_u8(torch.zeros(4, 4))


@pytest.mark.gpu
def test_u8_quantisation_exact_on_the_given_floats():
    g = torch.Generator().manual_seed(1)
    for n in [0, 1, 3, 4, 5, 1023, 375 * 1024 * 3 + 2]:
        t = torch.rand(n, generator=g)
        if n >= 5:
            t[:5] = torch.tensor([0.0, 1.0, 0.5 / 255, 1.5 / 255, 2.5 / 255])     # ties: half to even like np.round
        buf = torch.empty(n + 4, dtype=torch.float32, device="cuda")[:n]           # 16-byte aligned base
        buf.copy_(t)
        got = dataset.tiles_to_u8(buf).cpu().numpy()
        want = np.round(t.numpy() * np.float32(255)).astype(np.uint8)
        assert np.array_equal(got, want), n


@pytest.mark.gpu
def test_prepare_dataset_layout_and_images(tmp_path):
    from oracle import frontend_oracle as fo
    src = tmp_path / "night_site"
    src.mkdir()
    pcms = {"a#1": synth.synth_pcm(7.0, 900), "b": synth.synth_pcm(2.0, 901)}
    for name, pcm in pcms.items():
        synth.write_wav(str(src / f"{name}.wav"), pcm)
    out = tmp_path / "out"
    n = dataset.prepare_dataset(str(src), str(out), annotations=False)
    assert n == 3 + 1
    assert sorted(os.listdir(out / "negative_files")) == ["night_site__a__1", "night_site__b"]
    assert not (out / "positive_files").exists()
    for name, pcm in pcms.items():
        stem = "night_site__" + name.replace("#", "__")
        files = sorted(glob.glob(str(out / "negative_files" / stem / "*.png")))
        ref = fo.process(pcm, fo.derive_params()).tiles
        assert [os.path.basename(f) for f in files] == [f"{stem}__{i:05d}.png" for i in range(len(ref))]
        for f, r in zip(files, ref):
            got = dataset.decode_png_gray8(open(f, "rb").read()).astype(np.int32)
            want = np.round(r * 255).astype(np.uint8).astype(np.int32)
            d = np.abs(got - want)
            # the float32 tile differs from the float64 image by ~2e-7 rms (<= 1e-3 anywhere): a level flips only where
            # img * 255 lies that close to a rounding boundary, and never by more than one level
            assert got.shape == (375, 1024) and d.max() <= 1 and (d != 0).mean() < 2e-3, (f, d.max(), (d != 0).mean())
    # a second call skips directories that exist (prepare_dataset.py:51-52)
    assert dataset.prepare_dataset(str(src), str(out), annotations=False) == 0


@pytest.mark.gpu
def test_prepare_dataset_with_annotations(tmp_path):
    """annotations=True with the label table: annotated tiles + annotations.csv under positive_files, the other tiles
    under negative_files, a recording the table does not know is skipped (prepare_dataset.py:33-89, 146-153)."""
    import ast
    import csv
    from . import helpers as H
    g = H.load("labels.npz")
    src = tmp_path / "site"
    src.mkdir()
    synth.write_wav(str(src / "lab_tail.wav"), synth.synth_pcm(13.0, 21))
    synth.write_wav(str(src / "lab_far.wav"), synth.synth_pcm(13.0, 23))
    synth.write_wav(str(src / "unknown.wav"), synth.synth_pcm(2.0, 5))
    out = tmp_path / "out"
    import pandas as pd
    table = pd.concat([H.label_table("lab_tail", 12.8827), H.label_table("lab_far", 6.0)], ignore_index=True)
    n = dataset.prepare_dataset(str(src), str(out), labels=table)
    assert n == 12
    for name in ("lab_tail", "lab_far"):
        want = H.annotations_from_gold(g, name)
        pos = sorted(os.listdir(out / "positive_files" / f"site__{name}"))
        assert pos == ["annotations.csv"] + [f"site__{name}__{i:05d}.png" for i, _, _ in want]
        neg_dir = out / "negative_files" / f"site__{name}"
        neg = sorted(os.listdir(neg_dir)) if neg_dir.exists() else []
        assert neg == [f"site__{name}__{i:05d}.png" for i in range(6) if i not in {w[0] for w in want}]
        assert len(pos) - 1 + len(neg) == 6
    assert len(os.listdir(out / "negative_files" / "site__lab_far")) == 2
    assert not (out / "positive_files" / "site__unknown").exists() and not (out / "negative_files" / "site__unknown").exists()
    want = H.annotations_from_gold(g, "lab_tail")
    with open(out / "positive_files" / "site__lab_tail" / "annotations.csv") as f:
        rows = list(csv.reader(f, delimiter=";"))
    assert rows[0] == ["index", "coord", "bird_id"]
    got = [(int(r[0]), [tuple(b) for b in ast.literal_eval(r[1])], ast.literal_eval(r[2])) for r in rows[1:]]
    assert got == want
    # the last tile's image carries the annotated padding: compare with the reference's rows
    last = dataset.decode_png_gray8(open(out / "positive_files" / "site__lab_tail" / "site__lab_tail__00005.png", "rb").read())
    ref_rows = np.round(g["lab_tail/last_rows"].astype(np.float64) * 255)
    assert np.abs(last[[0, 187, 374]].astype(np.float64) - ref_rows).max() <= 1
