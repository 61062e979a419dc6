"""Annotated recordings (the labelled branches of the reference's File_Processor, prepare_dataset.py:146-153, 280-292,
297-376) against vectors recorded from the reference itself (tests/golden/labels.npz, oracle/make_golden.py
labels_golden): the stepwise padding of the last tile and the per-tile box table."""
import os

import numpy as np
import pytest
import torch

from birdsoundclassif_b200 import frontend, labels, synth
from oracle import frontend_oracle as fo

from . import helpers as H

pd = pytest.importorskip("pandas")


def _consts():
    c = frontend.derive_constants()
    c["H_PIX"] = 375
    return c


@pytest.mark.parametrize("name", ["rand_wav", "rand_mp3"])
def test_box_table_random(name):
    g = H.load("labels.npz")
    table = pd.DataFrame({col: g[f"{name}/table_{col}"] for col in ("t_start", "t_end", "f_start", "f_end", "bird_id")})
    table["filename"] = name
    ext = name.split("_")[1]
    got = labels.merge_and_filter_labels(table, name, ext, int(g[name + "/n_img"]), _consts())
    want = H.annotations_from_gold(g, name)
    assert len(want) > 3 and H.annotations_from_frame(got) == want
    assert list(got.columns) == ["index", "coord", "bird_id"]


@pytest.mark.parametrize("case", H.LABEL_CASES, ids=[c[0] for c in H.LABEL_CASES])
def test_box_table_and_padding_vs_reference(case):
    """CPU: the oracle's tiles (bit-identical to the reference's) re-padded with labelled_pad_map reproduce the
    reference's annotated last tile exactly; the box table equals the reference's."""
    name, secs, seed, last_end = case
    g = H.load("labels.npz")
    pcm = synth.synth_pcm(secs, seed)
    table = H.label_table(name, last_end)
    res = fo.process(pcm)
    c = _consts()
    T = int(g[name + "/spectrogram_length"])
    assert res.spectrogram_length == T and len(res.tiles) == int(g[name + "/n_tiles"])
    w_last = T - (len(res.tiles) - 1) * c["HOP_SPECTRO"]
    ew = labels.empty_width_of(labels.file_rows(table, name), T, c["DT"])
    src = labels.labelled_pad_map(w_last, c["W_PIX"], ew)
    assert src.shape == (1024,) and (src[:w_last] == np.arange(w_last)).all() and src.max() < w_last
    last = np.asarray(res.tiles[-1])[:, :w_last][:, src]
    np.testing.assert_array_equal(last[[0, 187, 374]].astype(np.float32), g[name + "/last_rows"])
    assert abs(last.sum() - float(g[name + "/last_sum"])) < 1e-6
    if name == "lab_far":       # enough empty frames: the unannotated padding
        np.testing.assert_array_equal(last, np.asarray(res.tiles[-1]))
    else:
        assert not np.array_equal(last, np.asarray(res.tiles[-1]))
    got = labels.merge_and_filter_labels(table, name, "wav", len(res.tiles), c)
    assert H.annotations_from_frame(got) == H.annotations_from_gold(g, name)


def test_pad_map_properties():
    # an unannotated recording: one reflect step, the oracle's closed form
    for w in (1, 2, 5, 390, 1023):
        src = labels.labelled_pad_map(w, 1024, 1024)
        assert [int(s) for s in src] == [fo.reflect_index(j, w) if j >= w else j for j in range(1024)]
    # annotation right up to the end: columns are mirrored one at a time, then in doubling steps
    src = labels.labelled_pad_map(100, 1024, 0)
    assert src[100] == 98 and src[101] == 99 and len(src) == 1024


def test_no_rows_for_file():
    table = H.label_table("lab_tail", 12.0)
    with pytest.raises(labels.NoLabelsForFile):
        labels.merge_and_filter_labels(table, "unknown", "wav", 3, _consts())


def test_piece_labels():
    t = pd.DataFrame({"filename": "night", "t_start": [10.0, 3399.0, 3500.0, 7000.0], "t_end": [12.0, 3405.0, 3501.0, 7001.0],
                      "f_start": 1000.0, "f_end": 2000.0, "bird_id": [1, 2, 3, 4]})
    L = 3401.0
    p0 = labels.piece_labels(t, "night", 0, L)
    assert p0["t_start"].tolist() == [10.0, 3399.0] and p0["t_end"].tolist() == [12.0, 3401.0] and set(p0["filename"]) == {"temp0"}
    p1 = labels.piece_labels(t, "night", 1, L)
    assert p1["bird_id"].tolist() == [3] and abs(p1["t_start"].iloc[0] - 99.0) < 1e-9
    assert labels.piece_labels(t, "night", 3, L) is None


@pytest.mark.gpu
@pytest.mark.parametrize("case", H.LABEL_CASES, ids=[c[0] for c in H.LABEL_CASES])
def test_gpu_file_processor_with_labels(case, tmp_path):
    name, secs, seed, last_end = case
    g = H.load("labels.npz")
    path = synth.write_wav(str(tmp_path / (name + ".wav")), synth.synth_pcm(secs, seed))
    table = H.label_table(name, last_end)
    fp = frontend.File_Processor(path, "", table)
    tiles, ann = fp.process_file()
    assert len(tiles) == int(g[name + "/n_tiles"]) and fp.spectrogram_length == int(g[name + "/spectrogram_length"])
    got = tiles[-1][[0, 187, 374]].cpu().numpy()
    assert np.abs(got.astype(np.float64) - g[name + "/last_rows"]).max() <= 1e-4
    assert H.annotations_from_frame(ann) == H.annotations_from_gold(g, name)
    # the unannotated tiles of the same recording differ only in the last tile's padding
    plain, none = frontend.File_Processor(path).process_file()
    assert none is None and torch.equal(plain[:-1], tiles[:-1])
    w_last = fp.spectrogram_length - 5 * fp.HOP_SPECTRO
    assert torch.equal(plain[-1][:, :w_last], tiles[-1][:, :w_last])
    assert torch.equal(plain[-1], tiles[-1]) == (name == "lab_far")
    # a table without rows for this recording: the reference prints and answers (None, None)
    other = table.loc[table["filename"] != name]
    assert frontend.File_Processor(path, "", other).process_file() == (None, None)


@pytest.mark.gpu
def test_gpu_long_recording_with_labels(monkeypatch):
    """prepare_dataset.py:201-222 at a small max_l: every piece is annotated against the calls that START inside it, on
    its own time axis; a piece without any is processed as unannotated and contributes no table.  Each annotated piece
    equals that piece processed as a labelled recording of its own."""
    L = 4 * 44100
    monkeypatch.setattr(frontend, "LONG_FILE_SAMPLES", L)
    pcm = synth.synth_pcm(11.0, 71)                                  # pieces of 4 s, 4 s and 3 s: 2 tiles each
    table = pd.DataFrame({"filename": "night", "t_start": [0.5, 3.9, 8.6], "t_end": [1.0, 4.6, 9.3],
                          "f_start": [2000.0, 1500.0, 3000.0], "f_end": [3000.0, 2500.0, 5000.0], "bird_id": [4, 5, 6]})
    fp = frontend.File_Processor("night.wav", "", table)
    fp.requantise_long = False
    img_db, ann = fp.process_pcm(torch.from_numpy(pcm).cuda())
    assert len(img_db) == 3 and len(ann) == 2                        # piece 1 (4..8 s) has no call starting in it
    for k, a in zip((0, 2), ann):
        rows = labels.piece_labels(table, "night", k, 4.0)
        alone = frontend.File_Processor(f"temp{k}.wav", "", rows)
        tiles, a1 = alone.process_pcm(torch.from_numpy(pcm[k * L:(k + 1) * L]).cuda())
        assert torch.equal(tiles, img_db[k]) and H.annotations_from_frame(a1) == H.annotations_from_frame(a)
    plain, _ = frontend.File_Processor("temp1.wav").process_pcm(torch.from_numpy(pcm[L:2 * L]).cuda())
    assert torch.equal(plain, img_db[1])
    # the call that starts at 3.9 s is clipped to the end of piece 0 (t_end 4.0 s -> the piece's last frames)
    a0 = H.annotations_from_frame(ann[0])
    assert any(b[2] == 1023 or b[2] >= 300 for _, boxes, _ in a0 for b in boxes)


@pytest.mark.parametrize("seed,ext,n_img,n", [(101, "wav", 12, 60), (102, "mp3", 30, 200), (103, "wav", 1, 8), (104, "wav", 300, 400)])
def test_box_table_against_the_live_reference(seed, ext, n_img, n):
    """Container only (needs the reference checkout): random annotation tables through the reference's own
    merge_and_filter_labels (a stub File_Processor: no audio involved) and through labels.merge_and_filter_labels."""
    from oracle import make_golden as mg, ref_shims
    if not ref_shims.have_reference():
        pytest.skip("no reference checkout")
    pd_mod = ref_shims.ref("nbm_model.nbm_datasets.prepare_dataset")
    c = _consts()
    rng = np.random.default_rng(seed)
    name = f"rec_{seed}"
    table = mg.random_label_table(rng, name, n, (n_img * 819 + 205) * c["DT"])
    fp = pd_mod.File_Processor(f"/nowhere/{name}.{ext}", "", table)
    for k, v in c.items():
        setattr(fp, k, v)
    want = fp.merge_and_filter_labels([None] * n_img)
    got = labels.merge_and_filter_labels(table, name, ext, n_img, c)
    assert H.annotations_from_frame(got) == H.annotations_from_frame(want)
    # and the padding steps, against the reference's split_power_spec on an index image (its reflect of a row of column
    # numbers IS the source map)
    T = (n_img - 1) * 819 + int(rng.integers(1, 1024))
    fp.spectrogram_length = T
    img = np.arange(T, dtype=np.float64)[None, :].repeat(2, axis=0)
    tiles = fp.split_power_spec([img])
    w_last = T - (len(tiles) - 1) * 819
    src = labels.labelled_pad_map(w_last, 1024, labels.empty_width_of(labels.file_rows(table, name), T, c["DT"]))
    np.testing.assert_array_equal(np.asarray(tiles[-1])[0], src + (len(tiles) - 1) * 819)
