"""End-to-end host driver on the GPU: wav -> front-end -> detector (test double) -> fused
post-processing -> per-file merge, against the same pipeline with the CPU oracle doing the
post-processing on the same head outputs; plus file sharding (union of ranks == one rank)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from birdsoundclassif_b200 import synth
from tests.standin_detector import StandInDetector

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def audio_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("audio")
    for i, secs in enumerate([9.0, 4.0, 12.5, 2.0, 7.7]):
        synth.write_wav(str(d / f"rec_{i:02d}.wav"), synth.synth_pcm(secs, 300 + i))
    with open(d / "bird_dict.json", "w") as f:
        json.dump({f"Species {i}": i for i in range(1, 151)}, f)
    return str(d)


def _oracle_run(wav, bird_dict, args, min_score, bs):
    from birdsoundclassif_b200.frontend import File_Processor
    from birdsoundclassif_b200.run_detection import detect_tiles
    from oracle import postproc_oracle as po
    fp = File_Processor(wav)
    tiles, _ = fp.process_file()
    model = StandInDetector(args, backend="oracle").cuda()
    outs = detect_tiles(model, tiles, min_score, bs)
    flat = [d for b in outs for d in b]
    flat_np = [{k: {kk: vv.numpy() for kk, vv in v.items()} for k, v in d.items()} for d in flat]
    merged = po.merge_images(flat_np, w_pix=fp.W_PIX, hop_spectro=fp.HOP_SPECTRO,
                             spectrogram_length=fp.spectrogram_length)
    names = {i: n for n, i in json.load(open(bird_dict)).items()}
    return {names[c]: {k: v.tolist() for k, v in merged[str(c)].items()} for c in range(1, 151)
            if len(merged[str(c)]["bbox_coord"])}


@pytest.mark.parametrize("bs,min_score", [(4, 0.2), (3, 0.05)])
def test_run_detection_matches_oracle_postproc(audio_dir, bs, min_score):
    from birdsoundclassif_b200 import run_detection as rd
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()
    bird = os.path.join(audio_dir, "bird_dict.json")
    total = 0
    for wav in sorted(glob.glob(os.path.join(audio_dir, "*.wav")))[:3]:
        tm = {}
        out = rd.run_detection(model, args, wav, bird, min_score=min_score, bs=bs, timings=tm)
        ref = _oracle_run(wav, bird, args, min_score, bs)
        assert out == ref
        total += tm["detections"]
    assert total > 0, "the stand-in should produce detections, otherwise the test is vacuous"


def test_sharded_directory_union_equals_single(audio_dir, tmp_path):
    from birdsoundclassif_b200 import nbm_detect, sharding
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()
    bird = os.path.join(audio_dir, "bird_dict.json")

    def run(world):
        for f in glob.glob(os.path.join(audio_dir, "*.txt")):
            os.remove(f)
        per_rank = [nbm_detect.detect_directory(model, args, audio_dir, bird, 0.2, 4, r, world, verbose=False)
                    for r in range(world)]
        texts = {os.path.basename(f): open(f).read() for f in sorted(glob.glob(os.path.join(audio_dir, "*.txt")))}
        return per_rank, texts

    one, t1 = run(1)
    two, t2 = run(2)
    assert t1 == t2 and len(t1) == 5
    for k in ("files", "tiles", "detections", "frames"):
        assert one[0][k] == sum(r[k] for r in two)
    assert sharding.totals(two)["files"] == 5


@pytest.mark.parametrize("group_tiles", [1, 9, 1024])
def test_pipelined_directory_equals_file_by_file(tmp_path, group_tiles):
    """pipeline.DetectionPipeline (reader threads -> batched front-end on its own stream -> detector, overlapped
    across groups of files) writes the same .txt as the reference-shaped one-file-at-a-time loop: ragged mono
    files, a stereo pair (its own group), a file shorter than one window, an unreadable file (skipped by both drivers),
    and a 48 kHz file and a 24-bit file (decoded / resampled on the one-file path by both drivers)."""
    from birdsoundclassif_b200 import nbm_detect
    d = tmp_path
    for i, secs in enumerate([9.0, 0.01, 12.5, 2.0, 7.7, 30.0, 3.3]):
        synth.write_wav(str(d / f"rec_{i:02d}.wav"), synth.synth_pcm(secs, 700 + i))
    for i in (0, 1):
        st = np.stack([synth.synth_pcm(5.0 + i, 720 + i), synth.synth_pcm(5.0 + i, 730 + i)], axis=1)
        synth.write_wav(str(d / f"rec_1{i}_stereo.wav"), st)
    raw = open(str(d / "rec_02.wav"), "rb").read()
    (d / "rec_05_truncated.wav").write_bytes(raw[:len(raw) // 3 + 1])      # processed up to where it ends, by both drivers
    (d / "rec_20_broken.wav").write_bytes(b"RIFFjunk")
    synth.write_wav(str(d / "rec_21_48k.wav"), synth.synth_pcm(4.0, 740), sample_rate=48000)
    import struct
    pcm24 = synth.synth_pcm(3.5, 741).astype(np.int32) * 256 + 77
    body = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in pcm24)
    hdr = b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 44100, 44100 * 3, 3, 24) + b"data" + struct.pack("<I", len(body))
    (d / "rec_22_s24.wav").write_bytes(b"RIFF" + struct.pack("<I", len(hdr) + len(body)) + hdr + body)
    bird = str(d / "bird_dict.json")
    with open(bird, "w") as f:
        json.dump({f"Species {i}": i for i in range(1, 151)}, f)
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()

    def run(pipelined):
        for f in glob.glob(str(d / "*.txt")):
            os.remove(f)
        c = nbm_detect.detect_directory(model, args, str(d), bird, 0.05, 4, verbose=False, pipelined=pipelined,
                                        group_tiles=group_tiles)
        return c, {os.path.basename(f): open(f).read() for f in sorted(glob.glob(str(d / "*.txt")))}

    c_seq, t_seq = run(False)
    c_pipe, t_pipe = run(True)
    assert len(t_seq) == 12 and t_pipe == t_seq and "rec_21_48k.txt" in t_seq and "rec_22_s24.txt" in t_seq
    assert c_seq["failed"] == c_pipe["failed"] == 1
    for k in ("files", "tiles", "detections", "frames"):
        assert c_pipe[k] == c_seq[k], k
    assert c_pipe["detections"] > 0


def test_long_recording_detection(tmp_path, monkeypatch):
    """A recording longer than max_l (made small here): run_detection detects each piece as a file of its own and
    shifts its boxes to the recording's time axis; the pipelined driver defers such files to that path."""
    from birdsoundclassif_b200 import frontend, nbm_detect, pipeline
    from birdsoundclassif_b200 import run_detection as rd
    L = 4 * 44100
    monkeypatch.setattr(frontend, "LONG_FILE_SAMPLES", L)
    monkeypatch.setattr(pipeline, "LONG_FILE_SAMPLES", L)
    pcm = synth.synth_pcm(10.5, 810)
    synth.write_wav(str(tmp_path / "a_long.wav"), pcm)
    synth.write_wav(str(tmp_path / "b_short.wav"), synth.synth_pcm(3.0, 811))
    pieces = tmp_path / "pieces"
    pieces.mkdir()
    from tests import helpers as H
    for k in range(3):
        synth.write_wav(str(pieces / f"p{k}.wav"), H.requant_piece(pcm[k * L:(k + 1) * L]))
    bird = str(tmp_path / "bird_dict.json")
    with open(bird, "w") as f:
        json.dump({f"Species {i}": i for i in range(1, 151)}, f)
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()
    tm = {}
    out = rd.run_detection(model, args, str(tmp_path / "a_long.wav"), bird, min_score=0.05, bs=4, timings=tm)
    want = {}
    for k in range(3):
        shift = float(round(k * L / 132))
        for name, v in rd.run_detection(model, args, str(pieces / f"p{k}.wav"), bird, min_score=0.05, bs=4).items():
            e = want.setdefault(name, {"bbox_coord": [], "scores": []})
            e["bbox_coord"] += [[b[0] + shift, b[1], b[2] + shift, b[3]] for b in v["bbox_coord"]]
            e["scores"] += v["scores"]
    assert out.keys() == want.keys() and all(out[k] == want[k] for k in out) and tm["detections"] > 0
    ids = [int(n.split()[1]) for n in out]
    assert ids == sorted(ids)
    texts = {}
    for pipelined in (False, True):
        for f in glob.glob(str(tmp_path / "*.txt")):
            os.remove(f)
        c = nbm_detect.detect_directory(model, args, str(tmp_path), bird, 0.05, 4, verbose=False, pipelined=pipelined,
                                        json_sidecar=True)
        assert c["files"] == 2
        import ast
        for f in glob.glob(str(tmp_path / "*.txt")):                 # the side-car holds the same dictionary
            assert json.load(open(f.replace(".txt", ".json"))) == ast.literal_eval(open(f).read())
        texts[pipelined] = {os.path.basename(f): open(f).read() for f in sorted(glob.glob(str(tmp_path / "*.txt")))}
    assert texts[True] == texts[False] and set(texts[True]) == {"a_long.txt", "b_short.txt"}
    assert texts[True]["a_long.txt"] == str(out)


def test_pipelined_driver_isolates_a_file_the_detector_fails_on(tmp_path):
    """A recording the detector raises on (upstream: "RPN failed" and a crash, layers.py:288-290 -- digital silence
    normalises to 0/0) is reported and has no .txt; every other file of the same front-end group is detected as if
    it were not there."""
    from birdsoundclassif_b200 import nbm_detect
    d = tmp_path
    for i, secs in enumerate([5.0, 9.0, 3.0, 12.5, 2.0]):
        pcm = synth.synth_pcm(secs, 760 + i)
        if i in (0, 3):
            pcm[:] = 0
        synth.write_wav(str(d / f"rec_{i:02d}.wav"), pcm)
    bird = str(d / "bird_dict.json")
    with open(bird, "w") as f:
        json.dump({f"Species {i}": i for i in range(1, 151)}, f)
    args = synth.default_args("cuda")
    inner = StandInDetector(args, backend="nbm").cuda()

    def model(batch, min_score=0.5):
        if torch.isnan(batch).any():
            raise RuntimeError("RPN produced fewer than rcnn_batch_size candidate boxes")
        return inner(batch, min_score=min_score)

    def run(pipelined):
        for f in glob.glob(str(d / "*.txt")):
            os.remove(f)
        c = nbm_detect.detect_directory(model, args, str(d), bird, 0.05, 4, verbose=False, pipelined=pipelined)
        return c, {os.path.basename(f): open(f).read() for f in sorted(glob.glob(str(d / "*.txt")))}

    c_seq, t_seq = run(False)
    c_pipe, t_pipe = run(True)
    assert sorted(t_seq) == ["rec_01.txt", "rec_02.txt", "rec_04.txt"] and t_pipe == t_seq
    assert c_seq["failed"] == c_pipe["failed"] == 2 and c_pipe["files"] == 3
