"""Host logic of the pipelined multi-file driver: frame / tile arithmetic against the
oracle's restatement of prepare_dataset.py:234-266, grouping invariants, wav probing and decoding
into a shared buffer.  No GPU."""
import wave

import numpy as np
import pytest

from birdsoundclassif_b200 import pipeline as pl
from birdsoundclassif_b200 import synth
from birdsoundclassif_b200.frontend import LONG_FILE_SAMPLES, derive_constants
from oracle import frontend_oracle as fo

CONST = derive_constants()


@pytest.mark.parametrize("n,frames,tiles", [(441_000, 3341, 4), (1_323_000, 10_023, 12), (2_646_000, 20_046, 25),
                                            (26_460_000, 200_455, 245), (149_984_100, 1_136_244, 1388)])
def test_counts_match_survey_table(n, frames, tiles):
    assert pl.count_frames_tiles(n, CONST["HOP_LENGTH"], CONST["W_PIX"], CONST["HOP_SPECTRO"]) == (frames, tiles)


def test_counts_match_oracle_on_ragged_sizes():
    p = fo.derive_params()
    rng = np.random.default_rng(5)
    sizes = [0, 1, 131, 132, 133, 1323, 1324, 135_167, 135_168, 135_300, 243_275, 243_276, 243_407, 243_408] + \
        rng.integers(0, 3_000_000, 40).tolist()
    for n in sizes:
        T = fo.n_frames(n, p)
        assert pl.count_frames_tiles(n, p.hop, p.w_pix, p.hop_spectro) == (T, fo.n_tiles(T, p)), n
    # several STFT chunks per file (prepare_dataset.py:234-237), at a small chunk size
    for n in [99_999, 100_000, 100_001, 250_000, 300_000]:
        frames = sum(1 + max(0, min(n, (c + 1) * 100_000) - c * 100_000) // p.hop for c in range(n // 100_000 + 1))
        assert pl.count_frames_tiles(n, p.hop, p.w_pix, p.hop_spectro, stft_chunk=100_000)[0] == frames


def _info(i, seconds, ch=1, sr=44100, error=None):
    return pl.WavInfo(f"f{i:03d}.wav", int(seconds * 44100), ch, sr, error)


def test_plan_groups_invariants():
    rng = np.random.default_rng(9)
    infos = [_info(i, s) for i, s in enumerate(rng.uniform(0.5, 90, 60))]
    for budget in (1, 7, 40, 10_000):
        groups, rejected = pl.plan_groups(infos, CONST, budget)
        assert not rejected
        assert [f.path for g in groups for f in g.files] == [f.path for f in infos]        # order kept, nothing split or lost
        for g in groups:
            assert g.n_tiles <= budget or len(g.files) == 1                                  # an over-budget file stands alone
            assert g.n_values == sum(f.n_samples for f in g.files)
            for f, nt, fr in zip(g.files, g.tiles, g.frames):
                assert (fr, nt) == pl.count_frames_tiles(f.n_samples, CONST["HOP_LENGTH"], CONST["W_PIX"], CONST["HOP_SPECTRO"])
        if budget == 10_000:
            assert len(groups) == 1
        # greedy: a group is closed only when the next file would not fit
        for a, b in zip(groups, groups[1:]):
            assert a.n_tiles + b.tiles[0] > budget


def test_plan_groups_ramp():
    infos = [_info(i, 30) for i in range(200)]                      # 12 tiles each
    groups, _ = pl.plan_groups(infos, CONST, 1024, first_group_tiles=64)
    assert [f.path for g in groups for f in g.files] == [f.path for f in infos]
    budgets = [pl.group_budget(i, 1024, 64) for i in range(len(groups))]
    assert budgets[:6] == [64, 128, 256, 512, 1024, 1024]
    assert all(g.n_tiles <= b for g, b in zip(groups, budgets))
    assert [len(g.files) for g in groups[:4]] == [5, 10, 21, 42]
    assert pl.group_budget(3, 100, None) == 100 and pl.group_budget(50, 1 << 20, 64) == 1 << 20


def test_plan_groups_channels_and_rejects():
    infos = [_info(0, 3), _info(1, 3), _info(2, 3, ch=2), _info(3, 3, ch=2), _info(4, 3),
             _info(5, 3, sr=48000), _info(6, 0, error="File loading failed (x)"),
             pl.WavInfo("long.wav", LONG_FILE_SAMPLES + 1, 1, 44100)]
    groups, rejected = pl.plan_groups(infos, CONST, 1000)
    assert [[f.path for f in g.files] for g in groups] == [["f000.wav", "f001.wav"], ["f002.wav", "f003.wav"], ["f004.wav"]]
    assert [g.channels for g in groups] == [1, 2, 1]
    assert groups[1].n_values == 2 * 2 * 3 * 44100
    assert [r[0].path for r in rejected] == ["f005.wav", "f006.wav", "long.wav"]
    assert rejected[0][1] == pl.SOLO_REASON and rejected[2][1] == pl.LONG_REASON
    # another sample format goes the same way as another rate: decoded on the one-file path
    g2, r2 = pl.plan_groups([pl.WavInfo("f24.wav", 44100, 1, 44100, pcm16=False), _info(1, 3)], CONST, 1000)
    assert [f.path for g in g2 for f in g.files] == ["f001.wav"] and r2[0][1] == pl.SOLO_REASON


def test_probe_and_read_into(tmp_path):
    mono = synth.synth_pcm(1.5, 3)
    stereo = np.stack([synth.synth_pcm(0.7, 4), synth.synth_pcm(0.7, 5)], axis=1)
    pm, ps = str(tmp_path / "m.wav"), str(tmp_path / "s.wav")
    synth.write_wav(pm, mono)
    synth.write_wav(ps, stereo)
    im, is_ = pl.probe_wav(pm), pl.probe_wav(ps)
    assert (im.n_samples, im.channels, im.sample_rate, im.error) == (len(mono), 1, 44100, None)
    assert (is_.n_samples, is_.channels) == (len(stereo), 2)
    buf = np.full(len(mono) + 2 * len(stereo) + 8, 7, dtype=np.int16)
    pl.read_into(im, buf[:len(mono)])
    pl.read_into(is_, buf[len(mono):len(mono) + 2 * len(stereo)])
    assert np.array_equal(buf[:len(mono)], mono)
    assert np.array_equal(buf[len(mono):len(mono) + 2 * len(stereo)].reshape(-1, 2), stereo)
    assert (buf[-8:] == 7).all()
    # failures: not a wav, 8-bit wav, truncated data
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wav at all")
    assert "File loading failed" in pl.probe_wav(str(bad)).error
    with wave.open(str(tmp_path / "u8.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(44100); w.writeframes(bytes(100))
    iu = pl.probe_wav(str(tmp_path / "u8.wav"))
    assert iu.error is None and not iu.pcm16 and iu.n_samples == 100      # decodable, but not on the batched PCM16 path
    # truncated data (header promises more than the file holds, cut in the middle of a frame): both readers yield the
    # whole frames that are there
    raw = open(pm, "rb").read()
    (tmp_path / "cut.wav").write_bytes(raw[:44 + 2 * 1000 + 1])
    ic = pl.probe_wav(str(tmp_path / "cut.wav"))
    assert ic.error is None and ic.n_samples == 1000
    got = np.zeros(1000, dtype=np.int16)
    pl.read_into(ic, got)
    seq, _ = synth.read_wav_pcm16(str(tmp_path / "cut.wav"))
    from birdsoundclassif_b200 import audio_io
    seq2, sr2 = audio_io.read_wav(str(tmp_path / "cut.wav"))
    assert np.array_equal(got, mono[:1000]) and np.array_equal(seq, mono[:1000]) and np.array_equal(seq2, mono[:1000]) and sr2 == 44100


def test_pipeline_needs_cuda(tmp_path):
    import json
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    d = tmp_path / "bird_dict.json"
    d.write_text(json.dumps({"A": 1}))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pl.DetectionPipeline(None, synth.default_args("cpu"), str(d))


@pytest.mark.parametrize("n_lanes", [1, 2, 3])
def test_detect_stream_schedule(n_lanes):
    """The software pipeline of GraphedDetector.detect_stream (host logic only, steps stubbed): every recording's
    batches come back in order; a lane's static buffers are reused only after its previous batch went through the
    second graph; and the host copy a second graph writes is read (`_finish`) before that lane's next second graph."""
    import torch
    from birdsoundclassif_b200.graphed import GraphedDetector, _Lane
    det = GraphedDetector.__new__(GraphedDetector)
    det._lanes = [_Lane() for _ in range(n_lanes)]
    log, state = [], {"n": 0}
    lane_of = {}

    def stage1(samples, nms_thresh, min_score, lane):
        b = state["n"]; state["n"] += 1
        li = det._lanes.index(lane)
        prev = [x for x in lane_of if lane_of[x] == li]
        assert all(("s2", x) in log for x in prev), "lane reused before its previous batch's second graph"
        lane_of[b] = li
        log.append(("s1", b))
        return ("started", b, int(samples.shape[0]))

    def stage2(st):
        _, b, n = st
        assert ("s1", b) in log
        earlier_same_lane = [x for x in lane_of if lane_of[x] == lane_of[b] and x < b]
        assert all(("fin", x) in log for x in earlier_same_lane), "host copy overwritten before it was read"
        log.append(("s2", b))
        return ("pending", b, n)

    def finish(p):
        _, b, n = p
        log.append(("fin", b))
        return [("tile", b)] * n

    det._stage1, det._stage2, det._finish = stage1, stage2, finish
    sizes = [1, 9, 0, 6, 4, 13]
    files = [torch.zeros((n, 2, 2)) for n in sizes]
    got = list(det.detect_stream(iter(files), 0.2, 4))
    assert [len(g) for g in got] == [(n + 3) // 4 for n in sizes]
    flat = [b for g in got for out in g for (_, b) in out[:1]]
    assert flat == list(range(sum((n + 3) // 4 for n in sizes)))           # batches in order, none lost
    assert [len(out) for g in got for out in g] == [min(4, n - s) for n in sizes for s in range(0, n, 4)]
    assert det.detect_tiles(files[1], 0.2, 4) is not None
    # at most n_lanes first graphs are in flight at any time
    inflight = 0
    for ev, _ in log:
        inflight += ev == "s1"
        inflight -= ev == "s2"
        assert inflight <= n_lanes


def test_detect_stream_names_the_failing_recording():
    """The stream runs ahead of the recording the caller waits for: an exception carries the ordinal of the recording
    whose batch raised (`nbm_file_index`), here the third one while the caller is still collecting the second."""
    import torch
    from birdsoundclassif_b200.graphed import GraphedDetector, _Lane
    det = GraphedDetector.__new__(GraphedDetector)
    det._lanes = [_Lane(), _Lane()]
    n = {"b": 0}

    def stage1(samples, nms_thresh, min_score, lane):
        n["b"] += 1
        return n["b"] - 1

    def stage2(b):
        if b == 3:                          # recordings of 2, 1, 2 batches: batch 3 is the first of the third
            raise RuntimeError("RPN failed")
        return b

    det._stage1, det._stage2, det._finish = stage1, stage2, lambda b: [b]
    files = [torch.zeros((8, 2, 2)), torch.zeros((3, 2, 2)), torch.zeros((5, 2, 2))]
    got = []
    with pytest.raises(RuntimeError) as ei:
        for out in det.detect_stream(iter(files), 0.2, 4):
            got.append(out)
    assert ei.value.nbm_file_index == 2 and got == [[[0], [1]]]       # the second recording was still in flight: the caller re-runs it


def test_gc_paused_restores_state():
    import gc
    from birdsoundclassif_b200 import postproc
    assert gc.isenabled()
    with postproc.gc_paused():
        assert not gc.isenabled()
        with postproc.gc_paused():          # nested: the inner exit must not switch it back on
            assert not gc.isenabled()
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with postproc.gc_paused():
            pass
        assert not gc.isenabled()           # it was off before: stays off
    finally:
        gc.enable()
    with pytest.raises(ValueError):
        with postproc.gc_paused():
            raise ValueError("x")
    assert gc.isenabled()


def test_pinned_pool_take_and_give():
    import torch
    from birdsoundclassif_b200 import pipeline
    if not torch.cuda.is_available():
        pytest.skip("pinned host memory needs the CUDA runtime")
    a = pipeline._take_pinned(1000)
    assert a.is_pinned() and a.numel() >= 1000 and a.dtype == torch.int16
    pipeline._give_pinned([a])
    b = pipeline._take_pinned(10)
    assert b.data_ptr() == a.data_ptr()      # reused
    pipeline._give_pinned([b])
