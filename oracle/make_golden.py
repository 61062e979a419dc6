"""Generate tests/golden/*.npz by executing the REFERENCE's own code (container only).

    python -m oracle.make_golden            # from the repo root; needs /root/reference

Front-end vectors come from the reference ``File_Processor.process_file`` run unmodified
through ``oracle.ref_shims`` (only ``librosa.load/stft`` are supplied, see that module);
post-processing vectors come from the reference ``nms``, ``bbox_reg_to_coord``,
``ProposalLayer``, the inference branch of ``FastRCNN.forward``, ``merge_images`` and ``ROIPooling``.
The fixtures are small (strided pixel samples, not full tiles) and are regenerated
bit-identically by this script (seeded inputs).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from birdsoundclassif_b200 import synth          # noqa: E402
from oracle import ref_shims                      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (name, seconds, seed, process_file kwargs)
FRONTEND_CASES = [
    ("fe_2s", 2.0, 11, {}),                     # one tile, single reflect pad
    ("fe_half_s", 0.5, 12, {}),                 # one tile, iterated reflect pad (pad > w-1)
    ("fe_6s", 6.2, 13, {}),                     # three tiles, last padded
    ("fe_exact", (1024 + 819) * 132 / 44100.0 - 0.001, 14, {}),   # T = 1843 -> 2 tiles, no pad
    ("fe_stress", 1.0, 15, dict(freq_accuracy=10.0, dt=0.001)),   # n_fft 4410, hop 44
]
ROW_STRIDE, COL_STRIDE = 5, 7


def frontend_golden():
    pd = ref_shims.ref("nbm_model.nbm_datasets.prepare_dataset")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, secs, seed, kw in FRONTEND_CASES:
            pcm = synth.synth_pcm(secs, seed)
            path = synth.write_wav(os.path.join(d, name + ".wav"), pcm)
            fp = pd.File_Processor(path)
            tiles, _ = fp.process_file(**kw)
            tiles32 = np.stack([np.asarray(t) for t in tiles])              # float64 [n,375,1024]
            out[name + "/pcm_crc"] = np.uint32(zlib.crc32(pcm.tobytes()))
            out[name + "/n_samples"] = np.int64(len(pcm))
            out[name + "/n_tiles"] = np.int64(len(tiles))
            out[name + "/spectrogram_length"] = np.int64(fp.spectrogram_length)
            out[name + "/sample"] = tiles32[:, ::ROW_STRIDE, ::COL_STRIDE].astype(np.float32)
            out[name + "/tile_sum"] = tiles32.sum(axis=(1, 2))
            out[name + "/last_col"] = tiles32[:, :, -1].astype(np.float32)
            out[name + "/consts"] = np.array([fp.W_PIX, fp.HOP_SPECTRO, fp.WIN_LENGTH, fp.HOP_LENGTH,
                                              fp.LOW_IDX, fp.HIGH_IDX], dtype=np.int64)
            out[name + "/fconsts"] = np.array([fp.FREQ_ACCURACY, fp.DT, fp.LOW_FREQ, fp.HIGH_FREQ])
            print(name, len(tiles), fp.spectrogram_length)
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **out)


def _rand_boxes(rng, n, w=1024, h=375, max_side=200):
    x1 = rng.integers(0, w - 6, n)
    y1 = rng.integers(0, h - 6, n)
    bw = rng.integers(5, max_side, n)
    bh = rng.integers(5, max_side // 2, n)
    return np.stack([x1, y1, np.minimum(x1 + bw, w - 1), np.minimum(y1 + bh, h - 1)], 1).astype(np.float32)


def nms_golden():
    nu = ref_shims.ref("nbm_model.nets.util.nets_utils")
    rng = np.random.default_rng(100)
    out, cases = {}, []
    # (name, B, N, thresh, topN)
    for name, B, N, th, top in [("rpn", 4, 500, 0.7, 50), ("final", 1, 50, 0.3, 50), ("one", 2, 1, 0.3, 300),
                                ("dense", 2, 300, 0.3, 300), ("big", 1, 2000, 0.3, 2000), ("wide", 3, 97, 0.5, 20)]:
        boxes = np.stack([_rand_boxes(rng, N, max_side=60 if name == "dense" else 200) for _ in range(B)])
        if name == "dense":         # cluster boxes so most get suppressed; add exact duplicates
            boxes[:, :, [0, 2]] = boxes[:, :, [0, 2]] % 120 + np.array([0, 0])
            boxes[:, :, 2] = np.maximum(boxes[:, :, 2], boxes[:, :, 0] + 5)
            boxes[:, 10] = boxes[:, 3]
            boxes[:, 200] = boxes[:, 3]
        scores = -np.sort(-rng.random((B, N)).astype(np.float32), axis=1)
        cases.append((name, boxes, scores, th, top))
    # IoU exactly on the threshold: 10x10 boxes shifted to give inter/union = 3/10 .. and 7/10
    b = np.array([[0, 0, 9, 9], [0, 0, 9, 5], [0, 4, 9, 9], [20, 20, 29, 29], [20, 20, 29, 26],
                  [50, 50, 59, 59], [50, 50, 59, 59], [100, 100, 99, 99], [100, 100, 99, 99],
                  [5, 5, 2, 2], [0, 0, 9, 9]], dtype=np.float32)[None]
    s = np.linspace(1, 0.1, b.shape[1], dtype=np.float32)[None]
    cases += [("edge03", b, s, 0.3, 300), ("edge07", b, s, 0.7, 300), ("edge06", b, s, 0.6, 300)]
    for name, boxes, scores, th, top in cases:
        rb, rs, idx = nu.nms(torch.from_numpy(boxes), torch.from_numpy(scores), nms_thresh=th,
                             post_nms_topN=top, return_idx=True)
        out[f"{name}/boxes"], out[f"{name}/scores"] = boxes, scores
        out[f"{name}/param"] = np.array([th, top], dtype=np.float64)
        out[f"{name}/out_boxes"], out[f"{name}/out_scores"] = rb.numpy(), rs.numpy()
        out[f"{name}/keep_len"] = np.array([len(k) for k in idx], dtype=np.int64)
        out[f"{name}/keep_flat"] = np.concatenate([np.asarray(k, dtype=np.int64) for k in idx])
        print("nms", name, [len(k) for k in idx])
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(GOLD, "postproc_nms.npz"), **out)


def decode_golden():
    nu = ref_shims.ref("nbm_model.nets.util.nets_utils")
    layers = ref_shims.ref("nbm_model.nets.layers")
    rng = np.random.default_rng(200)
    anchors = (layers.generate_anchors(base_size=16, ratios=[0.5, 1, 2], scales=2 ** np.arange(5))
               + layers.get_anchor_shifts(64, 24, 16)).reshape(-1, 4)
    deltas = (rng.standard_normal((1, anchors.shape[0], 4)) * np.array([0.3, 0.3, 0.4, 0.4])).astype(np.float32)
    boxes = nu.bbox_reg_to_coord(torch.from_numpy(deltas), torch.Tensor(anchors)).numpy()
    np.savez_compressed(os.path.join(GOLD, "postproc_decode.npz"), anchors=anchors.astype(np.float32),
                        deltas=deltas, boxes=boxes)
    print("decode", boxes.shape)


def _args(device="cpu"):
    return synth.default_args(device)


def proposal_golden():
    layers = ref_shims.ref("nbm_model.nets.layers")
    rng = np.random.default_rng(300)
    args = _args()
    pl = layers.ProposalLayer(args, args.n_layers).eval()
    out = {}
    for name, B, sigma in [("p3", 3, 0.25), ("p1", 1, 0.5), ("p2_few", 2, 0.1)]:
        while True:     # tie-free foreground scores: the reference's argsort is unspecified on ties
            logits = rng.standard_normal((B, 15, 2, 24, 64)).astype(np.float32)
            cls = torch.from_numpy(logits).softmax(2).reshape(B, 30, 24, 64)
            fg = -np.sort(-cls.permute(0, 2, 3, 1).reshape(B, -1, 2)[..., 1].numpy(), axis=1)[:, :2000]
            if all(len(np.unique(fg[b])) == fg.shape[1] for b in range(B)):
                break
        reg = (rng.standard_normal((B, 60, 24, 64)) * sigma).astype(np.float32)
        if name == "p2_few":     # collapse almost every box below the 5 px minimum -> "RPN failed" branch
            reg = reg.reshape(B, 15, 4, 24, 64)
            reg[:, :, 2:] = -8.0
            reg = reg.reshape(B, 60, 24, 64)
        rois, sc = pl(cls, torch.from_numpy(reg))
        out[f"{name}/cls"], out[f"{name}/reg"] = cls.numpy(), reg
        out[f"{name}/rois"], out[f"{name}/scores"] = rois.numpy(), sc.numpy()
        print("proposal", name, tuple(rois.shape))
    out["names"] = np.array(["p3", "p1", "p2_few"])
    np.savez_compressed(os.path.join(GOLD, "postproc_proposal.npz"), **out)


def _ref_tail(layers, args, bbox_reg, probs, rois, nms_thresh, min_score):
    """Run the inference branch of the reference FastRCNN.forward on given head outputs."""
    fr = layers.FastRCNN.__new__(layers.FastRCNN)
    torch.nn.Module.__init__(fr)
    fr.config = args
    fr.roi_pooling = lambda rois_, conv: (None, None, None)
    fr.rcnn = lambda a, b: (torch.from_numpy(bbox_reg), torch.from_numpy(probs))
    fr.eval()
    with torch.no_grad():
        return fr.forward(None, torch.from_numpy(rois), nms_thresh=nms_thresh, min_score=min_score)


def _flatten_dets(dets, num_classes):
    """list(B) of dict -> arrays: counts [B, C], boxes [sum,4], scores [sum] (class-major per image)."""
    counts = np.zeros((len(dets), num_classes), dtype=np.int64)
    bb, ss = [], []
    for b, d in enumerate(dets):
        for c in range(1, num_classes + 1):
            e = d[str(c)]
            n = len(e["bbox_coord"])
            counts[b, c - 1] = n
            if n:
                bb.append(np.asarray(e["bbox_coord"], np.float32).reshape(-1, 4))
                ss.append(np.asarray(e["scores"], np.float32).reshape(-1))
    return counts, (np.concatenate(bb) if bb else np.zeros((0, 4), np.float32)), \
        (np.concatenate(ss) if ss else np.zeros((0,), np.float32))


def tail_golden():
    layers = ref_shims.ref("nbm_model.nets.layers")
    rng = np.random.default_rng(400)
    args = _args()
    C = args.num_classes
    out, names = {}, []
    for name, B, R, n_hot, min_score in [("t4", 4, 50, 12, 0.2), ("t1", 1, 50, 5, 0.5), ("t_bg", 2, 16, 0, 0.0)]:
        logits = rng.standard_normal((B * R, C + 1)).astype(np.float32)
        hot = rng.integers(1, n_hot + 1, B * R) if n_hot else np.zeros(B * R, dtype=np.int64)
        logits[np.arange(B * R), hot] += rng.uniform(3, 9, B * R).astype(np.float32)
        probs = torch.from_numpy(logits).softmax(1).numpy()
        bbox_reg = (rng.standard_normal((B * R, 4 * (C + 1))) * 0.2).astype(np.float32)
        centres = rng.integers(0, 6, (B, R))
        rois = np.zeros((B, R, 4), dtype=np.float32)
        for b in range(B):
            base = _rand_boxes(rng, 6)
            jit = rng.integers(-6, 7, (R, 4)).astype(np.float32)
            rois[b] = np.clip(base[centres[b]] + jit, 0, [1023, 374, 1023, 374])
            rois[b, :, 2] = np.maximum(rois[b, :, 2], rois[b, :, 0] + 5)
            rois[b, :, 3] = np.maximum(rois[b, :, 3], rois[b, :, 1] + 5)
        dets = _ref_tail(layers, args, bbox_reg, probs, rois, 0.3, min_score)
        counts, bb, ss = _flatten_dets(dets, C)
        out[f"{name}/bbox_reg"], out[f"{name}/probs"], out[f"{name}/rois"] = bbox_reg, probs, rois
        out[f"{name}/min_score"] = np.float64(min_score)
        out[f"{name}/counts"], out[f"{name}/boxes"], out[f"{name}/scores"] = counts, bb, ss
        names.append(name)
        print("tail", name, int(counts.sum()))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "postproc_tail.npz"), **out)


def merge_golden():
    rd = ref_shims.ref("nbm_model.run_detection")
    rng = np.random.default_rng(500)
    C = 150
    out, names = {}, []
    for name, n_tiles, per_tile, spec_len in [("m5", 5, 14, 5 * 819 + 150), ("m1", 1, 6, 700), ("m_empty", 3, 0, 3000)]:
        tiles, flat_counts, flat_b, flat_s = [], np.zeros((n_tiles, C), np.int64), [], []
        for i in range(n_tiles):
            d = {}
            cls = rng.integers(1, 9, per_tile)
            boxes = _rand_boxes(rng, per_tile, max_side=260)
            if per_tile:
                boxes[0, 0], boxes[0, 2] = 0, 40            # left-border, short -> dropped unless first tile
                boxes[1, 0], boxes[1, 2] = 980, 1023        # right-border, short
                boxes[2, 0], boxes[2, 2] = 2, 400           # left-border but wide -> kept
            sc = rng.random(per_tile).astype(np.float32)
            for c in range(1, C + 1):
                w = np.nonzero(cls == c)[0]
                if len(w) == 0:
                    d[str(c)] = dict(bbox_coord=torch.Tensor(), scores=torch.Tensor())
                else:
                    d[str(c)] = dict(bbox_coord=torch.from_numpy(boxes[w].copy()),
                                     scores=torch.from_numpy(sc[w].copy())[None])
                    flat_counts[i, c - 1] = len(w)
                    flat_b.append(boxes[w]); flat_s.append(sc[w])
            tiles.append(d)
        fp = types.SimpleNamespace(W_PIX=1024, HOP_SPECTRO=819, spectrogram_length=spec_len)
        # outputs is a list of batches (bs=2 here) of per-tile dicts
        batches = [tiles[i:i + 2] for i in range(0, n_tiles, 2)]
        merged = rd.merge_images(fp, batches, C)
        mc = np.zeros((C,), np.int64)
        mb, ms = [], []
        for c in range(1, C + 1):
            e = merged[str(c)]
            mc[c - 1] = len(e["bbox_coord"])
            if mc[c - 1]:
                mb.append(e["bbox_coord"].numpy().reshape(-1, 4)); ms.append(e["scores"].numpy().reshape(-1))
        out[f"{name}/in_counts"] = flat_counts
        out[f"{name}/in_boxes"] = np.concatenate(flat_b) if flat_b else np.zeros((0, 4), np.float32)
        out[f"{name}/in_scores"] = np.concatenate(flat_s) if flat_s else np.zeros((0,), np.float32)
        out[f"{name}/spec_len"] = np.int64(spec_len)
        out[f"{name}/out_counts"] = mc
        out[f"{name}/out_boxes"] = np.concatenate(mb) if mb else np.zeros((0, 4), np.float32)
        out[f"{name}/out_scores"] = np.concatenate(ms) if ms else np.zeros((0,), np.float32)
        names.append(name)
        print("merge", name, int(mc.sum()))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "postproc_merge.npz"), **out)


def roipool_golden():
    """Reference ROIPooling.forward (layers.py:399-497) on seeded feature maps (8 channels; the maps are
    regenerated from the seed by the tests, only their CRC is stored)."""
    layers = ref_shims.ref("nbm_model.nets.layers")
    pos = ref_shims.ref("nbm_model.nets.position_encoding")
    rng = np.random.default_rng(600)
    args = _args()
    args.out_fpn_chan = 8
    B, R = 2, 14
    feats = synth.fpn_features(601, B, 8, args.n_layers)
    rois = np.stack([_rand_boxes(rng, R, max_side=400) for _ in range(B)])
    rois[0, 0] = [100, 50, 105, 55]          # tiny: grown to 2 x 2 cells
    rois[0, 1] = [0, 0, 1023, 374]           # whole image: x2 rounds to the map width (python slice clamps)
    rois[0, 2] = [1000, 360, 1023, 374]      # bottom-right corner
    rois[1, 0] = [10, 10, 30, 30]            # side 20 -> log2(2.0): level boundary
    pool, pe, lvl = layers.ROIPooling(args)(torch.from_numpy(rois), [torch.from_numpy(f) for f in feats])
    out = {"rois": rois, "pool": pool.numpy(), "pe": pe.numpy(), "lvl": np.asarray(lvl, dtype=np.int32),
           "pe_freq": pos.one_dimension_positional_encoding(args.img_height, 4).numpy(),
           "pe_time": pos.one_dimension_positional_encoding(args.img_width, 4).numpy(),
           "feat_crc": np.uint32(zlib.crc32(b"".join(f.tobytes() for f in feats)))}
    np.savez_compressed(os.path.join(GOLD, "postproc_roipool.npz"), **out)
    print("roipool", pool.shape, np.bincount(np.asarray(lvl).ravel()))


# ------------------------------------------------------------------------------ annotated recordings ----
from tests.helpers import LABEL_CASES, label_table          # noqa: E402  (shared with the tests that read the fixture)


def random_label_table(rng, filename, n, seconds):
    import pandas as pd
    t0 = rng.uniform(0, seconds, n)
    dur = np.where(rng.random(n) < 0.2, rng.uniform(2.0, 8.0, n), rng.uniform(0.01, 0.8, n))
    f0 = rng.uniform(0, 12000, n)
    return pd.DataFrame({"filename": filename, "t_start": t0, "t_end": t0 + dur, "f_start": f0,
                         "f_end": f0 + np.where(rng.random(n) < 0.1, 5.0, rng.uniform(100, 6000, n)),
                         "bird_id": rng.choice([-1, 1, 2, 3, 77, 150], n)})


def _pack_annotations(out, key, ann):
    out[key + "/index"] = np.asarray(ann["index"].values, dtype=np.int64)
    out[key + "/count"] = np.asarray([len(c) for c in ann["coord"]], dtype=np.int64)
    out[key + "/coord"] = np.asarray([list(map(int, b)) for c in ann["coord"] for b in c], dtype=np.int64).reshape(-1, 4)
    out[key + "/bird_id"] = np.asarray([int(b) for c in ann["bird_id"] for b in c], dtype=np.int64)


def labels_golden():
    """The reference's File_Processor WITH a label table (prepare_dataset.py:146-153, 280-292, 297-376)."""
    pd_mod = ref_shims.ref("nbm_model.nbm_datasets.prepare_dataset")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, secs, seed, last_end in LABEL_CASES:
            pcm = synth.synth_pcm(secs, seed)
            path = synth.write_wav(os.path.join(d, name + ".wav"), pcm)
            fp = pd_mod.File_Processor(path, "", label_table(name, last_end))
            tiles, ann = fp.process_file()
            last = np.asarray(tiles[-1])
            out[name + "/pcm_crc"] = np.uint32(zlib.crc32(pcm.tobytes()))
            out[name + "/n_tiles"] = np.int64(len(tiles))
            out[name + "/spectrogram_length"] = np.int64(fp.spectrogram_length)
            out[name + "/last_rows"] = last[[0, 187, 374], :].astype(np.float32)       # three full rows of the padded tile
            out[name + "/last_sum"] = np.float64(last.sum())
            _pack_annotations(out, name, ann)
            print(name, len(tiles), fp.spectrogram_length, len(ann))
    # merge_and_filter_labels alone on random tables (a stub instance: no audio needed), wav and mp3 naming
    rng = np.random.default_rng(5)
    for name, ext, n_img, n in (("rand_wav", "wav", 25, 120), ("rand_mp3", "mp3", 7, 40)):
        fp = pd_mod.File_Processor(f"/nowhere/{name}.{ext}", "", None)
        c = synth_consts()
        for k, v in c.items():
            setattr(fp, k, v)
        table = random_label_table(rng, name, n, (n_img * 819 + 205) * c["DT"])
        fp.labels = table
        ann = fp.merge_and_filter_labels([None] * n_img)
        for col in ("t_start", "t_end", "f_start", "f_end", "bird_id"):
            out[f"{name}/table_{col}"] = table[col].to_numpy()
        out[name + "/n_img"] = np.int64(n_img)
        _pack_annotations(out, name, ann)
        print(name, len(ann))
    np.savez_compressed(os.path.join(GOLD, "labels.npz"), **out)


def synth_consts():
    from birdsoundclassif_b200.frontend import derive_constants
    c = derive_constants()
    c["H_PIX"] = 375
    return c


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    frontend_golden()
    nms_golden()
    decode_golden()
    proposal_golden()
    tail_golden()
    merge_golden()
    roipool_golden()
    labels_golden()


if __name__ == "__main__":
    main()
