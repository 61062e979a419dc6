"""GPU post-processing (libnbm_b200 through the reference-named Python symbols) against the
golden vectors recorded from the reference and against the CPU oracle on larger random cases.
NMS keep indices, proposals, final detections and merged outputs must be BIT-EXACT; decode is
exact except where the pre-round value sits on a .5 tie (expf differs by an ulp between libms)."""
import types

import numpy as np
import pytest
import torch

from birdsoundclassif_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp():
    from birdsoundclassif_b200 import postproc
    return postproc


def _names(g):
    return [str(n) for n in g["names"]]


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_nms_golden(pp):
    g = H.load("postproc_nms.npz")
    for name in _names(g):
        th, top = g[f"{name}/param"]
        ob, os_, keep = pp.nms(_cuda(g[f"{name}/boxes"]), _cuda(g[f"{name}/scores"]), float(th), int(top), True)
        assert keep == H.split_keep(g, name), name
        np.testing.assert_array_equal(ob.cpu().numpy(), g[f"{name}/out_boxes"])
        np.testing.assert_array_equal(os_.cpu().numpy(), g[f"{name}/out_scores"])
        ob2, os2 = pp.nms(_cuda(g[f"{name}/boxes"]), _cuda(g[f"{name}/scores"]), float(th), int(top))
        assert torch.equal(ob, ob2) and torch.equal(os_, os2)


@pytest.mark.parametrize("B,N,th,span", [(1, 1, 0.3, 100), (2, 63, 0.5, 200), (3, 64, 0.5, 200), (2, 65, 0.7, 150),
                                         (4, 500, 0.7, 600), (1, 5000, 0.3, 900), (1, 20000, 0.3, 1000),
                                         (2, 777, 0.05, 60)])
def test_nms_random_vs_oracle(pp, B, N, th, span):
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(B * 100003 + N)
    x1 = rng.integers(0, span, (B, N)); y1 = rng.integers(0, 375, (B, N))
    boxes = np.stack([x1, y1, x1 + rng.integers(0, 90, (B, N)), y1 + rng.integers(0, 60, (B, N))], -1).astype(np.float32)
    if N > 10:      # duplicates, degenerate (zero/negative area) and a non-integer box
        boxes[:, 5] = boxes[:, 2]
        boxes[:, 7] = [10, 10, 9, 9]
        boxes[:, 8] = [10, 10, 9, 9]
        boxes[:, 9] = [3.25, 4.5, 40.75, 44.125]
    scores = rng.random((B, N)).astype(np.float32)
    ob, os_, keep = pp.nms(_cuda(boxes), _cuda(scores), th, N, True)
    rb, rs, rkeep = po.nms(boxes, scores, th, N, True)
    assert keep == rkeep
    np.testing.assert_array_equal(ob.cpu().numpy(), rb)
    np.testing.assert_array_equal(os_.cpu().numpy(), rs)


def test_nms_valid_counts(pp):
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(3)
    B, N = 3, 200
    x1 = rng.integers(0, 100, (B, N)); y1 = rng.integers(0, 100, (B, N))
    boxes = np.stack([x1, y1, x1 + 30, y1 + 30], -1).astype(np.float32)
    nv = np.array([200, 17, 0], dtype=np.int32)
    idx, cnt = pp.nms_keep(_cuda(boxes), 0.3, _cuda(nv))
    for b in range(B):
        ref = po.greedy(boxes[b, :nv[b]], 0.3)
        assert idx[b, :cnt[b]].cpu().tolist() == ref


def test_decode_golden(pp):
    from oracle import postproc_oracle as po
    g = H.load("postproc_decode.npz")
    boxes = pp.bbox_reg_to_coord(_cuda(g["deltas"]), _cuda(g["anchors"])).cpu().numpy()
    assert boxes.shape == g["boxes"].shape
    bad = (boxes != g["boxes"]).any(-1)
    ties = po.decode_tie_mask(g["deltas"], g["anchors"])
    assert (bad & ~ties).sum() == 0, f"{(bad & ~ties).sum()} decode mismatches away from .5 ties"
    assert np.abs(boxes - g["boxes"]).max() <= 1
    # per-image anchors (the RoI decode of the tail, layers.py:719)
    b2 = pp.decode_boxes(_cuda(g["deltas"]), _cuda(g["anchors"][None].copy())).cpu().numpy()
    np.testing.assert_array_equal(b2, boxes)
    # anchors built by the library == reference table
    a = pp.make_anchors(16, [0.5, 1, 2], 2 ** np.arange(5), 64, 24, 16, "cuda").cpu().numpy()
    np.testing.assert_array_equal(a, g["anchors"])


def test_proposal_layer_golden(pp):
    g = H.load("postproc_proposal.npz")
    args = synth.default_args("cuda")
    layer = pp.ProposalLayer(args, args.n_layers).eval()
    for name in _names(g):
        rois, sc = layer(_cuda(g[f"{name}/cls"]), _cuda(g[f"{name}/reg"]))
        assert tuple(rois.shape) == g[f"{name}/rois"].shape, name
        np.testing.assert_array_equal(sc.cpu().numpy(), g[f"{name}/scores"])
        np.testing.assert_array_equal(rois.cpu().numpy(), g[f"{name}/rois"])


def test_proposal_layer_random_vs_oracle(pp):
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(11)
    args = synth.default_args("cuda")
    layer = pp.ProposalLayer(args, args.n_layers).eval()
    for B in (1, 4):
        logits = rng.standard_normal((B, 15, 2, 24, 64)).astype(np.float32)
        cls = torch.from_numpy(logits).softmax(2).reshape(B, 30, 24, 64).numpy()
        reg = (rng.standard_normal((B, 60, 24, 64)) * 0.3).astype(np.float32)
        rois, sc = layer(_cuda(cls), _cuda(reg))
        rr, rs = po.proposal_layer(cls, reg)
        # stable order on both sides, so ties are harmless here; decode ties could differ in principle
        np.testing.assert_array_equal(sc.cpu().numpy(), rs)
        np.testing.assert_array_equal(rois.cpu().numpy(), rr)


def test_final_detections_golden(pp):
    g = H.load("postproc_tail.npz")
    args = synth.default_args("cuda")
    for name in _names(g):
        dets = pp.fastrcnn_inference_tail(_cuda(g[f"{name}/bbox_reg"]), _cuda(g[f"{name}/probs"]),
                                          _cuda(g[f"{name}/rois"]), args, 0.3, float(g[f"{name}/min_score"]))
        counts, bb, ss = H.dets_to_flat(dets, 150)
        np.testing.assert_array_equal(counts, g[f"{name}/counts"])
        np.testing.assert_array_equal(ss, g[f"{name}/scores"])
        np.testing.assert_array_equal(bb, g[f"{name}/boxes"])
        d0 = dets[0]
        for c in range(1, 151):     # container conventions of layers.py:753-775
            e = d0[str(c)]
            if len(e["bbox_coord"]):
                assert e["bbox_coord"].dim() == 2 and e["scores"].dim() == 2 and e["scores"].shape[0] == 1
            else:
                assert not e["bbox_coord"].is_cuda and e["bbox_coord"].numel() == 0


def _tiles_from_flat(counts, boxes, scores, device="cuda"):
    tiles, o = [], 0
    for i in range(counts.shape[0]):
        d = {}
        for c in range(counts.shape[1]):
            n = int(counts[i, c])
            if n:
                d[str(c + 1)] = dict(bbox_coord=_cuda(boxes[o:o + n]), scores=_cuda(scores[o:o + n])[None])
            else:
                d[str(c + 1)] = dict(bbox_coord=torch.Tensor(), scores=torch.Tensor())
            o += n
        tiles.append(d)
    return tiles


def test_merge_images_golden(pp):
    g = H.load("postproc_merge.npz")
    for name in _names(g):
        tiles = _tiles_from_flat(g[f"{name}/in_counts"], g[f"{name}/in_boxes"], g[f"{name}/in_scores"])
        fp = types.SimpleNamespace(W_PIX=1024, HOP_SPECTRO=819, spectrogram_length=int(g[f"{name}/spec_len"]))
        batches = [tiles[i:i + 2] for i in range(0, len(tiles), 2)]
        merged = pp.merge_images(fp, batches, 150)
        counts, bb, ss = H.dets_to_flat([merged], 150)
        np.testing.assert_array_equal(counts[0], g[f"{name}/out_counts"])
        np.testing.assert_array_equal(bb, g[f"{name}/out_boxes"])
        np.testing.assert_array_equal(ss, g[f"{name}/out_scores"])


def test_merge_stress_vs_oracle(pp):
    """Dense bursts: thousands of candidates in one file (BASELINE config 5)."""
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(21)
    n_tiles, per = 60, 80
    boxes = np.zeros((n_tiles * per, 4), np.float32)
    x1 = rng.integers(0, 960, n_tiles * per); y1 = rng.integers(0, 330, n_tiles * per)
    boxes[:, 0], boxes[:, 1] = x1, y1
    boxes[:, 2] = np.minimum(x1 + rng.integers(5, 250, n_tiles * per), 1023)
    boxes[:, 3] = np.minimum(y1 + rng.integers(5, 60, n_tiles * per), 374)
    scores = rng.random(n_tiles * per).astype(np.float32)
    classes = rng.integers(1, 12, n_tiles * per).astype(np.int32)
    tiles = np.repeat(np.arange(n_tiles), per).astype(np.int32)
    spec_len = (n_tiles - 1) * 819 + 600
    kb, ks, kc = pp.merge_flat(_cuda(boxes), _cuda(scores), _cuda(classes), _cuda(tiles), n_tiles, 1024, 819, spec_len)
    # oracle on the dict form
    tdicts = []
    for i in range(n_tiles):
        d = {}
        sl = slice(i * per, (i + 1) * per)
        for c in range(1, 151):
            w = np.nonzero(classes[sl] == c)[0]
            d[str(c)] = dict(bbox_coord=boxes[sl][w], scores=scores[sl][w][None]) if len(w) else \
                dict(bbox_coord=np.zeros((0,), np.float32), scores=np.zeros((0,), np.float32))
        tdicts.append(d)
    ref = po.merge_images(tdicts, spectrogram_length=spec_len)
    got = pp.survivors_to_class_dict(kb, ks, kc, 150)
    c1, b1, s1 = H.dets_to_flat([got], 150)
    c2, b2, s2 = H.dets_to_flat([ref], 150)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(b1, b2)
    np.testing.assert_array_equal(s1, s2)


def test_cpu_tensors_are_rejected(pp):
    from birdsoundclassif_b200 import _lib
    with pytest.raises(_lib.NbmError):
        pp.nms(torch.zeros(1, 4, 4), torch.zeros(1, 4))


def test_roi_pooling_golden_and_oracle(pp):
    """ROIPooling.forward (layers.py:399-497) in one launch: pyramid levels, pooled features and pooled
    positional encoding bit-exact against the vectors recorded from the reference class, then a larger
    random case against the oracle."""
    import zlib
    from oracle import postproc_oracle as po
    g = H.load("postproc_roipool.npz")
    feats = synth.fpn_features(601, 2, 8, 5)
    assert np.uint32(zlib.crc32(b"".join(f.tobytes() for f in feats))) == g["feat_crc"]
    cfg = synth.default_args("cuda")
    cfg.out_fpn_chan = 8
    layer = pp.ROIPooling(cfg)
    np.testing.assert_array_equal(layer._tables(torch.device("cuda"))[0].cpu().numpy(), g["pe_freq"])
    pool, pe, lvl = layer(_cuda(g["rois"]), [_cuda(f) for f in feats])
    np.testing.assert_array_equal(lvl, g["lvl"])
    np.testing.assert_array_equal(pool.cpu().numpy(), g["pool"])
    np.testing.assert_array_equal(pe.cpu().numpy(), g["pe"])
    # detector-sized random case: bs 4 x 50 RoIs, 16 channels
    rng = np.random.default_rng(77)
    feats = synth.fpn_features(78, 4, 16, 5)
    x1 = rng.integers(0, 900, (4, 50)); y1 = rng.integers(0, 330, (4, 50))
    rois = np.stack([x1, y1, np.minimum(x1 + rng.integers(5, 300, (4, 50)), 1023),
                     np.minimum(y1 + rng.integers(5, 120, (4, 50)), 374)], -1).astype(np.float32)
    cfg.out_fpn_chan = 16
    layer = pp.ROIPooling(cfg)
    pe_f, pe_t = [t.cpu().numpy() for t in layer._tables(torch.device("cuda"))]
    pool, pe, lvl = layer(_cuda(rois), [_cuda(f) for f in feats])
    rp, rpe, rl = po.roi_pool(rois, feats, pe_f, pe_t)
    np.testing.assert_array_equal(lvl, rl)
    np.testing.assert_array_equal(pool.cpu().numpy(), rp)
    np.testing.assert_array_equal(pe.cpu().numpy(), rpe)


def test_flat_record_fast_paths_equal_dictionary_paths(pp):
    """records_to_dicts attaches the flat record to each per-image dictionary; the per-file merge taking those records
    (three concatenations) and merge_to_output (one D2H copy) must equal the dictionary walk + the reference's output
    comprehension (run_detection.py:69-77) bit for bit."""
    rng = np.random.default_rng(33)
    B, R, ncls = 7, 50, 150
    x1 = rng.integers(0, 900, (B, R)); y1 = rng.integers(0, 300, (B, R))
    boxes = np.stack([x1, y1, np.minimum(x1 + rng.integers(6, 220, (B, R)), 1023),
                      np.minimum(y1 + rng.integers(6, 70, (B, R)), 374)], -1).astype(np.float32)
    scores = np.sort(rng.random((B, R)).astype(np.float32), axis=1)[:, ::-1].copy()
    classes = rng.integers(1, 9, (B, R)).astype(np.int32)
    counts = np.array([50, 0, 17, 50, 1, 33, 50], np.int32)
    dicts = pp.records_to_dicts(_cuda(boxes), _cuda(scores), _cuda(classes), _cuda(counts), ncls, 50)
    assert all(isinstance(d, pp.TileDetections) and d.flat is not None for d in dicts)
    dev = torch.device("cuda")
    fast = pp.flatten_tile_dicts(dicts, ncls, dev)
    plain = [dict(d) for d in dicts]                                   # ordinary dicts: the 150 x n_tiles walk
    slow = pp.flatten_tile_dicts(plain, ncls, dev)
    # same multiset per (tile, class) in the same relative order -> compare after the stable class sort merge applies
    def canon(t):
        b, s, c, ti = [x.cpu().numpy() for x in t]
        o = np.lexsort((np.arange(len(c)), c, ti))                     # tile-major, class-major, original order
        return b[o], s[o], c[o], ti[o]
    for a, b in zip(canon(fast), canon(slow)):
        np.testing.assert_array_equal(a, b)
    fp = types.SimpleNamespace(W_PIX=1024, HOP_SPECTRO=819, spectrogram_length=(B - 1) * 819 + 700)
    batches_fast = [dicts[i:i + 3] for i in range(0, B, 3)]
    batches_slow = [plain[i:i + 3] for i in range(0, B, 3)]
    rev = {i: f"Species {i}" for i in range(0, ncls + 1)}
    class_bbox = pp.merge_images(fp, batches_slow, ncls)
    want = {rev[idx]: {k: v.cpu().numpy().tolist() for k, v in class_bbox[str(idx)].items()}
            for idx in range(1, len(class_bbox) + 1) if len(class_bbox[str(idx)]["bbox_coord"]) > 0}
    got = pp.merge_to_output(fp, batches_fast, ncls, rev)
    assert got == want and list(got) == list(want) and len(got) > 0
    # a per-class truncation (more than proposal_number rows of one class) invalidates the record
    d2 = pp.records_to_dicts(_cuda(boxes), _cuda(scores), _cuda(np.ones_like(classes)), _cuda(counts), ncls, 10)
    assert d2[0].flat is None and len(d2[0]["1"]["bbox_coord"]) == 10
    assert d2[4].flat is not None
