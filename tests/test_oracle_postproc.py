"""The post-processing oracle against golden vectors recorded from the reference's own
nms / bbox_reg_to_coord / ProposalLayer / FastRCNN tail / merge_images, and (container
only) against those functions live on fresh random inputs."""
import types

import os

import numpy as np
import pytest

from oracle import postproc_oracle as po
from tests import helpers as H


def _names(g):
    return [str(n) for n in g["names"]]


def test_nms_golden():
    g = H.load("postproc_nms.npz")
    for name in _names(g):
        th, top = g[f"{name}/param"]
        ob, os_, keep = po.nms(g[f"{name}/boxes"], g[f"{name}/scores"], float(th), int(top), True)
        assert keep == H.split_keep(g, name), name
        np.testing.assert_array_equal(ob, g[f"{name}/out_boxes"])
        np.testing.assert_array_equal(os_, g[f"{name}/out_scores"])


def test_nms_threshold_is_float32_and_inclusive():
    # IoU 3/10 computed in float32 equals float32(0.3) -> suppressed (>=)
    b = np.array([[0, 0, 9, 9], [0, 0, 9, 2]], dtype=np.float32)
    assert po.greedy(b, 0.3) == [0]
    assert po.greedy(b, np.nextafter(np.float32(0.3), np.float32(1))) == [0, 1]
    # degenerate boxes: union 0 -> NaN -> never suppressed
    z = np.array([[10, 10, 9, 9], [10, 10, 9, 9]], dtype=np.float32)
    assert po.greedy(z, 0.3) == [0, 1]


def test_anchors_and_decode_golden():
    g = H.load("postproc_decode.npz")
    a = po.make_anchors()
    np.testing.assert_array_equal(a, g["anchors"])
    assert a[0].tolist() == [2, -3, 13, 19] and a.shape == (23040, 4)
    boxes = po.decode(g["deltas"], a)
    bad = (boxes != g["boxes"]).any(-1)
    # numpy's and torch's expf may differ by an ulp, flipping round() only next to a .5 tie
    assert bad.sum() <= po.decode_tie_mask(g["deltas"], a).sum()
    assert (bad & ~po.decode_tie_mask(g["deltas"], a)).sum() == 0
    assert np.abs(boxes - g["boxes"]).max() <= 1


def test_proposal_layer_golden():
    g = H.load("postproc_proposal.npz")
    for name in _names(g):
        rois, sc = po.proposal_layer(g[f"{name}/cls"], g[f"{name}/reg"])
        assert rois.shape == g[f"{name}/rois"].shape, name
        np.testing.assert_array_equal(sc, g[f"{name}/scores"])
        np.testing.assert_array_equal(rois, g[f"{name}/rois"])


def test_final_detections_golden():
    g = H.load("postproc_tail.npz")
    for name in _names(g):
        dets = po.final_detections(g[f"{name}/bbox_reg"], g[f"{name}/probs"], g[f"{name}/rois"],
                                   min_score=float(g[f"{name}/min_score"]))
        counts, bb, ss = H.dets_to_flat(dets, 150)
        np.testing.assert_array_equal(counts, g[f"{name}/counts"])
        np.testing.assert_array_equal(ss, g[f"{name}/scores"])
        np.testing.assert_array_equal(bb, g[f"{name}/boxes"])


def _tiles_from_flat(counts, boxes, scores):
    tiles, o = [], 0
    for i in range(counts.shape[0]):
        d = {}
        for c in range(counts.shape[1]):
            n = int(counts[i, c])
            d[str(c + 1)] = dict(bbox_coord=boxes[o:o + n] if n else np.zeros((0,), np.float32),
                                 scores=scores[o:o + n][None] if n else np.zeros((0,), np.float32))
            o += n
        tiles.append(d)
    return tiles


def test_merge_images_golden():
    g = H.load("postproc_merge.npz")
    for name in _names(g):
        tiles = _tiles_from_flat(g[f"{name}/in_counts"], g[f"{name}/in_boxes"], g[f"{name}/in_scores"])
        merged = po.merge_images(tiles, spectrogram_length=int(g[f"{name}/spec_len"]))
        counts, bb, ss = H.dets_to_flat([merged], 150)
        np.testing.assert_array_equal(counts[0], g[f"{name}/out_counts"])
        np.testing.assert_array_equal(bb, g[f"{name}/out_boxes"])
        np.testing.assert_array_equal(ss, g[f"{name}/out_scores"])


@pytest.mark.reference
def test_nms_live_against_reference():
    import torch
    from oracle import ref_shims
    nu = ref_shims.ref("nbm_model.nets.util.nets_utils")
    rng = np.random.default_rng(7)
    for B, N, th in [(1, 0 + 3, 0.3), (2, 64, 0.5), (3, 257, 0.7), (1, 700, 0.3)]:
        x1 = rng.integers(0, 300, (B, N)); y1 = rng.integers(0, 200, (B, N))
        boxes = np.stack([x1, y1, x1 + rng.integers(0, 80, (B, N)), y1 + rng.integers(0, 80, (B, N))], -1).astype(np.float32)
        scores = rng.random((B, N)).astype(np.float32)
        rb, rs, idx = nu.nms(torch.from_numpy(boxes), torch.from_numpy(scores), th, 100, True)
        ob, os_, keep = po.nms(boxes, scores, th, 100, True)
        assert keep == [list(k) for k in idx]
        np.testing.assert_array_equal(ob, rb.numpy())
        np.testing.assert_array_equal(os_, rs.numpy())


@pytest.mark.reference
def test_default_args_match_reference_parser():
    """synth.DEFAULT_ARGS is the reference training parser's defaults (the schema of `args`)."""
    import ast
    from birdsoundclassif_b200 import synth
    from oracle import ref_shims
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "nbm_model", "train.py")).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            name = node.args[0].value.lstrip("-")
            kw = {k.arg: k.value for k in node.keywords}
            if "default" in kw:
                found[name] = ast.literal_eval(kw["default"])
            elif "action" in kw and ast.literal_eval(kw["action"]) == "store_true":
                found[name] = False
    assert found == synth.DEFAULT_ARGS


def _roipool_case():
    import zlib
    from birdsoundclassif_b200 import synth
    g = H.load("postproc_roipool.npz")
    feats = synth.fpn_features(601, 2, 8, 5)
    assert np.uint32(zlib.crc32(b"".join(f.tobytes() for f in feats))) == g["feat_crc"], "seeded feature maps not reproducible"
    return g, feats


def test_roi_pool_oracle_matches_reference_golden():
    """ROIPooling.forward (layers.py:399-497): levels, pooled features and pooled positional encoding bit-exact."""
    from oracle import postproc_oracle as po
    g, feats = _roipool_case()
    pool, pe, lvl = po.roi_pool(g["rois"], feats, g["pe_freq"], g["pe_time"])
    np.testing.assert_array_equal(lvl, g["lvl"])
    np.testing.assert_array_equal(pool, g["pool"])
    np.testing.assert_array_equal(pe, g["pe"])
