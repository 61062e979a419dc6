"""Latency of the detector post-processing kernels (decode, greedy NMS, fused ProposalLayer) on the GPU, next to
the CPU oracle port of the reference algorithm on the same inputs (checker timed as a reported baseline only).
These kernels are latency-bound (N <= 500 per image in the detector), so the unit is microseconds per call.

    python scripts/bench_postproc.py            # prints one JSON object
"""
import json, os, sys, time, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from birdsoundclassif_b200 import postproc as pp, synth
from oracle import postproc_oracle as po


def gpu_us(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def cpu_us(fn, iters=3):
    fn()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t) * 1e6 / iters


def boxes(rng, B, N, span):
    x1 = rng.integers(0, span, (B, N)); y1 = rng.integers(0, 375, (B, N))
    return np.stack([x1, y1, x1 + rng.integers(5, 90, (B, N)), y1 + rng.integers(5, 60, (B, N))], -1).astype(np.float32)


def main():
    rng = np.random.default_rng(7)
    out = {}
    for name, B, N, th, span in (("nms_rpn_bs4_n500_t0.7", 4, 500, 0.7, 900), ("nms_final_bs4_n50_t0.3", 4, 50, 0.3, 900),
                                 ("nms_merge_n5000_t0.3", 1, 5000, 0.3, 200000), ("nms_stress_n20000_t0.3", 1, 20000, 0.3, 200000)):
        b = boxes(rng, B, N, span); s = rng.random((B, N)).astype(np.float32)
        db, ds = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
        g = gpu_us(lambda: pp.nms_keep(db, th))
        c = cpu_us(lambda: [po.greedy(b[i], th) for i in range(B)], iters=1 if N > 1000 else 3)
        k_g = pp.nms(db, ds, th, N, True)[2]; k_c = po.nms(b, s, th, N, True)[2]
        out[name] = {"gpu_us": round(g, 1), "cpu_oracle_us": round(c, 1), "keep_lists_identical": k_g == k_c}
    # anchor decode, RPN size
    A = po.make_anchors()
    d = (rng.standard_normal((4, A.shape[0], 4)) * 0.3).astype(np.float32)
    dd, da = torch.from_numpy(d).cuda(), torch.from_numpy(A.astype(np.float32)).cuda()
    out["decode_bs4_n23040"] = {"gpu_us": round(gpu_us(lambda: pp.bbox_reg_to_coord(dd, da)), 1),
                                "cpu_oracle_us": round(cpu_us(lambda: po.decode(d, A)), 1)}
    # fused ProposalLayer (decode 23 040 anchors/img, clamp, min-size, stable sort, top-500, NMS .7, keep 50)
    cfg = types.SimpleNamespace(**synth.DEFAULT_ARGS)
    cfg.ratios = [0.5, 1, 2]; cfg.device = "cuda"
    layer = pp.ProposalLayer(cfg, 5).eval()
    logits = rng.standard_normal((4, 15, 2, 24, 64)).astype(np.float32)
    e = np.exp(logits - logits.max(2, keepdims=True)); prob = (e / e.sum(2, keepdims=True))
    cls = np.ascontiguousarray(np.transpose(prob, (0, 1, 2, 3, 4)).reshape(4, 30, 24, 64)).astype(np.float32)
    reg = (rng.standard_normal((4, 60, 24, 64)) * 0.2).astype(np.float32)
    dc, dr = torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda()
    with torch.no_grad():
        out["proposal_layer_bs4"] = {"gpu_us": round(gpu_us(lambda: layer(dc, dr), iters=20), 1),
                                     "cpu_oracle_us": round(cpu_us(lambda: po.proposal_layer(cls, reg), iters=2), 1)}
    # second-stage RoI pooling + positional encoding: bs 4 x 50 RoIs, 256 channels, 2 x 2 bins
    cfg.n_layers, cfg.out_fpn_chan = 5, 256
    feats = synth.fpn_features(5, 4, 256, 5)
    x1 = rng.integers(0, 900, (4, 50)); y1 = rng.integers(0, 330, (4, 50))
    rois = np.stack([x1, y1, np.minimum(x1 + rng.integers(5, 300, (4, 50)), 1023),
                     np.minimum(y1 + rng.integers(5, 120, (4, 50)), 374)], -1).astype(np.float32)
    rp = pp.ROIPooling(cfg)
    dro, dfe = torch.from_numpy(rois).cuda(), [torch.from_numpy(f).cuda() for f in feats]
    small = [f[:, :16] for f in feats]
    pe_f, pe_t = po.positional_encoding_1d(375, 8), po.positional_encoding_1d(1024, 8)
    c16 = cpu_us(lambda: po.roi_pool(rois, small, pe_f, pe_t), iters=1)
    out["roi_pooling_bs4_r50_c256"] = {"gpu_us": round(gpu_us(lambda: rp(dro, dfe), iters=20), 1),
                                       "cpu_oracle_us": round(c16 * 16, 1), "cpu_note": "oracle timed on 16 of 256 channels, x16"}
    print(json.dumps({"postproc_latency": out, "cpu_threads": 1,
                      "note": "GPU: CUDA events over 20-30 calls through the Python mirror (includes its host overhead); "
                              "CPU: oracle port of the reference algorithm, single thread"}))


if __name__ == "__main__":
    main()
