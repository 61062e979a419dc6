"""nbm_detect-shaped throughput (wav files on disk -> .txt files): the pipelined multi-file driver against the
reference-shaped one-file-at-a-time loop, same library kernels, same stand-in detector (the reference CNN does not
travel to the GPU box).  BASELINE configs[0] scaled up: N synthetic 30 s mono wavs.

    python scripts/bench_pipeline.py [--files 64] [--seconds 30] [--bs 4]
"""
import argparse, glob, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from birdsoundclassif_b200 import nbm_detect, synth
from tests.standin_detector import StandInDetector

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=64)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--bs", type=int, default=4)
ap.add_argument("--group_tiles", type=int, default=1024)
a = ap.parse_args()
with tempfile.TemporaryDirectory() as d:
    for i in range(a.files):
        synth.write_wav(os.path.join(d, f"rec_{i:04d}.wav"), synth.synth_pcm(a.seconds, 1000 + i))
    bird = os.path.join(d, "bird_dict.json")
    json.dump({f"Species {i}": i for i in range(1, 151)}, open(bird, "w"))
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()
    hours = a.files * a.seconds / 3600
    res = {}
    for name, pipelined in (("file_by_file", False), ("pipelined", True), ("file_by_file", False), ("pipelined", True)):
        for f in glob.glob(os.path.join(d, "*.txt")):
            os.remove(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = nbm_detect.detect_directory(model, args, d, bird, 0.2, a.bs, verbose=False, pipelined=pipelined,
                                        group_tiles=a.group_tiles)
        dt = time.perf_counter() - t0
        res[name] = dict(wall_s=round(dt, 3), audio_h_per_s=round(hours / dt, 2), files=c["files"], tiles=c["tiles"],
                         detections=c["detections"], t_front_ms=c["t_front_us"] / 1e3, t_model_ms=c["t_model_us"] / 1e3,
                         t_post_ms=c["t_post_us"] / 1e3)
    print(json.dumps({"files": a.files, "seconds": a.seconds, "bs": a.bs, "detector": "stand-in (tests/standin_detector.py)",
                      **res}))
