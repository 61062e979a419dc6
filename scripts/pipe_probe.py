"""Diagnostic: where the pipelined driver's detector loop loses time against the file-by-file loop (stand-in detector)."""
import glob, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from birdsoundclassif_b200 import synth
from birdsoundclassif_b200.pipeline import DetectionPipeline
from birdsoundclassif_b200 import run_detection as rd
from tests.standin_detector import StandInDetector
d = tempfile.mkdtemp()
pcm = synth.synth_pcm(600.0, 1)
for i in range(12):
    synth.write_wav(os.path.join(d, f"r{i:02d}.wav"), pcm)
bird = os.path.join(d, "b.json"); json.dump({f"S{i}": i for i in range(1, 151)}, open(bird, "w"))
args = synth.default_args("cuda"); model = StandInDetector(args, backend="nbm").cuda()
files = sorted(glob.glob(os.path.join(d, "*.wav")))
def seq():
    t = time.perf_counter(); tm_all = dict(frontend_s=0, model_s=0, post_s=0)
    for f in files:
        tm = {}; rd.run_detection(model, args, f, bird, 0.2, 4, timings=tm)
        for k in tm_all: tm_all[k] += tm[k]
    return time.perf_counter() - t, tm_all
def pipe(**kw):
    p = DetectionPipeline(model, args, bird, 0.2, 4, **kw)
    t = time.perf_counter()
    for _ in p.run(files): pass
    return time.perf_counter() - t, {k: v / 1e6 for k, v in p.counts.items() if k.startswith("t_")}
seq()
print("seq", seq())
for kw in (dict(), dict(readers=1), dict(max_group_tiles=245, first_group_tiles=None), dict(max_group_tiles=4096), dict(readers=1, max_group_tiles=100000, first_group_tiles=None)):
    print("pipe", kw, pipe(**kw))
