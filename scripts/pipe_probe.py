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
print("pipe one group", pipe(readers=1, max_group_tiles=100000, first_group_tiles=None))
from concurrent.futures import ThreadPoolExecutor
pool = ThreadPoolExecutor(4); list(pool.map(lambda i: time.sleep(0.05), range(4)))
print("seq with 4 idle pool threads", seq())
pool.shutdown()
pin = [torch.empty(320_000_000, dtype=torch.int16).pin_memory() for _ in range(3)]
print("seq with 1.9 GB pinned held", seq())
del pin
big = torch.empty((2940 * 2, 1, 375, 1024), device="cuda")
print("seq with 9 GB device buffer held", seq())
del big
# the pipeline's detection phase alone, on tiles produced up front (no reader, no second stream)
from birdsoundclassif_b200.frontend import File_Processor
from birdsoundclassif_b200 import postproc
from types import SimpleNamespace
fps = []
for f in files:
    fp = File_Processor(f); tiles, _ = fp.process_file(); fps.append((fp, tiles))
torch.cuda.synchronize()
rev = {i: f"S{i}" for i in range(151)}
t = time.perf_counter(); tm = 0.0
for fp, tiles in fps:
    t0 = time.perf_counter()
    outs = rd.detect_tiles(model, tiles, 0.2, 4)
    tm += time.perf_counter() - t0
    postproc.merge_to_output(fp, outs, 150, rev)
print("detector + merge over resident tiles", time.perf_counter() - t, "model", tm)
