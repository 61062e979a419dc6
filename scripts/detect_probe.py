"""Where a detector batch spends its time on the GPU box: eager reference flow vs patched+accelerated eager vs CUDA-graph
replay.  Host wall per batch, device time per batch (CUDA events), kernel launches per batch (torch profiler).

    python scripts/detect_probe.py [n_batches]
"""
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
from birdsoundclassif_b200 import frontend, run_detection as rd, synth  # noqa: E402
from birdsoundclassif_b200.graphed import GraphedDetector  # noqa: E402

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 30
d = tempfile.mkdtemp()
synth.write_standin_checkpoint(d, seed=0, sharpen=400.0)
pcm = synth.synth_pcm(4 * n_batches * 2.46 + 3, 77, calls_per_s=6.0)
fp = frontend.File_Processor("x.wav")
tiles, _ = fp.process_pcm(torch.from_numpy(pcm).cuda())
tiles = tiles[:4 * n_batches]
print("tiles", tuple(tiles.shape))


def timed(model, label, n=n_batches):
    rd.detect_tiles(model, tiles[:8], 0.2, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record()
    rd.detect_tiles(model, tiles[:4 * n], 0.2, 4)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t) / n * 1e3
    print(f"{label:34s} wall {wall:7.2f} ms/batch   device span {e0.elapsed_time(e1) / n:7.2f} ms/batch")
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            rd.detect_tiles(model, tiles[:8], 0.2, 4)
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        busy = sum(e.device_time for e in ev) / 2 / 1e3
        print(f"{'':34s} {len(ev) // 2} device kernels/batch, device busy {busy:.2f} ms/batch")
        top = {}
        for e in ev:
            top[e.name] = top.get(e.name, 0) + e.device_time / 2 / 1e3
        for k, v in sorted(top.items(), key=lambda kv: -kv[1])[:8]:
            print(f"{'':38s}{v:6.2f} ms  {k[:90]}")
    except Exception as e:
        print("profiler failed:", e)


m_ref, _ = ref_shims.ref("nbm_model.run_detection").load_model(d)
timed(m_ref, "reference model, unpatched", n=min(n_batches, 6))
model, _ = rd.load_model(d)
rd.patch_reference()
rd.accelerate_model(model)
timed(model, "patched + accelerated, eager")
g1 = GraphedDetector(model, lanes=1)
timed(g1, "CUDA graphs, one lane")
g = GraphedDetector(model, lanes=2)
timed(g, "CUDA graphs, two lanes")
print("eager fallback:", g1._eager_only, g._eager_only)
if os.environ.get("PROBE_NOGC"):
    import gc
    gc.collect()
    gc.disable()
    timed(g, "two lanes, cyclic GC disabled")
    timed(model, "eager, cyclic GC disabled")
    gc.enable()
if os.environ.get("PROBE_LANES3"):
    timed(GraphedDetector(model, lanes=3), "CUDA graphs, three lanes")
if not os.environ.get("PROBE_CUDNN_BENCH"):
    sys.exit(0)
torch.backends.cudnn.benchmark = True
model2, _ = rd.load_model(d)
rd.accelerate_model(model2)
g2 = GraphedDetector(model2)
timed(g2, "graphs + cudnn.benchmark")
