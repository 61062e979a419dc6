"""Diagnostic: per-chunk timeline of run_batch_from_host (copy vs compute), re-implemented with timing events."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from birdsoundclassif_b200 import frontend
n, clips, fpc = 2646000, 512, 64
plan = frontend.get_plan()
host = torch.empty(clips * n, dtype=torch.int16).pin_memory()
host.view(torch.int32)[:] = torch.randint(-2000, 2000, (clips * n // 2,), dtype=torch.int32)
offs = [i * n for i in range(clips + 1)]
_, tile_off, _ = plan.query_batch([n] * clips)
tiles = torch.empty((tile_off[-1], 1, 375, 1024), dtype=torch.float32, device='cuda')
mm = torch.empty((clips, 2), dtype=torch.float32, device='cuda')
buf = [torch.empty(fpc * n, dtype=torch.int16, device='cuda') for _ in range(2)]
cs = torch.cuda.Stream()
cur = torch.cuda.current_stream()
chunks = [(f, f + fpc) for f in range(0, clips, fpc)]
def run(record):
    ev = []
    ready = [torch.cuda.Event() for _ in range(2)]; done = [torch.cuda.Event() for _ in range(2)]
    t0 = torch.cuda.Event(enable_timing=True); t0.record(cur)
    cs.wait_stream(cur)
    for ci, (fa, fb) in enumerate(chunks):
        b = ci & 1
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(cs):
            if ci >= 2: cs.wait_event(done[b])
            e[0].record(cs)
            buf[b].copy_(host[offs[fa]:offs[fb]], non_blocking=True)
            e[1].record(cs); ready[b].record(cs)
        cur.wait_event(ready[b])
        e[2].record(cur)
        plan.run_batch(buf[b], [o - offs[fa] for o in offs[fa:fb + 1]], out=tiles[tile_off[fa]:tile_off[fb]], minmax_out=mm[fa:fb])
        e[3].record(cur); done[b].record(cur)
        ev.append(e)
    torch.cuda.synchronize()
    if record:
        for ci, e in enumerate(ev):
            print("chunk %d: copy %.1f..%.1f  compute %.1f..%.1f" % (ci, t0.elapsed_time(e[0]), t0.elapsed_time(e[1]), t0.elapsed_time(e[2]), t0.elapsed_time(e[3])))
run(False); run(True)
