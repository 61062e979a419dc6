"""Diagnostic: PCIe copy rate alone vs the pipelined pinned-host entry at several chunk sizes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from birdsoundclassif_b200 import frontend
n, clips = 2646000, 1024
plan = frontend.get_plan()
host = torch.empty(clips * n, dtype=torch.int16).pin_memory()
host.view(torch.int32)[:] = torch.randint(-2000, 2000, (clips * n // 2,), dtype=torch.int32)
dev = torch.empty_like(host, device='cuda')
offs = [i * n for i in range(clips + 1)]
_, tile_off, _ = plan.query_batch([n] * clips)
tiles = torch.empty((tile_off[-1], 1, 375, 1024), dtype=torch.float32, device='cuda')
def timeit(fn, k=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms = timeit(lambda: dev.copy_(host, non_blocking=True))
print("plain H2D %.1f ms  %.1f GB/s" % (ms, host.numel() * 2 / ms / 1e6))
ms = timeit(lambda: plan.run_batch(dev, offs, out=tiles))
print("device-resident run_batch %.1f ms" % ms)
for fpc in (16, 32, 64, 128, 256, 512):
    ms = timeit(lambda: plan.run_batch_from_host(host, offs, out=tiles, files_per_chunk=fpc))
    print("from_host chunk %4d files: %.1f ms  -> %.1f audio-h/s" % (fpc, ms, clips * 60 / 3600 / ms * 1e3))
