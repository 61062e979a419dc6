"""Host-side ceiling of the end-to-end leg: N processes (one per GPU) copy pinned host memory to their device at the same
time, no kernels.  What bench.py's e2e reaches at N GPUs is a fraction of this, not of N x the single-GPU PCIe rate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/h2d_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
out = {}
for bind in (False, True):
    node = bench.bind_near_gpu(local) if bind else None
    n = 2_709_504_000                          # int16 values = bench.py's 1024 x 60 s clips (5.4 GB)
    host = torch.empty(n, dtype=torch.int16).pin_memory()
    host.zero_()
    d = torch.empty(n, dtype=torch.int16, device=dev)
    for _ in range(2):
        d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([5 * n * 2 / (e0.elapsed_time(e1) / 1e3) / 1e9], dtype=torch.float64, device=dev)
    if world > 1:
        allg = [torch.empty_like(gbs) for _ in range(world)]
        dist.all_gather(allg, gbs)
        vals = [float(v) for v in allg]
    else:
        vals = [float(gbs)]
    out["numa_bound" if bind else "unbound"] = {"per_gpu_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(sum(vals), 1),
                                                 "numa_node_rank0": node}
    del host, d
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_pinned_concurrent": out, "cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
