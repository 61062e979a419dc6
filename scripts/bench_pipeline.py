"""nbm_detect-shaped throughput (wav files on disk -> .txt files): the pipelined multi-file driver against the
reference-shaped one-file-at-a-time loop, same library kernels, same stand-in detector (the reference CNN does not
travel to the GPU box).  BASELINE configs[0] scaled up: N synthetic 30 s mono wavs.

    python scripts/bench_pipeline.py [--files 64] [--seconds 30] [--bs 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_pipeline.py ...

Under torchrun the files are sharded over the ranks (sharding.shard_files), every rank runs its share, the per-rank count /
timing vectors are all-gathered over NCCL (the path's only collective) and rank 0 reports totals over the slowest rank's wall time.
"""
import argparse, glob, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from birdsoundclassif_b200 import nbm_detect, sharding, synth
from tests.standin_detector import StandInDetector

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=64)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--bs", type=int, default=4)
ap.add_argument("--group_tiles", type=int, default=1024)
a = ap.parse_args()
rank, world, local = sharding.env_rank_world()
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
shared = [tempfile.mkdtemp(prefix="nbm_bench_") if rank == 0 else None]
if world > 1:
    torch.distributed.broadcast_object_list(shared, src=0)
d = shared[0]
try:
    bird = os.path.join(d, "bird_dict.json")
    if rank == 0:
        for i in range(a.files):
            synth.write_wav(os.path.join(d, f"rec_{i:04d}.wav"), synth.synth_pcm(a.seconds, 1000 + i))
        json.dump({f"Species {i}": i for i in range(1, 151)}, open(bird, "w"))
    if world > 1:
        torch.distributed.barrier()
    args = synth.default_args("cuda")
    model = StandInDetector(args, backend="nbm").cuda()
    hours = a.files * a.seconds / 3600
    res = {}
    for name, pipelined in (("file_by_file", False), ("pipelined", True), ("file_by_file", False), ("pipelined", True)):
        if rank == 0:
            for f in glob.glob(os.path.join(d, "*.txt")):
                os.remove(f)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        c = nbm_detect.detect_directory(model, args, d, bird, 0.2, a.bs, rank, world, verbose=False, pipelined=pipelined,
                                        group_tiles=a.group_tiles)
        per_rank = sharding.gather_counts(c, device=torch.device("cuda", local))
        tot = sharding.totals(per_rank)
        dt = tot["t_wall_us_max"] / 1e6
        res[name] = dict(wall_s=round(dt, 3), audio_h_per_s=round(hours / dt, 2), files=tot["files"], tiles=tot["tiles"],
                         detections=tot["detections"], per_rank_files=[r["files"] for r in per_rank],
                         t_front_ms=sum(r["t_front_us"] for r in per_rank) / 1e3,
                         t_model_ms=sum(r["t_model_us"] for r in per_rank) / 1e3,
                         t_post_ms=sum(r["t_post_us"] for r in per_rank) / 1e3)
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        n_txt = len(glob.glob(os.path.join(d, "*.txt")))
        print(json.dumps({"files": a.files, "seconds": a.seconds, "bs": a.bs, "n_gpus": world, "txt_written": n_txt,
                          "detector": "stand-in (tests/standin_detector.py)", **res}))
finally:
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank == 0:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
