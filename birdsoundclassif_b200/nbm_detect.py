"""CLI with the reference's ``nbm_detect.py`` contract (nbm_model/nbm_detect.py:6-29):

    python -m birdsoundclassif_b200.nbm_detect --ckpt model_weights --audio_dir D [--min_score .2] [--batch 4]
    torchrun --nproc-per-node N -m birdsoundclassif_b200.nbm_detect ...     # files sharded over N GPUs

For every ``D/*.wav`` it writes ``str(output_dict)`` to the sibling ``.txt``.  The detector is the
reference's own model, loaded by the reference's ``load_model`` (``nbm_model`` must be importable,
e.g. PYTHONPATH=<reference checkout>); its hot-path symbols are patched to libnbm_b200."""
from __future__ import annotations

import argparse
import glob
import json
import os
import time

import torch

from . import run_detection as rd
from . import sharding


def _size(path: str) -> int:
    try:
        return os.path.getsize(path)
    except OSError:
        return 0


def detect_directory(model, model_args, audio_dir, bird_dict="bird_dict.json", min_score=0.2, bs=4,
                     rank=0, world=1, skip_done=False, verbose=True, pipelined=True, group_tiles=1024,
                     json_sidecar=False, balance="duration") -> dict:
    """This rank's share of ``audio_dir/*.wav`` -> sibling ``.txt`` files (nbm_detect.py:23-29).  ``pipelined``: wav
    decoding, the batched front-end and the detector overlap across files (pipeline.DetectionPipeline); otherwise the
    reference's one-file-at-a-time loop through ``run_detection``.  Same outputs either way.  ``json_sidecar``: also
    write ``<wav>.json`` (the same dictionary as JSON; the ``.txt`` is ``str(dict)``, readable only by ``ast.literal_eval``).
    ``balance``: how the directory is split over ``world`` ranks -- "duration" (longest file first onto the least loaded
    rank) or "name" (``sorted(files)[rank::world]``); the union over the ranks is the directory either way."""
    paths = glob.glob(os.path.join(audio_dir, "*.wav"))
    if world > 1 and balance == "duration":
        # longest first onto the least loaded rank, by file size (PCM: bytes ~ duration); every rank computes the same
        # partition from the same directory listing.  Equal-sized files fall back to the name-order round robin.
        files = sharding.shard_by_duration([(p, _size(p)) for p in paths], rank, world)
    else:
        files = sharding.shard_files(paths, rank, world)
    if skip_done:
        files = [f for f in files if not os.path.exists(f.replace(".wav", ".txt"))]
    counts = dict.fromkeys(sharding.COUNT_FIELDS, 0)
    t_wall = time.perf_counter()

    def done(wav_path, output):
        with open(wav_path.replace(".wav", ".txt"), "w") as f:
            f.write(str(output))
        if json_sidecar:
            with open(wav_path.replace(".wav", ".json"), "w") as f:
                json.dump(output, f)
        if verbose:
            print(f"~~~~~ File {os.path.basename(wav_path).replace('.wav', '')} done ~~~~~")

    if pipelined:
        from .pipeline import DetectionPipeline
        pipe = DetectionPipeline(model, model_args, bird_dict, min_score=min_score, bs=bs, max_group_tiles=group_tiles)
        for wav_path, output in pipe.run(files):
            done(wav_path, output)
        for wav_path, why in pipe.failed:       # per-file failure isolation (the reference crashes)
            print(f"[rank {rank}] {wav_path}: FAILED: {why}")
        counts.update(pipe.counts)
    else:
        for k, wav_path in enumerate(files):
            tm = {}
            try:
                output = rd.run_detection(model, model_args, wav_path, bird_dicts_path=bird_dict, min_score=min_score,
                                          bs=bs, timings=tm)
            except Exception as e:
                print(f"[rank {rank}] {wav_path}: FAILED: {e}")
                continue
            done(wav_path, output)
            counts["files"] += 1
            counts["tiles"] += tm["tiles"]; counts["frames"] += tm["frames"]; counts["detections"] += tm["detections"]
            counts["t_front_us"] += int(tm["frontend_s"] * 1e6); counts["t_model_us"] += int(tm["model_s"] * 1e6)
            counts["t_post_us"] += int(tm["post_s"] * 1e6)
    torch.cuda.synchronize()
    counts["t_wall_us"] = int((time.perf_counter() - t_wall) * 1e6)
    counts["failed"] = len(files) - counts["files"]
    if counts["failed"]:
        print(f"[rank {rank}] {counts['failed']} of {len(files)} files FAILED and have no .txt (see above)")
    return counts


def main(argv=None):
    parser = argparse.ArgumentParser("Bird call detection with NBM model (B200 hot path)")
    parser.add_argument("--ckpt", dest="model_dirp", type=str, default="model_weights")
    parser.add_argument("--audio_dir", dest="audio_dirp", type=str, required=True)
    parser.add_argument("--min_score", type=float, default=0.2)
    parser.add_argument("--batch", dest="bs", type=int, default=4)
    parser.add_argument("--bird_dict", type=str, default="bird_dict.json")
    parser.add_argument("--skip_done", action="store_true", help="skip wavs that already have a .txt")
    parser.add_argument("--json", action="store_true", help="also write <wav>.json next to the reference's <wav>.txt")
    parser.add_argument("--no_pipeline", action="store_true", help="one file at a time, as the reference loops")
    parser.add_argument("--no_graphs", action="store_true", help="run the detector eagerly instead of replaying CUDA graphs")
    parser.add_argument("--group_tiles", type=int, default=1024, help="detector tiles per front-end batch (pipelined)")
    parser.add_argument("--balance", choices=["duration", "name"], default="duration",
                        help="multi-GPU split of the directory: longest file first onto the least loaded rank, or name order round robin")
    args = parser.parse_args(argv)
    assert os.path.isfile(args.bird_dict), "Missing dictionary of bird species names --> bird_dict.json."

    rank, world, local = sharding.env_rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        model, model_args = rd.load_model(args.model_dirp)      # the reference's network classes, unchanged
    except ImportError as e:
        raise SystemExit(str(e))
    rd.patch_reference()
    rd.accelerate_model(model)
    if not args.no_graphs:
        from .graphed import GraphedDetector
        model = GraphedDetector(model)
    counts = detect_directory(model, model_args, args.audio_dirp, args.bird_dict, args.min_score, args.bs,
                              rank, world, args.skip_done, pipelined=not args.no_pipeline, group_tiles=args.group_tiles,
                              json_sidecar=args.json, balance=args.balance)
    per_rank = sharding.gather_counts(counts, device=torch.device("cuda", local))
    if rank == 0:
        print(json.dumps({"per_rank": per_rank, "totals": sharding.totals(per_rank)}))
    if world > 1:
        torch.distributed.destroy_process_group()
    if counts.get("failed"):
        raise SystemExit(3)         # some of this rank's files have no output


if __name__ == "__main__":
    main()
