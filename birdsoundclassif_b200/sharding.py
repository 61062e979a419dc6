"""File-level data parallelism: one process per GPU, files sharded, no data-path collective.

Every stage of the hot path is per file (normalisation min/max prepare_dataset.py:248-250,
batching run_detection.py:49-67, cross-window NMS :163-249, output file nbm_detect.py:27), and
the reference's nms/ProposalLayer are batch-coupled, so a file is never split across ranks.
The only communication is one all_gather of a small int64 vector of per-rank counts and timings
(NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

COUNT_FIELDS = ("files", "tiles", "detections", "frames", "t_front_us", "t_model_us", "t_post_us", "t_wall_us")


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_files(files, rank: int, world: int) -> list:
    """Deterministic static partition: rank r takes sorted(files)[r::world].  The union over ranks
    is the whole list and shards are disjoint."""
    return sorted(files)[rank::world]


def shard_by_duration(files_with_len, rank: int, world: int) -> list:
    """Longest-first greedy partition by duration (ties -> name), deterministic on every rank.
    `files_with_len`: iterable of (path, n_samples)."""
    loads = [0] * world
    mine = []
    for path, n in sorted(files_with_len, key=lambda t: (-t[1], t[0])):
        r = min(range(world), key=lambda i: (loads[i], i))
        loads[r] += n
        if r == rank:
            mine.append(path)
    return sorted(mine)


def gather_counts(counts: dict, device=None) -> list[dict]:
    """all_gather of the COUNT_FIELDS vector; returns one dict per rank (on every rank)."""
    vec = torch.tensor([int(counts.get(k, 0)) for k in COUNT_FIELDS], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [dict(zip(COUNT_FIELDS, vec.tolist()))]
    out = [torch.empty_like(vec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, vec)
    return [dict(zip(COUNT_FIELDS, o.tolist())) for o in out]


def totals(per_rank: list[dict]) -> dict:
    t = {k: sum(r[k] for r in per_rank) for k in COUNT_FIELDS if not k.startswith("t_")}
    t["t_wall_us_max"] = max(r["t_wall_us"] for r in per_rank)
    return t
