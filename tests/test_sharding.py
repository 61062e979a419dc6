"""Host logic of the multi-GPU path on CPU: partition properties and the count gather over a
world_size-2 gloo group."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings, strategies as st

from birdsoundclassif_b200 import sharding


@settings(max_examples=50, deadline=None)
@given(st.lists(st.text(alphabet="abcdef0123", min_size=1, max_size=6), unique=True, max_size=40), st.integers(1, 8))
def test_shards_partition_the_file_list(files, world):
    shards = [sharding.shard_files(files, r, world) for r in range(world)]
    assert sorted(sum(shards, [])) == sorted(files)
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.text(alphabet="abcdef0123", min_size=1, max_size=6), st.integers(0, 10 ** 8)),
                unique_by=lambda t: t[0], max_size=30), st.integers(1, 8))
def test_duration_balanced_shards_partition(files, world):
    shards = [sharding.shard_by_duration(files, r, world) for r in range(world)]
    assert sorted(sum(shards, [])) == sorted(f for f, _ in files)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = [f"f{i:03d}.wav" for i in range(11)]
    mine = sharding.shard_files(files, rank, world)
    counts = dict(files=len(mine), tiles=10 * len(mine), detections=rank + 1, frames=1000 * len(mine),
                  t_front_us=5, t_model_us=6, t_post_us=7, t_wall_us=100 + rank)
    per_rank = sharding.gather_counts(counts)
    q.put((rank, per_rank, sharding.totals(per_rank)))
    dist.destroy_process_group()


def test_count_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, per_rank, tot in got:
        assert [r["files"] for r in per_rank] == [6, 5]
        assert tot["files"] == 11 and tot["tiles"] == 110 and tot["detections"] == 3 and tot["t_wall_us_max"] == 101


def test_single_process_gather_is_identity():
    out = sharding.gather_counts(dict(files=3, tiles=7))
    assert out[0]["files"] == 3 and out[0]["tiles"] == 7 and out[0]["frames"] == 0


def test_duration_balance_quality_and_equal_sizes():
    # equal sizes: exactly the name-order round robin (the bench's 16 x 30 s directory keeps its partition)
    files = [(f"rec_{i:04d}.wav", 2646044) for i in range(16)]
    for world in (1, 2, 4, 8):
        for r in range(world):
            assert sharding.shard_by_duration(files, r, world) == sharding.shard_files([f for f, _ in files], r, world)
    # a night of mixed lengths: one long recording does not pile up with others on the same rank
    mixed = [("long.wav", 3000)] + [(f"s{i:02d}.wav", 100) for i in range(30)]
    loads = [sum(dict(mixed)[p] for p in sharding.shard_by_duration(mixed, r, 4)) for r in range(4)]
    assert max(loads) == 3000 and sorted(loads)[:3] == [1000, 1000, 1000]
    rr = [sum(dict(mixed)[p] for p in sharding.shard_files([f for f, _ in mixed], r, 4)) for r in range(4)]
    assert max(rr) > max(loads)


def test_detect_directory_partitions_by_duration(tmp_path, monkeypatch):
    """nbm_detect.detect_directory's file selection (no GPU: the loop body is stubbed): the ranks' shares are disjoint,
    cover the directory, and follow the duration balance unless balance='name'."""
    from birdsoundclassif_b200 import nbm_detect
    sizes = {"a.wav": 5000, "b.wav": 100, "c.wav": 100, "d.wav": 100, "e.wav": 4000}
    for name, n in sizes.items():
        (tmp_path / name).write_bytes(b"\\0" * n)
    seen = {}

    def fake_run_detection(model, args, wav_path, **kw):
        seen.setdefault(kw["_rank"], []).append(os.path.basename(wav_path))
        raise RuntimeError("stub")                       # counted as failed, no output written

    for balance, want in (("duration", [["a.wav"], ["b.wav", "c.wav", "d.wav", "e.wav"]]),
                          ("name", [["a.wav", "c.wav", "e.wav"], ["b.wav", "d.wav"]])):
        seen.clear()
        for rank in range(2):
            monkeypatch.setattr(nbm_detect.rd, "run_detection",
                                lambda *a, _r=rank, **kw: fake_run_detection(*a, _rank=_r, **kw))
            monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
            c = nbm_detect.detect_directory(None, None, str(tmp_path), rank=rank, world=2, verbose=False,
                                            pipelined=False, balance=balance)
            assert c["files"] == 0 and c["failed"] == len(want[rank])
        assert [sorted(seen[r]) for r in range(2)] == want
