"""`detect` leg of bench.py: audio-hours/s THROUGH ``nbm_detect`` (BASELINE.json metric, first half).

wav files on disk -> ``birdsoundclassif_b200.nbm_detect.detect_directory`` (reader threads, batched front-end,
the reference's own CNN with the library-backed post-processing, per-file merge) -> ``.txt`` files, wall clock
from the first file read to the last ``.txt`` written, files sharded over the ranks, max over ranks.

  * ``cfg0``  BASELINE configs[0]: 16 synthetic 30 s mono wavs, ONE directory sharded over all ranks (strong).
  * ``night`` BASELINE configs[2]/[3] shape: ten-minute wavs (an 8-hour night is 48 of them; configs[3] shards
              10 000 of them), ``--detect-files`` per GPU (weak: every rank brings its own files).
  * ``reference`` (N = 1 only): the reference's UNPATCHED flow on the same GPU and the same cfg0 files -- its
              ``run_detection`` (CPU front-end through the librosa restatement of oracle/ref_shims.py, its own
              ProposalLayer / ROIPooling / FastRCNN tail / merge_images Python loops) -- as the baseline.

The detector network is the reference's (``nbm_model.nets``), found at $NBM_REFERENCE_ROOT, /root/reference or
the oracle/_ref copy (oracle/build_ref.py), with a seeded stand-in checkpoint (the shipped one is a Git-LFS
stub): ``synth.write_standin_checkpoint(seed=0, sharpen=400)``.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def find_reference_root():
    for cand in (os.environ.get("NBM_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "nbm_model", "nets", "nbm_model.py")):
            return cand
    return None


def _write_files(dirpath, n_files, seconds, seed0, distinct=4):
    """n_files wavs of `seconds` s; `distinct` different synthetic recordings, the rest circular shifts of them."""
    from birdsoundclassif_b200 import synth
    os.makedirs(dirpath, exist_ok=True)
    base = [synth.synth_pcm(seconds, seed0 + i) for i in range(min(distinct, n_files))]
    paths = []
    for i in range(n_files):
        pcm = base[i % len(base)]
        if i >= len(base):
            pcm = np.roll(pcm, (i // len(base)) * 44100 * 7)
        paths.append(synth.write_wav(os.path.join(dirpath, f"rec_{i:04d}.wav"), pcm))
    return paths


def _rm_outputs(dirpath):
    for f in os.listdir(dirpath):
        if f.endswith(".txt"):
            try:
                os.remove(os.path.join(dirpath, f))
            except FileNotFoundError:
                pass


def run(a, rank, world, local, dist):
    """Returns the `detect` dict on rank 0 (None elsewhere).  Exceptions never escape between two collectives: a rank that
    fails keeps taking part in them and every rank learns about the failure at the next `all_ok` (a rank that ran away
    would leave the others in a barrier until the NCCL watchdog fires)."""
    import traceback
    import torch
    ref_root = find_reference_root()
    if ref_root is None:
        return {"unavailable": "no reference checkout (nbm_model.nets) on this machine"} if rank == 0 else None
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from birdsoundclassif_b200 import nbm_detect, run_detection as rd, sharding, synth
    dev = torch.device("cuda", local)
    errors = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_ok(ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def guarded(fn, *args, **kw):
        try:
            return True, fn(*args, **kw)
        except Exception as e:
            errors.append(f"rank {rank}: {type(e).__name__}: {e} | {traceback.format_exc()[-600:]}")
            return False, None

    def fail(what):
        return {"error": what, "details": errors[:2]} if rank == 0 else None

    base = os.environ.get("NBM_BENCH_TMP") or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    work = os.path.join(base, f"nbm_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}")
    bird_dict = os.path.join(ref_root, "bird_dict.json")
    ckpt = os.path.join(work, "model_weights")
    cfg0_dir = os.path.join(work, "cfg0")
    night_dir = os.path.join(work, f"night_r{rank}")
    out = {}
    state = {}

    def setup_shared():
        os.makedirs(work, exist_ok=True)
        synth.write_standin_checkpoint(ckpt, seed=0, sharpen=400.0)
        _write_files(cfg0_dir, 16, 30.0, 1000 * 0)

    def setup_rank():
        from birdsoundclassif_b200.graphed import GraphedDetector
        _write_files(night_dir, a.detect_files, 600.0, 1000 * 2 + 10 * rank, distinct=2)
        model, margs = rd.load_model(ckpt)
        rd.patch_reference()
        rd.accelerate_model(model)
        state.update(model=model, margs=margs, graphed=GraphedDetector(model))

    def leg(dirpath, r, w, repeats, det=None, pipelined=True):
        """Best of `repeats` passes over the directory: (wall = max over ranks, per-rank counts) or None if a rank failed."""
        best = None
        for _ in range(repeats):
            barrier()                                      # every rank has finished the previous pass over this directory
            if w == 1 or rank == 0:                        # a shared directory is cleaned by one rank only
                _rm_outputs(dirpath)
            barrier()
            ok, c = guarded(nbm_detect.detect_directory, det or state["graphed"], state["margs"], dirpath, bird_dict,
                            min_score=0.2, bs=4, rank=r, world=w, verbose=False, pipelined=pipelined)
            if ok and c.get("failed"):
                errors.append(f"rank {rank}: {c['failed']} files failed in {dirpath}")
                ok = False
            per_rank = sharding.gather_counts(c if ok else {}, device=dev)
            if not all_ok(ok):
                return None
            wall = max(x["t_wall_us"] for x in per_rank) / 1e6
            if best is None or wall < best[0]:
                best = (wall, per_rank)
        return best

    def stages(per_rank):
        return {k: sum(x[k] for x in per_rank) / 1e6 for k in ("t_front_us", "t_model_us", "t_post_us")}

    try:
        ok = guarded(setup_shared)[0] if rank == 0 else True
        if not all_ok(ok):
            return fail("setting up the stand-in checkpoint / cfg0 files failed")
        if not all_ok(guarded(setup_rank)[0]):
            return fail("loading the detector failed")
        model = state.get("model")
        # warm-up: cuDNN heuristics, the allocator's pools, the front-end plan, graph capture
        if leg(cfg0_dir, rank, world, 1) is None or leg(cfg0_dir, rank, world, 1, det=model) is None:
            return fail("detect_directory failed on a rank")
        hours = 16 * 30.0 / 3600.0
        legs = [("cfg0", None, True), ("cfg0_eager", model, True)]
        if world == 1:
            legs.append(("cfg0_file_by_file", None, False))
        for name, det, pipelined in legs:
            res = leg(cfg0_dir, rank, world, 2, det=det, pipelined=pipelined)
            if res is None:
                return fail("detect_directory failed on a rank")
            wall, per_rank = res
            tot = sharding.totals(per_rank)
            out[name] = {"workload": "BASELINE configs[0]: 16 x 30 s wavs -> .txt, one directory sharded over the ranks, "
                                     "min_score 0.2, bs 4, reference CNN + stand-in checkpoint, detector forward "
                                     + ("launched eagerly" if det is not None else "replayed from CUDA graphs")
                                     + ("" if pipelined else "; one run_detection call per file, as the reference loops "
                                        "(no reader threads, no batched front-end, replay lanes drained at every file)"),
                         "audio_hours_per_s": hours / wall, "wall_s": wall, "files": tot["files"], "tiles": tot["tiles"],
                         "detections": tot["detections"], "scaling": "strong", "stage_s_sum_over_ranks": stages(per_rank)}
        # two passes, the faster one counts: the first also records a CUDA graph of the second stage for every RoI count M it
        # meets (tens of milliseconds each, once per process) -- in the regime this slice stands for (thousands of files per
        # GPU) those are long amortised
        res = leg(night_dir, 0, 1, 2)
        if res is None:
            return fail("detect_directory failed on a rank (night slice)")
        wall, per_rank = res
        tot = sharding.totals(per_rank)
        hours = tot["files"] * 600.0 / 3600.0
        out["night"] = {"workload": f"BASELINE configs[2]/[3] shape: {a.detect_files} x 10-min wavs per GPU -> .txt, "
                                    "min_score 0.2, bs 4, reference CNN + stand-in checkpoint, CUDA graphs",
                        "audio_hours_per_s": hours / wall, "wall_s": wall, "files": tot["files"], "tiles": tot["tiles"],
                        "detections": tot["detections"], "scaling": "weak", "stage_s_sum_over_ranks": stages(per_rank),
                        "per_rank_wall_s": [x["t_wall_us"] / 1e6 for x in per_rank]}
        rd.unpatch_reference()
        if world == 1 and not a.no_detect_reference:
            ok, ref = guarded(_reference_flow, ckpt, cfg0_dir, bird_dict, a.detect_ref_files)
            out["reference"] = ref if ok else {"error": errors[-1]}
        barrier()                                          # nobody reads the shared directory any more
        if rank == 0:
            shutil.rmtree(work, ignore_errors=True)
    finally:
        shutil.rmtree(night_dir, ignore_errors=True)       # no collective here
    return out if rank == 0 else None


def _reference_flow(ckpt, cfg0_dir, bird_dict, n_files):
    """The reference's unpatched nbm_detect loop (nbm_detect.py:23-29) on this GPU, first `n_files` of cfg0."""
    import glob
    import torch
    from oracle import ref_shims            # the librosa / matplotlib stand-ins the reference's imports need here
    ref_shims.install()
    ref_rd = ref_shims.ref("nbm_model.run_detection")
    model, margs = ref_rd.load_model(ckpt)
    files = sorted(glob.glob(os.path.join(cfg0_dir, "*.wav")))[:n_files]
    ref_rd.run_detection(model, margs, files[0], bird_dicts_path=bird_dict, min_score=0.2, bs=4)     # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    n_det = 0
    for w in files:
        o = ref_rd.run_detection(model, margs, w, bird_dicts_path=bird_dict, min_score=0.2, bs=4)
        with open(w.replace(".wav", ".ref.txt"), "w") as f:
            f.write(str(o))
        n_det += sum(len(v["scores"]) for v in o.values())
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    return {"workload": f"the reference's own run_detection loop, unpatched, same GPU, {len(files)} of the cfg0 files "
                        "(CPU front-end = librosa restatement, 1 host thread, as upstream)",
            "audio_hours_per_s": len(files) * 30.0 / 3600.0 / wall, "wall_s": wall, "files": len(files), "detections": n_det}
