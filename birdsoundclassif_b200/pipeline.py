"""Pipelined multi-file host driver around the reference's per-file loop.

The reference handles a directory one file at a time (nbm_detect.py:23-29 -> run_detection.py:40-67):
decode the wav, transform it, run the detector batch by batch, merge, write -- every stage waiting
for the one before.  Here the same per-file results come out of three overlapped stages:

  reader threads   wav -> int16 straight into a slice of a PINNED group buffer (no float conversion,
                   no per-file allocation); groups of whole files, sized by detector tiles
  front-end stream one H2D copy + ONE batched front-end call per group (nbm_frontend_run_batch: the
                   kernels are sized for thousands of tiles, a single 30 s file leaves them
                   launch-bound), into one of two device tile buffers
  caller's stream  the detector over the previous group's tiles in the reference's batching (bs, order,
                   partial last batch PER FILE: nms / ProposalLayer are batch-coupled),
                   the per-file merge, the output dictionary

Group g+1 is read and transformed while group g is in the detector; events hand the tile buffers back
and forth.  A file is never split across groups (its normalisation is file-global), so every tile, and
therefore every box, equals what ``run_detection.run_detection`` returns for that file alone (tested).

The detector itself is whatever ``model`` is: the reference's NbmModel, eager, or ``graphed.GraphedDetector``
(its forward replayed from CUDA graphs on two lanes); a group's files go to it as ONE stream
(``run_detection.detect_stream``), so a graphed detector stays busy across file boundaries.

Start-up matters for short directories (BASELINE configs[0] is 16 files, 0.6 s): the pipeline's private front-end
plan (twiddle tables, workspace) and its pinned host buffers are taken from process-wide pools and given back
at the end of ``run`` (the device buffers come from torch's caching allocator), so a second directory in the same process does not pay for them again.
"""
from __future__ import annotations

import gc
import json
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np
import torch

from . import audio_io, postproc
from .frontend import LONG_FILE_SAMPLES, STFT_CHUNK, FrontendPlan, derive_constants
from .run_detection import detect_stream, run_detection


# ------------------------------------------------------------------ pure host arithmetic ------
def count_frames_tiles(n_samples: int, hop_length: int, w_pix: int, hop_spectro: int, stft_chunk: int = STFT_CHUNK):
    """(frames, tiles) of a file: prepare_dataset.py:234-237 (one STFT per <= stft_chunk samples, each
    1 + len // hop frames, `range(int(len / max_l) + 1)` chunks) and :255-266 (tiles)."""
    frames = 0
    for c in range(n_samples // stft_chunk + 1):
        length = max(0, min(n_samples, (c + 1) * stft_chunk) - c * stft_chunk)
        frames += 1 + length // hop_length
    tiles = 1 if frames <= w_pix else 1 + -(-(frames - w_pix) // hop_spectro)
    return frames, tiles


@dataclass
class WavInfo:
    path: str
    n_samples: int = 0          # per channel
    channels: int = 1
    sample_rate: int = 0
    error: str | None = None
    pcm16: bool = True          # False: another PCM width or float samples (decoded on the one-file path)


def probe_wav(path: str) -> WavInfo:
    """Header only (no sample data is read).  n_samples counts the whole frames actually present: a truncated file is
    processed up to where it ends, as the one-file path (audio_io.read_wav) and libsndfile do."""
    try:
        with open(path, "rb") as f:
            h = audio_io.parse_wav_header(f)
        ch, bits = h["channels"], h["bits"]
        if ch < 1 or bits % 8 or bits == 0:
            return WavInfo(path, error="File loading failed (bad fmt chunk)")
        if (h["format"], bits) not in ((1, 8), (1, 16), (1, 24), (1, 32), (3, 32), (3, 64)):
            return WavInfo(path, error=f"File loading failed (unsupported wav encoding: format {h['format']}, {bits} bits)")
        held = max(0, min(h["data_bytes"], os.path.getsize(path) - h["data_offset"])) // (bits // 8 * ch)
        return WavInfo(path, held, ch, h["sample_rate"], pcm16=(h["format"] == 1 and bits == 16))
    except Exception as e:          # the reference prints 'File loading failed' and returns None (prepare_dataset.py:163-165)
        return WavInfo(path, error=f"File loading failed ({e})")


@dataclass
class Group:
    files: list = field(default_factory=list)      # WavInfo
    tiles: list = field(default_factory=list)      # tiles per file
    frames: list = field(default_factory=list)
    channels: int = 1

    @property
    def n_tiles(self):
        return sum(self.tiles)

    @property
    def n_values(self):                             # int16 values in the group buffer
        return sum(f.n_samples * f.channels for f in self.files)


def group_budget(index: int, max_group_tiles: int, first_group_tiles: int | None) -> int:
    """Tile budget of the index-th group: the first group is small so that the detector starts after a few files
    have been read rather than after a whole batch (nothing overlaps the first read), then budgets double."""
    if not first_group_tiles:
        return max_group_tiles
    return min(max_group_tiles, first_group_tiles << min(index, 30))


LONG_REASON = "long recording (> 3401 s): processed on its own after the batched groups"
SOLO_REASON = "not 44.1 kHz PCM16: decoded / resampled on the one-file path after the batched groups"


def plan_groups(infos, const, max_group_tiles: int, stft_chunk: int = STFT_CHUNK, first_group_tiles: int | None = None):
    """Contiguous groups of whole files in the given order, group i holding at most
    `group_budget(i, ...)` detector tiles (a single larger file forms its own group) and one channel count per group.
    Returns (groups, rejected) -- rejected = [(WavInfo, reason)] for unreadable / unsupported files."""
    groups, rejected = [], []
    cur = None
    for info in infos:
        if info.error:
            rejected.append((info, info.error)); continue
        if info.sample_rate != 44100 or not info.pcm16:
            rejected.append((info, SOLO_REASON)); continue
        if info.n_samples > LONG_FILE_SAMPLES:
            rejected.append((info, LONG_REASON)); continue
        fr, nt = count_frames_tiles(info.n_samples, const["HOP_LENGTH"], const["W_PIX"], const["HOP_SPECTRO"], stft_chunk)
        if cur is None or cur.channels != info.channels or \
                (cur.files and cur.n_tiles + nt > group_budget(len(groups) - 1, max_group_tiles, first_group_tiles)):
            cur = Group(channels=info.channels)
            groups.append(cur)
        cur.files.append(info); cur.tiles.append(nt); cur.frames.append(fr)
    return groups, rejected


def read_into(info: WavInfo, dst: np.ndarray) -> None:
    """Decode one wav's int16 samples (interleaved if multi-channel) into `dst` (a slice of the pinned buffer)."""
    with open(info.path, "rb") as f:
        h = audio_io.parse_wav_header(f)
        raw = f.read(info.n_samples * info.channels * 2)
    a = np.frombuffer(raw, dtype="<i2")
    if a.size != dst.size:
        raise IOError(f"{info.path}: header promised {dst.size} values, file holds {a.size}")
    dst[:] = a


# --------------------------------------------------------------------- process-wide pools -----
_POOL_LOCK = threading.Lock()
_PLAN_POOL: dict = {}           # front-end parameters -> [idle FrontendPlan, ...]
_PINNED_POOL: list = []         # idle pinned int16 buffers


def _take_plan(fe_args):
    """A front-end plan nobody else is using (a plan owns one workspace and one side stream: one caller at a time), on the
    current device."""
    key = (torch.cuda.current_device(),) + tuple(fe_args)
    with _POOL_LOCK:
        idle = _PLAN_POOL.get(key)
        if idle:
            return idle.pop()
    return FrontendPlan(*fe_args)


def _give_plan(fe_args, plan):
    key = (plan.device.index if plan.device.index is not None else torch.cuda.current_device(),) + tuple(fe_args)
    with _POOL_LOCK:
        _PLAN_POOL.setdefault(key, []).append(plan)


def _take_pinned(n_values: int) -> torch.Tensor:
    with _POOL_LOCK:
        for i, t in enumerate(_PINNED_POOL):
            if t.numel() >= n_values:
                return _PINNED_POOL.pop(i)
    return torch.empty(max(n_values, 1), dtype=torch.int16, pin_memory=True)


def _give_pinned(bufs):
    with _POOL_LOCK:
        _PINNED_POOL.extend(bufs)
        _PINNED_POOL.sort(key=lambda t: t.numel())
        del _PINNED_POOL[:-6]                       # keep the six largest


# ------------------------------------------------------------------------- the pipeline -------
class DetectionPipeline:
    """``run(paths)`` yields ``(path, output_dict)`` in input order (recordings longer than 3401 s last), the
    dictionaries equal to ``run_detection.run_detection(model, config, path, ...)``."""

    def __init__(self, model, config, bird_dicts_path, min_score=0.5, bs=10, max_group_tiles=1024, readers=4, first_group_tiles=64,
                 freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024):
        if not torch.cuda.is_available():
            raise RuntimeError("the detection pipeline needs a CUDA device (no CPU fallback)")
        self.model, self.config, self.min_score, self.bs = model, config, min_score, bs
        self.max_group_tiles, self.readers, self.first_group_tiles = int(max_group_tiles), int(readers), first_group_tiles
        self.fe_args = (freq_accuracy, dt, overlap_spectro, w_pix)
        self.bird_dicts_path = bird_dicts_path
        self.const = derive_constants(*self.fe_args)
        with open(bird_dicts_path, "r") as f:
            birds = json.load(f)
        birds.update({"Non bird sound": 0})                       # run_detection.py:71
        self.reverse_dict = {idx: name for name, idx in birds.items()}
        self.counts = dict(files=0, tiles=0, detections=0, frames=0, t_front_us=0, t_model_us=0, t_post_us=0)
        self.failed: list = []
        # A plan owns ONE workspace and one side stream: it serves one caller at a time.  The pipeline runs its front-end
        # on a stream of its own while the caller's stream (and possibly File_Processor / run_detection calls on the
        # process-wide plan of frontend.get_plan) is busy with the detector, so it gets a private plan.
        self._plan = None
        self._trace = None

    # reader side: fills pinned buffers, two groups ahead at most
    def _reader(self, groups, out_q: queue.Queue, free_q: queue.Queue, pool: ThreadPoolExecutor):
        try:
            for g in groups:
                buf = free_q.get()
                if buf is None:
                    return
                arr = buf.numpy()
                offs, futs, o = [0], [], 0
                for info in g.files:
                    n = info.n_samples * info.channels
                    futs.append(pool.submit(read_into, info, arr[o:o + n]))
                    o += n
                    offs.append(o)
                bad = {}
                for i, fu in enumerate(futs):
                    try:
                        fu.result()
                    except Exception as e:
                        arr[offs[i]:offs[i + 1]] = 0
                        bad[i] = f"File loading failed ({e})"
                out_q.put((g, buf, offs, bad))
            out_q.put(None)
        except BaseException as e:          # surfaces in run(), never a silent short result
            out_q.put(e)

    def _mark(self, label):
        if self._trace is not None:
            self._trace.append((label, time.perf_counter()))

    def run(self, paths):
        self._trace = [] if os.environ.get("NBM_PIPE_TRACE") else None       # start-up timeline on stderr
        self._mark("start")
        infos = [probe_wav(p) for p in paths]
        groups, rejected = plan_groups(infos, self.const, self.max_group_tiles, first_group_tiles=self.first_group_tiles)
        long_files = [info.path for info, why in rejected if why in (LONG_REASON, SOLO_REASON)]
        self.failed += [(info.path, why) for info, why in rejected if why not in (LONG_REASON, SOLO_REASON)]
        self._mark(f"probed {len(infos)} files, {len(groups)} groups")
        if groups:
            plan = self._plan = _take_plan(self.fe_args)
            self._mark("front-end plan")
            try:
                yield from self._run_groups(plan, groups)
            finally:
                self._plan = None
                _give_plan(self.fe_args, plan)          # _run_groups has synchronised the device: nothing of ours is in flight
        # recordings longer than 3401 s: File_Processor cuts them into pieces that are one front-end batch already
        # (frontend.File_Processor._process_long); files at another sample rate or sample format need the host decode /
        # resample stage of File_Processor.load.  Both go through run_detection one at a time, after the groups
        for path in long_files:
            tm = {}
            try:
                output = run_detection(self.model, self.config, path, self.bird_dicts_path, min_score=self.min_score,
                                       bs=self.bs, timings=tm)
            except Exception as e:
                self.failed.append((path, str(e))); continue
            c = self.counts
            c["files"] += 1; c["tiles"] += tm["tiles"]; c["frames"] += tm["frames"]; c["detections"] += tm["detections"]
            c["t_front_us"] += int(tm["frontend_s"] * 1e6); c["t_model_us"] += int(tm["model_s"] * 1e6)
            c["t_post_us"] += int(tm["post_s"] * 1e6)
            yield path, output

    def _run_groups(self, plan, groups):
        dev = plan.device
        cap_tiles = max(g.n_tiles for g in groups)
        cap_vals = max(g.n_values for g in groups)
        n_buf = 2
        tiles_buf = [torch.empty((cap_tiles, 1, plan.n_bins, plan.w_pix), dtype=torch.float32, device=dev) for _ in range(n_buf)]
        pcm_dev = [torch.empty(cap_vals, dtype=torch.int16, device=dev) for _ in range(n_buf)]
        fe_stream = torch.cuda.Stream(device=dev)
        fe_start = [torch.cuda.Event(enable_timing=True) for _ in range(n_buf)]
        fe_done = [torch.cuda.Event(enable_timing=True) for _ in range(n_buf)]
        det_done = [None] * n_buf
        out_q: queue.Queue = queue.Queue()
        free_q: queue.Queue = queue.Queue()
        pinned = [_take_pinned(cap_vals) for _ in range(n_buf + 1)]        # one being filled, one in flight, one being consumed
        for t in pinned:
            free_q.put(t)
        pool = ThreadPoolExecutor(max_workers=self.readers)
        th = threading.Thread(target=self._reader, args=(groups, out_q, free_q, pool), daemon=True)
        th.start()
        self._mark("buffers, reader started")
        cur = torch.cuda.current_stream(dev)

        def enqueue_frontend(slot, item):
            g, host, offs, bad = item
            with torch.cuda.stream(fe_stream):
                if det_done[slot] is not None:
                    fe_stream.wait_event(det_done[slot])            # the detector has finished with this tile buffer
                fe_start[slot].record(fe_stream)
                d = pcm_dev[slot][:g.n_values]
                d.copy_(host[:g.n_values], non_blocking=True)
                ch = g.channels                                     # interleaved; offsets are per-channel sample indices
                _, tile_off, _ = plan.run_batch(d, [o // ch for o in offs], channels=ch, stream=fe_stream,
                                                out=tiles_buf[slot][:g.n_tiles])
                fe_done[slot].record(fe_stream)
            assert [tile_off[i + 1] - tile_off[i] for i in range(len(g.files))] == g.tiles, "host tile count != library"
            return tile_off

        def next_group():
            item = out_q.get()
            if isinstance(item, BaseException):
                raise item
            return item

        gc_was_on, last_gc = gc.isenabled(), time.perf_counter()
        gc.disable()                                                # postproc.gc_paused explains; restored in `finally`
        try:
            nxt = next_group()
            pending = None if nxt is None else (0, nxt, enqueue_frontend(0, nxt))
            while pending is not None:
                slot, item, tile_off = pending
                g, host, offs, bad = item
                nxt = next_group()                                  # next group's PCM (read while we were busy)
                pending = None if nxt is None else ((slot + 1) % n_buf, nxt, enqueue_frontend((slot + 1) % n_buf, nxt))
                cur.wait_event(fe_done[slot])
                fe_done[slot].synchronize()
                self._mark(f"group of {len(g.files)} files transformed")
                free_q.put(host)                                    # H2D done: the pinned buffer can be refilled
                self.counts["t_front_us"] += int(fe_start[slot].elapsed_time(fe_done[slot]) * 1e3)
                good = []
                for i, info in enumerate(g.files):
                    if i in bad:
                        self.failed.append((info.path, bad[i]))
                    else:
                        good.append(i)
                # one detector stream per group: the first batches of file i+1 are already running on the detector's
                # replay lanes while file i's boxes are merged and written.  A file the detector fails on (the reference
                # crashes there, e.g. "RPN failed" on a silent recording) is reported and left out; the stream is
                # restarted behind it.
                while good:
                    stream = detect_stream(self.model, (tiles_buf[slot][tile_off[i]:tile_off[i + 1], 0] for i in good),
                                           self.min_score, self.bs)
                    n_done = 0
                    t0 = time.perf_counter()
                    try:
                        for i, outputs in zip(good, stream):
                            info = g.files[i]
                            t1 = time.perf_counter()
                            fp = SimpleNamespace(W_PIX=self.const["W_PIX"], HOP_SPECTRO=self.const["HOP_SPECTRO"],
                                                 spectrogram_length=g.frames[i])
                            output = postproc.merge_to_output(fp, outputs, self.config.num_classes, self.reverse_dict)
                            t2 = time.perf_counter()
                            c = self.counts
                            c["files"] += 1; c["tiles"] += g.tiles[i]; c["frames"] += g.frames[i]
                            c["detections"] += sum(len(v["scores"]) for v in output.values())
                            c["t_model_us"] += int((t1 - t0) * 1e6); c["t_post_us"] += int((t2 - t1) * 1e6)
                            n_done += 1
                            if self.counts["files"] <= 2:
                                self._mark(f"file {self.counts['files']} merged")
                            yield info.path, output
                            t0 = time.perf_counter()
                        good = []
                    except Exception as e:
                        k = getattr(e, "nbm_file_index", None)
                        if k is None:                               # a detector without attribution: the file it was asked for
                            k = n_done
                        k = max(k, n_done)                          # (merge errors surface at the file being merged)
                        self.failed.append((g.files[good[k]].path, f"{type(e).__name__}: {e}"))
                        torch.cuda.synchronize(dev)
                        good = good[n_done:k] + good[k + 1:]
                det_done[slot] = torch.cuda.Event()
                det_done[slot].record(cur)
                if gc_was_on and time.perf_counter() - last_gc > 120.0:
                    # between groups the heap is small; a full collection still walks the interpreter's whole heap
                    # (~0.1 s), so only every two minutes: cycles (tracebacks of failed files) go here
                    gc.collect()
                    last_gc = time.perf_counter()
        finally:
            if gc_was_on:
                gc.enable()
            free_q.put(None)                                        # unblock the reader if we stop early
            pool.shutdown(wait=False, cancel_futures=True)
            torch.cuda.synchronize(dev)
            th.join(timeout=10.0)                                   # the reader no longer writes into the pinned buffers
            if not th.is_alive():
                _give_pinned(pinned)
            self._mark("done")
            if self._trace:
                import sys
                t0 = self._trace[0][1]
                print("[pipeline] " + "; ".join(f"{lab} +{(t - t0) * 1e3:.1f} ms" for lab, t in self._trace[1:]), file=sys.stderr)
