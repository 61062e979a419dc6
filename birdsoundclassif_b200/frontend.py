"""GPU front-end behind the reference's ``File_Processor`` interface.

Mirrors ``nbm_model/nbm_datasets/prepare_dataset.py`` ``File_Processor`` (:92-294) for the
inference path: same constructor, same ``process_file`` keyword defaults, same attributes
set on the object (``W_PIX, HOP_SPECTRO, WIN_LENGTH, HOP_LENGTH, FREQ_ACCURACY, DT, LOW_IDX,
HIGH_IDX, LOW_FREQ, HIGH_FREQ, spectrogram_length``) and the same ``(None, None)`` failure
convention.  The arithmetic runs in libnbm_b200.so on the current CUDA device; the returned
``img_db`` is a float32 CUDA tensor ``[n_tiles, H_PIX, W_PIX]`` (``len()``-able and
indexable like the reference's list of arrays), already laid out as the detector batch
``run_detection.py:53-55`` builds, so no host round trip happens.

PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib, audio_io, labels as labels_mod

STFT_CHUNK = int(5e7)                                   # prepare_dataset.py:234
LONG_FILE_SAMPLES = int(15e7) - int(15e7) % 44100       # prepare_dataset.py:194


def derive_constants(freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024,
                     sample_rate=44100, h_pix=375, low_freq=500) -> dict:
    """The scalar derivations of process_file (prepare_dataset.py:114-138), same float/int
    arithmetic, as a dict keyed by the reference's attribute names."""
    W_PIX = w_pix
    HOP_SPECTRO = int((1 - overlap_spectro) * W_PIX)
    WIN_LENGTH = int(sample_rate / freq_accuracy)
    HOP_LENGTH = int(sample_rate * dt)
    overlap_fft = float(np.round(1 - HOP_LENGTH / WIN_LENGTH, 3))
    FREQ_ACCURACY = sample_rate / WIN_LENGTH
    DT = int((1 - overlap_fft) * WIN_LENGTH) / sample_rate
    LOW_IDX = 1 + int(low_freq / FREQ_ACCURACY)
    HIGH_IDX = LOW_IDX + h_pix
    return dict(W_PIX=W_PIX, HOP_SPECTRO=HOP_SPECTRO, WIN_LENGTH=WIN_LENGTH, HOP_LENGTH=HOP_LENGTH,
                FREQ_ACCURACY=FREQ_ACCURACY, DT=DT, LOW_IDX=LOW_IDX, HIGH_IDX=HIGH_IDX,
                LOW_FREQ=(LOW_IDX - 1) * FREQ_ACCURACY, HIGH_FREQ=(HIGH_IDX - 1) * FREQ_ACCURACY)


def _stream_ptr(stream) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


class FrontendPlan:
    """Owns an ``nbm_frontend_plan`` (device twiddle tables) plus a reusable torch workspace.

    A plan is SINGLE-CALLER: its workspace (descriptors, dB band, candidate list), its side stream and its events are
    shared by every run, and a run only orders itself after the stream it is given.  Two runs of the same plan must
    therefore be issued to the same stream (or be separated by an event / synchronisation); callers that overlap
    front-end work with other work on another stream create a plan of their own (``pipeline.DetectionPipeline`` does).
    ``get_plan`` returns the process-wide plan used by ``File_Processor`` on the current stream."""

    def __init__(self, freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024,
                 sample_rate=44100, h_pix=375, low_freq=500, device=None, stft_chunk=STFT_CHUNK):
        if not torch.cuda.is_available():
            raise _lib.NbmError("the NBM front-end needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.const = derive_constants(freq_accuracy, dt, overlap_spectro, w_pix, sample_rate, h_pix, low_freq)
        c = self.const
        self.n_bins, self.w_pix = h_pix, w_pix
        self.params = _lib.FrontendParams(
            sample_rate=sample_rate, n_fft=c["WIN_LENGTH"], hop=c["HOP_LENGTH"], low_idx=c["LOW_IDX"],
            n_bins=h_pix, w_pix=w_pix, hop_spectro=c["HOP_SPECTRO"], pad_mode=0, stft_chunk=int(stft_chunk),
            min_level=float(np.exp(-100 / 20 * np.log(10))))                # prepare_dataset.py:229
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbm_frontend_plan_create(C.byref(self.params), C.byref(self._h)),
                       "nbm_frontend_plan_create")
        self._ws = None
        self.impl = "tcgen05" if _lib.lib().nbm_frontend_impl(self._h) == 1 else "cuda-core"

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nbm_frontend_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host arithmetic -------------------------------------------------------------------
    def query(self, n_samples: int):
        nf, nt, ws = C.c_int64(), C.c_int64(), C.c_size_t()
        _lib.check(_lib.lib().nbm_frontend_query(self._h, int(n_samples), C.byref(nf), C.byref(nt), C.byref(ws)),
                   "nbm_frontend_query")
        return nf.value, nt.value, ws.value

    def query_batch(self, n_samples):
        n = len(n_samples)
        ns = (C.c_int64 * n)(*[int(v) for v in n_samples])
        nf = (C.c_int64 * n)()
        to = (C.c_int64 * (n + 1))()
        ws = C.c_size_t()
        _lib.check(_lib.lib().nbm_frontend_query_batch(self._h, ns, n, nf, to, C.byref(ws)), "nbm_frontend_query_batch")
        return list(nf), list(to), ws.value

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    # -- device work --------------------------------------------------------------------------
    @staticmethod
    def _pcm_dtype(pcm: torch.Tensor) -> int:
        if pcm.dtype == torch.int16:
            return 0
        if pcm.dtype == torch.float32:
            return 1
        raise TypeError("pcm must be int16 or float32")

    def run(self, pcm: torch.Tensor, stream=None, out: torch.Tensor | None = None):
        """pcm: CUDA tensor [n] or [n, channels] -> (tiles [n_tiles,1,n_bins,w_pix], minmax [2])."""
        assert pcm.is_cuda and pcm.is_contiguous()
        n = pcm.shape[0]
        ch = 1 if pcm.dim() == 1 else pcm.shape[1]
        tiles, _, minmax = self.run_batch(pcm, [0, n], channels=ch, stream=stream, out=out)
        return tiles, minmax[0]

    def run_batch(self, pcm: torch.Tensor, sample_offsets, channels=1, stream=None, out=None, minmax_out=None):
        """pcm: flat CUDA tensor holding every file back to back; sample_offsets: n_files+1
        per-channel sample indices.  Returns (tiles [total_tiles,1,n_bins,w_pix],
        tile_offsets list[n_files+1], minmax [n_files, 2])."""
        assert pcm.is_cuda and pcm.is_contiguous()
        n_files = len(sample_offsets) - 1
        sizes = [int(sample_offsets[i + 1] - sample_offsets[i]) for i in range(n_files)]
        _, tile_off, ws_bytes = self.query_batch(sizes)
        total = tile_off[-1]
        if out is None:
            out = torch.empty((total, 1, self.n_bins, self.w_pix), dtype=torch.float32, device=self.device)
        else:
            assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and \
                out.numel() >= total * self.n_bins * self.w_pix
        minmax = minmax_out if minmax_out is not None else torch.empty((n_files, 2), dtype=torch.float32, device=self.device)
        assert minmax.is_contiguous() and minmax.shape == (n_files, 2)
        ws = self._workspace(ws_bytes)
        offs = (C.c_int64 * (n_files + 1))(*[int(v) for v in sample_offsets])
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbm_frontend_run_batch(
                self._h, pcm.data_ptr(), self._pcm_dtype(pcm), int(channels), offs, n_files, out.data_ptr(),
                minmax.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(stream)), "nbm_frontend_run_batch")
        self._last_sizes = sizes
        return out, tile_off, minmax

    def set_profiling(self, enable: bool):
        _lib.check(_lib.lib().nbm_frontend_set_profiling(self._h, int(bool(enable))), "nbm_frontend_set_profiling")

    def get_profile(self):
        """(stft_ms, tile_ms, runs) accumulated since profiling was enabled; waits for the last run."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(_lib.lib().nbm_frontend_get_profile(self._h, C.byref(a), C.byref(b), C.byref(n)),
                   "nbm_frontend_get_profile")
        return a.value, b.value, n.value

    def get_profile_kernels(self):
        """({'anchor': ms, 'stft': ms, 'minmax': ms, 'tile': ms}, runs) accumulated since profiling was enabled."""
        ms, n = (C.c_double * 4)(), C.c_int64()
        _lib.check(_lib.lib().nbm_frontend_get_profile_kernels(self._h, ms, C.byref(n)), "nbm_frontend_get_profile_kernels")
        return dict(zip(("anchor", "stft", "minmax", "tile"), list(ms))), n.value

    def last_listed(self):
        """(blocks listed for float64 refinement by the last run, capacity of the list, pixels recomputed); waits for
        that run."""
        n, cap, px = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().nbm_frontend_last_listed(self._h, C.byref(n), C.byref(cap), C.byref(px)), "nbm_frontend_last_listed")
        return n.value, cap.value, px.value

    def run_batch_from_host(self, host_pcm: torch.Tensor, sample_offsets, out=None, files_per_chunk=64):
        """run_batch for mono PCM16 in PINNED host memory: the files are copied to the device in chunks on a
        side stream while the previous chunk is being transformed (two device staging buffers), so the
        call is bound by max(PCIe, front-end) rather than their sum.  Returns like run_batch."""
        assert not host_pcm.is_cuda and host_pcm.is_pinned() and host_pcm.dim() == 1
        n_files = len(sample_offsets) - 1
        offs = [int(v) for v in sample_offsets]
        _, tile_off, _ = self.query_batch([offs[i + 1] - offs[i] for i in range(n_files)])
        if out is None:
            out = torch.empty((tile_off[-1], 1, self.n_bins, self.w_pix), dtype=torch.float32, device=self.device)
        minmax = torch.empty((n_files, 2), dtype=torch.float32, device=self.device)
        chunks = [(f, min(n_files, f + files_per_chunk)) for f in range(0, n_files, files_per_chunk)]
        need = max(offs[b] - offs[a] for a, b in chunks)
        st = getattr(self, "_h2d", None)
        if st is None or st["buf"][0].numel() < need or st["buf"][0].dtype != host_pcm.dtype:
            st = {"buf": [torch.empty(need, dtype=host_pcm.dtype, device=self.device) for _ in range(2)],
                  "stream": torch.cuda.Stream(device=self.device),
                  "ready": [torch.cuda.Event() for _ in range(2)], "done": [torch.cuda.Event() for _ in range(2)]}
            self._h2d = st
        cur = torch.cuda.current_stream(self.device)
        st["stream"].wait_stream(cur)
        for ci, (fa, fb) in enumerate(chunks):
            b = ci & 1
            s0, s1 = offs[fa], offs[fb]
            with torch.cuda.stream(st["stream"]):
                if ci >= 2:
                    st["stream"].wait_event(st["done"][b])          # chunk ci-2 no longer reads this buffer
                st["buf"][b][:s1 - s0].copy_(host_pcm[s0:s1], non_blocking=True)
                st["ready"][b].record(st["stream"])
            cur.wait_event(st["ready"][b])
            self.run_batch(st["buf"][b][:s1 - s0], [o - s0 for o in offs[fa:fb + 1]],
                           out=out[tile_off[fa]:tile_off[fb]], minmax_out=minmax[fa:fb])
            st["done"][b].record(cur)
        self._last_sizes = None
        return out, tile_off, minmax

    def spectrogram_view(self, file_index: int = 0) -> torch.Tensor:
        """Un-normalised dB band [n_bins, n_frames] of a file from the LAST run (a view into the
        workspace; for tests and diagnostics)."""
        sizes = self._last_sizes
        n = len(sizes)
        ns = (C.c_int64 * n)(*sizes)
        off, stride = C.c_size_t(), C.c_int64()
        _lib.check(_lib.lib().nbm_frontend_spectrogram_view(self._h, ns, n, file_index, C.byref(off), C.byref(stride)),
                   "nbm_frontend_spectrogram_view")
        nf = self.query(sizes[file_index])[0]
        flat = self._ws[off.value: off.value + self.n_bins * stride.value * 4].view(torch.float32)
        return flat.view(self.n_bins, stride.value)[:, :nf]


_PLANS: dict = {}


def get_plan(freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024, sample_rate=44100,
             h_pix=375, low_freq=500) -> FrontendPlan:
    key = (freq_accuracy, dt, overlap_spectro, w_pix, sample_rate, h_pix, low_freq, torch.cuda.current_device())
    if key not in _PLANS:
        _PLANS[key] = FrontendPlan(freq_accuracy, dt, overlap_spectro, w_pix, sample_rate, h_pix, low_freq)
    return _PLANS[key]


class File_Processor:
    """Drop-in for the reference class (prepare_dataset.py:92-376).  Without ``labels`` -- the inference path --
    ``process_file`` answers ``(tiles, None)``; with the reference's label table it answers ``(tiles, annotations)``:
    the last tile padded by the annotated rule and one row of boxes per annotated tile (labels.py)."""

    H_PIX = 375      # px          prepare_dataset.py:96
    LOW_FREQ = 500   # hz          prepare_dataset.py:97
    FREQ = 44100     # hz          prepare_dataset.py:98
    requantise_long = True      # long recordings: the pieces pass through PCM16 temp files upstream (see _process_long)

    def __init__(self, filepath, extra_str_label="", labels=None):
        self.labels = labels
        self.ext = os.path.basename(filepath).split(".")[-1]
        self.filename = os.path.basename(filepath).replace("." + self.ext, "").replace(extra_str_label, "")
        self.filepath = filepath

    def load(self):
        """wav -> pinned host tensor: int16 for PCM16 (the /32768 scaling and the mono mix happen on the device), float32
        with libsndfile's scaling for the other encodings (audio_io.read_wav).  Returns None on failure like the
        reference (prepare_dataset.py:160-165).  A file that is not at 44.1 kHz is converted the way the reference's
        ffmpeg call converts it -- mono, 44.1 kHz, PCM16 (prepare_dataset.py:166-182) -- by audio_io.resample_pcm16."""
        try:
            pcm, sr = audio_io.read_wav(self.filepath)
        except Exception:
            print("File loading failed")
            return None
        if sr != self.FREQ:
            print(f"{self.filepath}: {sr} Hz -> {self.FREQ} Hz mono PCM16 (polyphase resampler in place of the reference's ffmpeg call)")
            pcm = audio_io.resample_pcm16(pcm, sr, self.FREQ)
        t = torch.empty(pcm.shape, dtype=torch.int16 if pcm.dtype == np.int16 else torch.float32,
                        pin_memory=torch.cuda.is_available())
        t.numpy()[...] = pcm            # one copy, decoded samples -> pinned memory
        return t

    def process_file(self, freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024):
        data = self.load()
        if data is None:
            return None, None
        return self.process_pcm(data, freq_accuracy, dt, overlap_spectro, w_pix)

    def process_pcm(self, data: torch.Tensor, freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024):
        """Same as process_file for PCM already in memory (host pinned or CUDA, int16/float32)."""
        n = data.shape[0]
        plan = get_plan(freq_accuracy, dt, overlap_spectro, w_pix, self.FREQ, self.H_PIX, type(self).LOW_FREQ)
        for k, v in plan.const.items():
            setattr(self, k, v)
        dev = data if data.is_cuda else data.to(plan.device, non_blocking=True)
        if n > LONG_FILE_SAMPLES:
            return self._process_long(plan, dev, n)
        tiles, minmax = plan.run(dev)
        self.spectrogram_length, _, _ = plan.query(n)
        self.s_min_max = minmax
        if self.labels is None:
            return tiles[:, 0], None
        return self._annotate(tiles[:, 0], self.labels, self.filename, self.spectrogram_length)

    def _annotate(self, tiles, table, filename, spectrogram_length):
        """The labelled branches of split_power_spec (prepare_dataset.py:280-292: the last tile is padded in steps that
        never mirror an annotated call into the padding) and process_file (:146-153: the per-tile box table, or
        ``(None, None)`` when the table has no row for this recording)."""
        rows = labels_mod.file_rows(table, filename)
        w_last = spectrogram_length - (len(tiles) - 1) * self.HOP_SPECTRO
        if len(rows) > 0 and w_last < self.W_PIX:
            src = labels_mod.labelled_pad_map(w_last, self.W_PIX, labels_mod.empty_width_of(rows, spectrogram_length, self.DT))
            pad = torch.from_numpy(src[w_last:]).to(tiles.device)
            tiles[-1][:, w_last:] = tiles[-1][:, :w_last].index_select(1, pad)
        c = dict(DT=self.DT, FREQ_ACCURACY=self.FREQ_ACCURACY, LOW_FREQ=self.LOW_FREQ, HIGH_FREQ=self.HIGH_FREQ,
                 W_PIX=self.W_PIX, HOP_SPECTRO=self.HOP_SPECTRO, H_PIX=self.H_PIX)
        try:
            return tiles, labels_mod.merge_and_filter_labels(table, filename, self.ext, len(tiles), c)
        except labels_mod.NoLabelsForFile:
            print("Something went wrong with the annotation file, skipping~~")
            return None, None

    def _process_long(self, plan, dev: torch.Tensor, n: int):
        """prepare_dataset.py:187-225: a recording longer than max_l = 3401 s is cut into max_l-sample pieces and every
        piece is processed as a file of its own (own STFT padding, own min/max, own tiling, own reflect-padded tail).
        The reference round-trips the pieces through temp wav files; here they are the `files` of one batched
        front-end call.  Returns the reference's ``(img_db, annotations)`` = ([tiles of piece 0, tiles of piece 1, ...],
        []) -- a structure only prepare_dataset() consumes upstream (run_detection.py:47-53 crashes on it; this
        package's run_detection handles it, see there).  An empty trailing piece (length an exact multiple of max_l)
        is dropped; the reference writes an empty temp wav there and gets one meaningless image."""
        L = LONG_FILE_SAMPLES
        cuts = [min(n, k * L) for k in range(int(n / L) + 2)]
        if cuts[-1] == cuts[-2]:
            cuts.pop()
        ch = 1 if dev.dim() == 1 else dev.shape[1]
        if self.requantise_long:
            # The reference writes every piece with soundfile.write(outp, float32 data, 44100) -- for .wav that is subtype
            # PCM_16, libsndfile's float -> short conversion lrintf(x * 0x7FFF) (pcm.c, norm_float on) -- and reads it back
            # as int16 / 32768 (prepare_dataset.py:199, 162).  x = s / 32768 for a PCM16 source, so samples with
            # |s| >= 16384 come back one LSB smaller.  Restated here in the same float32 arithmetic (libsndfile is not in
            # this image, so this line is unpinned; requantise_long = False keeps the source samples).
            x = dev.to(torch.float32) / 32768.0 if dev.dtype == torch.int16 else dev
            if ch > 1:
                x = x.mean(dim=1)                                   # librosa.load mixes to mono before the split
                ch = 1
            dev = torch.round(x * 32767.0).clamp_(-32768, 32767).to(torch.int16)
        tiles, tile_off, minmax = plan.run_batch(dev.reshape(-1), cuts, channels=ch)
        self.piece_spectrogram_lengths = [plan.query(cuts[k + 1] - cuts[k])[0] for k in range(len(cuts) - 1)]
        self.piece_samples = L
        self.s_min_max = minmax
        img_db = [tiles[tile_off[k]:tile_off[k + 1], 0] for k in range(len(cuts) - 1)]
        annotations = []
        if self.labels is not None:
            # every piece against the annotations that start inside it, on its own time axis (:205-219); a piece
            # without any is processed as unannotated, one whose table comes back empty answers (None, None) upstream
            for k in range(len(img_db)):
                rows = labels_mod.piece_labels(self.labels, self.filename, k, L / self.FREQ)
                if rows is None:
                    continue
                img_db[k], ann = self._annotate(img_db[k], rows, f"temp{k}", self.piece_spectrogram_lengths[k])
                if ann is not None:
                    annotations.append(ann)
        return img_db, annotations
