"""CPU oracle (numpy, float64) for the waveform -> detector-tile front-end.

TEST INFRASTRUCTURE: checker only, never on the product path (see
``oracle/__init__.py``).  Restates, stage by stage,
``/root/reference/nbm_model/nbm_datasets/prepare_dataset.py`` ``File_Processor``:

  derive_params      <- process_file            prepare_dataset.py:108-138
  load_pcm16 / to_float <- load (librosa.core.load, sr=None)   :160-165
  stft               <- librosa.stft(y, n_fft, hop_length)     :237
                        (third-party, version un-pinned; librosa>=0.10
                        defaults: center=True, pad_mode='constant', periodic
                        Hann in float64, rfft in float64, stored complex64)
  db_spectrogram     <- amp_to_db(np.abs(.)) + band crop       :228-230,240,247
  normalise          <- whole-file min/max                     :248-250
  tile               <- split_power_spec (no-label branch)     :255-294
  process            <- process_file end to end -> float64 tiles, exactly as
                        run_detection.py:53 sees them before torch.Tensor().

Parity status: checked bit-for-bit against the reference's own code run through
``oracle/ref_shims.py`` (tests/test_oracle_frontend.py, container only) and
against committed fixtures in tests/golden/.  The librosa layer itself is
**parity unpinned** (no librosa, no reference fixtures).
"""
from __future__ import annotations

import dataclasses
import wave

import numpy as np
import scipy.signal

STFT_CHUNK = int(5e7)          # prepare_dataset.py:234
LONG_FILE_SAMPLES = int(15e7) - int(15e7) % 44100   # prepare_dataset.py:194


@dataclasses.dataclass(frozen=True)
class FrontendParams:
    """The constants process_file() derives (prepare_dataset.py:114-138)."""
    sample_rate: int
    n_fft: int
    hop: int
    low_idx: int
    high_idx: int
    h_pix: int
    w_pix: int
    hop_spectro: int
    freq_accuracy: float
    dt: float
    low_freq: float
    high_freq: float
    min_level: float

    @property
    def n_bins(self) -> int:
        return self.high_idx - self.low_idx


def derive_params(freq_accuracy=33.3, dt=0.003, overlap_spectro=0.2, w_pix=1024,
                  sample_rate=44100, h_pix=375, low_freq=500) -> FrontendParams:
    hop_spectro = int((1 - overlap_spectro) * w_pix)                 # :115
    n_fft = int(sample_rate / freq_accuracy)                         # :125
    hop = int(sample_rate * dt)                                      # :126
    overlap_fft = np.round(1 - hop / n_fft, 3)                       # :127
    fa = sample_rate / n_fft                                         # :130
    dt_real = int((1 - overlap_fft) * n_fft) / sample_rate           # :131
    low_idx = 1 + int(low_freq / fa)                                 # :134
    high_idx = low_idx + h_pix                                       # :135
    min_level = float(np.exp(-100 / 20 * np.log(10)))                # :229
    return FrontendParams(sample_rate, n_fft, hop, low_idx, high_idx, h_pix, w_pix,
                          hop_spectro, fa, dt_real, (low_idx - 1) * fa,
                          (high_idx - 1) * fa, min_level)


def n_frames(n_samples: int, p: FrontendParams) -> int:
    """Total STFT columns of a file, summed over its <=5e7-sample STFT chunks."""
    total = 0
    for k in range(int(n_samples / STFT_CHUNK) + 1):                 # :236
        seg = max(0, min(n_samples, (k + 1) * STFT_CHUNK) - k * STFT_CHUNK)
        total += 1 + seg // p.hop
    return total


def n_tiles(total_frames: int, p: FrontendParams) -> int:
    return max(1, int(1 + np.ceil((total_frames - p.w_pix) / p.hop_spectro)))   # :266


def load_pcm16(path: str) -> tuple[np.ndarray, int]:
    """PCM16 wav -> int16 [n, channels]."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("oracle reads PCM16 wavs only")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch)
    return pcm, sr


def to_float(pcm: np.ndarray) -> np.ndarray:
    """soundfile's int16 -> float32 scaling (x / 2**15) and librosa's to_mono (mean)."""
    x = pcm.astype(np.float32) / np.float32(32768.0)
    if x.ndim == 2:
        x = x[:, 0] if x.shape[1] == 1 else np.mean(x.T, axis=0)
    return np.ascontiguousarray(x, dtype=np.float32)


def stft(y: np.ndarray, n_fft: int, hop: int, pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(y, n_fft=n_fft, hop_length=hop) with librosa>=0.10 defaults.

    float32 signal, centre padding of n_fft//2 both sides, frames t = 0..len(y)//hop,
    float64 periodic Hann, float64 rfft, result stored as complex64 [1+n_fft//2, T].
    """
    y = np.asarray(y, dtype=np.float32)
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)       # float64
    half = n_fft // 2
    ypad = np.pad(y, (half, half), mode=pad_mode)
    T = 1 + len(y) // hop
    out = np.empty((1 + half, T), dtype=np.complex64, order="F")
    step = 4096
    for t0 in range(0, T, step):
        t1 = min(T, t0 + step)
        idx = (np.arange(t0, t1) * hop)[:, None] + np.arange(n_fft)[None, :]
        frames = ypad[idx]                                           # [t, n_fft] float32
        out[:, t0:t1] = np.fft.rfft(win[None, :] * frames, axis=1).T
    return out


def db_spectrogram(y: np.ndarray, p: FrontendParams, pad_mode="constant") -> list[np.ndarray]:
    """Un-normalised dB band, one float64 [n_bins, T_k] array per STFT chunk (:233-247)."""
    out = []
    for k in range(int(len(y) / STFT_CHUNK) + 1):
        z = stft(y[k * STFT_CHUNK:(k + 1) * STFT_CHUNK], p.n_fft, p.hop, pad_mode)
        mag = np.abs(z)                                              # float32
        db = 20 * np.log10(np.maximum(np.float64(p.min_level), mag))  # float64 (NEP 50)
        out.append(db[p.low_idx:p.high_idx, :])
    return out


def normalise(chunks: list[np.ndarray]) -> tuple[list[np.ndarray], float, float]:
    s_max = max(c.max() for c in chunks)
    s_min = min(c.min() for c in chunks)
    return [(c - s_min) / (s_max - s_min) for c in chunks], float(s_min), float(s_max)


def tile_plan(chunk_frames: list[int], p: FrontendParams) -> list[list[tuple[int, int, int]]]:
    """For every detector window, the (chunk, first column, end column) pieces the
    reference concatenates (:261-278).  Follows its bin search literally, including the
    quirk that a window which starts in chunk c and runs past the END OF THE FILE inside
    chunk c+1 keeps only chunk c's columns (e_bin is the sentinel, so `next_bin` is False)."""
    edges = np.cumsum([0] + list(chunk_frames))
    total = int(edges[-1])
    plan = []
    for k in range(n_tiles(total, p)):
        start, end = k * p.hop_spectro, k * p.hop_spectro + p.w_pix
        s_bin = int((start >= edges).sum()) - 1
        e_bin = int((end > edges).sum()) - 1
        inside = e_bin < len(edges) - 1
        s_off = start - int(edges[s_bin])
        if e_bin > s_bin and inside:
            plan.append([(s_bin, s_off, chunk_frames[s_bin]), (e_bin, 0, end - int(edges[e_bin]))])
        else:
            stop = end - int(edges[e_bin]) if inside else chunk_frames[s_bin]
            plan.append([(s_bin, s_off, min(stop, chunk_frames[s_bin]))])
    return plan


def tile(chunks: list[np.ndarray], p: FrontendParams) -> list[np.ndarray]:
    """split_power_spec, no-label branch: windows of w_pix at hop_spectro over the
    concatenation of the chunks; the last one right-padded by numpy 'reflect'."""
    tiles = [np.concatenate([chunks[c][:, a:b] for (c, a, b) in segs], axis=1)
             for segs in tile_plan([c.shape[1] for c in chunks], p)]
    last = tiles[-1]
    if last.shape[1] < p.w_pix:
        # :283-292 with labels None: empty_width = w_pix, so a single pad call of w_pix - w
        tiles[-1] = np.pad(last, ((0, 0), (0, p.w_pix - last.shape[1])), mode="reflect")
    return tiles


def reflect_index(j: int, w: int) -> int:
    """Source column of padded column j (>= w) under numpy's iterated 'reflect' pad of
    a width-w row; used by tests to check padded columns are exact copies."""
    if w == 1:
        return 0
    period = 2 * (w - 1)
    r = j % period
    return r if r < w else period - r


@dataclasses.dataclass
class FrontendResult:
    tiles: list            # float64 [n_bins, w_pix] each
    s_min: float
    s_max: float
    spectrogram_length: int
    params: FrontendParams


def process(pcm_or_float: np.ndarray, p: FrontendParams | None = None,
            pad_mode: str = "constant") -> FrontendResult:
    p = p or derive_params()
    y = to_float(pcm_or_float) if pcm_or_float.dtype == np.int16 else np.asarray(pcm_or_float, np.float32)
    if len(y) > LONG_FILE_SAMPLES:
        raise ValueError("files longer than 3401 s take the reference's (broken) long-file branch")
    chunks = db_spectrogram(y, p, pad_mode)
    norm, s_min, s_max = normalise(chunks)
    total = int(sum(c.shape[1] for c in norm))
    return FrontendResult(tile(norm, p), s_min, s_max, total, p)


def process_file(path: str, **kw) -> FrontendResult:
    pcm, sr = load_pcm16(path)
    if sr != 44100:
        raise ValueError("oracle handles 44.1 kHz input only (reference shells out to ffmpeg)")
    return process(pcm, derive_params(**kw))
