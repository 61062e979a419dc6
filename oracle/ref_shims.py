"""Run the UNMODIFIED reference: ``/root/reference`` in the build container, or the byte-identical
copy of its ``nbm_model`` package that ``oracle/build_ref.py`` leaves in git-ignored ``oracle/_ref/``
(which travels to the GPU box with the snapshot).

TEST INFRASTRUCTURE: used by ``oracle/make_golden.py``, by the tests that pin the oracle
restatements and the system-level parity tests, and by ``bench.py``'s reference legs.  Callers
skip when ``have_reference()`` is False.

The reference imports third-party modules that are not in this image.  We install
``sys.modules`` stand-ins for them:

* ``librosa`` / ``librosa.core``: ``load`` (PCM16 wav -> float32 mono, like
  soundfile + to_mono) and ``stft`` (librosa>=0.10 defaults, written the way
  librosa does it: centre pad, strided frames, float64 window product, rfft,
  complex64 result in Fortran order).  This is the only arithmetic we supply; the
  rest of ``File_Processor`` is the reference's own code.
* ``soundfile``, ``imageio``, ``imageio.v2``, ``ffmpeg``: empty modules.
* ``matplotlib*``: catch-all stubs (must raise AttributeError for dunder lookups or
  torchvision's ``inspect`` calls crash).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import wave

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "nbm_model", "run_detection.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nbm_model", "run_detection.py"))


# ----------------------------------------------------------------------------- librosa
def _load(path, sr=None, **_):
    with wave.open(path, "rb") as w:
        rate, ch, n, width = w.getframerate(), w.getnchannels(), w.getnframes(), w.getsampwidth()
        raw = w.readframes(n)
    if width != 2:
        raise ValueError("shim reads PCM16 only")
    x = np.frombuffer(raw, dtype="<i2").reshape(-1, ch).astype(np.float32) / np.float32(32768.0)
    y = x[:, 0].copy() if ch == 1 else np.mean(x.T, axis=0)
    if sr is not None and sr != rate:
        raise ValueError("shim cannot resample")
    return y.astype(np.float32), rate


def _stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
          dtype=None, pad_mode="constant"):
    import scipy.signal
    hop_length = hop_length or n_fft // 4
    assert win_length in (None, n_fft) and window == "hann" and center
    fft_window = scipy.signal.get_window("hann", n_fft, fftbins=True).reshape(-1, 1)
    ypad = np.pad(np.asarray(y), n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(ypad) - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        ypad, shape=(n_fft, n_frames), strides=(ypad.itemsize, hop_length * ypad.itemsize))
    out = np.empty((1 + n_fft // 2, n_frames), dtype=np.complex64, order="F")
    cols = max(1, 2 ** 22 // n_fft)
    for s in range(0, n_frames, cols):
        t = min(n_frames, s + cols)
        out[:, s:t] = np.fft.rfft(fft_window * frames[:, s:t], axis=0)
    return out


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


_installed = False


def install() -> None:
    """Idempotently install the shims and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not have_reference():
        raise RuntimeError("reference tree not present; container-only facility")
    librosa = types.ModuleType("librosa")
    core = types.ModuleType("librosa.core")
    core.load = _load
    core.stft = _stft
    librosa.core = core
    librosa.load = _load
    librosa.stft = _stft
    sys.modules.setdefault("librosa", librosa)
    sys.modules.setdefault("librosa.core", core)
    for name in ("soundfile", "imageio", "imageio.v2", "ffmpeg"):
        sys.modules.setdefault(name, types.ModuleType(name))
    for name in ("matplotlib", "matplotlib.ticker", "matplotlib.patches", "matplotlib.pyplot",
                 "seaborn"):
        sys.modules.setdefault(name, _StubModule(name))
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "nbm_model")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def ref(module: str):
    """Import a reference module, e.g. ref('nbm_model.nets.util.nets_utils')."""
    install()
    return importlib.import_module(module)


def default_args(device: str = "cpu"):
    """Namespace equal to the defaults of the reference training parser
    (train.py:21-168), extended as load_model does (run_detection.py:95-99)."""
    install()
    train = ref("nbm_model.train")
    parser = train.get_args_parser() if hasattr(train, "get_args_parser") else None
    if parser is None:
        raise RuntimeError("reference train.py has no get_args_parser()")
    ns = parser.parse_args([])
    ns.device = device
    ref("nbm_model.nets.util.nets_utils").setattr_others(ns)
    return ns
