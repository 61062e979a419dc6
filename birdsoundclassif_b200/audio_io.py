"""wav decoding and sample-rate conversion for ``File_Processor.load`` (prepare_dataset.py:160-184).

The reference decodes with ``librosa.core.load(path, sr=None)`` (libsndfile: any PCM width or float, scaled to
float32 in [-1, 1), channels averaged) and, when the rate is not 44.1 kHz, shells out to ffmpeg
(``-ac 1 -acodec pcm_s16le -ar 44100`` for wav files, prepare_dataset.py:174-177) and decodes the result.
Neither libsndfile nor ffmpeg is a dependency here:

* ``read_wav`` parses RIFF/WAVE itself: PCM 8/16/24/32-bit, IEEE float 32/64, plain or WAVE_FORMAT_EXTENSIBLE.
  PCM16 comes back as int16 (the tensor-core front-end's input, uploaded as is); every other format as float32
  with libsndfile's scaling (u8: (x-128)/128, s24: x/2^23, s32: x/2^31), which the front-end's float path takes.
* ``resample_pcm16`` is the stand-in for the ffmpeg call: polyphase FIR resampling (scipy.signal.resample_poly,
  Kaiser window) of the channel mean to 44.1 kHz, rounded to PCM16 like ``-acodec pcm_s16le``.  It is NOT ffmpeg's
  resampler: spectrograms of resampled files agree with upstream's only as far as two good resamplers agree
  (upstream's own result depends on the ffmpeg build).  44.1 kHz files -- the reference's native rate -- never
  take this path.
"""
from __future__ import annotations

import struct
from math import gcd

import numpy as np

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 1, 3, 0xFFFE


class WavFormatError(ValueError):
    pass


def parse_wav_header(f):
    """-> dict(format, channels, sample_rate, bits, block_align, data_offset, data_bytes) from an open binary file."""
    head = f.read(12)
    if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
        raise WavFormatError("not a RIFF/WAVE file")
    fmt = None
    while True:
        h = f.read(8)
        if len(h) < 8:
            raise WavFormatError("no data chunk")
        tag, size = h[:4], struct.unpack("<I", h[4:])[0]
        if tag == b"fmt ":
            body = f.read(size + (size & 1))
            if size < 16:
                raise WavFormatError("short fmt chunk")
            code, ch, sr, _, align, bits = struct.unpack("<HHIIHH", body[:16])
            if code == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                code = struct.unpack("<H", body[24:26])[0]          # first two bytes of the sub-format GUID
            fmt = dict(format=code, channels=ch, sample_rate=sr, bits=bits, block_align=align)
        elif tag == b"data":
            if fmt is None:
                raise WavFormatError("data chunk before fmt chunk")
            fmt.update(data_offset=f.tell(), data_bytes=size)
            return fmt
        else:
            f.seek(size + (size & 1), 1)


def read_wav(path: str):
    """-> (samples, sample_rate): int16 [n] / [n, ch] for PCM16, float32 otherwise (libsndfile scaling).  A truncated
    file yields the whole frames it holds."""
    with open(path, "rb") as f:
        h = parse_wav_header(f)
        raw = f.read(h["data_bytes"])
    ch, bits, code = h["channels"], h["bits"], h["format"]
    if ch < 1:
        raise WavFormatError("no channels")
    bps = bits // 8
    raw = raw[:len(raw) - len(raw) % (bps * ch)] if bps else b""
    if code == WAVE_FORMAT_PCM and bits == 16:
        x = np.frombuffer(raw, dtype="<i2")
    elif code == WAVE_FORMAT_PCM and bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif code == WAVE_FORMAT_PCM and bits == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = (v.astype(np.float64) / 8388608.0).astype(np.float32)
    elif code == WAVE_FORMAT_PCM and bits == 32:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif code == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
        x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    elif code == WAVE_FORMAT_IEEE_FLOAT and bits == 64:
        x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    else:
        raise WavFormatError(f"unsupported wav encoding (format {code}, {bits} bits)")
    return (x if ch == 1 else x.reshape(-1, ch)), h["sample_rate"]


def to_float_mono(x: np.ndarray) -> np.ndarray:
    """float32 mono as librosa.load hands it on (int16 / 32768, channel mean)."""
    y = x.astype(np.float32) / np.float32(32768.0) if x.dtype == np.int16 else x.astype(np.float32)
    return y if y.ndim == 1 else np.mean(y.T, axis=0).astype(np.float32)


def resample_pcm16(x: np.ndarray, sr: int, target: int = 44100) -> np.ndarray:
    """Mono mix, polyphase resampling to `target` Hz, PCM16 rounding (stand-in for the reference's ffmpeg call)."""
    from scipy.signal import resample_poly
    y = to_float_mono(x).astype(np.float64)
    g = gcd(int(sr), int(target))
    z = resample_poly(y, target // g, sr // g, window=("kaiser", 8.6)) if len(y) else y
    return np.clip(np.rint(z * 32768.0), -32768, 32767).astype(np.int16)
