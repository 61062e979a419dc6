"""wav decoding (audio_io.read_wav) against libsndfile's documented conventions -- what librosa.core.load(sr=None) hands
the reference (prepare_dataset.py:160-165) -- and the resampling stand-in for the reference's ffmpeg call (:166-182)."""
import struct

import numpy as np
import pytest

from birdsoundclassif_b200 import audio_io, synth


def _riff(fmt_code, channels, sr, bits, payload, extensible=False):
    align = channels * bits // 8
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, sr, sr * align, align, bits, 22, bits, 0, fmt_code) + guid_tail
    else:
        fmt = struct.pack("<HHIIHH", fmt_code, channels, sr, sr * align, align, bits)
    junk = b"LIST" + struct.pack("<I", 5) + b"hello" + b"\0"          # an odd-sized chunk before the data (padded)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_pcm16_is_returned_as_int16(tmp_path):
    pcm = synth.synth_pcm(0.3, 1)
    p = synth.write_wav(str(tmp_path / "a.wav"), pcm)
    x, sr = audio_io.read_wav(p)
    assert x.dtype == np.int16 and sr == 44100 and np.array_equal(x, pcm)
    st = np.stack([pcm, pcm[::-1]], axis=1)
    x2, _ = audio_io.read_wav(synth.write_wav(str(tmp_path / "s.wav"), st))
    assert x2.shape == st.shape and np.array_equal(x2, st)
    np.testing.assert_array_equal(audio_io.to_float_mono(x2), np.mean((st.astype(np.float32) / np.float32(32768)).T, axis=0))


@pytest.mark.parametrize("ext", [False, True])
def test_other_encodings_follow_libsndfile_scaling(tmp_path, ext):
    rng = np.random.default_rng(5)
    v24 = rng.integers(-2 ** 23, 2 ** 23, 300)
    b24 = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in v24)
    cases = {
        "u8": (1, 8, rng.integers(0, 256, 300, dtype=np.uint8).tobytes(), lambda raw: (np.frombuffer(raw, np.uint8).astype(np.float64) - 128) / 128),
        "s24": (1, 24, b24, lambda raw: v24 / 2.0 ** 23),
        "s32": (1, 32, rng.integers(-2 ** 31, 2 ** 31, 300, dtype=np.int64).astype("<i4").tobytes(), lambda raw: np.frombuffer(raw, "<i4") / 2.0 ** 31),
        "f32": (3, 32, rng.uniform(-1, 1, 300).astype("<f4").tobytes(), lambda raw: np.frombuffer(raw, "<f4").astype(np.float64)),
        "f64": (3, 64, rng.uniform(-1, 1, 300).astype("<f8").tobytes(), lambda raw: np.frombuffer(raw, "<f8")),
    }
    for name, (code, bits, raw, want) in cases.items():
        f = tmp_path / f"{name}.wav"
        f.write_bytes(_riff(code, 1, 22050, bits, raw, extensible=ext))
        x, sr = audio_io.read_wav(str(f))
        assert sr == 22050 and x.dtype == np.float32 and x.shape == (300,), name
        np.testing.assert_array_equal(x, want(raw).astype(np.float32), err_msg=name)


def test_bad_files_raise(tmp_path):
    (tmp_path / "x.wav").write_bytes(b"RIFFxxxxWAVEdata")
    with pytest.raises(Exception):
        audio_io.read_wav(str(tmp_path / "x.wav"))
    (tmp_path / "adpcm.wav").write_bytes(_riff(2, 1, 44100, 4, bytes(64)))
    with pytest.raises(audio_io.WavFormatError):
        audio_io.read_wav(str(tmp_path / "adpcm.wav"))


def test_resample_48k_to_44k1():
    """A 3 kHz tone at 48 kHz comes out as a 3 kHz tone at 44.1 kHz: length ratio 147/160, same amplitude, PCM16."""
    sr, f0, secs = 48000, 3000.0, 0.5
    t = np.arange(int(sr * secs)) / sr
    x = np.round(0.5 * 32767 * np.sin(2 * np.pi * f0 * t)).astype(np.int16)
    y = audio_io.resample_pcm16(x, sr, 44100)
    assert y.dtype == np.int16 and abs(len(y) - len(x) * 147 / 160) <= 1
    t2 = np.arange(len(y)) / 44100.0
    ref = 0.5 * 32767 * np.sin(2 * np.pi * f0 * t2)
    core = slice(500, len(y) - 500)                                   # away from the filter's edge transients
    assert np.abs(y[core] - ref[core]).max() <= 0.002 * 32767
    # stereo input is mixed down first (ffmpeg -ac 1)
    y2 = audio_io.resample_pcm16(np.stack([x, x], axis=1), sr, 44100)
    assert np.abs(y2.astype(np.int32) - y.astype(np.int32)).max() <= 1
