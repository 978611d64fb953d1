"""Host-side FLAC codec (include/vfx_hostio.h, libvfx_hostio.so) behind the file API.

The reference's acceptance test is FLAC in / FLAC out through librosa.load and soundfile.write
(test/test.py:48-57,85-89; voicefixer/base.py:47-49; tools/wav.py:37).  The decoder is pinned by
(a) a libFLAC-encoded excerpt of the reference's own test input with the PCM of its sibling .wav
(tests/golden/flac_libflac_excerpt.npz, made by tests/golden/make_flac_golden.py), (b) when
/root/reference is present, every .flac the reference ships, against the MD5 signature libFLAC
stored in STREAMINFO, and (c) hashlib for the MD5 itself.  The encoder is pinned by the decoder."""
import glob
import hashlib
import os
import re
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_UTT = "/root/reference/test/utterance"


def test_hostio_library_exports_every_declared_symbol():
    from voicefixer_b200 import _hostio
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vfx_hostio.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(vfx_[a-z_0-9]+)\s*\(", hdr))
    lib = _hostio.load()
    assert declared and all(hasattr(lib, n) for n in declared)
    assert declared == set(_hostio.SIGNATURES)


def test_hostio_header_is_valid_c99_and_struct_layout_matches(tmp_path):
    import ctypes, shutil, subprocess
    from voicefixer_b200 import _hostio
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "vfx_hostio.h"\nint main(void){printf("%zu %zu %zu",sizeof(vfx_flac_info),'
                   '__builtin_offsetof(vfx_flac_info,md5),__builtin_offsetof(vfx_flac_info,audio_offset));return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run([shutil.which("gcc"), "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    size, off_md5, off_audio = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert size == ctypes.sizeof(_hostio.FlacInfo)
    assert off_md5 == _hostio.FlacInfo.md5.offset and off_audio == _hostio.FlacInfo.audio_offset.offset


def test_md5_matches_hashlib():
    import ctypes
    from voicefixer_b200 import _hostio
    lib = _hostio.load()
    rng = np.random.default_rng(3)
    for n in (0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 1000, 70001):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        out = (ctypes.c_uint8 * 16)()
        lib.vfx_md5(data, n, out)
        assert bytes(out) == hashlib.md5(data).digest(), n


def test_decode_libflac_excerpt_golden():
    """LPC subframes + Rice partitions written by libFLAC, checked against the sibling WAV's PCM."""
    from voicefixer_b200 import _hostio
    g = np.load(os.path.join(ROOT, "tests", "golden", "flac_libflac_excerpt.npz"))
    data = g["flac"].tobytes()
    info = _hostio.flac_info(data)
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.total_samples) == (44100, 1, 16, g["pcm"].shape[0])
    assert bytes(info.md5) == hashlib.md5(g["pcm"].astype("<i2").tobytes()).digest()
    pcm, sr, bps = _hostio.flac_decode(data)
    assert sr == 44100 and bps == 16 and np.array_equal(pcm[:, 0], g["pcm"])


def test_decode_detects_corruption():
    from voicefixer_b200 import _hostio
    g = np.load(os.path.join(ROOT, "tests", "golden", "flac_libflac_excerpt.npz"))
    good = bytearray(g["flac"].tobytes())
    for pos, what in ((len(good) // 2, "CRC-16"), (42 + 2, "CRC-8"), (len(good) - 1, "CRC-16")):
        bad = bytearray(good)
        bad[pos] ^= 0x10
        with pytest.raises(RuntimeError, match=what):
            _hostio.flac_decode(bytes(bad))
    with pytest.raises(RuntimeError, match="stream ends after|truncated|bitstream ended|CRC"):
        _hostio.flac_decode(bytes(good[: len(good) - 700]))
    bad = bytearray(good)
    bad[30] ^= 0xFF                                               # inside the MD5 signature
    with pytest.raises(RuntimeError, match="MD5"):
        _hostio.flac_decode(bytes(bad))
    with pytest.raises(RuntimeError, match="fLaC"):
        _hostio.flac_decode(b"RIFF" + bytes(100))


SIGNALS = {
    "noise": lambda rng, n, ch: rng.integers(-32768, 32768, (n, ch)),
    "tone": lambda rng, n, ch: (12000 * np.sin(2 * np.pi * 220 * np.arange(n)[:, None] / 44100) + rng.normal(0, 20, (n, ch))).astype(np.int64),
    "silence": lambda rng, n, ch: np.zeros((n, ch), dtype=np.int64),
    "dc": lambda rng, n, ch: np.full((n, ch), -7, dtype=np.int64),
    "rails": lambda rng, n, ch: np.where(rng.random((n, ch)) < 0.5, -32768, 32767),
    "quiet": lambda rng, n, ch: rng.integers(-3, 4, (n, ch)),
}


@pytest.mark.parametrize("n", [0, 1, 4, 5, 255, 256, 257, 4095, 4096, 4097, 3 * 4096 + 17])
@pytest.mark.parametrize("ch", [1, 2])
def test_encode_decode_roundtrip_is_bit_exact(n, ch):
    """Empty, ragged last block, every subframe choice (constant / verbatim / fixed 0-4), stereo."""
    from voicefixer_b200 import _hostio
    rng = np.random.default_rng(n * 2 + ch)
    for name, make in SIGNALS.items():
        x = make(rng, n, ch).astype(np.int16)
        data = _hostio.flac_encode_int16(x, 44100 if name != "quiet" else 12345)
        info = _hostio.flac_info(data)
        assert info.total_samples == n and info.channels == ch and info.bits_per_sample == 16
        assert bytes(info.md5) == hashlib.md5(x.astype("<i2").tobytes()).digest()
        y, sr, bps = _hostio.flac_decode(data)
        assert y.shape == (n, ch) and np.array_equal(y, x), name
        if n >= 4096 and name in ("tone", "silence", "quiet"):
            assert len(data) < 0.6 * x.nbytes                      # it does compress


def test_long_stream_frame_numbers_use_multibyte_coding():
    """> 2048 frames: the UTF-8 style frame number takes 1, 2 and 3 bytes along the stream."""
    from voicefixer_b200 import _hostio
    x = (np.arange(4096 * 2100) % 251 - 125).astype(np.int16)
    y, _, _ = _hostio.flac_decode(_hostio.flac_encode_int16(x, 44100))
    assert np.array_equal(y[:, 0], x)


def test_wavio_flac_follows_the_file_extension(tmp_path):
    """save_wave -> .flac has tools/wav.py:9-37 semantics (x 2^15, int16 truncation); load_mono reads it back
    like librosa.load (float32 in [-1, 1), mono mix), and a .wav of the same frames loads identically."""
    from voicefixer_b200 import wavio
    rng = np.random.default_rng(0)
    x = (0.3 * rng.standard_normal((1, 30000))).clip(-1, 1).astype(np.float32)
    wavio.save_wave(x, str(tmp_path / "a.flac"), 44100)
    wavio.save_wave(x, str(tmp_path / "a.wav"), 44100)
    assert (tmp_path / "a.flac").read_bytes()[:4] == b"fLaC"
    yf, yw = wavio.load_mono(str(tmp_path / "a.flac")), wavio.load_mono(str(tmp_path / "a.wav"))
    assert yf.dtype == np.float32 and np.array_equal(yf, yw)
    assert np.array_equal((yf * 32768).astype(np.int16), (x[0] * 2 ** 15).astype(np.short))
    stereo = np.stack([x[0], -0.5 * x[0]], axis=1)                   # [samples, 2] -> two channels
    wavio.save_wave(stereo, str(tmp_path / "s.flac"), 22050)
    r = wavio.read_wave(str(tmp_path / "s.flac"), 22050)
    assert r.shape == (30000, 2)
    m = wavio.load_mono(str(tmp_path / "s.flac"), 44100)             # resampled mono mix, like librosa.load(sr=44100)
    assert m.shape == (60000,)


@pytest.mark.skipif(not os.path.isdir(REF_UTT), reason="reference checkout not present (GPU box)")
def test_decode_every_reference_flac_against_its_libflac_md5():
    """All FLAC files the reference ships (inputs, targets, outputs): header fields, frame CRCs and the MD5
    signature written by libFLAC; sample counts are the ones BASELINE.md quotes; original.flac == original.wav."""
    from voicefixer_b200 import _hostio, wavio
    files = sorted(glob.glob(os.path.join(REF_UTT, "*", "*.flac")))
    assert len(files) >= 6
    expect = {"original.flac": 132300, "p360_001_mic1.flac": 96076, "oracle.flac": 97902, "output_mode_0.flac": 132300,
              "output_mode_1.flac": 132096, "output_mode_2.flac": 132300}
    for f in files:
        data = open(f, "rb").read()
        info = _hostio.flac_info(data)
        assert any(info.md5), f
        pcm, sr, bps = _hostio.flac_decode(data)                     # raises on CRC / MD5 mismatch
        assert (sr, bps, pcm.shape) == (44100, 16, (expect[os.path.basename(f)], 1)), f
        assert hashlib.md5(pcm.astype("<i2").tobytes()).digest() == bytes(info.md5)
    with wave.open(os.path.join(REF_UTT, "original", "original.wav"), "rb") as w:
        ref = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    got = wavio.load_mono(os.path.join(REF_UTT, "original", "original.flac"))
    assert np.array_equal(got, ref.astype(np.float32) / 32768.0)
