"""ctypes binding of include/vfx_hostio.h (libvfx_hostio.so: FLAC reader / writer, plain C).

The reference gets FLAC through librosa.load / soundfile.write (voicefixer/base.py:47-49,
tools/wav.py:37; its test/test.py is FLAC in, FLAC out); neither is in this image."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvfx_hostio.so")


class FlacInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("total_samples", C.c_uint64),
                ("md5", C.c_uint8 * 16), ("audio_offset", C.c_uint64)]


SIGNATURES = {
    "vfx_flac_probe": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(FlacInfo)]),
    "vfx_flac_decode": (C.c_longlong, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "vfx_flac_encode_bound": (C.c_size_t, [C.c_size_t, C.c_int]),
    "vfx_flac_encode": (C.c_longlong, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "vfx_md5": (None, [C.c_char_p, C.c_size_t, C.c_void_p]),
    "vfx_hostio_last_error": (C.c_char_p, []),
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m voicefixer_b200.build` "
                               "(gcc, no GPU needed); FLAC files cannot be read or written without it")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _err(lib, what):
    return RuntimeError(f"{what}: {lib.vfx_hostio_last_error().decode()}")


def flac_info(data):
    lib = load()
    info = FlacInfo()
    if lib.vfx_flac_probe(data, len(data), C.byref(info)) != 0:
        raise _err(lib, "not a readable FLAC stream")
    return info


def flac_decode(data):
    """bytes -> (int32 [samples, channels] at the stream's bit depth, sample_rate, bits_per_sample).
    A stream whose STREAMINFO MD5 does not match the decoded audio is rejected."""
    lib = load()
    info = flac_info(data)
    cap = int(info.total_samples) if info.total_samples else max(1, (len(data) * 16) // max(1, info.channels))
    pcm = np.empty((cap, info.channels), dtype=np.int32)
    md5_state = C.c_int(0)
    n = lib.vfx_flac_decode(data, len(data), pcm.ctypes.data, cap, C.byref(md5_state))
    if n < 0:
        raise _err(lib, "FLAC decode failed")
    if md5_state.value < 0:
        raise RuntimeError("FLAC decode failed: audio does not match the MD5 signature in STREAMINFO")
    return pcm[:n], int(info.sample_rate), int(info.bits_per_sample)


def flac_encode_int16(frames, sample_rate):
    """int16 [samples, channels] -> bytes of a 16-bit FLAC stream (what soundfile writes for int16 input)."""
    lib = load()
    x = np.ascontiguousarray(frames, dtype=np.int32)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    cap = lib.vfx_flac_encode_bound(n, ch)
    out = np.empty(cap, dtype=np.uint8)
    nb = lib.vfx_flac_encode(x.ctypes.data, n, ch, int(sample_rate), out.ctypes.data, cap)
    if nb < 0:
        raise _err(lib, "FLAC encode failed")
    return out[:nb].tobytes()
