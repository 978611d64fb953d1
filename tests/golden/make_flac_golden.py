"""Builds tests/golden/flac_libflac_excerpt.npz: a short libFLAC-encoded stream (LPC subframes) that
travels to the GPU box, cut from the reference's own test input, plus the PCM it must decode to.

    python tests/golden/make_flac_golden.py        # needs /root/reference (this container only)

Source: test/utterance/original/original.flac (the input of the reference's test/test.py:48-57) and its
sibling original.wav, which holds the same 132300 samples as plain PCM -- an independent decode.
FLAC frames are self-contained, so STREAMINFO + the first NFRAMES frames are a valid stream once
total_samples and the MD5 signature in STREAMINFO are rewritten for the excerpt.
"""
import hashlib
import os
import wave

import numpy as np

REF = "/root/reference/test/utterance/original"
NFRAMES = 2
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flac_libflac_excerpt.npz")


def crc8(b):
    c = 0
    for x in b:
        c ^= x
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def frame_offset(data, start, number):
    """Offset of the fixed-blocksize frame with the given frame number (< 128): sync, number byte, header CRC-8."""
    i = start
    while True:
        i = data.index(b"\xff\xf8", i)
        bs_code, sr_code = data[i + 2] >> 4, data[i + 2] & 15
        n = 5 + {6: 1, 7: 2}.get(bs_code, 0) + {12: 1, 13: 2, 14: 2}.get(sr_code, 0)
        if data[i + 4] == number and crc8(data[i:i + n]) == data[i + n]:
            return i
        i += 1


def main():
    data = open(os.path.join(REF, "original.flac"), "rb").read()
    assert data[:4] == b"fLaC" and data[4] & 0x7F == 0 and data[5:8] == b"\x00\x00\x22"
    off, last = 4, False
    while not last:                                  # walk the metadata chain to the first frame
        last, ln = bool(data[off] >> 7), int.from_bytes(data[off + 1:off + 4], "big")
        off += 4 + ln
    first, end = frame_offset(data, off, 0), frame_offset(data, off, NFRAMES)
    assert first == off
    blocksize = int.from_bytes(data[8:10], "big")
    nsamp = NFRAMES * blocksize
    with wave.open(os.path.join(REF, "original.wav"), "rb") as w:
        pcm = np.frombuffer(w.readframes(nsamp), dtype="<i2").copy()
    info = bytearray(data[8:42])
    info[13] = (info[13] & 0xF0) | ((nsamp >> 32) & 0x0F)
    info[14:18] = (nsamp & 0xFFFFFFFF).to_bytes(4, "big")
    info[18:34] = hashlib.md5(pcm.astype("<i2").tobytes()).digest()
    stream = b"fLaC" + bytes([0x80, 0, 0, 34]) + bytes(info) + data[first:end]
    np.savez_compressed(OUT, flac=np.frombuffer(stream, dtype=np.uint8), pcm=pcm,
                        source="test/utterance/original/original.flac frames 0..%d" % (NFRAMES - 1))
    print(OUT, len(stream), "bytes,", nsamp, "samples")


if __name__ == "__main__":
    main()
