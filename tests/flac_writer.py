"""Independent, deliberately simple FLAC *writer* used only by tests/test_hostio.py to reach decoder paths that no
available file exercises (the reference's fixtures are 16-bit mono; the in-tree encoder writes independent channels).
Pure Python integers, written from the format definition: frame header fields, subframe types, residual coding,
CRC-8 (poly 0x07) / CRC-16 (poly 0x8005), STREAMINFO.  Forward transforms (left/right -> mid/side, prediction ->
residual) are the definitions; the decoder under test has to invert them."""
import hashlib


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, nbits):
        if nbits:
            self.v = (self.v << nbits) | (value & ((1 << nbits) - 1))
            self.n += nbits

    def signed(self, value, nbits):
        assert -(1 << (nbits - 1)) <= value < (1 << (nbits - 1)), (value, nbits)
        self.put(value, nbits)

    def unary(self, q):                      # q zeros, then a one
        self.put(1, q + 1)

    def align(self):
        self.put(0, -self.n % 8)

    def bytes(self):
        assert self.n % 8 == 0
        return self.v.to_bytes(self.n // 8, "big")


def _crc(data, poly, width):
    c, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for b in data:
        c ^= b << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def utf8_number(v):
    if v < 0x80:
        return bytes([v])
    n = 1
    while v >> (5 * n + 6):
        n += 1
    out = [((0xFF << (7 - n)) & 0xFF) | (v >> (6 * n))]
    out += [0x80 | ((v >> (6 * i)) & 0x3F) for i in range(n - 1, -1, -1)]
    return bytes(out)


def rice(bits, residual, k):
    for e in residual:
        u = (e << 1) ^ (e >> 63) if e >= 0 else ((-e) << 1) - 1
        bits.unary(u >> k)
        bits.put(u, k)


def sub_verbatim(bits, x, bps, wasted=0):
    bits.put(0, 1); bits.put(1, 6)
    if wasted:
        bits.put(1, 1); bits.unary(wasted - 1)
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
    else:
        bits.put(0, 1)
    for v in x:
        bits.signed(v, bps - wasted)


def sub_constant(bits, x, bps):
    assert len(set(x)) == 1
    bits.put(0, 8)
    bits.signed(x[0], bps)


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def sub_fixed(bits, x, bps, order, part_order, ks, method=0):
    """ks: one Rice parameter per partition, or ("esc", nbits) for an escape partition of raw nbits-bit residuals."""
    bits.put(0, 1); bits.put(8 + order, 6); bits.put(0, 1)
    for v in x[:order]:
        bits.signed(v, bps)
    res = [x[i] - sum(c * x[i - 1 - j] for j, c in enumerate(FIXED[order])) for i in range(order, len(x))]
    _residual(bits, res, len(x), order, part_order, ks, method)


def sub_lpc(bits, x, bps, coefs, precision, shift, part_order, ks, method=0):
    order = len(coefs)
    bits.put(0, 1); bits.put(31 + order, 6); bits.put(0, 1)
    for v in x[:order]:
        bits.signed(v, bps)
    bits.put(precision - 1, 4); bits.signed(shift, 5)
    for c in coefs:
        bits.signed(c, precision)
    res = [x[i] - (sum(c * x[i - 1 - j] for j, c in enumerate(coefs)) >> shift) for i in range(order, len(x))]
    _residual(bits, res, len(x), order, part_order, ks, method)


def _residual(bits, res, blocksize, order, part_order, ks, method):
    bits.put(method, 2); bits.put(part_order, 4)
    pbits, pos = (5 if method else 4), 0
    for p in range(1 << part_order):
        count = (blocksize >> part_order) - (order if p == 0 else 0)
        chunk = res[pos:pos + count]; pos += count
        if isinstance(ks[p], tuple):
            bits.put((1 << pbits) - 1, pbits); bits.put(ks[p][1], 5)
            for e in chunk:
                if ks[p][1]:
                    bits.signed(e, ks[p][1])
                else:
                    assert e == 0
        else:
            bits.put(ks[p], pbits); rice(bits, chunk, ks[p])
    assert pos == len(res)


BS_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13}
SS_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def frame(number, blocksize, ch_code, bps, subframes, sr_code=0, sr_extra=b"", variable=False, ss_from_info=False,
          bs_explicit=None):
    """subframes: callables(bits) in channel order.  number: frame number (fixed) or first sample number (variable)."""
    hdr = Bits()
    hdr.put(0xFFF8 | (1 if variable else 0), 16)
    if bs_explicit == 8:
        bs_code = 6
    elif bs_explicit == 16:
        bs_code = 7
    else:
        bs_code = BS_CODES[blocksize]
    hdr.put(bs_code, 4); hdr.put(sr_code, 4); hdr.put(ch_code, 4)
    hdr.put(0 if ss_from_info else SS_CODES[bps], 3); hdr.put(0, 1)
    head = hdr.bytes() + utf8_number(number)
    if bs_code == 6:
        head += bytes([blocksize - 1])
    elif bs_code == 7:
        head += (blocksize - 1).to_bytes(2, "big")
    head += sr_extra
    head += bytes([_crc(head, 0x07, 8)])
    body = Bits()
    for s in subframes:
        s(body)
    body.align()
    data = head + body.bytes()
    return data + _crc(data, 0x8005, 16).to_bytes(2, "big")


def stream(frames, sample_rate, channels, bps, total, pcm_interleaved=None, min_bs=16, max_bs=65535):
    info = Bits()
    info.put(min_bs, 16); info.put(max_bs, 16); info.put(0, 24); info.put(0, 24)
    info.put(sample_rate, 20); info.put(channels - 1, 3); info.put(bps - 1, 5); info.put(total, 36)
    if pcm_interleaved is None:
        md5 = bytes(16)
    else:
        nb = (bps + 7) // 8
        md5 = hashlib.md5(b"".join(int(v).to_bytes(nb, "little", signed=True) for v in pcm_interleaved)).digest()
    return b"fLaC" + bytes([0x80, 0, 0, 34]) + info.bytes() + md5 + b"".join(frames)
