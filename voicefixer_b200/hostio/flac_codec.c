/* flac_codec.c -- host-side FLAC reader / writer behind include/vfx_hostio.h (plain C11, no CUDA).
 *
 * Why it exists: the reference's file contract is "whatever librosa.load reads / soundfile.write
 * writes" (voicefixer/base.py:47-49, tools/wav.py:37) and its acceptance test is FLAC in, FLAC out
 * (test/test.py:48-57,85-89).  This file restates the published FLAC format (frame header, subframe
 * types, residual coding, CRCs, STREAMINFO + MD5 signature) so the mirrored API accepts the same
 * files.  It is host I/O only: nothing here is on the GPU hot path.
 */
#include "../../include/vfx_hostio.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

const char* vfx_hostio_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------ CRC tables */
static uint8_t crc8_tab[256];
static uint16_t crc16_tab[256];

__attribute__((constructor)) static void init_crc(void) {
  for (int i = 0; i < 256; ++i) {
    uint8_t c8 = (uint8_t)i;
    uint16_t c16 = (uint16_t)(i << 8);
    for (int b = 0; b < 8; ++b) {
      c8 = (uint8_t)((c8 & 0x80) ? (c8 << 1) ^ 0x07 : c8 << 1);          /* x^8 + x^2 + x + 1   */
      c16 = (uint16_t)((c16 & 0x8000) ? (c16 << 1) ^ 0x8005 : c16 << 1); /* x^16 + x^15 + x^2 + 1 */
    }
    crc8_tab[i] = c8;
    crc16_tab[i] = c16;
  }
}

static uint8_t crc8(const uint8_t* p, size_t n) {
  uint8_t c = 0;
  while (n--) c = crc8_tab[c ^ *p++];
  return c;
}

static uint16_t crc16(const uint8_t* p, size_t n) {
  uint16_t c = 0;
  while (n--) c = (uint16_t)((c << 8) ^ crc16_tab[(c >> 8) ^ *p++]);
  return c;
}

/* ------------------------------------------------------------------------------------------- MD5 */
typedef struct {
  uint32_t h[4];
  uint64_t nbytes;
  uint8_t buf[64];
  size_t fill;
} Md5;

static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t md5_s[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_init(Md5* m) {
  m->h[0] = 0x67452301; m->h[1] = 0xefcdab89; m->h[2] = 0x98badcfe; m->h[3] = 0x10325476;
  m->nbytes = 0;
  m->fill = 0;
}

static void md5_block(Md5* m, const uint8_t* p) {
  uint32_t w[16];
  for (int i = 0; i < 16; ++i)
    w[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
  uint32_t a = m->h[0], b = m->h[1], c = m->h[2], d = m->h[3];
  for (int i = 0; i < 64; ++i) {
    uint32_t f;
    int g;
    if (i < 16)      { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d;          g = (3 * i + 5) & 15; }
    else             { f = c ^ (b | ~d);       g = (7 * i) & 15; }
    uint32_t t = a + f + md5_k[i] + w[g];
    a = d; d = c; c = b;
    b = b + ((t << md5_s[i]) | (t >> (32 - md5_s[i])));
  }
  m->h[0] += a; m->h[1] += b; m->h[2] += c; m->h[3] += d;
}

static void md5_update(Md5* m, const uint8_t* p, size_t n) {
  m->nbytes += n;
  if (m->fill) {
    size_t take = 64 - m->fill < n ? 64 - m->fill : n;
    memcpy(m->buf + m->fill, p, take);
    m->fill += take; p += take; n -= take;
    if (m->fill < 64) return;
    md5_block(m, m->buf);
    m->fill = 0;
  }
  for (; n >= 64; p += 64, n -= 64) md5_block(m, p);
  if (n) { memcpy(m->buf, p, n); m->fill = n; }
}

static void md5_final(Md5* m, uint8_t out[16]) {
  uint64_t bits = m->nbytes * 8;
  uint8_t pad[72] = {0x80};
  size_t padn = (m->fill < 56 ? 56 : 120) - m->fill;
  for (int i = 0; i < 8; ++i) pad[padn + i] = (uint8_t)(bits >> (8 * i));
  md5_update(m, pad, padn + 8);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(m->h[i] >> (8 * j));
}

void vfx_md5(const uint8_t* data, size_t nbytes, uint8_t digest[16]) {
  Md5 m;
  md5_init(&m);
  md5_update(&m, data, nbytes);
  md5_final(&m, digest);
}

/* Feeds interleaved samples to the signature: ceil(bps / 8) little-endian bytes per sample. */
static void md5_samples(Md5* m, const int32_t* pcm, size_t count, int bytes_per_sample) {
  uint8_t tmp[4096];
  size_t fill = 0;
  for (size_t i = 0; i < count; ++i) {
    uint32_t v = (uint32_t)pcm[i];
    for (int b = 0; b < bytes_per_sample; ++b) tmp[fill++] = (uint8_t)(v >> (8 * b));
    if (fill > sizeof tmp - 4) { md5_update(m, tmp, fill); fill = 0; }
  }
  if (fill) md5_update(m, tmp, fill);
}

/* ------------------------------------------------------------------------------------ bit reader */
typedef struct {
  const uint8_t* d;
  size_t n, pos;   /* pos = next byte to load into acc */
  uint64_t acc;    /* unread bits, left-aligned        */
  int cnt;         /* number of valid bits in acc      */
  int err;
} BitReader;

static inline void br_fill(BitReader* b) {
  while (b->cnt <= 56 && b->pos < b->n) {
    b->acc |= (uint64_t)b->d[b->pos++] << (56 - b->cnt);
    b->cnt += 8;
  }
}

static inline uint32_t br_bits(BitReader* b, int n) { /* 0 <= n <= 32 */
  if (n == 0) return 0;
  br_fill(b);
  if (b->cnt < n) { b->err = 1; b->cnt = 0; b->acc = 0; return 0; }
  uint32_t v = (uint32_t)(b->acc >> (64 - n));
  b->acc <<= n;
  b->cnt -= n;
  return v;
}

static inline int64_t br_signed(BitReader* b, int n) { /* 1 <= n <= 33 */
  uint64_t v;
  if (n > 32) v = ((uint64_t)br_bits(b, n - 32) << 32) | br_bits(b, 32);
  else v = br_bits(b, n);
  const uint64_t sign = 1ull << (n - 1);
  return (int64_t)((v ^ sign) - sign);
}

static inline uint32_t br_unary(BitReader* b) { /* number of 0 bits before the next 1 */
  uint32_t q = 0;
  for (;;) {
    br_fill(b);
    if (b->cnt == 0) { b->err = 1; return q; }
    if (b->acc == 0) { q += (uint32_t)b->cnt; b->cnt = 0; continue; }
    int z = __builtin_clzll(b->acc);
    q += (uint32_t)z;
    b->acc <<= z;        /* z <= 63 */
    b->acc <<= 1;
    b->cnt -= z + 1;
    return q;
  }
}

static inline size_t br_byte_pos(const BitReader* b) { return b->pos - (size_t)(b->cnt / 8); }

static inline void br_align(BitReader* b) {
  int drop = b->cnt & 7;
  b->acc <<= drop;
  b->cnt -= drop;
}

/* ------------------------------------------------------------------------------------- metadata */
int vfx_flac_probe(const uint8_t* data, size_t nbytes, vfx_flac_info* info) {
  g_err[0] = 0;
  if (!data || !info) return fail(VFX_IO_ERR_ARGUMENT, "flac_probe: null argument");
  size_t off = 0;
  if (nbytes >= 10 && memcmp(data, "ID3", 3) == 0) { /* skip an ID3v2 tag if someone prepended one */
    size_t sz = ((size_t)(data[6] & 0x7f) << 21) | ((size_t)(data[7] & 0x7f) << 14) | ((size_t)(data[8] & 0x7f) << 7) | (data[9] & 0x7f);
    off = 10 + sz;
  }
  if (nbytes < off + 4 + 4 + 34 || memcmp(data + off, "fLaC", 4) != 0)
    return fail(VFX_IO_ERR_FORMAT, "flac_probe: missing fLaC marker");
  off += 4;
  int have_info = 0;
  for (;;) {
    if (off + 4 > nbytes) return fail(VFX_IO_ERR_FORMAT, "flac_probe: truncated metadata chain");
    const int last = data[off] >> 7, type = data[off] & 0x7f;
    const size_t len = ((size_t)data[off + 1] << 16) | ((size_t)data[off + 2] << 8) | data[off + 3];
    off += 4;
    if (off + len > nbytes) return fail(VFX_IO_ERR_FORMAT, "flac_probe: metadata block overruns the file");
    if (type == 0) {
      if (len != 34 || have_info) return fail(VFX_IO_ERR_FORMAT, "flac_probe: bad STREAMINFO block");
      const uint8_t* p = data + off;
      info->min_blocksize = (uint32_t)p[0] << 8 | p[1];
      info->max_blocksize = (uint32_t)p[2] << 8 | p[3];
      info->sample_rate = (uint32_t)p[10] << 12 | (uint32_t)p[11] << 4 | p[12] >> 4;
      info->channels = ((p[12] >> 1) & 7u) + 1;
      info->bits_per_sample = (((uint32_t)p[12] & 1u) << 4 | p[13] >> 4) + 1;
      info->total_samples = ((uint64_t)(p[13] & 0x0f) << 32) | (uint64_t)p[14] << 24 | (uint64_t)p[15] << 16 | (uint64_t)p[16] << 8 | p[17];
      memcpy(info->md5, p + 18, 16);
      have_info = 1;
    } else if (type == 127) {
      return fail(VFX_IO_ERR_FORMAT, "flac_probe: invalid metadata block type 127");
    }
    off += len;
    if (last) break;
  }
  if (!have_info) return fail(VFX_IO_ERR_FORMAT, "flac_probe: no STREAMINFO block");
  if (info->sample_rate == 0 || info->bits_per_sample < 4)
    return fail(VFX_IO_ERR_FORMAT, "flac_probe: invalid STREAMINFO (rate %u, %u bits)", info->sample_rate, info->bits_per_sample);
  info->audio_offset = off;
  return VFX_IO_OK;
}

/* -------------------------------------------------------------------------------------- decoder */
static int decode_residual(BitReader* br, int64_t* x, int blocksize, int order) {
  const int method = (int)br_bits(br, 2);
  if (method > 1) return fail(VFX_IO_ERR_FORMAT, "flac: reserved residual coding method %d", method);
  const int pbits = method ? 5 : 4, escape = method ? 31 : 15;
  const int po = (int)br_bits(br, 4);
  const int parts = 1 << po;
  if ((blocksize >> po) << po != blocksize && po > 0) return fail(VFX_IO_ERR_FORMAT, "flac: blocksize %d not divisible by 2^%d", blocksize, po);
  if ((blocksize >> po) < order) return fail(VFX_IO_ERR_FORMAT, "flac: partition shorter than predictor order");
  int i = order;
  for (int p = 0; p < parts; ++p) {
    const int count = (blocksize >> po) - (p == 0 ? order : 0);
    const int k = (int)br_bits(br, pbits);
    if (k == escape) {
      const int raw = (int)br_bits(br, 5);
      for (int j = 0; j < count; ++j) x[i++] = raw ? br_signed(br, raw) : 0;
    } else {
      for (int j = 0; j < count; ++j) {
        const uint32_t q = br_unary(br);
        const uint64_t u = ((uint64_t)q << k) | br_bits(br, k);
        x[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
      }
    }
    if (br->err) return fail(VFX_IO_ERR_FORMAT, "flac: bitstream ended inside a residual partition");
  }
  return VFX_IO_OK;
}

#define U(v) ((uint64_t)(v))

static int decode_subframe(BitReader* br, int64_t* x, int blocksize, int bps) {
  if (br_bits(br, 1)) return fail(VFX_IO_ERR_FORMAT, "flac: subframe padding bit set");
  const int type = (int)br_bits(br, 6);
  int wasted = 0;
  if (br_bits(br, 1)) wasted = (int)br_unary(br) + 1;
  if (wasted >= bps) return fail(VFX_IO_ERR_FORMAT, "flac: %d wasted bits in a %d-bit subframe", wasted, bps);
  bps -= wasted;
  int rc = VFX_IO_OK;
  if (type == 0) {                                  /* CONSTANT */
    const int64_t v = br_signed(br, bps);
    for (int i = 0; i < blocksize; ++i) x[i] = v;
  } else if (type == 1) {                           /* VERBATIM */
    for (int i = 0; i < blocksize; ++i) x[i] = br_signed(br, bps);
  } else if (type >= 8 && type <= 12) {             /* FIXED, order = type - 8 */
    const int order = type - 8;
    if (order > blocksize) return fail(VFX_IO_ERR_FORMAT, "flac: fixed order %d > blocksize %d", order, blocksize);
    for (int i = 0; i < order; ++i) x[i] = br_signed(br, bps);
    if ((rc = decode_residual(br, x, blocksize, order)) != VFX_IO_OK) return rc;
    /* the recursions run in unsigned arithmetic: a corrupt (but CRC-valid) stream may drive them past 2^63, which
     * must wrap, not be undefined; for a valid stream every intermediate fits and the result is identical */
    switch (order) {
      case 1: for (int i = 1; i < blocksize; ++i) x[i] = (int64_t)(U(x[i]) + U(x[i - 1])); break;
      case 2: for (int i = 2; i < blocksize; ++i) x[i] = (int64_t)(U(x[i]) + 2 * U(x[i - 1]) - U(x[i - 2])); break;
      case 3: for (int i = 3; i < blocksize; ++i) x[i] = (int64_t)(U(x[i]) + 3 * U(x[i - 1]) - 3 * U(x[i - 2]) + U(x[i - 3])); break;
      case 4: for (int i = 4; i < blocksize; ++i) x[i] = (int64_t)(U(x[i]) + 4 * U(x[i - 1]) - 6 * U(x[i - 2]) + 4 * U(x[i - 3]) - U(x[i - 4])); break;
      default: break;
    }
  } else if (type >= 32) {                          /* LPC, order = type - 31 */
    const int order = type - 31;
    if (order > blocksize) return fail(VFX_IO_ERR_FORMAT, "flac: LPC order %d > blocksize %d", order, blocksize);
    for (int i = 0; i < order; ++i) x[i] = br_signed(br, bps);
    const int prec = (int)br_bits(br, 4) + 1;
    if (prec == 16) return fail(VFX_IO_ERR_FORMAT, "flac: invalid LPC precision");
    const int shift = (int)br_signed(br, 5);
    if (shift < 0) return fail(VFX_IO_ERR_FORMAT, "flac: negative LPC shift");
    int64_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = br_signed(br, prec);
    if ((rc = decode_residual(br, x, blocksize, order)) != VFX_IO_OK) return rc;
    for (int i = order; i < blocksize; ++i) {
      uint64_t acc = 0;
      for (int j = 0; j < order; ++j) acc += U(coef[j]) * U(x[i - 1 - j]);
      x[i] = (int64_t)(U(x[i]) + U((int64_t)acc >> shift));
    }
  } else {
    return fail(VFX_IO_ERR_FORMAT, "flac: reserved subframe type %d", type);
  }
  if (br->err) return fail(VFX_IO_ERR_FORMAT, "flac: bitstream ended inside a subframe");
  if (wasted)
    for (int i = 0; i < blocksize; ++i) x[i] = (int64_t)((uint64_t)x[i] << wasted);
  return VFX_IO_OK;
}

static const uint32_t k_rates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
static const int k_depths[8] = {0, 8, 12, -1, 16, 20, 24, 32};

long long vfx_flac_decode(const uint8_t* data, size_t nbytes, int32_t* pcm, size_t capacity_frames, int* md5_state) {
  vfx_flac_info info;
  int rc = vfx_flac_probe(data, nbytes, &info);
  if (rc != VFX_IO_OK) return rc;
  if (!pcm && capacity_frames) return fail(VFX_IO_ERR_ARGUMENT, "flac_decode: null output buffer");
  if (info.bits_per_sample > 32) return fail(VFX_IO_ERR_UNSUPPORTED, "flac_decode: %u bits per sample", info.bits_per_sample);
  const int nch = (int)info.channels;
  int64_t* sub = (int64_t*)malloc(sizeof(int64_t) * 65536u * (size_t)nch);
  if (!sub) return fail(VFX_IO_ERR_CAPACITY, "flac_decode: out of memory");
  Md5 md5;
  md5_init(&md5);
  const int md5_bytes = ((int)info.bits_per_sample + 7) / 8;
  size_t off = (size_t)info.audio_offset;
  uint64_t done = 0;
  long long result = 0;
  while (off + 2 <= nbytes) {
    if (info.total_samples && done >= info.total_samples) break;
    if (data[off] != 0xFF || (data[off + 1] & 0xFE) != 0xF8) {
      if (info.total_samples == 0 || done == info.total_samples) break;  /* trailing tag / padding */
      result = fail(VFX_IO_ERR_FORMAT, "flac_decode: lost frame sync at byte %zu", off);
      goto out;
    }
    BitReader br = {data + off, nbytes - off, 0, 0, 0, 0};
    br_bits(&br, 16);
    const int bs_code = (int)br_bits(&br, 4), sr_code = (int)br_bits(&br, 4);
    const int ch_code = (int)br_bits(&br, 4), ss_code = (int)br_bits(&br, 3);
    if (br_bits(&br, 1)) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: reserved header bit set"); goto out; }
    { /* UTF-8 style frame / sample number: the value is not needed, only its length */
      const uint32_t first = br_bits(&br, 8);
      int extra = 0;
      if (first & 0x80) {
        int ones = 0;
        while (ones < 8 && (first & (0x80u >> ones))) ++ones;
        if (ones < 2 || ones > 7) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: bad frame number coding"); goto out; }
        extra = ones - 1;
      }
      for (int i = 0; i < extra; ++i)
        if ((br_bits(&br, 8) & 0xC0) != 0x80) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: bad frame number continuation"); goto out; }
    }
    int blocksize;
    if (bs_code == 0) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: reserved blocksize code"); goto out; }
    else if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br_bits(&br, 8) + 1;
    else if (bs_code == 7) blocksize = (int)br_bits(&br, 16) + 1;
    else blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) br_bits(&br, 8);
    else if (sr_code == 13 || sr_code == 14) br_bits(&br, 16);
    else if (sr_code == 15) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: invalid sample rate code"); goto out; }
    int bps = ss_code == 0 ? (int)info.bits_per_sample : k_depths[ss_code];
    if (bps < 0) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: reserved sample size code"); goto out; }
    if (bps != (int)info.bits_per_sample) { result = fail(VFX_IO_ERR_UNSUPPORTED, "flac_decode: bit depth changes mid-stream"); goto out; }
    const size_t hdr_len = br_byte_pos(&br);
    const uint8_t hdr_crc = (uint8_t)br_bits(&br, 8);
    if (br.err) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: truncated frame header"); goto out; }
    if (crc8(data + off, hdr_len) != hdr_crc) { result = fail(VFX_IO_ERR_CRC, "flac_decode: header CRC-8 mismatch at byte %zu", off); goto out; }
    int frame_ch;
    if (ch_code < 8) frame_ch = ch_code + 1;
    else if (ch_code <= 10) frame_ch = 2;
    else { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: reserved channel assignment"); goto out; }
    if (frame_ch != nch) { result = fail(VFX_IO_ERR_UNSUPPORTED, "flac_decode: channel count changes mid-stream"); goto out; }
    for (int c = 0; c < nch; ++c) {
      const int side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      rc = decode_subframe(&br, sub + (size_t)c * 65536u, blocksize, bps + side);
      if (rc != VFX_IO_OK) { result = rc; goto out; }
    }
    br_align(&br);
    const size_t body_len = br_byte_pos(&br);
    const uint16_t frame_crc = (uint16_t)br_bits(&br, 16);
    if (br.err) { result = fail(VFX_IO_ERR_FORMAT, "flac_decode: truncated frame at byte %zu", off); goto out; }
    if (crc16(data + off, body_len) != frame_crc) { result = fail(VFX_IO_ERR_CRC, "flac_decode: frame CRC-16 mismatch at byte %zu", off); goto out; }
    off += body_len + 2;

    int64_t *c0 = sub, *c1 = sub + 65536u;
    if (ch_code == 8) for (int i = 0; i < blocksize; ++i) c1[i] = (int64_t)(U(c0[i]) - U(c1[i]));   /* left, side  */
    else if (ch_code == 9) for (int i = 0; i < blocksize; ++i) c0[i] = (int64_t)(U(c0[i]) + U(c1[i])); /* side, right */
    else if (ch_code == 10)
      for (int i = 0; i < blocksize; ++i) {                                                  /* mid, side   */
        const int64_t s = c1[i], m = (int64_t)(U(c0[i]) << 1) | (s & 1);
        c0[i] = (int64_t)(U(m) + U(s)) >> 1;
        c1[i] = (int64_t)(U(m) - U(s)) >> 1;
      }
    size_t take = (size_t)blocksize;
    if (info.total_samples && done + take > info.total_samples) take = (size_t)(info.total_samples - done);
    if (done + take > capacity_frames) { result = fail(VFX_IO_ERR_CAPACITY, "flac_decode: output buffer holds %zu frames, stream has more", capacity_frames); goto out; }
    int32_t* dst = pcm + done * (size_t)nch;
    for (size_t i = 0; i < take; ++i)
      for (int c = 0; c < nch; ++c) dst[i * (size_t)nch + c] = (int32_t)sub[(size_t)c * 65536u + i];
    md5_samples(&md5, dst, take * (size_t)nch, md5_bytes);
    done += take;
  }
  if (info.total_samples && done != info.total_samples) {
    result = fail(VFX_IO_ERR_FORMAT, "flac_decode: stream ends after %llu of %llu samples", (unsigned long long)done, (unsigned long long)info.total_samples);
    goto out;
  }
  if (md5_state) {
    static const uint8_t zero[16] = {0};
    uint8_t digest[16];
    md5_final(&md5, digest);
    *md5_state = memcmp(info.md5, zero, 16) == 0 ? 0 : (memcmp(info.md5, digest, 16) == 0 ? 1 : -1);
  }
  result = (long long)done;
out:
  free(sub);
  return result;
}

/* -------------------------------------------------------------------------------------- encoder */
typedef struct {
  uint8_t* d;
  size_t cap, pos;
  uint64_t acc;
  int cnt;   /* bits held in acc (right-aligned) */
  int err;
} BitWriter;

static inline void bw_flush_bytes(BitWriter* w) {
  while (w->cnt >= 8) {
    if (w->pos >= w->cap) { w->err = 1; w->cnt = 0; return; }
    w->d[w->pos++] = (uint8_t)(w->acc >> (w->cnt - 8));
    w->cnt -= 8;
  }
}

static inline void bw_bits(BitWriter* w, uint32_t v, int n) { /* 0 <= n <= 32 */
  if (n == 0) return;
  w->acc = (w->acc << n) | (n == 32 ? v : (v & ((1u << n) - 1u)));
  w->cnt += n;
  bw_flush_bytes(w);
}

static inline void bw_unary(BitWriter* w, uint32_t q) {
  while (q >= 32) { bw_bits(w, 0, 32); q -= 32; }
  bw_bits(w, 1, (int)q + 1);
}

static inline void bw_align(BitWriter* w) {
  if (w->cnt & 7) bw_bits(w, 0, 8 - (w->cnt & 7));
}

static inline uint32_t zigzag(int32_t e) { return ((uint32_t)e << 1) ^ (uint32_t)(e >> 31); }

/* Bits needed to Rice-code u[0..n) with parameter k. */
static uint64_t rice_bits(const uint32_t* u, int n, int k) {
  uint64_t bits = (uint64_t)n * (uint64_t)(k + 1);
  for (int i = 0; i < n; ++i) bits += u[i] >> k;
  return bits;
}

static int best_rice_param(const uint32_t* u, int n, uint64_t* bits_out) {
  uint64_t sum = 0;
  for (int i = 0; i < n; ++i) sum += u[i];
  int k0 = 0;
  if (n > 0) while (k0 < 30 && ((uint64_t)n << (k0 + 1)) <= sum) ++k0;   /* ~ floor(log2(mean)) */
  int best = k0;
  uint64_t best_bits = rice_bits(u, n, k0);
  for (int k = k0 - 1; k <= k0 + 1; k += 2) {
    if (k < 0 || k > 30) continue;
    const uint64_t b = rice_bits(u, n, k);
    if (b < best_bits) { best_bits = b; best = k; }
  }
  *bits_out = best_bits;
  return best;
}

/* One 16-bit subframe: CONSTANT, FIXED(0..4) + partitioned Rice, or VERBATIM, whichever is smallest. */
static void encode_subframe(BitWriter* w, const int32_t* x, int n, uint32_t* u /* scratch, n */) {
  int constant = 1;
  for (int i = 1; i < n; ++i) if (x[i] != x[0]) { constant = 0; break; }
  if (constant) {
    bw_bits(w, 0x00, 8);
    bw_bits(w, (uint32_t)x[0], 16);
    return;
  }
  int order = -1;
  if (n > 4) {                                   /* pick the fixed predictor with the least |residual| */
    uint64_t err[5] = {0, 0, 0, 0, 0};
    for (int i = 4; i < n; ++i) {
      const int64_t e0 = x[i], e1 = e0 - x[i - 1], e2 = e1 - ((int64_t)x[i - 1] - x[i - 2]);
      const int64_t d2p = (int64_t)x[i - 1] - 2 * (int64_t)x[i - 2] + x[i - 3];
      const int64_t e3 = e2 - d2p;
      const int64_t d3p = (int64_t)x[i - 1] - 3 * (int64_t)x[i - 2] + 3 * (int64_t)x[i - 3] - x[i - 4];
      const int64_t e4 = e3 - d3p;
      err[0] += (uint64_t)(e0 < 0 ? -e0 : e0); err[1] += (uint64_t)(e1 < 0 ? -e1 : e1);
      err[2] += (uint64_t)(e2 < 0 ? -e2 : e2); err[3] += (uint64_t)(e3 < 0 ? -e3 : e3);
      err[4] += (uint64_t)(e4 < 0 ? -e4 : e4);
    }
    order = 0;
    for (int o = 1; o <= 4; ++o) if (err[o] < err[order]) order = o;
  }
  uint64_t best_total = ~0ull;
  int best_po = 0;
  if (order >= 0) {
    for (int i = order; i < n; ++i) {
      int64_t e;
      switch (order) {
        case 0: e = x[i]; break;
        case 1: e = (int64_t)x[i] - x[i - 1]; break;
        case 2: e = (int64_t)x[i] - 2 * (int64_t)x[i - 1] + x[i - 2]; break;
        case 3: e = (int64_t)x[i] - 3 * (int64_t)x[i - 1] + 3 * (int64_t)x[i - 2] - x[i - 3]; break;
        default: e = (int64_t)x[i] - 4 * (int64_t)x[i - 1] + 6 * (int64_t)x[i - 2] - 4 * (int64_t)x[i - 3] + x[i - 4]; break;
      }
      u[i] = zigzag((int32_t)e);                 /* |e| <= 16 * 2^15 for 16-bit input */
    }
    for (int po = 0; po <= 6; ++po) {
      if ((n >> po) << po != n || (n >> po) <= order) break;
      uint64_t total = 0;
      for (int p = 0; p < (1 << po); ++p) {
        const int start = p == 0 ? order : p * (n >> po), count = (n >> po) - (p == 0 ? order : 0);
        uint64_t b;
        best_rice_param(u + start, count, &b);
        total += b + 5;
      }
      if (total < best_total) { best_total = total; best_po = po; }
    }
  }
  if (order < 0 || best_total + 6 + 16 * (uint64_t)order >= 16ull * (uint64_t)n) {   /* VERBATIM */
    bw_bits(w, 0x02, 8);
    for (int i = 0; i < n; ++i) bw_bits(w, (uint32_t)x[i], 16);
    return;
  }
  bw_bits(w, (uint32_t)((8 + order) << 1), 8);
  for (int i = 0; i < order; ++i) bw_bits(w, (uint32_t)x[i], 16);
  int ks[64], kmax = 0;
  for (int p = 0; p < (1 << best_po); ++p) {
    const int start = p == 0 ? order : p * (n >> best_po), count = (n >> best_po) - (p == 0 ? order : 0);
    uint64_t b;
    ks[p] = best_rice_param(u + start, count, &b);
    if (ks[p] > kmax) kmax = ks[p];
  }
  const int method = kmax > 14;
  bw_bits(w, (uint32_t)method, 2);
  bw_bits(w, (uint32_t)best_po, 4);
  for (int p = 0; p < (1 << best_po); ++p) {
    const int start = p == 0 ? order : p * (n >> best_po), count = (n >> best_po) - (p == 0 ? order : 0);
    const int k = ks[p];
    bw_bits(w, (uint32_t)k, method ? 5 : 4);
    for (int i = start; i < start + count; ++i) {
      bw_unary(w, u[i] >> k);
      bw_bits(w, u[i], k);
    }
  }
}

size_t vfx_flac_encode_bound(size_t nframes, int channels) {
  const size_t blocks = nframes / 4096 + 1;
  return 42 + blocks * (16 + (size_t)channels) + nframes * (size_t)channels * 2 + 64;
}

long long vfx_flac_encode(const int32_t* pcm, size_t nframes, int channels, int sample_rate, uint8_t* out, size_t capacity) {
  g_err[0] = 0;
  if (!pcm && nframes) return fail(VFX_IO_ERR_ARGUMENT, "flac_encode: null input");
  if (!out) return fail(VFX_IO_ERR_ARGUMENT, "flac_encode: null output");
  if (channels < 1 || channels > 8) return fail(VFX_IO_ERR_UNSUPPORTED, "flac_encode: %d channels", channels);
  if (sample_rate < 1 || sample_rate >= (1 << 20)) return fail(VFX_IO_ERR_ARGUMENT, "flac_encode: sample rate %d", sample_rate);
  if ((uint64_t)nframes >= (1ull << 36)) return fail(VFX_IO_ERR_UNSUPPORTED, "flac_encode: too many samples");
  if (capacity < 42) return fail(VFX_IO_ERR_CAPACITY, "flac_encode: output buffer too small");
  for (size_t i = 0; i < nframes * (size_t)channels; ++i)
    if (pcm[i] < -32768 || pcm[i] > 32767) return fail(VFX_IO_ERR_ARGUMENT, "flac_encode: sample %zu = %d outside the 16-bit range", i, pcm[i]);
  enum { BS = 4096 };
  int sr_code = 0;
  for (int i = 1; i < 12; ++i) if ((int)k_rates[i] == sample_rate) sr_code = i;
  int32_t* chan = (int32_t*)malloc(sizeof(int32_t) * BS);
  uint32_t* scratch = (uint32_t*)malloc(sizeof(uint32_t) * BS);
  if (!chan || !scratch) { free(chan); free(scratch); return fail(VFX_IO_ERR_CAPACITY, "flac_encode: out of memory"); }
  size_t pos = 42, min_frame = ~(size_t)0, max_frame = 0;
  long long result = 0;
  uint32_t frame_no = 0;
  for (size_t start = 0; start < nframes; start += BS, ++frame_no) {
    const int n = (int)(nframes - start < BS ? nframes - start : BS);
    BitWriter w = {out + pos, capacity - pos, 0, 0, 0, 0};
    bw_bits(&w, 0xFFF8, 16);
    const int bs_code = n == BS ? 12 : (n <= 256 ? 6 : 7);
    bw_bits(&w, (uint32_t)bs_code, 4);
    bw_bits(&w, (uint32_t)sr_code, 4);
    bw_bits(&w, (uint32_t)(channels - 1), 4);
    bw_bits(&w, 4, 3);                                  /* 16 bits per sample */
    bw_bits(&w, 0, 1);
    if (frame_no < 0x80) bw_bits(&w, frame_no, 8);       /* UTF-8 style frame number */
    else {
      int extra = 1;                                     /* continuation bytes: lead byte carries 6 - extra bits */
      while (extra < 5 && (frame_no >> (5 * extra + 6)) != 0) ++extra;
      bw_bits(&w, ((0xFFu << (7 - extra)) & 0xFFu) | (frame_no >> (6 * extra)), 8);
      for (int i = extra - 1; i >= 0; --i) bw_bits(&w, 0x80u | ((frame_no >> (6 * i)) & 0x3Fu), 8);
    }
    if (bs_code == 6) bw_bits(&w, (uint32_t)(n - 1), 8);
    else if (bs_code == 7) bw_bits(&w, (uint32_t)(n - 1), 16);
    if (w.err) { result = fail(VFX_IO_ERR_CAPACITY, "flac_encode: output buffer too small"); goto out; }
    bw_bits(&w, crc8(w.d, w.pos), 8);
    for (int c = 0; c < channels; ++c) {
      for (int i = 0; i < n; ++i) chan[i] = pcm[(start + (size_t)i) * (size_t)channels + (size_t)c];
      encode_subframe(&w, chan, n, scratch);
    }
    bw_align(&w);
    if (w.err) { result = fail(VFX_IO_ERR_CAPACITY, "flac_encode: output buffer too small"); goto out; }
    bw_bits(&w, crc16(w.d, w.pos), 16);
    if (w.err) { result = fail(VFX_IO_ERR_CAPACITY, "flac_encode: output buffer too small"); goto out; }
    if (w.pos < min_frame) min_frame = w.pos;
    if (w.pos > max_frame) max_frame = w.pos;
    pos += w.pos;
  }
  if (nframes == 0) min_frame = 0;
  {
    Md5 md5;
    md5_init(&md5);
    md5_samples(&md5, pcm, nframes * (size_t)channels, 2);
    uint8_t* p = out;
    memcpy(p, "fLaC", 4);
    p[4] = 0x80; p[5] = 0; p[6] = 0; p[7] = 34;          /* last-block flag | STREAMINFO, length 34 */
    p += 8;
    p[0] = BS >> 8; p[1] = BS & 0xFF; p[2] = BS >> 8; p[3] = BS & 0xFF;
    p[4] = (uint8_t)(min_frame >> 16); p[5] = (uint8_t)(min_frame >> 8); p[6] = (uint8_t)min_frame;
    p[7] = (uint8_t)(max_frame >> 16); p[8] = (uint8_t)(max_frame >> 8); p[9] = (uint8_t)max_frame;
    const uint64_t total = nframes;
    p[10] = (uint8_t)(sample_rate >> 12); p[11] = (uint8_t)(sample_rate >> 4);
    p[12] = (uint8_t)(((sample_rate & 0xF) << 4) | ((channels - 1) << 1) | 0 /* (16-1) >> 4 */);
    p[13] = (uint8_t)(((16 - 1) & 0xF) << 4 | (uint8_t)(total >> 32));
    p[14] = (uint8_t)(total >> 24); p[15] = (uint8_t)(total >> 16); p[16] = (uint8_t)(total >> 8); p[17] = (uint8_t)total;
    md5_final(&md5, p + 18);
  }
  result = (long long)pos;
out:
  free(chan);
  free(scratch);
  return result;
}
