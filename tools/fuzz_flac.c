/* Fuzzer for the FLAC decoder (include/vfx_hostio.h): encodes random signals, flips bits (every third case with the
 * header CRC-8 and frame CRC-16 of the first frame recomputed so the corruption reaches the subframe decoder), truncates,
 * decodes.  Build and run under the sanitizers:
 *   gcc -O1 -g -fsanitize=address,undefined -std=gnu11 -o /tmp/fuzz_flac tools/fuzz_flac.c voicefixer_b200/hostio/flac_codec.c && /tmp/fuzz_flac
 * Round 1: 30000 cases clean (it found signed-overflow UB in the predictor recursions, now unsigned). */
#include "../include/vfx_hostio.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s = 88172645463325252ull;
static uint32_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
static uint8_t crc8(const uint8_t* p, size_t n) { uint8_t c = 0; while (n--) { c ^= *p++; for (int b = 0; b < 8; ++b) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 7) : (uint8_t)(c << 1); } return c; }
static uint16_t crc16(const uint8_t* p, size_t n) { uint16_t c = 0; while (n--) { c ^= (uint16_t)(*p++ << 8); for (int b = 0; b < 8; ++b) c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x8005) : (uint16_t)(c << 1); } return c; }
int main(void) {
  enum { N = 20000 };
  static int32_t pcm[N * 2], out[N * 2 + 8];
  static uint8_t buf[1 << 20], mut[1 << 20];
  long ok = 0, err = 0;
  for (int it = 0; it < 30000; ++it) {
    int n = rnd() % N, ch = 1 + rnd() % 2, v = 0, step = 1 + rnd() % 3000;
    for (int i = 0; i < n * ch; ++i) { v += (int)(rnd() % (2 * step + 1)) - step; if (v > 32767) v = 32767; if (v < -32768) v = -32768; pcm[i] = v; }
    long long nb = vfx_flac_encode(pcm, n, ch, 44100, buf, sizeof buf);
    if (nb < 0) { printf("encode failed %s\n", vfx_hostio_last_error()); return 1; }
    memcpy(mut, buf, nb);
    int flips = 1 + rnd() % 4;
    for (int f = 0; f < flips; ++f) { size_t pos = rnd() % nb; mut[pos] ^= (uint8_t)(1u << (rnd() % 8)); }
    if (it % 3 == 0 && nb > 60) {
      /* make the corruption pass the CRCs: recompute header CRC-8 and frame CRC-16 of the first frame */
      size_t off = 42; int bs_code = mut[off + 2] >> 4; size_t hl = 4 + 1 + (bs_code == 6 ? 1 : bs_code == 7 ? 2 : 0);
      if ((mut[off + 4] & 0x80) == 0) { mut[off + hl] = crc8(mut + off, hl);
        /* find the end of frame 0 = start of frame 1 in the ORIGINAL (same length) */
        size_t end = (size_t)nb; for (size_t p = off + 8; p + 4 < (size_t)nb; ++p) if (buf[p] == 0xFF && buf[p + 1] == 0xF8 && buf[p + 4] == 1) { end = p; break; }
        if (end >= off + 4) { uint16_t c = crc16(mut + off, end - 2 - off); mut[end - 2] = (uint8_t)(c >> 8); mut[end - 1] = (uint8_t)c; } }
    }
    int md5 = 0;
    size_t len = (it % 5 == 0) ? (size_t)(rnd() % (nb + 1)) : (size_t)nb;
    long long r = vfx_flac_decode(mut, len, out, N, &md5);
    if (r >= 0) ++ok; else ++err;
  }
  printf("decoded %ld, rejected %ld\n", ok, err);
  return 0;
}
