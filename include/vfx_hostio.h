/* vfx_hostio.h -- C ABI of the host-side audio file codec (libvfx_hostio.so, plain C, no CUDA).
 *
 * The reference reads its inputs with librosa.load (voicefixer/base.py:47-49, tools/wav.py:136,144)
 * and writes its outputs with soundfile.write (tools/wav.py:37); both pick the container from the
 * file extension, and the reference's own acceptance test runs on FLAC files in and out
 * (test/test.py:48-57,85-89: original.flac -> output_mode_N.flac, p360_001_mic1.flac -> oracle.flac).
 * Neither library exists in this image, so the FLAC side of that file contract is provided here:
 * a complete decoder for the FLAC subset format (CONSTANT / VERBATIM / FIXED / LPC subframes, Rice
 * and Rice2 residuals with escape partitions, all stereo decorrelation modes, wasted bits, header
 * CRC-8, frame CRC-16, STREAMINFO MD5 verification) and an encoder that writes what soundfile writes
 * for this path: 16-bit PCM (FIXED predictors 0..4 + partitioned Rice, 4096-sample blocks).
 *
 * All functions are thread-safe except for the per-thread error string.  No allocation crosses the
 * boundary: the caller provides every buffer.  Return values < 0 are VFX_IO_* error codes.
 */
#ifndef VFX_HOSTIO_H
#define VFX_HOSTIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  VFX_IO_OK = 0,
  VFX_IO_ERR_FORMAT = -1,      /* not a FLAC stream / reserved field / corrupt bitstream         */
  VFX_IO_ERR_CRC = -2,         /* header CRC-8 or frame CRC-16 mismatch                          */
  VFX_IO_ERR_CAPACITY = -3,    /* caller buffer too small                                        */
  VFX_IO_ERR_UNSUPPORTED = -4, /* valid FLAC outside what this codec handles (e.g. > 8 channels) */
  VFX_IO_ERR_ARGUMENT = -5
};

typedef struct vfx_flac_info {
  uint32_t sample_rate;
  uint32_t channels;
  uint32_t bits_per_sample;
  uint32_t min_blocksize;
  uint32_t max_blocksize;
  uint64_t total_samples;      /* per channel; 0 = unknown (decode still works)                  */
  uint8_t md5[16];             /* of the interleaved little-endian PCM; all zero = not set        */
  uint64_t audio_offset;       /* byte offset of the first frame                                 */
} vfx_flac_info;

/* Parses the "fLaC" marker and the metadata chain (what librosa.load / soundfile.info see first). */
int vfx_flac_probe(const uint8_t* data, size_t nbytes, vfx_flac_info* info);

/* Decodes the whole stream into interleaved int32 PCM (sample values at the stream's bit depth,
 * not scaled).  capacity_frames = room in `pcm` in samples per channel.  Returns the number of
 * samples per channel decoded.  *md5_state (optional): 1 = signature verified, 0 = stream carries
 * no signature, -1 = mismatch (decoded data is still returned). */
long long vfx_flac_decode(const uint8_t* data, size_t nbytes, int32_t* pcm, size_t capacity_frames,
                          int* md5_state);

/* Upper bound of the encoded size for vfx_flac_encode. */
size_t vfx_flac_encode_bound(size_t nframes, int channels);

/* Encodes interleaved 16-bit-range PCM (int32 container, values in [-32768, 32767]) the way
 * soundfile.write(fname.flac, int16 frames, sr) does for tools/wav.py:37: subtype PCM_16, fixed
 * 4096-sample blocks, STREAMINFO with total samples and MD5.  Returns bytes written. */
long long vfx_flac_encode(const int32_t* pcm, size_t nframes, int channels, int sample_rate,
                          uint8_t* out, size_t capacity);

/* MD5 of a byte buffer (RFC 1321), exposed because the stream signature is part of the contract. */
void vfx_md5(const uint8_t* data, size_t nbytes, uint8_t digest[16]);

/* Message for the last error on the calling thread ("" if none). */
const char* vfx_hostio_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VFX_HOSTIO_H */
