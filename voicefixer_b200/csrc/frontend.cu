// Framed real FFT (2048, hop 441, periodic Hann, reflect pad) -> magnitude -> 128-bin mel.
// Replaces the reference's DFT-matrix conv1d STFT (torchlibrosa STFT built at
// voicefixer/tools/modules/fDomainHelper.py:23-31, used by spectrogram_phase :81-86 with
// eps = 1e-8 from wav_to_spectrogram_phase :88) and MelScale.forward
// (voicefixer/tools/mel_scale.py:63-77), as called by VoiceFixer._pre voicefixer/base.py:78-85.
// The unused cos/sin phase outputs are not computed.
#include "vfx_common.cuh"

namespace vfx {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place-in-shared-memory complex FFT of 1024 points (radix-2 Stockham autosort, 10 passes,
// 256 threads, two butterflies per thread per pass).  tw[i] = exp(-2*pi*i*I/2048), i < 1024.
// Returns the buffer holding the result.
__device__ float2* fft1024(float2* a, float2* b, const float2* __restrict__ tw, int tid) {
  float2* src = a; float2* dst = b;
#pragma unroll 1
  for (int Ns = 1; Ns < 1024; Ns <<= 1) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = tid + q * 256;          // butterfly index 0..511
      const int k = j & (Ns - 1);
      const float2 w = tw[k * (1024 / Ns)];
      const float2 u = src[j];
      const float2 v = cmul(w, src[j + 512]);
      const int j0 = ((j - k) << 1) + k;
      dst[j0] = make_float2(u.x + v.x, u.y + v.y);
      dst[j0 + Ns] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

__global__ void __launch_bounds__(256) stft_mel_kernel(const float* __restrict__ wav, int L, int T,
                                                       const float* __restrict__ window,
                                                       const float2* __restrict__ tw,
                                                       const float* __restrict__ fbT,
                                                       const int* __restrict__ fb_start,
                                                       const int* __restrict__ fb_len,
                                                       float* __restrict__ mel, float* __restrict__ sp) {
  __shared__ float2 bufA[1024], bufB[1024];
  __shared__ float mag[1025];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* x = wav + (long long)b * L;
  const int base = t * 441 - 1024;
  // z[n] = xw[2n] + i*xw[2n+1], reflect padding at both ends (F.pad(..., mode='reflect'))
  for (int n = tid; n < 1024; n += 256) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int i = base + 2 * n + h;
      if (i < 0) i = -i;
      if (i >= L) i = 2 * (L - 1) - i;
      v[h] = x[i] * window[2 * n + h];
    }
    bufA[n] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  const float2* Z = fft1024(bufA, bufB, tw, tid);
  // untangle the packed real transform, magnitude with the reference's clamp (eps 1e-8)
  for (int k = tid; k <= 1024; k += 256) {
    float re, im;
    if (k == 1024) {
      re = Z[0].x - Z[0].y; im = 0.f;
    } else {
      const float2 zk = Z[k], zn = Z[(1024 - k) & 1023];
      const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // -i/2 (zk - conj zn)
      const float2 wo = cmul(tw[k], o);
      re = e.x + wo.x; im = e.y + wo.y;
    }
    const float m = sqrtf(fmaxf(re * re + im * im, 1e-8f));
    mag[k] = m;
    if (sp) sp[((long long)b * T + t) * 1025 + k] = m;
  }
  __syncthreads();
  if (tid < 128) {
    const int s = fb_start[tid], n = fb_len[tid];
    const float* f = fbT + (long long)tid * 1025 + s;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(mag[s + j], f[j], acc);
    mel[((long long)b * T + t) * 128 + tid] = acc;
  }
}

}  // namespace

int stft_mel(const float* wav, int B, int L, int T, const float* window, const float2* tw,
             const float* fbT, const int* fb_start, const int* fb_len, float* mel, float* sp,
             cudaStream_t st) {
  VFX_REQUIRE(L > 1024, "frontend: L=%d must exceed 1024 (reflect padding of n_fft/2)", L);
  VFX_REQUIRE(T == 1 + L / 441, "frontend: T=%d inconsistent with L=%d", T, L);
  dim3 grid(T, B);
  stft_mel_kernel<<<grid, 256, 0, st>>>(wav, L, T, window, tw, fbT, fb_start, fb_len, mel, sp);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
