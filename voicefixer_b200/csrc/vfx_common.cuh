// Shared declarations of the vfx_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/vfx_b200.h"

namespace vfx {

void set_error(const char* fmt, ...);

#define VFX_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      vfx::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,           \
                     cudaGetErrorString(_e));                                         \
      return VFX_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)
#define VFX_LAUNCH_CHECK()                                                            \
  do { ++vfx::g_launches; VFX_CUDA_CHECK(cudaGetLastError()); } while (0)

#define VFX_REQUIRE(cond, ...)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      vfx::set_error(__VA_ARGS__);                                                    \
      return VFX_ERR_INVALID;                                                         \
    }                                                                                 \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
// VFX_PREC_TF32 operand element: fp32 storage holding a value already rounded to tf32 (10-bit mantissa,
// round-to-nearest) by its producer, so that the tensor core's truncation of the low 13 bits is exact.
// (Un-rounded operands make kind::tf32 truncate: a CPU simulation of the whole path gave 1.4e-2 waveform
// rel-RMS with truncation against 1.7e-3 with round-to-nearest, tools/sim_precision.py.)
struct tf32_t { float v; };
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float to_f(tf32_t v) { return __uint_as_float(__float_as_uint(v.v) & 0xFFFFE000u); }   // as the tensor core reads it
template <> __device__ __forceinline__ tf32_t from_f<tf32_t>(float v) { return tf32_t{round_tf32(v)}; }

// Encoded tf32 stream (vfx_conv_desc.res_enc / raw_enc): S = bits(lrelu(x)) + 0x1000 -- see include/vfx_b200.h.
// (lrelu and its inverse as max / min with the scaled value: two instructions, the same result for 0 < slope < 1)
__device__ __forceinline__ float stream_enc(float x, float slope) {
  const float y = fmaxf(x, x * slope);
  return __uint_as_float(__float_as_uint(y) + 0x1000u);
}
__device__ __forceinline__ float stream_dec(float s, float inv_slope) {
  const float y = __uint_as_float(__float_as_uint(s) - 0x1000u);
  return fminf(y, y * inv_slope);
}

// element size of a GEMM operand / weight in the given vfx_precision
static inline size_t prec_esz(int precision) { return (precision == VFX_PREC_BF16 || precision == VFX_PREC_FP16) ? 2 : 4; }
// two fp32 -> one packed pair of 16-bit operands (bf16 or fp16), round-to-nearest
template <bool FP16> __device__ __forceinline__ uint32_t pack16(float lo, float hi) {
  if (FP16) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float apply_act(float v, int act, float p) {
  switch (act) {
    case VFX_ACT_LRELU: return v > 0.f ? v : v * p;
    case VFX_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case VFX_ACT_LRELU_XSINX: {
      float u = v > 0.f ? v : v * p;
      return u + sinf(u);
    }
    case VFX_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ------------------------------------------------------------------ kernels (host launchers)
int conv_gemm_simt(int precision, const vfx_conv_desc& d, cudaStream_t st);
int resstack_pair_tc(const vfx_pair_desc& d, cudaStream_t st);     // fused ResStack pair (bf16, C = 64), one CTA per tile
int resstack_pair2_tc(const vfx_pair_desc& d, cudaStream_t st);    // ... two-CTA cluster pipeline (bf16 C = 128, tf32 C = 64)
size_t resstack_pair2_scratch_bytes();
int resstack_pair3_tc(const vfx_pair_desc& d, cudaStream_t st);    // ... tf32 C = 64 on one SM, residual stashed in TMEM
int conv_gemm_tc(int precision, const vfx_conv_desc& d, cudaStream_t st);   // tcgen05: bf16 (kind::f16) or tf32 (kind::tf32)

// y = act(scale[b][c]*x + shift[b][c]); x fp32 [B][P][C] (row pitch ldx), y operand type.
// bn_C == 1: single-channel BN (scale[b][0]).  stat_sB: element stride between items (0 = shared).
int bn_act(int precision, const float* x, long long x_sB, long long ldx, int B, long long P, int C,
           const float* scale, const float* shift, int bn_C, int stat_sB, int act, float act_param,
           void* y, long long y_sB, long long ldy, cudaStream_t st);
// per-item biased batch statistics -> scale/shift [B][C] (train-mode BN, eps 1e-5)
// acc: scratch of 2*B*bn_C doubles.
int bn_stats(const float* x, long long x_sB, long long ldx, int B, long long P, int C, int bn_C,
             const float* gamma, const float* beta, float* scale, float* shift, double* acc,
             cudaStream_t st);
int dropout_apply(float* x, const uint8_t* keep, long long n, cudaStream_t st);
int avgpool2x2(const float* x, long long x_sB, long long x_sH, long long x_sW, int B, int H, int W,
               int C, float* y, cudaStream_t st);
int stft_mel(const float* wav, int B, int L, int T, const float* window, const float2* tw,
             const float* fbT, const int* fb_start, const int* fb_len, float* mel, float* sp,
             cudaStream_t st);
int gru_layer(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out,
              cudaStream_t st);
// denoiser tail: clean = sigmoid(lin)*mel ; x = log10(clip(clean)); unet_in[b][t][f<127][2]
int mask_log_pack(const float* lin_sig, const float* mel, int B, int T, int Tp, float* xlog,
                  float* unet_in, cudaStream_t st);
// unet head: out[b][t][f] = (f<127 ? bias + sum_c w[c]*x[b][t][f][c] : 0) + xlog[b][t][f]
int unet_head(const float* x, int B, int T, int Tp, const float* w, const float* bias,
              const float* xlog, float* out, cudaStream_t st);
// vocoder conditions: (from_log) / weight, dB, normalise, tail pad of -4 -> cond [B][Tc][128]
// tab = mel_weight[128] ++ [min_level] (fp32, computed by the host exactly as the reference does)
int voc_normalize(const float* mel, int B, int T, int Tc, int input_is_log, const float* tab,
                  void* cond, int precision, cudaStream_t st);
int cast_rows(const float* x, long long n, void* y, int precision, cudaStream_t st);
// rows [0,3) and [3+L, L+6) of a [B][L+6][C] operand buffer <- reflection of the interior
int reflect_pad3(void* buf, int B, int L, int C, int precision, cudaStream_t st);
// final conv: lrelu0.2 -> reflect pad 3 -> conv k7 (64->1) -> tanh -> trim -> *scale
int voc_post(const float* x, int B, int L, const float* w, const float* bias, int lo, int out_len,
             float scale, float* out, cudaStream_t st);
int hf_cut(const float* wav, int B, int L, float ratio, const float* window, const float2* tw,
           float* out, int* cut_bins, void* ws, size_t ws_bytes, cudaStream_t st);
size_t hf_cut_workspace(int B, int L);

}  // namespace vfx
