// HBM-bound elementwise / reduction kernels of the path (all channels-last).
#include "vfx_common.cuh"

namespace vfx {

namespace {

// ------------------------------------------------------------------ bn_act
// y = act(scale*x + shift).  Reference: F.leaky_relu_(bn(x)) voicefixer/restorer/modules.py:70-71,
// F.relu_(bn1(x)) modules.py:150, BatchNorm2d(1)/ReLU members of the denoiser restorer/model.py:69-99.
template <typename T>
__global__ void bn_act_kernel(const float* __restrict__ x, long long x_sB, long long ldx, int B,
                              long long P, int C, const float* __restrict__ scale,
                              const float* __restrict__ shift, int bn_C, int stat_sB, int act,
                              float act_param, T* __restrict__ y, long long y_sB, long long ldy) {
  const int C4 = C >> 2;
  const long long total = (long long)B * P * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long r = i / C4;
    const long long p = r % P;
    const int b = (int)(r / P);
    const float4 v = *reinterpret_cast<const float4*>(x + b * x_sB + p * ldx + c);
    float in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sc = b * stat_sB + (bn_C == 1 ? 0 : c + j);
      const float s = scale ? scale[sc] : 1.f, t = shift ? shift[sc] : 0.f;
      o[j] = apply_act(fmaf(in[j], s, t), act, act_param);
    }
    T* yp = y + b * y_sB + p * ldy + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) yp[j] = from_f<T>(o[j]);
  }
}

template <typename T>
__global__ void bn_act_scalar_kernel(const float* __restrict__ x, long long x_sB, long long ldx, int B,
                                     long long P, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, int bn_C, int stat_sB, int act,
                                     float act_param, T* __restrict__ y, long long y_sB, long long ldy) {
  const long long total = (long long)B * P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const long long p = r % P;
    const int b = (int)(r / P);
    const int sc = b * stat_sB + (bn_C == 1 ? 0 : c);
    const float s = scale ? scale[sc] : 1.f, t = shift ? shift[sc] : 0.f;
    y[b * y_sB + p * ldy + c] = from_f<T>(apply_act(fmaf(x[b * x_sB + p * ldx + c], s, t), act, act_param));
  }
}

// ------------------------------------------------------------------ bn_stats (mode 2)
// Biased batch statistics per item (reference batch is 1: SURVEY D5), F.batch_norm training=True.
__global__ void bn_stats_accum_kernel(const float* __restrict__ x, long long x_sB, long long ldx,
                                      long long P, int C, int bn_C, long long rows_per_block,
                                      double* __restrict__ acc /*[B][bn_C][2]*/) {
  const int b = blockIdx.z;
  const int c = blockIdx.y * 32 + threadIdx.x;
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  const long long p1 = min(P, p0 + rows_per_block);
  double s = 0.0, s2 = 0.0;
  if (c < C) {
    const float* xp = x + b * x_sB + c;
    for (long long p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
      const float v = xp[p * ldx];
      s += v; s2 += (double)v * v;
    }
  }
  __shared__ double sh[2][8][33];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int j = 1; j < 8; ++j) { s += sh[0][j][threadIdx.x]; s2 += sh[1][j][threadIdx.x]; }
    if (bn_C == 1) {
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
      }
      if (threadIdx.x == 0) { atomicAdd(&acc[(long long)b * 2], s); atomicAdd(&acc[(long long)b * 2 + 1], s2); }
    } else if (c < C) {
      atomicAdd(&acc[((long long)b * C + c) * 2], s);
      atomicAdd(&acc[((long long)b * C + c) * 2 + 1], s2);
    }
  }
}

__global__ void bn_stats_final_kernel(const double* __restrict__ acc, int B, int bn_C, double count,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * bn_C) return;
  const int c = i % bn_C;
  const double mean = acc[2 * i] / count;
  double var = acc[2 * i + 1] / count - mean * mean;
  if (var < 0) var = 0;
  const float inv = (float)(1.0 / sqrt(var + 1e-5));
  const float s = gamma[c] * inv;
  scale[i] = s;
  shift[i] = beta[c] - (float)mean * s;
}

__global__ void dropout_kernel(float* __restrict__ x, const uint8_t* __restrict__ keep, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    x[i] = keep[i] ? x[i] * 2.0f : 0.f;     // nn.Dropout(0.5): scale 1/(1-p)
}

// F.avg_pool2d(x, (2,2)) voicefixer/restorer/modules.py:103 (floor: odd last column dropped)
__global__ void avgpool_kernel(const float* __restrict__ x, long long x_sB, long long x_sH,
                               long long x_sW, int B, int Ho, int Wo, int C, float* __restrict__ y) {
  const int C4 = C >> 2;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    long long r = i / C4;
    const int w = (int)(r % Wo); r /= Wo;
    const int h = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const float* p = x + b * x_sB + (2LL * h) * x_sH + (2LL * w) * x_sW + c;
    const float4 a = *reinterpret_cast<const float4*>(p), bq = *reinterpret_cast<const float4*>(p + x_sW);
    const float4 cq = *reinterpret_cast<const float4*>(p + x_sH),
                 dq = *reinterpret_cast<const float4*>(p + x_sH + x_sW);
    float4 o;
    o.x = (a.x + bq.x + cq.x + dq.x) * 0.25f; o.y = (a.y + bq.y + cq.y + dq.y) * 0.25f;
    o.z = (a.z + bq.z + cq.z + dq.z) * 0.25f; o.w = (a.w + bq.w + cq.w + dq.w) * 0.25f;
    *reinterpret_cast<float4*>(y + (((long long)b * Ho + h) * Wo + w) * C + c) = o;
  }
}

// restorer Generator.forward voicefixer/restorer/model.py:105-109 + UNet input pad/crop
// restorer/model_kqq_bn.py:144-151: clean = sigmoid(lin)*mel; x = to_log(clean);
// unet_in[b][t][f][0..1] = (to_log(mel), x) for f < 127, zero rows for T <= t < Tp.
__global__ void mask_log_pack_kernel(const float* __restrict__ lin, const float* __restrict__ mel,
                                     int B, int T, int Tp, float* __restrict__ xlog,
                                     float* __restrict__ unet_in) {
  const long long total = (long long)B * Tp * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i & 127);
    const long long r = i >> 7;
    const int t = (int)(r % Tp);
    const int b = (int)(r / Tp);
    float lm = 0.f, lx = 0.f;
    if (t < T) {
      const long long src = ((long long)b * T + t) * 128 + f;
      const float m = mel[src];
      const float s = 1.f / (1.f + expf(-lin[src]));
      lx = log10f(fmaxf(s * m, 1e-8f));
      lm = log10f(fmaxf(m, 1e-8f));
      xlog[src] = lx;
    }
    if (f < 127) {
      float2 o; o.x = lm; o.y = lx;
      *reinterpret_cast<float2*>(unet_in + (((long long)b * Tp + t) * 127 + f) * 2) = o;
    }
  }
}

// after_conv2 (1x1, 32->1, bias) + F.pad(0,1) + crop + residual: restorer/model_kqq_bn.py:174-178,
// restorer/model.py:113.
__global__ void unet_head_kernel(const float* __restrict__ x, int B, int T, int Tp,
                                 const float* __restrict__ w, const float* __restrict__ bias,
                                 const float* __restrict__ xlog, float* __restrict__ out) {
  const long long total = (long long)B * T * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i & 127);
    const long long r = i >> 7;
    const int t = (int)(r % T);
    const int b = (int)(r / T);
    float v = 0.f;
    if (f < 127) {
      const float4* xp = reinterpret_cast<const float4*>(x + (((long long)b * Tp + t) * 127 + f) * 32);
      v = bias[0];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 a = xp[q];
        v = fmaf(a.x, w[4 * q], v); v = fmaf(a.y, w[4 * q + 1], v);
        v = fmaf(a.z, w[4 * q + 2], v); v = fmaf(a.w, w[4 * q + 3], v);
      }
    }
    out[i] = v + xlog[i];
  }
}

// from_log tools/pytorch_util.py:25-27; Vocoder.forward vocoder/base.py:51-54; tr_amp_to_db,
// tr_normalize, tr_pre vocoder/model/util.py:8-36,69-80.  tab = mel_weight[128] ++ [min_level].
template <typename T>
__global__ void voc_normalize_kernel(const float* __restrict__ mel, int B, int Tn, int Tc,
                                     int input_is_log, const float* __restrict__ tab,
                                     T* __restrict__ cond) {
  const long long total = (long long)B * Tc * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i & 127);
    const long long r = i >> 7;
    const int t = (int)(r % Tc);
    const int b = (int)(r / Tc);
    float c = -4.0f;
    if (t < Tn) {
      float v = mel[((long long)b * Tn + t) * 128 + k];
      if (input_is_log) v = exp10f(fminf(v, 5.0f));
      v = v / tab[k];
      const float S = 20.0f * log10f(fmaxf(tab[128], fabsf(v))) - 20.0f;
      c = fminf(fmaxf(8.0f * ((S + 115.0f) / 115.0f) - 4.0f, -4.0f), 4.0f);
    }
    cond[i] = from_f<T>(c);
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, long long n, T* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(x[i]);
}

// nn.ReflectionPad1d(3) voicefixer/vocoder/model/generator.py:74,94 on a [B][L+6][C] buffer
template <typename T>
__global__ void reflect_pad3_kernel(T* __restrict__ buf, int B, int L, int C) {
  const long long total = (long long)B * 6 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int j = (int)((i / C) % 6);
    const int b = (int)(i / (6LL * C));
    T* base = buf + (long long)b * (L + 6) * C;
    const int dst = j < 3 ? j : L + j;            // rows 0,1,2 and L+3,L+4,L+5
    const int src = j < 3 ? 6 - j : L + 4 - j;    // x[3-j] at row 6-j ; x[L-2-(j-3)] at row L+4-j
    base[(long long)dst * C + c] = base[(long long)src * C + c];
  }
}

// generator[14..17]: LeakyReLU(0.2) -> ReflectionPad1d(3) -> Conv1d(64,1,7) -> Tanh
// (voicefixer/vocoder/model/generator.py:93-99) fused with _trim_center (voicefixer/base.py:63-76).
//
// A 64 -> 1 channel reduction over 7 taps: 448 MACs per output sample on a tensor that is read once (HBM-bound when done
// right).  16 lanes share an output: lane q owns channels [4q, 4q+4) and its 7 x 4 weights in registers; a group walks a
// run of 64 consecutive outputs, loading each input row ONCE as one coalesced 256-byte request (16 lanes x float4) and
// feeding it to the 7 outputs it contributes to (7 rotating accumulators, statically indexed by unrolling the row loop
// by 7); a finished output is summed across the 16 lanes by a 4-step butterfly.  No shared memory.  (The previous
// version staged rows in shared memory and issued two LDS per FMA: 2.0 ms for 3.6 GB = 28 % of HBM.)
constexpr int POST_RUN = 64;                 // outputs per 16-lane group (70 input rows = 10 x 7)
constexpr int POST_GROUPS = 16;              // groups per block (256 threads)

__device__ __forceinline__ float4 post_load_row(const float* __restrict__ xb, int t, int L, int q) {
  if (t < 0) t = -t;                         // ReflectionPad1d(3)
  if (t >= L) t = 2 * (L - 1) - t;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t >= 0 && t < L) v = __ldcs(reinterpret_cast<const float4*>(xb + (long long)t * 64) + q);
  v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
  v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
  return v;
}

__global__ void __launch_bounds__(16 * POST_GROUPS) voc_post_kernel(const float* __restrict__ x, int L,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias, int lo,
                                                                     int out_len, float scale, float* __restrict__ out) {
  const int q = threadIdx.x & 15, g = threadIdx.x >> 4;
  const int b = blockIdx.y;
  const int o0 = (blockIdx.x * POST_GROUPS + g) * POST_RUN;       // first output of this group's run (may be >= out_len)
  const float* xb = x + (long long)b * L * 64;
  float4 wk[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) wk[k] = *reinterpret_cast<const float4*>(w + k * 64 + 4 * q);
  const float bv = bias[0];
  float acc[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = 0.f;
  const int t0 = lo + o0 - 3;                                     // input position of row 0 of the run
  float4 cur[7], nxt[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) cur[j] = post_load_row(xb, t0 + j, L, q);
  float res = 0.f;
#pragma unroll 1
  for (int it = 0; it < (POST_RUN + 6) / 7; ++it) {
    if (it + 1 < (POST_RUN + 6) / 7) {
#pragma unroll
      for (int j = 0; j < 7; ++j) nxt[j] = post_load_row(xb, t0 + 7 * (it + 1) + j, L, q);   // in flight during this block of rows
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4 xv = cur[j];                                   // row r = 7 it + j feeds outputs r - k, k = 0..6
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        float a = acc[(j - k + 7) % 7];
        a = fmaf(xv.x, wk[k].x, a); a = fmaf(xv.y, wk[k].y, a); a = fmaf(xv.z, wk[k].z, a); a = fmaf(xv.w, wk[k].w, a);
        acc[(j - k + 7) % 7] = a;
      }
      // output o = r - 6 is complete: its slot is (j + 1) % 7
      float sum = acc[(j + 1) % 7];
      acc[(j + 1) % 7] = 0.f;
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const int o = 7 * it + j - 6;
      if ((o & 15) == q) res = sum;                               // lane q keeps outputs o = q (mod 16)
      if (o >= 0 && (o & 15) == 15) {                             // 16 outputs gathered: one coalesced 64-byte store
        const int oo = o0 + o - 15 + q;
        if (oo < out_len) out[(long long)b * out_len + oo] = tanhf(res + bv) * scale;
      }
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) cur[j] = nxt[j];
  }
}

inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  if (g > 148LL * 32) g = 148LL * 32;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int bn_act(int precision, const float* x, long long x_sB, long long ldx, int B, long long P, int C,
           const float* scale, const float* shift, int bn_C, int stat_sB, int act, float act_param,
           void* y, long long y_sB, long long ldy, cudaStream_t st) {
  const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (x_sB % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const long long total = (long long)B * P * (vec ? C / 4 : C);
  const int g = grid_for(total);
#define VFX_BN_LAUNCH(K, T)                                                                      \
  K<T><<<g, 256, 0, st>>>(x, x_sB, ldx, B, P, C, scale, shift, bn_C, stat_sB, act, act_param,   \
                          reinterpret_cast<T*>(y), y_sB, ldy)
  if (precision == VFX_PREC_BF16) {
    if (vec) VFX_BN_LAUNCH(bn_act_kernel, __nv_bfloat16); else VFX_BN_LAUNCH(bn_act_scalar_kernel, __nv_bfloat16);
  } else if (precision == VFX_PREC_FP16) {
    if (vec) VFX_BN_LAUNCH(bn_act_kernel, __half); else VFX_BN_LAUNCH(bn_act_scalar_kernel, __half);
  } else if (precision == VFX_PREC_TF32) {
    if (vec) VFX_BN_LAUNCH(bn_act_kernel, tf32_t); else VFX_BN_LAUNCH(bn_act_scalar_kernel, tf32_t);
  } else {
    if (vec) VFX_BN_LAUNCH(bn_act_kernel, float); else VFX_BN_LAUNCH(bn_act_scalar_kernel, float);
  }
#undef VFX_BN_LAUNCH
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int bn_stats(const float* x, long long x_sB, long long ldx, int B, long long P, int C, int bn_C,
             const float* gamma, const float* beta, float* scale, float* shift, double* acc,
             cudaStream_t st) {
  VFX_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * B * bn_C, st));
  const long long rows_per_block = 256;
  dim3 grid(ceil_div(P, rows_per_block), ceil_div(C, 32), B), block(32, 8);
  bn_stats_accum_kernel<<<grid, block, 0, st>>>(x, x_sB, ldx, P, C, bn_C, rows_per_block, acc);
  VFX_LAUNCH_CHECK();
  const double count = bn_C == 1 ? (double)P * C : (double)P;
  bn_stats_final_kernel<<<ceil_div(B * bn_C, 128), 128, 0, st>>>(acc, B, bn_C, count, gamma, beta, scale, shift);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int dropout_apply(float* x, const uint8_t* keep, long long n, cudaStream_t st) {
  dropout_kernel<<<grid_for(n), 256, 0, st>>>(x, keep, n);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int avgpool2x2(const float* x, long long x_sB, long long x_sH, long long x_sW, int B, int H, int W,
               int C, float* y, cudaStream_t st) {
  VFX_REQUIRE(C % 4 == 0 && x_sW % 4 == 0 && x_sH % 4 == 0 && x_sB % 4 == 0, "avgpool: C/strides must be multiples of 4");
  const int Ho = H / 2, Wo = W / 2;
  avgpool_kernel<<<grid_for((long long)B * Ho * Wo * C / 4), 256, 0, st>>>(x, x_sB, x_sH, x_sW, B, Ho, Wo, C, y);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int mask_log_pack(const float* lin, const float* mel, int B, int T, int Tp, float* xlog,
                  float* unet_in, cudaStream_t st) {
  mask_log_pack_kernel<<<grid_for((long long)B * Tp * 128), 256, 0, st>>>(lin, mel, B, T, Tp, xlog, unet_in);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int unet_head(const float* x, int B, int T, int Tp, const float* w, const float* bias,
              const float* xlog, float* out, cudaStream_t st) {
  unet_head_kernel<<<grid_for((long long)B * T * 128), 256, 0, st>>>(x, B, T, Tp, w, bias, xlog, out);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int voc_normalize(const float* mel, int B, int T, int Tc, int input_is_log, const float* tab,
                      void* cond, int precision, cudaStream_t st) {
  const int g = grid_for((long long)B * Tc * 128);
  if (precision == VFX_PREC_BF16)
    voc_normalize_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(mel, B, T, Tc, input_is_log, tab,
                                                           reinterpret_cast<__nv_bfloat16*>(cond));
  else if (precision == VFX_PREC_FP16)
    voc_normalize_kernel<__half><<<g, 256, 0, st>>>(mel, B, T, Tc, input_is_log, tab, reinterpret_cast<__half*>(cond));
  else if (precision == VFX_PREC_TF32)
    voc_normalize_kernel<tf32_t><<<g, 256, 0, st>>>(mel, B, T, Tc, input_is_log, tab, reinterpret_cast<tf32_t*>(cond));
  else
    voc_normalize_kernel<float><<<g, 256, 0, st>>>(mel, B, T, Tc, input_is_log, tab,
                                                   reinterpret_cast<float*>(cond));
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int cast_rows(const float* x, long long n, void* y, int precision, cudaStream_t st) {
  if (precision == VFX_PREC_BF16)
    cast_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>(x, n, reinterpret_cast<__nv_bfloat16*>(y));
  else if (precision == VFX_PREC_FP16)
    cast_kernel<__half><<<grid_for(n), 256, 0, st>>>(x, n, reinterpret_cast<__half*>(y));
  else if (precision == VFX_PREC_TF32)
    cast_kernel<tf32_t><<<grid_for(n), 256, 0, st>>>(x, n, reinterpret_cast<tf32_t*>(y));
  else
    cast_kernel<float><<<grid_for(n), 256, 0, st>>>(x, n, reinterpret_cast<float*>(y));
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int reflect_pad3(void* buf, int B, int L, int C, int precision, cudaStream_t st) {
  VFX_REQUIRE(L >= 4, "reflect_pad3: length %d too short for reflection pad 3", L);
  const int g = grid_for((long long)B * 6 * C);
  if (precision == VFX_PREC_BF16 || precision == VFX_PREC_FP16)      // a 2-byte copy either way
    reflect_pad3_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(buf), B, L, C);
  else
    reflect_pad3_kernel<float><<<g, 256, 0, st>>>(reinterpret_cast<float*>(buf), B, L, C);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

int voc_post(const float* x, int B, int L, const float* w, const float* bias, int lo, int out_len,
             float scale, float* out, cudaStream_t st) {
  VFX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "voc_post: x and w must be 16-byte aligned");
  dim3 grid(ceil_div(out_len, POST_RUN * POST_GROUPS), B);
  voc_post_kernel<<<grid, 16 * POST_GROUPS, 0, st>>>(x, L, w, bias, lo, out_len, scale, out);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
