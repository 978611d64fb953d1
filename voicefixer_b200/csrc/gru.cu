// Bidirectional GRU layer recurrence (hidden 256), torch.nn.GRU semantics:
//   r = sigmoid(gi_r + W_hr h + b_hr); z = sigmoid(gi_z + W_hz h + b_hz)
//   n = tanh(gi_n + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h
// Reference: BN_GRU voicefixer/restorer/model.py:22-62 (nn.GRU(512, 256, num_layers=2,
// bidirectional=True, batch_first=True)); state is zero at the start of every segment.
// The input projections gi = x W_ih^T + b_ih are one batched GEMM (conv_gemm); this kernel is the
// sequential part: one CTA per (direction, group of G items), 768 threads = one per gate row,
// W_hh^T streamed from L2 each step and shared by the G items of the CTA.
#include "vfx_common.cuh"

namespace vfx {

namespace {

constexpr int H = 256, G3 = 768, GI = 4;   // GI items per CTA

__global__ void __launch_bounds__(G3) gru_layer_kernel(const float* __restrict__ gi,
                                                       const float* __restrict__ whh_t,
                                                       const float* __restrict__ bhh, int B, int T,
                                                       float* __restrict__ out) {
  __shared__ float h_s[GI][H];
  __shared__ float gh_s[GI][G3];
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * GI;
  const int j = threadIdx.x;
  const float* __restrict__ W = whh_t + (long long)dir * H * G3;   // [256][768]
  const float bj = bhh[dir * G3 + j];
  for (int i = j; i < GI * H; i += G3) (&h_s[0][0])[i] = 0.f;
  __syncthreads();

  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    // prefetch this step's input projections for the (item, unit) pairs this thread finalises
    float gir[2], giz[2], gin[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = j + q * G3;            // 0 .. GI*H-1 (= 1024) in two rounds
      const int g = idx >> 8, u = idx & 255;
      if (idx < GI * H && b0 + g < B) {
        const float* p = gi + (((long long)(b0 + g) * T + t) * 2 + dir) * G3;
        gir[q] = p[u]; giz[q] = p[H + u]; gin[q] = p[2 * H + u];
      } else { gir[q] = giz[q] = gin[q] = 0.f; }
    }
    float acc[GI];
#pragma unroll
    for (int g = 0; g < GI; ++g) acc[g] = bj;
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float w = W[(long long)k * G3 + j];
#pragma unroll
      for (int g = 0; g < GI; ++g) acc[g] = fmaf(w, h_s[g][k], acc[g]);
    }
#pragma unroll
    for (int g = 0; g < GI; ++g) gh_s[g][j] = acc[g];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = j + q * G3;
      const int g = idx >> 8, u = idx & 255;
      if (idx < GI * H && b0 + g < B) {
        const float r = 1.f / (1.f + expf(-(gir[q] + gh_s[g][u])));
        const float z = 1.f / (1.f + expf(-(giz[q] + gh_s[g][H + u])));
        const float n = tanhf(gin[q] + r * gh_s[g][2 * H + u]);
        const float hn = (1.f - z) * n + z * h_s[g][u];
        h_s[g][u] = hn;
        out[((long long)(b0 + g) * T + t) * (2 * H) + dir * H + u] = hn;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int gru_layer(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out,
              cudaStream_t st) {
  VFX_REQUIRE(B > 0 && T > 0, "gru_layer: empty problem");
  dim3 grid(ceil_div(B, GI), 2);
  gru_layer_kernel<<<grid, G3, 0, st>>>(gi, whh_t, bhh, B, T, out);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
