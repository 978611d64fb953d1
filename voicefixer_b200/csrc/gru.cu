// Bidirectional GRU layer recurrence (hidden 256), torch.nn.GRU semantics:
//   r = sigmoid(gi_r + W_hr h + b_hr); z = sigmoid(gi_z + W_hz h + b_hz)
//   n = tanh(gi_n + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h
// Reference: BN_GRU voicefixer/restorer/model.py:22-62 (nn.GRU(512, 256, num_layers=2,
// bidirectional=True, batch_first=True)); state is zero at the start of every segment.
// The input projections gi = x W_ih^T + b_ih are one batched GEMM (conv_gemm); this kernel is the
// sequential part and is latency-bound (T = 1001 .. 3001 dependent steps).
//
// Design (sm_100a): one thread-block CLUSTER of 8 CTAs per (direction, group of G sequences).
//   * W_hh (768 x 256 fp32 = 786 KB) never fits one SM: CTA c owns hidden units [32c, 32c+32), i.e.
//     96 gate rows, and keeps its 96 x 256 slice in REGISTERS (64 per thread) for all T steps, so a
//     step reads no weights from shared memory, L2 or HBM.
//   * every CTA holds a replica of h (double-buffered, G x 256); after computing its 32 x G new
//     values a CTA stores them into all 8 replicas through distributed shared memory, then one
//     cluster barrier (release/acquire) ends the step.
//   * fp32 FMA throughout (the recurrence is precision-sensitive); gi for step t+1 is prefetched
//     into registers during step t.
#include <cooperative_groups.h>
#include "vfx_common.cuh"

namespace cg = cooperative_groups;

namespace vfx {

namespace {

constexpr int H = 256, G3 = 768;
constexpr int CL = 8;            // CTAs per cluster
constexpr int UPC = H / CL;      // hidden units per CTA = 32
constexpr int RPC = 3 * UPC;     // gate rows per CTA = 96
constexpr int KQ = 4;            // K split: 4 quarters of 64
constexpr int NT = RPC * KQ;     // 384 threads

template <int G>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1)
gru_cluster_kernel(const float* __restrict__ gi, const float* __restrict__ whh_t,
                   const float* __restrict__ bhh, int B, int T, float* __restrict__ out) {
  __shared__ __align__(16) float h_s[2][G][H];
  __shared__ float part[KQ][G][RPC];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cl = blockIdx.x / CL;
  const int dir = cl & 1;
  const int b0 = (cl >> 1) * G;
  const int tid = threadIdx.x;
  const int q = tid / RPC, r = tid % RPC;
  const int gate = r / UPC, u = r % UPC;
  const int j = gate * H + rank * UPC + u;               // row of W_hh this thread serves

  // this thread's 64 weights: W_hh[j][q*64 .. q*64+63] (whh_t is [dir][k][768]: coalesced over j)
  float w[64];
  {
    const float* W = whh_t + (long long)dir * H * G3 + (long long)(q * 64) * G3 + j;
#pragma unroll
    for (int i = 0; i < 64; ++i) w[i] = W[(long long)i * G3];
  }
  // finaliser role: thread (fg, fu) produces h'[fg][rank*32 + fu]
  const bool fin = tid < UPC * G;
  const int fu = tid % UPC, fg = tid / UPC;
  const bool fvalid = fin && (b0 + fg) < B;
  float b_r = 0.f, b_z = 0.f, b_n = 0.f;
  if (fin) {
    const float* bb = bhh + dir * G3 + rank * UPC + fu;
    b_r = bb[0]; b_z = bb[H]; b_n = bb[2 * H];
  }
  float* remote_h[CL];
#pragma unroll
  for (int c = 0; c < CL; ++c) remote_h[c] = cluster.map_shared_rank(&h_s[0][0][0], c);

  for (int i = tid; i < 2 * G * H; i += NT) (&h_s[0][0][0])[i] = 0.f;
  cluster.sync();

  // input projections are prefetched two steps ahead (register ring): an HBM round trip is longer
  // than one step
  const long long gstride = 2LL * G3;                     // floats per (b, t)
  const float* gbase = gi + ((long long)(b0 + fg) * T) * gstride + dir * G3 + rank * UPC + fu;
  float pr[2] = {0.f, 0.f}, pz[2] = {0.f, 0.f}, pn[2] = {0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (fvalid && s < T) {
      const int ts = dir == 0 ? s : T - 1 - s;
      const float* p = gbase + (long long)ts * gstride;
      pr[s] = p[0]; pz[s] = p[H]; pn[s] = p[2 * H];
    }

  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    const int cur = step & 1;
    const float gir = pr[0], giz = pz[0], gin = pn[0];
    pr[0] = pr[1]; pz[0] = pz[1]; pn[0] = pn[1];
    if (fvalid && step + 2 < T) {
      const int tn = dir == 0 ? t + 2 : t - 2;
      const float* p = gbase + (long long)tn * gstride;
      pr[1] = p[0]; pz[1] = p[H]; pn[1] = p[2 * H];
    }
    // ---- phase A: partial dot products over this thread's K quarter, all G sequences
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_s[cur][g][q * 64 + i]);
        acc[g] = fmaf(w[i], hv.x, acc[g]);
        acc[g] = fmaf(w[i + 1], hv.y, acc[g]);
        acc[g] = fmaf(w[i + 2], hv.z, acc[g]);
        acc[g] = fmaf(w[i + 3], hv.w, acc[g]);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) part[q][g][r] = acc[g];
    __syncthreads();
    // ---- phase B: gates for 32 units x G sequences, publish h' to all 8 replicas
    float hn = 0.f;
    const int hu = rank * UPC + fu;
    if (fin) {
      float gr = b_r, gz = b_z, gn = b_n;
#pragma unroll
      for (int k = 0; k < KQ; ++k) {
        gr += part[k][fg][fu]; gz += part[k][fg][UPC + fu]; gn += part[k][fg][2 * UPC + fu];
      }
      const float rg = 1.f / (1.f + expf(-(gir + gr)));
      const float zg = 1.f / (1.f + expf(-(giz + gz)));
      const float ng = tanhf(gin + rg * gn);
      hn = (1.f - zg) * ng + zg * h_s[cur][fg][hu];
      const int dst = ((cur ^ 1) * G + fg) * H + hu;
#pragma unroll
      for (int c = 0; c < CL; ++c) remote_h[c][dst] = hn;
    }
    // release only has the DSMEM stores to publish; the global store of this step's output is
    // issued between arrive and wait so no later fence ever waits on a fresh L2 round trip
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (fvalid) out[((long long)(b0 + fg) * T + t) * (2 * H) + dir * H + hu] = hn;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int G>
int launch(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out, cudaStream_t st) {
  const int groups = (B + G - 1) / G;
  gru_cluster_kernel<G><<<2 * groups * CL, NT, 0, st>>>(gi, whh_t, bhh, B, T, out);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace

int gru_layer(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out,
              cudaStream_t st) {
  VFX_REQUIRE(B > 0 && T > 0, "gru_layer: empty problem");
  // one wave: 8 GPCs x 2 clusters of 8 CTAs = 16 co-resident clusters on the 148 SMs
  if (2 * ((B + 1) / 2) <= 16) return launch<2>(gi, whh_t, bhh, B, T, out, st);
  if (2 * ((B + 3) / 4) <= 16) return launch<4>(gi, whh_t, bhh, B, T, out, st);
  return launch<8>(gi, whh_t, bhh, B, T, out, st);
}

}  // namespace vfx
