// Fused ResStack pair as a TWO-CTA pipeline (sm_100a thread-block cluster of 2): the shapes whose two weight sets do not
// fit one SM next to the operand stages --
//     bf16, C = 128   (2 x 96 KB of weights)
//     tf32, C = 64    (2 x 48 KB of weights, fp32-sized operand / intermediate tiles)
// -- run conv1 on one SM and conv2 on its cluster neighbour:
//
//     front CTA (cluster rank 0)                                   back CTA (cluster rank 1)
//     TMA operand boxes -> tcgen05 conv1 (W1 resident) -> TMEM     loader warp: TMA-loads the h tile it is told about
//     epilogue 1: +b1, lrelu, zero outside [0, L), round ->        tcgen05 conv2 (W2 resident) on that tile
//        h tile in a local staging buffer                          epilogue 2: + b2 + residual (TMA ring) -> x' / act(x')
//     store warp: TMA-stores the tile into a small global            (or the encoded tf32 stream), TMA stores
//        scratch ring (L2-resident) and signals the neighbour
//
//     x' = x + conv2_{k3,d=1}( lrelu_0.01( conv1_{k3,d}( lrelu_0.01(x) ) + b1 ) ) + b2
//     (ResStack.forward voicefixer/vocoder/model/modules.py:592-595, layers :550-576)
//
// The 32 KB intermediate tile travels SM -> L2 -> SM through a per-cluster ring of four scratch slots (74 clusters x 128 KB =
// 9.5 MB, rewritten every few microseconds, so it lives in the 126 MB L2 and HBM sees only x in and x' out); the two CTAs
// synchronise through each other's mbarriers (mapa + mbarrier.arrive.release.cluster).  A first version wrote the tile straight
// into the neighbour's shared memory (st.shared::cluster): numerically right but the SM-to-SM fabric moved it at only ~7-9
// bytes per cycle -- 4.1 ms per tf32 pair, 2.8 ms per bf16 C = 128 pair (that kernel: git history, commit 77349f4).
// MEASURED (B200, B = 32, tools/bench_pair.py --impl 2):  bf16 / fp16 C = 128: 1.62 ms per pair against 1.85-1.95 ms for the two
// launches it replaces -> used by the engine (option fuse_pair2 = 1).  tf32 C = 64: 3.46 ms against 2.98 (small dilations) /
// 3.49 ms (large): no gain -- each SM of the pair is limited by its own L2 -> SM bandwidth (front: 96 KB of operand boxes +
// 32 KB of h out per tile; back: 32 KB of h + residual + outputs), not by HBM -> left off in tf32 (fuse_pair2 = 2 enables it).
// Both variants have the same byte geometry: an operand / h row is C * elem = 256 bytes = two swizzled 128-byte K chunks, one
// MMA advances 32 bytes along K, only N (= C) and the operand format differ.  Tiling, the 126-of-128 row trick, the residual
// ring of epilogue 2 and the in-place rules are those of resstack_pair_tc.cu.
//   bf16: a (bf16 operand copy) in, x (fp32) residual in / x' out in place, a' (bf16) out to the other ping-pong buffer.
//   tf32: S (encoded stream, include/vfx_b200.h) in as operand AND residual, S' (or plain x' for the last pair) out to the
//         OTHER buffer (conv1 reads S with a halo that the in-place update of neighbouring tiles would destroy).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>
#include "vfx_common.cuh"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace vfx {

namespace {

constexpr int QTILE = 126;                      // output positions per tile
constexpr int QNACC = 4;                        // TMEM accumulator stages
constexpr int Q_THREADS = 224;                  // warp 0: TMA loads, warp 1: MMA, warps 2-5: epilogue, warp 6: h-tile store (front)
constexpr int Q_SG = 4;                         // global scratch slots per cluster
constexpr int Q_HTILE = 32 * 1024;              // h tile as stored: 2 K chunks x 128 rows x 128 B
constexpr int Q_STAGE = 32 * 1024;              // one tap box: 2 K chunks x 128 rows x 128 B
constexpr int Q_HSLOT = 2 * 17 * 1024;          // h tile: 2 K chunks x (130 rows x 128 B rounded to swizzle atoms)
constexpr int Q_MAX_STAGES = 6;

struct Pair2Params {
  int B, L, d, n_t;
  uint32_t total_tiles, n_clusters;
  int d_b, d_it;                  // digits of n_clusters in (n_t, B): per-iteration tile increment
  uint32_t stages;                // front: operand stages (one tap box each)
  uint32_t h_slots;               // back: h tiles in shared memory (1 .. 3)
  uint32_t stg_bufs;              // front: h staging buffers (1 or 2)
  const float* bias1; const float* bias2;
  uint32_t has_raw, has_act, res_enc, raw_enc;
  float act_param, enc_slope, enc_inv_slope;
  uint32_t epi_warp_bytes;
  uint32_t idesc;
  uint32_t w_bytes;               // one convolution's weights in shared memory
};

struct QTileIter {
  int b, it;
  __device__ __forceinline__ void init(const Pair2Params& p, uint32_t tile) {
    b = (int)(tile / (uint32_t)p.n_t); it = (int)(tile % (uint32_t)p.n_t);
  }
  __device__ __forceinline__ void next(const Pair2Params& p) {
    it += p.d_it; const int c = it >= p.n_t; it -= c ? p.n_t : 0;
    b += p.d_b + c;
  }
  __device__ __forceinline__ int p0() const { return it * QTILE; }
};

// ---- cluster-scope helpers (distributed shared memory)
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CWAIT_DONE;\n\t"
      "bra CWAIT_LOOP;\n\t"
      "CWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int ACT, bool TF32, bool FP16, int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q_THREADS, 1)
resstack_pair2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                      const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO30,
                      const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmT30,
                      const __grid_constant__ CUtensorMap tmS, const __grid_constant__ Pair2Params p) {
  constexpr int NCH = C / 32;                    // 32-column chunks of an accumulator row
  constexpr int KCH = TF32 ? 32 : 64;            // channels per 128-byte K chunk
  constexpr uint32_t WBLK = C * 128;             // one (tap, K chunk) weight block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // Both CTAs place bias and barriers at the SAME offset so that each can address the other's barriers (mapa):
  //   [W (this CTA's convolution) | region: front = h staging + operand stages, back = h slots + epilogue-2 staging | bias | barriers]
  uint8_t* const ws = smem;
  uint8_t* const region = smem + p.w_bytes;
  uint8_t* const hstg = region;                                        // front: h staging buffers (as stored: 32 KB each)
  uint8_t* const stages0 = region + (size_t)p.stg_bufs * Q_HTILE;      // front: operand stages
  uint8_t* const hs = region;                                          // back: h slots
  uint8_t* const staging = region + (size_t)p.h_slots * Q_HSLOT;       // back: epilogue-2 staging
  const uint32_t front_bytes = p.stg_bufs * (uint32_t)Q_HTILE + p.stages * (uint32_t)Q_STAGE,
                 back_bytes = p.h_slots * (uint32_t)Q_HSLOT + 4u * p.epi_warp_bytes;
  float* const bias_s = reinterpret_cast<float*>(region + (front_bytes > back_bytes ? front_bytes : back_bytes));
  uint64_t* const bars = reinterpret_cast<uint64_t*>(bias_s + C);
  uint64_t* const a_full = bars;                     // front [stages]
  uint64_t* const a_empty = a_full + Q_MAX_STAGES;   // front
  uint64_t* const acc_full = a_empty + Q_MAX_STAGES; // [QNACC] (front: conv1, back: conv2)
  uint64_t* const acc_empty = acc_full + QNACC;
  uint64_t* const h_full = acc_empty + QNACC;        // back [h_slots <= 4]: the TMA load of the tile has landed (32 KB of tx)
  uint64_t* const h_free = h_full + 4;               // back: conv2's MMAs have read the slot (tcgen05.commit)
  uint64_t* const h_ready = h_free + 4;              // back [Q_SG]: the front CTA's store of scratch slot g is complete (remote arrive)
  uint64_t* const sc_free = h_ready + Q_SG;          // front [Q_SG]: the back CTA has loaded scratch slot g (remote arrive)
  uint64_t* const stg_full = sc_free + Q_SG;         // front [2]: epilogue 1 has written staging buffer b (128 arrivals)
  uint64_t* const stg_free = stg_full + 2;           // front [2]: its TMA store has read it
  uint64_t* const wfull = stg_free + 2;
  uint64_t* const res_full = wfull + 1;              // back [4 warps][4 ring slots]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(res_full + 16);

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();        // 0 = front (conv1), 1 = back (conv2)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cl = blockIdx.x >> 1;
  const int sc_row0 = (int)cl * (Q_SG * 256);       // this cluster's first row in the scratch tensor (256 rows of 128 B per slot)

  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < QNACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    for (int a = 0; a < 4; ++a) { mbar_init(&h_full[a], 1); mbar_init(&h_free[a], 1); mbar_init(&h_ready[a], 1); mbar_init(&sc_free[a], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&stg_full[a], 128); mbar_init(&stg_free[a], 1); }
    mbar_init(wfull, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(QNACC * C) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C; i += Q_THREADS) bias_s[i] = rank == 0 ? p.bias1[i] : p.bias2[i];
  if (rank == 1) {     // h slots: rows 128/129 of a chunk are only read by the dropped accumulator rows but must stay finite
    for (uint32_t i = threadIdx.x; i < p.h_slots * (uint32_t)Q_HSLOT / 4; i += Q_THREADS) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();                                    // the peer's barriers are initialised and its h slots zeroed
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_my = cl < p.total_tiles ? (p.total_tiles - cl + p.n_clusters - 1) / p.n_clusters : 0;
  const uint32_t dhi = desc_hi(64u, 2u);             // SBO = 8 rows x 128 B, SWIZZLE_128B
  const uint32_t wa = smem_u32(ws);

  if (rank == 0) {
    // =============================================================== FRONT CTA: conv1
    if (warp == 0) {
      // ---- TMA producer
      if (elect_one()) {
        mbar_expect_tx(wfull, p.w_bytes);
#pragma unroll 1
        for (int blk = 0; blk < 6; ++blk)            // (tap, kc) blocks of [C rows][KCH channels]
          tma_load_2d(&tmW1, wfull, ws + blk * WBLK, (blk & 1) * KCH, (blk >> 1) * C);
      }
      __syncwarp();
      uint32_t s = 0, ph = 0;
      QTileIter it; it.init(p, cl);
      for (uint32_t i = 0; i < n_my; ++i) {
        const int p0 = it.p0(), b = it.b;
        it.next(p);
#pragma unroll 1
        for (int tap = 0; tap < 3; ++tap) {
          mbar_wait(&a_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&a_full[s], Q_STAGE);
            uint8_t* const dst = stages0 + (size_t)s * Q_STAGE;
            const int row = p0 - 1 + (tap - 1) * p.d;
            tma_load_4d(&tmA, &a_full[s], dst, 0, row, 0, b);
            tma_load_4d(&tmA, &a_full[s], dst + 16384, KCH, row, 0, b);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ---- MMA issuer: conv1
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t s = 0, ph = 0;
      for (uint32_t i = 0; i < n_my; ++i) {
        mbar_wait(&acc_empty[i & (QNACC - 1)], ((i / QNACC) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + (i & (QNACC - 1)) * C;
#pragma unroll 1
        for (int tap = 0; tap < 3; ++tap) {
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stages0 + (size_t)s * Q_STAGE);
          if (elect_one()) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              const uint32_t a_lo = desc_lo(sa + kc * 16384u), b_lo = desc_lo(wa + (uint32_t)(tap * 2 + kc) * WBLK);
              if (tap == 0 && kc == 0) tc_mma_lo<false, TF32>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, TF32>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, TF32>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
            }
            tc_commit(&a_empty[s]);
            if (tap == 2) tc_commit(&acc_full[i & (QNACC - 1)]);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    } else {
      // ---- warp 6: store warp.  Takes a finished staging buffer, TMA-stores it into scratch slot g of this cluster and,
      //      once the store is COMPLETE (not merely read), tells the back CTA; frees the staging buffer as soon as it is read.
      if (warp == 6) {
        const uint32_t hready_remote = map_to_rank(smem_u32(h_ready), 1);
        for (uint32_t i = 0; i < n_my; ++i) {
          const uint32_t b = i % p.stg_bufs, g = i & (Q_SG - 1);
          mbar_wait(&stg_full[b], (i / p.stg_bufs) & 1);
          mbar_wait(&sc_free[g], ((i / Q_SG) & 1) ^ 1);         // the back CTA has consumed what this slot held before
          if (lane == 0) {
            const uint8_t* src = hstg + (size_t)b * Q_HTILE;
            tma_store_2d(&tmS, src, 0, sc_row0 + (int)g * 256);
            tma_store_2d(&tmS, src + 16384, 0, sc_row0 + (int)g * 256 + 128);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (p.stg_bufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            // staging buffer of tile i - (stg_bufs - 1) has been read out
            if (i + 1 >= p.stg_bufs) mbar_arrive(&stg_free[(i + 1 - p.stg_bufs) % p.stg_bufs]);
            if (i >= 1) {                                          // tile i-1's store is complete: hand it over
              asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
              __threadfence();
              mbar_arrive_remote(hready_remote + ((i - 1) & (Q_SG - 1)) * 8u);
            }
          }
          __syncwarp();
        }
        if (lane == 0 && n_my > 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          __threadfence();
          mbar_arrive_remote(hready_remote + ((n_my - 1) & (Q_SG - 1)) * 8u);
        }
        __syncwarp();
      } else {
      // ---- epilogue 1 (warps 2-5): conv1 accumulator -> h tile in a local staging buffer (the layout TMA stores / loads)
      const int sub = warp & 3;
      const int r = sub * 32 + lane;                       // accumulator row = h row; position q = p0 - 1 + r
      const uint32_t swz = (uint32_t)(r & 7);
      QTileIter it; it.init(p, cl);
      for (uint32_t i = 0; i < n_my; ++i) {
        const int q = it.p0() - 1 + r;
        it.next(p);
        const bool inside = q >= 0 && q < p.L;
        const uint32_t b = i % p.stg_bufs;
        mbar_wait(&acc_full[i & (QNACC - 1)], (i / QNACC) & 1);
        tc_fence_after();
        mbar_wait(&stg_free[b], ((i / p.stg_bufs) & 1) ^ 1);     // the store that last used this staging buffer has read it
        uint8_t* const hrow = hstg + (size_t)b * Q_HTILE + (size_t)r * 128;
        const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (i & (QNACC - 1)) * C;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32];
          tc_ld32(t_row + c * 32, v);
          const float4* bp = reinterpret_cast<const float4*>(bias_s + c * 32);
          if (TF32) {                                      // 32 channels = one whole 128-byte row of K chunk c
            uint8_t* const base = hrow + (size_t)c * 16384;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bp[j];
              float f0 = __uint_as_float(v[4 * j]) + b4.x, f1 = __uint_as_float(v[4 * j + 1]) + b4.y;
              float f2 = __uint_as_float(v[4 * j + 2]) + b4.z, f3 = __uint_as_float(v[4 * j + 3]) + b4.w;
              f0 = f0 > 0.f ? f0 : f0 * 0.01f; f1 = f1 > 0.f ? f1 : f1 * 0.01f;
              f2 = f2 > 0.f ? f2 : f2 * 0.01f; f3 = f3 > 0.f ? f3 : f3 * 0.01f;
              *reinterpret_cast<float4*>(base + (((uint32_t)j ^ swz) << 4)) =
                  make_float4(inside ? round_tf32(f0) : 0.f, inside ? round_tf32(f1) : 0.f,
                              inside ? round_tf32(f2) : 0.f, inside ? round_tf32(f3) : 0.f);
            }
          } else {                                         // 32 channels = half a row of K chunk c / 2
            uint8_t* const base = hrow + (size_t)(c >> 1) * 16384;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = bp[2 * j], b1 = bp[2 * j + 1];
              float f[8];
              f[0] = __uint_as_float(v[8 * j]) + b0.x; f[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
              f[2] = __uint_as_float(v[8 * j + 2]) + b0.z; f[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
              f[4] = __uint_as_float(v[8 * j + 4]) + b1.x; f[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
              f[6] = __uint_as_float(v[8 * j + 6]) + b1.z; f[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo = f[2 * e], hi = f[2 * e + 1];
                lo = lo > 0.f ? lo : lo * 0.01f; hi = hi > 0.f ? hi : hi * 0.01f;
                w[e] = pack16<FP16>(inside ? lo : 0.f, inside ? hi : 0.f);
              }
              *reinterpret_cast<uint4*>(base + (((uint32_t)((c & 1) * 4 + j) ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[i & (QNACC - 1)]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> the TMA store's async-proxy reads
        mbar_arrive(&stg_full[b]);
      }
      }
    }
  } else {
    // =============================================================== BACK CTA: conv2 + final epilogue
    if (warp == 0) {
      // ---- weights, then the h-tile loader: scratch slot g (stored by the front CTA) -> shared-memory slot s, and
      //      "scratch slot g is free again" back to the front CTA once the load has landed
      if (elect_one()) {
        mbar_expect_tx(wfull, p.w_bytes);
#pragma unroll 1
        for (int blk = 0; blk < 6; ++blk)
          tma_load_2d(&tmW2, wfull, ws + blk * WBLK, (blk & 1) * KCH, (blk >> 1) * C);
      }
      __syncwarp();
      const uint32_t scfree_remote = map_to_rank(smem_u32(sc_free), 0);
      const bool pipelined = p.h_slots >= 2;             // two loads in flight need two shared-memory slots
      for (uint32_t i = 0; i < n_my; ++i) {
        const uint32_t g = i & (Q_SG - 1), sl = i % p.h_slots;
        mbar_wait_cluster(&h_ready[g], (i / Q_SG) & 1);                   // the front CTA's store of tile i is complete
        mbar_wait(&h_free[sl], ((i / p.h_slots) & 1) ^ 1);               // conv2 of the tile that used this slot has read it
        if (!pipelined && i >= 1) {                                        // one slot: conv2(i-1) done => load(i-1) landed long ago
          if (lane == 0) mbar_arrive_remote(scfree_remote + ((i - 1) & (Q_SG - 1)) * 8u);
        }
        if (lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_expect_tx(&h_full[sl], Q_HTILE);
          uint8_t* const dst = hs + (size_t)sl * Q_HSLOT;
          tma_load_2d(&tmS, &h_full[sl], dst, 0, sc_row0 + (int)g * 256);
          tma_load_2d(&tmS, &h_full[sl], dst + 17408, 0, sc_row0 + (int)g * 256 + 128);
        }
        __syncwarp();
        if (pipelined && i >= 1) {                                         // tile i-1's load: landed -> its scratch slot is free
          mbar_wait(&h_full[(i - 1) % p.h_slots], ((i - 1) / p.h_slots) & 1);
          if (lane == 0) mbar_arrive_remote(scfree_remote + ((i - 1) & (Q_SG - 1)) * 8u);
          __syncwarp();
        }
      }
      if (n_my > 0) {
        mbar_wait(&h_full[(n_my - 1) % p.h_slots], ((n_my - 1) / p.h_slots) & 1);
        if (lane == 0) mbar_arrive_remote(scfree_remote + ((n_my - 1) & (Q_SG - 1)) * 8u);
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---- MMA issuer: conv2 on the received h tiles (taps = row-shifted views at offsets 0 / 1 / 2)
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t slot = 0, sph = 0;
      for (uint32_t i = 0; i < n_my; ++i) {
        mbar_wait(&acc_empty[i & (QNACC - 1)], ((i / QNACC) & 1) ^ 1);
        mbar_wait(&h_full[slot], sph);                     // TMA transaction bytes: the tile is in this CTA's shared memory
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (i & (QNACC - 1)) * C;
        const uint32_t ha = smem_u32(hs + (size_t)slot * Q_HSLOT);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              const uint32_t a_lo = desc_lo(ha + kc * 17408u + tap * 128u), b_lo = desc_lo(wa + (uint32_t)(tap * 2 + kc) * WBLK);
              if (tap == 0 && kc == 0) tc_mma_lo<false, TF32>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, TF32>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, TF32>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
            }
          }
          tc_commit(&h_free[slot]);
          tc_commit(&acc_full[i & (QNACC - 1)]);
        }
        __syncwarp();
        if (++slot == p.h_slots) { slot = 0; sph ^= 1; }
      }
    } else if (warp < 6) {
      // ---- epilogue 2: conv2 accumulator + b2 + residual -> outputs (chunk stream with a 4-slot residual ring, as in
      //      resstack_pair_tc.cu)
      const int ew = warp - 2, sub = warp & 3;
      uint8_t* const stg = staging + ew * p.epi_warp_bytes;     // [RO0..RO3 (4 KB each)] [AT0 AT1 (2 KB each, bf16 act output)]
      uint8_t* const at_base = stg + 16384;
      uint64_t* const rfull = res_full + ew * 4;
      const int r0 = sub * 32;
      const CUtensorMap* const mO = sub == 3 ? &tmO30 : &tmO;   // rows 126/127 of a tile belong to the next tile
      const CUtensorMap* const mT = sub == 3 ? &tmT30 : &tmT;
      uint32_t rph = 0;
      QTileIter it; it.init(p, cl);
      QTileIter pit; pit.init(p, cl);
      uint32_t pn = 0, n = 0;
      int pc = 0;
      const uint32_t n_chunks = (uint32_t)NCH * n_my;
#pragma unroll 1
      for (; pn < 3 && pn < n_chunks; ++pn) {
        if (lane == 0) {
          mbar_expect_tx(&rfull[pn & 3], 4096);
          tma_load_4d(&tmR, &rfull[pn & 3], stg + (pn & 3) * 4096, pc * 32, pit.p0() + r0, 0, pit.b);
        }
        if (++pc == NCH) { pc = 0; pit.next(p); }
      }
      const uint32_t sw128 = (uint32_t)(lane & 7) << 4, sw64 = (uint32_t)((lane >> 1) & 3) << 4;
      for (uint32_t i = 0; i < n_my; ++i) {
        const int p0 = it.p0(), b = it.b;
        it.next(p);
        mbar_wait(&acc_full[i & (QNACC - 1)], (i / QNACC) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (i & (QNACC - 1)) * C;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c, ++n) {
          const uint32_t slot = n & 3, k = n & 1;
          uint8_t* const ro = stg + slot * 4096 + lane * 128;
          mbar_wait(&rfull[slot], (rph >> slot) & 1); rph ^= 1u << slot;
          uint32_t v[32];
          tc_ld32(t_row + c * 32, v);
          float f[32];
          {
            const float4* bp = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bp[j];
              float4 r4 = *reinterpret_cast<const float4*>(ro + ((uint32_t)(j << 4) ^ sw128));
              if (TF32 && p.res_enc) {
                r4.x = stream_dec(r4.x, p.enc_inv_slope); r4.y = stream_dec(r4.y, p.enc_inv_slope);
                r4.z = stream_dec(r4.z, p.enc_inv_slope); r4.w = stream_dec(r4.w, p.enc_inv_slope);
              }
              f[4 * j] = __uint_as_float(v[4 * j]) + b4.x + r4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y + r4.y;
              f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z + r4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w + r4.w;
            }
          }
          if (p.has_raw) {
            if (TF32 && p.raw_enc) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(ro + ((uint32_t)(j << 4) ^ sw128)) =
                    make_float4(stream_enc(f[4 * j], p.enc_slope), stream_enc(f[4 * j + 1], p.enc_slope),
                                stream_enc(f[4 * j + 2], p.enc_slope), stream_enc(f[4 * j + 3], p.enc_slope));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(ro + ((uint32_t)(j << 4) ^ sw128)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          }
          if (!TF32 && p.has_act) {
            uint8_t* const at = at_base + k * 2048 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int q = 0; q < 4; ++q)
                w[q] = pack16<FP16>(act_fast<ACT>(f[8 * j + 2 * q], p.act_param), act_fast<ACT>(f[8 * j + 2 * q + 1], p.act_param));
              *reinterpret_cast<uint4*>(at + ((uint32_t)(j << 4) ^ sw64)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (p.has_raw) tma_store_4d(mO, stg + slot * 4096, c * 32, p0 + r0, 0, b);
            if (!TF32 && p.has_act) tma_store_4d(mT, at_base + k * 2048, c * 32, p0 + r0, 0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (pn < n_chunks) {
              mbar_expect_tx(&rfull[pn & 3], 4096);
              tma_load_4d(&tmR, &rfull[pn & 3], stg + (pn & 3) * 4096, pc * 32, pit.p0() + r0, 0, pit.b);
            }
          }
          if (pn < n_chunks) { ++pn; if (++pc == NCH) { pc = 0; pit.next(p); } }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[i & (QNACC - 1)]);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster.sync();           // no CTA may exit while its peer could still write to its shared memory or barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(QNACC * C) : "memory");
  }
}

template <bool TF32, bool FP16, int C>
int launch_pair2(const vfx_pair_desc& d, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("resstack_pair2: cuTensorMapEncodeTiled not available"); return VFX_ERR_CUDA; }
  constexpr int E = TF32 ? 4 : 2;
  constexpr int KCH = TF32 ? 32 : 64;
  Pair2Params p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.L = d.L; p.d = d.dilation;
  p.n_t = ceil_div(d.L, QTILE);
  const long long total = (long long)d.B * p.n_t;
  if (total >= (1LL << 31)) return VFX_ERR_UNSUPPORTED;
  p.total_tiles = (uint32_t)total;
  p.bias1 = d.b1; p.bias2 = d.b2;
  p.has_raw = d.write_raw ? 1u : 0u; p.has_act = (!TF32 && d.out_act) ? 1u : 0u;
  p.res_enc = (TF32 && d.stream_enc) ? 1u : 0u;
  p.raw_enc = (TF32 && d.stream_enc && d.stream_enc_out) ? 1u : 0u;
  p.enc_slope = 0.01f; p.enc_inv_slope = 100.0f;
  p.act_param = d.act_param;
  p.w_bytes = 6u * C * 128u;
  p.epi_warp_bytes = 16384u + (p.has_act ? 4096u : 0u);
  const uint32_t fmt = TF32 ? 2u : FP16 ? 0u : 1u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t budget = 227u * 1024u - 1024u /*align*/ - 1024u /*bias + barriers*/ - p.w_bytes;
  uint32_t h_slots = (budget - 4u * p.epi_warp_bytes) / (uint32_t)Q_HSLOT;
  if (h_slots > 3) h_slots = 3;
  if (h_slots < 1) return VFX_ERR_UNSUPPORTED;
  // front: two staging buffers if three operand stages (one tile's tap boxes) still fit beside them, else one
  uint32_t stg_bufs = budget >= 2u * Q_HTILE + 3u * Q_STAGE ? 2u : 1u;
  uint32_t stages = (budget - stg_bufs * (uint32_t)Q_HTILE) / (uint32_t)Q_STAGE;
  if (stages > (uint32_t)Q_MAX_STAGES) stages = Q_MAX_STAGES;
  if (stages < 3) return VFX_ERR_UNSUPPORTED;
  p.h_slots = h_slots; p.stages = stages; p.stg_bufs = stg_bufs;
  const uint32_t front_bytes = stg_bufs * (uint32_t)Q_HTILE + stages * (uint32_t)Q_STAGE,
                 back_bytes = h_slots * (uint32_t)Q_HSLOT + 4u * p.epi_warp_bytes;
  const size_t smem_bytes = 1024 + p.w_bytes + (front_bytes > back_bytes ? front_bytes : back_bytes) + C * 4 + 512;

  float* const out_raw = d.x_out ? d.x_out : d.x;
  CUtensorMap tmA, tmW1, tmW2, tmR, tmO, tmO30, tmT, tmT30;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  auto enc4 = [&](CUtensorMap* tm, CUtensorMapDataType dt, int esz, const void* base, cuuint32_t box_c, cuuint32_t box_rows,
                  CUtensorMapSwizzle swz) -> CUresult {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d.L, 1, (cuuint64_t)d.B};
    cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)d.L * C * esz, (cuuint64_t)d.L * C * esz};
    cuuint32_t box[4] = {box_c, box_rows, 1, 1};
    return encode(tm, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  const CUtensorMapDataType op_dt = TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  auto encw = [&](CUtensorMap* tm, const void* base) -> CUresult {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)3 * C};
    cuuint64_t strides[1] = {(cuuint64_t)C * E};
    cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)C};
    cuuint32_t es[2] = {1, 1};
    return encode(tm, op_dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = enc4(&tmA, op_dt, E, d.a, KCH, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = encw(&tmW1, d.w1);
  if (r == CUDA_SUCCESS) r = encw(&tmW2, d.w2);
  if (r == CUDA_SUCCESS) r = enc4(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.x, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = enc4(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out_raw, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = enc4(&tmO30, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out_raw, 32, 30, CU_TENSOR_MAP_SWIZZLE_128B);
  tmT = tmO; tmT30 = tmO30;
  if (r == CUDA_SUCCESS && p.has_act) r = enc4(&tmT, op_dt, 2, d.out_act, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r == CUDA_SUCCESS && p.has_act) r = enc4(&tmT30, op_dt, 2, d.out_act, 32, 30, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) { set_error("resstack_pair2: cuTensorMapEncodeTiled failed with %d", (int)r); return VFX_ERR_CUDA; }

  int dev = 0, num_sms = 0;
  VFX_CUDA_CHECK(cudaGetDevice(&dev));
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev]) num_sms = sms_of[dev];
  else {
    VFX_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
#define VFX_P2_ATTR(A) VFX_CUDA_CHECK(cudaFuncSetAttribute(resstack_pair2_kernel<A, TF32, FP16, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    VFX_P2_ATTR(VFX_ACT_NONE); VFX_P2_ATTR(VFX_ACT_LRELU); VFX_P2_ATTR(VFX_ACT_LRELU_XSINX);
#undef VFX_P2_ATTR
    if (dev < 64) sms_of[dev] = num_sms;
  }
  uint32_t n_clusters = (uint32_t)num_sms / 2;
  if (n_clusters > p.total_tiles) n_clusters = p.total_tiles;
  p.n_clusters = n_clusters;
  // global scratch ring: [cluster][slot][2 K chunks][128 rows][128 B], addressed as a 2-D byte tensor of 128-byte rows
  const size_t scratch_need = (size_t)n_clusters * Q_SG * Q_HTILE;
  if (!d.scratch || d.scratch_bytes < scratch_need || ((uintptr_t)d.scratch & 127)) {
    set_error("resstack_pair2: needs %zu bytes of 128-byte aligned scratch (vfx_resstack_pair_scratch_bytes())", scratch_need);
    return VFX_ERR_WORKSPACE;
  }
  CUtensorMap tmS;
  {
    cuuint64_t dims[2] = {128, (cuuint64_t)n_clusters * Q_SG * 256};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult rs = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d.scratch, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rs != CUDA_SUCCESS) { set_error("resstack_pair2: cuTensorMapEncodeTiled(scratch) failed with %d", (int)rs); return VFX_ERR_CUDA; }
  }
  p.d_it = (int)(n_clusters % (uint32_t)p.n_t); p.d_b = (int)(n_clusters / (uint32_t)p.n_t);
  const int act = p.has_act ? d.act : VFX_ACT_NONE;
  switch (act) {
#define VFX_P2_LAUNCH(A) \
  case A: resstack_pair2_kernel<A, TF32, FP16, C><<<2 * n_clusters, Q_THREADS, smem_bytes, st>>>(tmA, tmW1, tmW2, tmR, tmO, tmO30, tmT, tmT30, tmS, p); break
    VFX_P2_LAUNCH(VFX_ACT_NONE); VFX_P2_LAUNCH(VFX_ACT_LRELU); VFX_P2_LAUNCH(VFX_ACT_LRELU_XSINX);
#undef VFX_P2_LAUNCH
    default: set_error("resstack_pair2: unsupported activation %d", act); return VFX_ERR_INVALID;
  }
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace

size_t resstack_pair2_scratch_bytes() { return (size_t)128 * Q_SG * Q_HTILE; }      // enough for 256 SMs

int resstack_pair2_tc(const vfx_pair_desc& d, cudaStream_t st) {
  VFX_REQUIRE(d.a && d.x && d.w1 && d.w2 && d.b1 && d.b2, "resstack_pair2: null argument");
  VFX_REQUIRE(d.B > 0 && d.L > 0 && d.dilation >= 1, "resstack_pair2: empty problem");
  const bool tf32 = d.precision == VFX_PREC_TF32;
  if (((uintptr_t)d.a & 15) || ((uintptr_t)d.x & 15) || ((uintptr_t)d.w1 & 15) || ((uintptr_t)d.w2 & 15) ||
      ((uintptr_t)d.out_act & 15) || ((uintptr_t)d.x_out & 15) || ((uintptr_t)d.b1 & 15) || ((uintptr_t)d.b2 & 15))
    return VFX_ERR_UNSUPPORTED;
  if (tf32) {
    VFX_REQUIRE(d.write_raw, "resstack_pair2: the tf32 form writes the stream tensor");
    VFX_REQUIRE(d.x_out && d.x_out != d.x && (const void*)d.x_out != d.a,
                "resstack_pair2: the tf32 form needs an output buffer that aliases neither input (halo reads)");
    if (d.C == 64) return launch_pair2<true, false, 64>(d, st);
    return VFX_ERR_UNSUPPORTED;
  }
  VFX_REQUIRE(d.out_act != d.a, "resstack_pair2: the activated output must not alias the operand input (halo reads)");
  VFX_REQUIRE(d.write_raw || d.out_act, "resstack_pair2: nothing to write");
  if (d.C == 128) return d.precision == VFX_PREC_FP16 ? launch_pair2<false, true, 128>(d, st) : launch_pair2<false, false, 128>(d, st);
  return VFX_ERR_UNSUPPORTED;        // bf16 C = 64 rows are one 128-byte K chunk: that shape is resstack_pair_tc.cu's
}

}  // namespace vfx
