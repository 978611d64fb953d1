// Fused ResStack pair on the 5th-gen tensor cores (sm_100a, 16-bit operands -- bf16 or fp16 --, C = 64):
//
//     x' = x + conv2_{k3,d=1}( lrelu_0.01( conv1_{k3,d}( lrelu_0.01(x) ) + b1 ) ) + b2
//
// Reference: ResStack.forward voicefixer/vocoder/model/modules.py:592-595 over the layers built at :550-576
// (SURVEY 2.3 K5).  The two convolutions of a pair used to be two launches of conv_gemm_tc with the
// intermediate h = lrelu(conv1 + b1) written to and re-read from HBM as a bf16 tensor; here h never leaves the
// SM: conv1's accumulator (TMEM) is turned into conv2's A operand (a SWIZZLE_128B K-major tile in shared
// memory) by four epilogue warps, and conv2 consumes it through row-shifted descriptor views.
//
//   HBM traffic per element of the pair: a in 2 B + x in 4 B + x' out 4 B + a' out 2 B = 12 B (was 16 B).
//
// * Tile = 126 output positions of one item.  conv1 is computed for the 128 positions q = p0-1 .. p0+126 (one
//   UMMA M = 128 tile); conv2's taps (-1, 0, +1) are the views of the h tile at row offsets 0, 1, 2, so its
//   accumulator rows 0..125 are the outputs p0 .. p0+125 (rows 126/127 read the two pad rows and are dropped).
//   h rows whose position lies outside [0, L) are zero (conv2's zero padding applies to h, not to x).
// * Both weight sets (2 x 3 x [64][64] bf16 = 48 KB) stay resident in shared memory.
// * Roles (persistent CTA per SM, 320 threads):
//     warp 0     TMA producer: one aligned 128-row box of the operand per tap, each its own pipeline stage (the three
//                boxes of a tile overlap in L2; a single halo box with row-shifted tap views measured slower, see below)
//     warp 1     MMA issuer, software-pipelined: conv1(i), then conv2(i-1) while the h tile of i is being produced
//     warp 2-5   epilogue 1: TMEM -> +b1 -> lrelu -> bf16 -> h tile (double-buffered)
//     warp 6-9   epilogue 2: TMA-staged like conv_gemm_tc's: residual tiles in by TMA (ring of four, prefetched three
//                32-column chunks ahead), fp32 result written back in place + activated bf16 operand tile, TMA stores
//                (30-row boxes for the last quarter of a tile)
//   TMEM: 4 accumulator stages of 64 columns for each convolution (512 columns).
// * x is updated in place (a tile's residual rows are its own output rows); the operand copy `a` is read with a halo
//   that neighbouring tiles overwrite, so the activated output goes to a different buffer (ping-pong in the engine).
#include <stdlib.h>
#include <string.h>
#include "vfx_common.cuh"
#include "tc_ptx.cuh"

namespace vfx {

namespace {

constexpr int PC = 64;                         // channels
constexpr int PTILE = 126;                     // output positions per tile
constexpr int PW_BYTES = 3 * PC * PC * 2;      // one convolution's weights: 3 taps x [64][64] bf16 = 24 KB
constexpr int PH_BYTES = 17 * 1024;            // h tile: 130 rows x 128 B, rounded up to whole swizzle atoms
constexpr int PNACC = 4;                       // TMEM accumulator stages per convolution
constexpr int P_THREADS = 320;
constexpr int P_MAX_STAGES = 12;

struct PairParams {
  int B, L, d, n_t;
  uint32_t total_tiles;
  int d_b, d_it;                // mixed-radix digits of gridDim.x in (n_t, B)
  uint32_t halo;                // 1: one box of halo_rows per tile; 0: one 128-row box per tap and stage
  uint32_t halo_rows, stage_bytes, stages;
  const float* bias1; const float* bias2;
  uint32_t has_raw, has_act;
  float act_param;
  uint32_t epi_warp_bytes;
  uint32_t idesc;
};

struct PTile { int b, p0; };
struct PTileIter {
  int b, it;
  __device__ __forceinline__ void init(const PairParams& p, uint32_t tile) {
    b = (int)(tile / (uint32_t)p.n_t); it = (int)(tile % (uint32_t)p.n_t);
  }
  __device__ __forceinline__ void next(const PairParams& p) {
    it += p.d_it; const int c = it >= p.n_t; it -= c ? p.n_t : 0;
    b += p.d_b + c;
  }
  __device__ __forceinline__ bool valid(const PairParams& p) const { return b < p.B; }
  __device__ __forceinline__ PTile coord() const { PTile t; t.b = b; t.p0 = it * PTILE; return t; }
};

template <int ACT, bool FP16>
__global__ void __launch_bounds__(P_THREADS, 1)
resstack_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO30,
                     const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmT30,
                     const __grid_constant__ PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // [W1 | W2 | a stages | h0 h1 | epilogue-2 staging | bias1 bias2 | barriers]
  uint8_t* const w1s = smem;
  uint8_t* const w2s = smem + PW_BYTES;
  uint8_t* const ast = smem + 2 * PW_BYTES;
  uint8_t* const hs = ast + (size_t)p.stages * p.stage_bytes;
  uint8_t* const staging = hs + 2 * PH_BYTES;
  float* const bias1_s = reinterpret_cast<float*>(staging + 4 * p.epi_warp_bytes);
  float* const bias2_s = bias1_s + PC;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(bias2_s + PC);
  uint64_t* const a_full = bars;
  uint64_t* const a_empty = a_full + P_MAX_STAGES;
  uint64_t* const acc1_full = a_empty + P_MAX_STAGES;
  uint64_t* const acc1_empty = acc1_full + PNACC;
  uint64_t* const acc2_full = acc1_empty + PNACC;
  uint64_t* const acc2_empty = acc2_full + PNACC;
  uint64_t* const h_full = acc2_empty + PNACC;       // [2]
  uint64_t* const h_empty = h_full + 2;              // [2]
  uint64_t* const wfull = h_empty + 2;
  uint64_t* const res_full = wfull + 1;              // [4 warps][4 ring slots]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(res_full + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
    for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < PNACC; ++a) {
      mbar_init(&acc1_full[a], 1); mbar_init(&acc1_empty[a], 128);
      mbar_init(&acc2_full[a], 1); mbar_init(&acc2_empty[a], 128);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&h_full[a], 128); mbar_init(&h_empty[a], 1); }
    mbar_init(wfull, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < PC; i += P_THREADS) { bias1_s[i] = p.bias1[i]; bias2_s[i] = p.bias2[i]; }
  // zero the h tiles (rows 128/129 are only ever read by the dropped accumulator rows, but must stay finite) and the
  // rows behind a halo box (read by the last tap's view, never written by TMA)
  for (int i = threadIdx.x; i < 2 * PH_BYTES / 4; i += P_THREADS) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
  if (p.halo) {
    const uint32_t used = p.halo_rows * 128u, padw = (p.stage_bytes - used) / 4;
    if (padw)
      for (uint32_t i = threadIdx.x; i < p.stages * padw; i += P_THREADS)
        reinterpret_cast<uint32_t*>(ast + (size_t)(i / padw) * p.stage_bytes + used)[i % padw] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_my = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, 2 * PW_BYTES);
#pragma unroll 1
      for (int tap = 0; tap < 3; ++tap) {
        tma_load_2d(&tmW1, wfull, w1s + tap * (PC * PC * 2), 0, tap * PC);
        tma_load_2d(&tmW2, wfull, w2s + tap * (PC * PC * 2), 0, tap * PC);
      }
    }
    __syncwarp();
    uint32_t s = 0, ph = 0;
    PTileIter it; it.init(p, blockIdx.x);
    for (uint32_t i = 0; i < n_my; ++i) {
      const PTile t = it.coord();
      it.next(p);
      if (p.halo) {
        mbar_wait(&a_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], p.halo_rows * 128u);
          tma_load_4d(&tmA, &a_full[s], ast + (size_t)s * p.stage_bytes, 0, t.p0 - 1 - p.d, 0, t.b);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      } else {
#pragma unroll 1
        for (int tap = 0; tap < 3; ++tap) {
          mbar_wait(&a_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&a_full[s], 128u * 128u);
            tma_load_4d(&tmA, &a_full[s], ast + (size_t)s * p.stage_bytes, 0, t.p0 - 1 + (tap - 1) * p.d, 0, t.b);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(wfull, 0);
    tc_fence_after();
    uint32_t s = 0, ph = 0;
    const uint32_t dhi = desc_hi(64u /* 8 rows x 128 B >> 4 */, 2u /* SWIZZLE_128B */);
    const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s);
    for (uint32_t i = 0; i <= n_my; ++i) {
      if (i < n_my) {                                     // ---- conv1 of tile i
        mbar_wait(&acc1_empty[i & (PNACC - 1)], ((i / PNACC) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + (i & (PNACC - 1)) * PC;
        if (p.halo) {
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(ast + (size_t)s * p.stage_bytes);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
              const uint32_t a_lo = desc_lo(sa + (uint32_t)(tap * p.d) * 128u), b_lo = desc_lo(w1a + tap * (PC * PC * 2));
              if (tap == 0) tc_mma_lo<false, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, false>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
            }
            tc_commit(&a_empty[s]);
            tc_commit(&acc1_full[i & (PNACC - 1)]);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        } else {
#pragma unroll 1
          for (int tap = 0; tap < 3; ++tap) {
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(ast + (size_t)s * p.stage_bytes);
            if (elect_one()) {
              const uint32_t a_lo = desc_lo(sa), b_lo = desc_lo(w1a + tap * (PC * PC * 2));
              if (tap == 0) tc_mma_lo<false, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, false>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
              tc_commit(&a_empty[s]);
              if (tap == 2) tc_commit(&acc1_full[i & (PNACC - 1)]);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
      if (i > 0) {                                        // ---- conv2 of tile i-1 (its h tile was produced meanwhile)
        const uint32_t j = i - 1;
        mbar_wait(&acc2_empty[j & (PNACC - 1)], ((j / PNACC) & 1) ^ 1);
        mbar_wait(&h_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + PNACC * PC + (j & (PNACC - 1)) * PC;
        const uint32_t ha = smem_u32(hs + (j & 1) * PH_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
            const uint32_t a_lo = desc_lo(ha + tap * 128u), b_lo = desc_lo(w2a + tap * (PC * PC * 2));
            if (tap == 0) tc_mma_lo<false, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
            else tc_mma_lo<true, false>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
            for (int k = 1; k < 4; ++k) tc_mma_lo<true, false>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
          }
          tc_commit(&h_empty[j & 1]);
          tc_commit(&acc2_full[j & (PNACC - 1)]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue 1: conv1 accumulator -> h tile (conv2's A operand) =====================
    const int sub = warp & 3;
    const int r = sub * 32 + lane;                        // accumulator row = h row; position q = p0 - 1 + r
    const uint32_t swz = (uint32_t)(r & 7);
    PTileIter it; it.init(p, blockIdx.x);
    for (uint32_t i = 0; i < n_my; ++i) {
      const PTile t = it.coord();
      it.next(p);
      const int q = t.p0 - 1 + r;
      const bool inside = q >= 0 && q < p.L;
      mbar_wait(&acc1_full[i & (PNACC - 1)], (i / PNACC) & 1);
      tc_fence_after();
      mbar_wait(&h_empty[i & 1], ((i >> 1) & 1) ^ 1);     // conv2 of tile i-2 has finished reading this buffer
      uint8_t* const hrow = hs + (i & 1) * PH_BYTES + r * 128;
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (i & (PNACC - 1)) * PC;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tc_ld32(t_row + c * 32, v);
        const float4* bp = reinterpret_cast<const float4*>(bias1_s + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {                     // 8 channels -> one 16-byte chunk of the swizzled row
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
          *reinterpret_cast<uint4*>(hrow + (((uint32_t)(c * 4 + j) ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&acc1_empty[i & (PNACC - 1)]);
      mbar_arrive(&h_full[i & 1]);
    }
  } else {
    // ===================== epilogue 2: conv2 accumulator + b2 + residual -> x' (fp32) and act(x') (bf16) ============
    // A warp walks the chunk stream n = 2 * tile + c (c = 32-column half).  Its residual tiles arrive by TMA in a ring of
    // four 4 KB buffers, prefetched THREE chunks ahead (a DRAM round trip is longer than one chunk's work; with a
    // one-chunk look-ahead the kernel ran at 4.4 TB/s, latency-bound in this role).  The fp32 result is written back in
    // place and stored from the same buffer; ring slot (n + 3) % 4 = (n - 1) % 4 is free again once the store group of chunk
    // n - 1 has been read out (wait_group.read 1 after committing chunk n's group).
    const int ew = warp - 6, sub = warp & 3;
    uint8_t* const stg = staging + ew * p.epi_warp_bytes;       // [RO0..RO3 (4 KB each)] [AT0 AT1 (2 KB each, optional)]
    uint8_t* const at_base = stg + 16384;
    uint64_t* const rfull = res_full + ew * 4;
    const int r0 = sub * 32;
    const CUtensorMap* const mO = sub == 3 ? &tmO30 : &tmO;     // rows 126/127 of a tile belong to the next tile
    const CUtensorMap* const mT = sub == 3 ? &tmT30 : &tmT;
    uint32_t rph = 0;                                           // bit b = phase of rfull[b]
    PTileIter it; it.init(p, blockIdx.x);
    PTileIter pit; pit.init(p, blockIdx.x);                     // tile of the next chunk to prefetch
    uint32_t pn = 0;                                            // next chunk index to prefetch
    const uint32_t n_chunks = 2 * n_my;
#pragma unroll 1
    for (; pn < 3 && pn < n_chunks; ++pn) {                     // bookkeeping is warp-uniform, lane 0 issues
      if (lane == 0) {
        const PTile t = pit.coord();
        mbar_expect_tx(&rfull[pn & 3], 4096);
        tma_load_4d(&tmR, &rfull[pn & 3], stg + (pn & 3) * 4096, (int)(pn & 1) * 32, t.p0 + r0, 0, t.b);
      }
      if (pn & 1) pit.next(p);
    }
    const uint32_t sw128 = (uint32_t)(lane & 7) << 4, sw64 = (uint32_t)((lane >> 1) & 3) << 4;
    for (uint32_t i = 0; i < n_my; ++i) {
      const PTile t = it.coord();
      it.next(p);
      mbar_wait(&acc2_full[i & (PNACC - 1)], (i / PNACC) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + PNACC * PC + (i & (PNACC - 1)) * PC;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const uint32_t n = 2 * i + (uint32_t)c, slot = n & 3, k = n & 1;
        uint8_t* const ro = stg + slot * 4096 + lane * 128;
        mbar_wait(&rfull[slot], (rph >> slot) & 1); rph ^= 1u << slot;
        uint32_t v[32];
        tc_ld32(t_row + c * 32, v);
        float f[32];
        {
          const float4* bp = reinterpret_cast<const float4*>(bias2_s + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bp[j];
            const float4 r4 = *reinterpret_cast<const float4*>(ro + ((uint32_t)(j << 4) ^ sw128));
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x + r4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y + r4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z + r4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w + r4.w;
          }
        }
        if (p.has_raw) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(ro + ((uint32_t)(j << 4) ^ sw128)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        if (p.has_act) {
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
          if (p.has_raw) tma_store_4d(mO, stg + slot * 4096, c * 32, t.p0 + r0, 0, t.b);
          if (p.has_act) tma_store_4d(mT, at_base + k * 2048, c * 32, t.p0 + r0, 0, t.b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the stores of chunk n-1 have left their buffers
          if (pn < n_chunks) {                                              // residual of chunk n+3 -> the slot chunk n-1 used
            const PTile tp = pit.coord();
            mbar_expect_tx(&rfull[pn & 3], 4096);
            tma_load_4d(&tmR, &rfull[pn & 3], stg + (pn & 3) * 4096, (int)(pn & 1) * 32, tp.p0 + r0, 0, tp.b);
          }
        }
        if (pn < n_chunks) { if (pn & 1) pit.next(p); ++pn; }                // warp-uniform bookkeeping of the prefetch stream
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&acc2_empty[i & (PNACC - 1)]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int resstack_pair_tc(const vfx_pair_desc& d, cudaStream_t st) {
  if (d.C != PC) return VFX_ERR_UNSUPPORTED;
  VFX_REQUIRE(d.a && d.x && d.w1 && d.w2 && d.b1 && d.b2, "resstack_pair: null argument");
  VFX_REQUIRE(d.B > 0 && d.L > 0 && d.dilation >= 1, "resstack_pair: empty problem");
  VFX_REQUIRE(d.out_act != d.a, "resstack_pair: the activated output must not alias the operand input (halo reads)");
  VFX_REQUIRE(d.write_raw || d.out_act, "resstack_pair: nothing to write");
  if (((uintptr_t)d.a & 15) || ((uintptr_t)d.x & 15) || ((uintptr_t)d.w1 & 15) || ((uintptr_t)d.w2 & 15) ||
      ((uintptr_t)d.out_act & 15) || ((uintptr_t)d.b1 & 15) || ((uintptr_t)d.b2 & 15))
    return VFX_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("resstack_pair: cuTensorMapEncodeTiled not available"); return VFX_ERR_CUDA; }

  PairParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.L = d.L; p.d = d.dilation;
  p.n_t = ceil_div(d.L, PTILE);
  const long long total = (long long)d.B * p.n_t;
  if (total >= (1LL << 31)) return VFX_ERR_UNSUPPORTED;
  p.total_tiles = (uint32_t)total;
  p.bias1 = d.b1; p.bias2 = d.b2;
  p.has_raw = d.write_raw ? 1u : 0u; p.has_act = d.out_act ? 1u : 0u;
  p.act_param = d.act_param;
  p.epi_warp_bytes = 16384u + (d.out_act ? 4096u : 0u);
  // c = F32, a = b = BF16 (1) or F16 (0), K-major, N = 64, M = 128
  const bool fp16 = d.precision == VFX_PREC_FP16;
  const uint32_t fmt = fp16 ? 0u : 1u;
  const CUtensorMapDataType op_dt = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(PC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // One box per tap by default: measured on B200 (tools/bench_pair.py, B = 32) the single halo box with row-shifted tap
  // views is SLOWER than three aligned 128-row boxes -- d = 1: 1.84 vs 1.72 ms, d = 27: 1.89 vs 1.68 ms (the L2 absorbs
  // the 3x operand re-reads; every MMA of conv1 then reads an 8-row-aligned tile).  VFX_PAIR_HALO=1 selects the halo box.
  static const bool want_halo = getenv("VFX_PAIR_HALO") != nullptr;
  p.halo = (want_halo && d.dilation <= 64) ? 1u : 0u;
  p.halo_rows = 128u + 2u * (uint32_t)d.dilation;
  p.stage_bytes = p.halo ? ((p.halo_rows + 2u) * 128u + 1023u) / 1024u * 1024u : 128u * 128u;
  const uint32_t fixed = 2u * PW_BYTES + 2u * PH_BYTES + 4u * p.epi_warp_bytes + 2u * PC * 4u + 1024u /*align*/ + 512u /*barriers*/;
  const uint32_t budget = 227u * 1024u;
  uint32_t stages = (budget - fixed) / p.stage_bytes;
  if (stages > (uint32_t)P_MAX_STAGES) stages = P_MAX_STAGES;
  if (stages < (p.halo ? 2u : 3u)) return VFX_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem_bytes = (size_t)fixed + (size_t)stages * p.stage_bytes;

  CUtensorMap tmA, tmW1, tmW2, tmR, tmO, tmO30, tmT, tmT30;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  auto enc4 = [&](CUtensorMap* tm, CUtensorMapDataType dt, int esz, const void* base, cuuint32_t box_c, cuuint32_t box_rows,
                  CUtensorMapSwizzle swz) -> CUresult {
    cuuint64_t dims[4] = {(cuuint64_t)PC, (cuuint64_t)d.L, 1, (cuuint64_t)d.B};
    cuuint64_t strides[3] = {(cuuint64_t)PC * esz, (cuuint64_t)d.L * PC * esz, (cuuint64_t)d.L * PC * esz};
    cuuint32_t box[4] = {box_c, box_rows, 1, 1};
    return encode(tm, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  auto encw = [&](CUtensorMap* tm, const void* base) -> CUresult {
    cuuint64_t dims[2] = {(cuuint64_t)PC, (cuuint64_t)3 * PC};
    cuuint64_t strides[1] = {(cuuint64_t)PC * 2};
    cuuint32_t box[2] = {(cuuint32_t)PC, (cuuint32_t)PC};
    cuuint32_t es[2] = {1, 1};
    return encode(tm, op_dt, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = enc4(&tmA, op_dt, 2, d.a, PC, p.halo ? p.halo_rows : 128u, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = encw(&tmW1, d.w1);
  if (r == CUDA_SUCCESS) r = encw(&tmW2, d.w2);
  if (r == CUDA_SUCCESS) r = enc4(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.x, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = enc4(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.x, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = enc4(&tmO30, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.x, 32, 30, CU_TENSOR_MAP_SWIZZLE_128B);
  tmT = tmO; tmT30 = tmO30;
  if (r == CUDA_SUCCESS && d.out_act) r = enc4(&tmT, op_dt, 2, d.out_act, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r == CUDA_SUCCESS && d.out_act) r = enc4(&tmT30, op_dt, 2, d.out_act, 32, 30, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) { set_error("resstack_pair: cuTensorMapEncodeTiled failed with %d", (int)r); return VFX_ERR_CUDA; }

  int dev = 0, num_sms = 0;
  VFX_CUDA_CHECK(cudaGetDevice(&dev));
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev]) num_sms = sms_of[dev];
  else {
    VFX_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
#define VFX_PAIR_ATTR(A)                                                                                                          \
  VFX_CUDA_CHECK(cudaFuncSetAttribute(resstack_pair_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
  VFX_CUDA_CHECK(cudaFuncSetAttribute(resstack_pair_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    VFX_PAIR_ATTR(VFX_ACT_NONE); VFX_PAIR_ATTR(VFX_ACT_LRELU); VFX_PAIR_ATTR(VFX_ACT_LRELU_XSINX);
#undef VFX_PAIR_ATTR
    if (dev < 64) sms_of[dev] = num_sms;
  }
  const int grid = (int)(p.total_tiles < (uint32_t)num_sms ? p.total_tiles : (uint32_t)num_sms);
  p.d_it = grid % p.n_t; p.d_b = grid / p.n_t;
  const int act = d.out_act ? d.act : VFX_ACT_NONE;
  switch (act) {
#define VFX_PAIR_LAUNCH(A)                                                                                                 \
  case A:                                                                                                                  \
    if (fp16) resstack_pair_kernel<A, true><<<grid, P_THREADS, smem_bytes, st>>>(tmA, tmW1, tmW2, tmR, tmO, tmO30, tmT, tmT30, p);  \
    else resstack_pair_kernel<A, false><<<grid, P_THREADS, smem_bytes, st>>>(tmA, tmW1, tmW2, tmR, tmO, tmO30, tmT, tmT30, p);      \
    break
    VFX_PAIR_LAUNCH(VFX_ACT_NONE); VFX_PAIR_LAUNCH(VFX_ACT_LRELU); VFX_PAIR_LAUNCH(VFX_ACT_LRELU_XSINX);
#undef VFX_PAIR_LAUNCH
    default: set_error("resstack_pair: unsupported activation %d", act); return VFX_ERR_INVALID;
  }
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
