// Fused ResStack pair in tf32 on ONE SM (sm_100a, tcgen05 kind::tf32, C = 64), encoded stream in and out:
//
//     x' = x + conv2_{k3,d=1}( lrelu_0.01( conv1_{k3,d}( lrelu_0.01(x) ) + b1 ) ) + b2
//     (ResStack.forward voicefixer/vocoder/model/modules.py:592-595, layers :550-576)
//
// The input is the encoded stream S = bits(lrelu(x)) + 0x1000 (include/vfx_b200.h): as a kind::tf32 operand the tensor core
// truncates it to round-to-nearest tf32, and as a 32-bit word it still carries x exactly.  The 16-bit pair kernel
// (resstack_pair_tc.cu) re-reads its fp32 residual through a TMA ring; with fp32-sized tiles there is no shared memory left
// for that (2 x 48 KB of weights + 34 KB of h + the operand ring), so here the residual never makes a second trip at all:
//
//   * the operand boxes that hold the tile's own rows (the centre tap, or the whole halo box) are loaded FIRST; four
//     "stash" warps copy the rows p0 .. p0+127 of those boxes from shared memory into spare TENSOR MEMORY columns
//     (tcgen05.st) while the MMAs read the same boxes, and release the ring slot;
//   * epilogue 2 reads the conv2 accumulator AND the stashed residual from TMEM (two tcgen05.ld), decodes, adds, encodes.
//
//   HBM traffic per element of the pair: S in 4 B + S' out 4 B = 8 B = SURVEY 8d's algorithmic figure (two launches: 20 B).
//
// Roles (persistent CTA per SM, 448 threads):
//     warp 0       TMA producer: ring of operand slots, one slot = one 128-byte K chunk (32 channels) of one box
//                    d <= 27: two halo boxes of 128 + 2d rows per tile (taps = row-shifted descriptor views);
//                    else:   six 128-row boxes per tile (tap 1 first -- it carries the residual --, then taps 0 and 2)
//     warp 1       MMA issuer, software-pipelined: conv1(i), then conv2(i-1) on the h tile epilogue 1 produced meanwhile
//     warps 2-5    epilogue 1: conv1 accumulator -> +b1 -> lrelu -> round to tf32 -> h tile (SWIZZLE_128B, two K chunks)
//     warps 6-9    epilogue 2: accumulator + b2 + decoded residual -> S' (or plain x' for the last pair), TMA stores
//     warps 10-13  residual stash: shared memory -> TMEM, then the slot's second "empty" arrival
//   TMEM (512 columns): conv1 accumulator x2 | conv2 accumulator x2 | residual stash x4.
// MEASURED (B200, B = 32 x 443 646 positions, tools/bench_pair.py --prec tf32 --impl 3): 1.55-1.60 ms per pair with halo boxes
// (d <= 27), 1.9-2.3 ms with aligned boxes (three times the TMA bytes into shared memory; run-to-run spread with the buffer
// addresses), against 2.93 / 3.26 ms for the two launches; DRAM traffic 7.22 GB per pair = the 8 B-per-element figure
// (profiles/r02_pair3_tf32.txt).  What paces it is shared-memory bandwidth: an N = 64 tf32 MMA fetches 6 KB of operands per
// 32 cycles, so the 48 MMAs of a tile plus TMA writes, the h tile, the stash read and the store staging keep the 128 B/cycle
// port busy for ~3 600 of the tile's ~4 000 cycles (tensor pipe 36 % active).  Tried without gain: eight warps for either
// epilogue, an L2 prefetch (cp.async.bulk.prefetch.tensor) of the tiles ahead.
// Tiling (126 outputs per 128-row MMA tile, h rows outside [0, L) zeroed, 30-row store boxes for the last quarter) is that of
// resstack_pair_tc.cu.  The output goes to a different buffer than the input (neighbouring tiles read S with a halo).
#include <stdlib.h>
#include <string.h>
#include "vfx_common.cuh"
#include "tc_ptx.cuh"

namespace vfx {

namespace {

constexpr int RC = 64;                          // channels
constexpr int RTILE = 126;                      // output positions per tile
constexpr int R_WBLK = RC * 128;                // one (tap, K chunk) weight block: [64 out channels][32 in channels] tf32
constexpr int R_WBYTES = 6 * R_WBLK;            // one convolution's weights: 48 KB
constexpr int R_HPANEL = 17 * 1024;             // one K chunk of the h tile: 130 rows x 128 B, rounded up to swizzle atoms
constexpr int R_MAX_SLOTS = 8;
constexpr int R_ACC1 = 0, R_ACC2 = 128, R_STASH = 256;   // TMEM column bases

struct Pair3Params {
  int B, L, d, n_t;
  uint32_t total_tiles;
  int d_b, d_it;                  // mixed-radix digits of gridDim.x in (n_t, B)
  uint32_t halo;                  // 1: two boxes of halo_rows per tile; 0: six 128-row boxes per tile
  uint32_t halo_rows, slot_bytes, slots;
  const float* bias1; const float* bias2;
  uint32_t raw_enc;               // 1: write S' (encoded), 0: write plain x'
  float enc_slope, enc_inv_slope;
  uint32_t idesc;
};

struct RTileIter {
  int b, it;
  __device__ __forceinline__ void init(const Pair3Params& p, uint32_t tile) {
    b = (int)(tile / (uint32_t)p.n_t); it = (int)(tile % (uint32_t)p.n_t);
  }
  __device__ __forceinline__ void next(const Pair3Params& p) {
    it += p.d_it; const int c = it >= p.n_t; it -= c ? p.n_t : 0;
    b += p.d_b + c;
  }
  __device__ __forceinline__ int p0() const { return it * RTILE; }
};

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// E1W / E2W: warps of epilogue 1 / epilogue 2 (4: one warp per TMEM lane quarter does both 32-column halves;
// 8: two warps per quarter, one half each)
template <int E1W, int E2W>
__global__ void __launch_bounds__((10 + E1W + E2W) * 32 - 128, 1)
resstack_pair3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                      const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmO30, const __grid_constant__ Pair3Params p) {
  // every byte of the 227 KB is spoken for: no slack to round the base up, so the declaration carries the alignment the
  // swizzled tiles need (and the kernel refuses to run if the driver did not honour it)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  // [W1 | W2 | h (2 K chunks) | operand ring | epilogue-2 staging (4 KB per warp) | bias1 bias2 | barriers]
  uint8_t* const w1s = smem;
  uint8_t* const w2s = smem + R_WBYTES;
  uint8_t* const hs = smem + 2 * R_WBYTES;
  uint8_t* const ring = hs + 2 * R_HPANEL;
  uint8_t* const staging = ring + (size_t)p.slots * p.slot_bytes;
  float* const bias1_s = reinterpret_cast<float*>(staging + E2W * 4096);
  constexpr int R_THREADS = (6 + E1W + E2W) * 32;
  constexpr int W_E2 = 2 + E1W, W_ST = 2 + E1W + E2W;      // first warp of epilogue 2 / of the stash role
  float* const bias2_s = bias1_s + RC;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(bias2_s + RC);
  uint64_t* const a_full = bars;                         // [slots] TMA transaction bytes
  uint64_t* const a_empty = a_full + R_MAX_SLOTS;        // [slots] 1 tcgen05.commit + 4 stash warps
  uint64_t* const acc1_full = a_empty + R_MAX_SLOTS;     // [2]
  uint64_t* const acc1_empty = acc1_full + 2;            // [2] 128 epilogue-1 threads
  uint64_t* const acc2_full = acc1_empty + 2;            // [2]
  uint64_t* const acc2_empty = acc2_full + 2;            // [2] 256 epilogue-2 threads
  uint64_t* const h_full = acc2_empty + 2;               // 128 epilogue-1 threads
  uint64_t* const h_empty = h_full + 1;                  // conv2's MMAs have read the h tile (tcgen05.commit)
  uint64_t* const st_full = h_empty + 1;                 // [4] 128 stash threads
  uint64_t* const st_empty = st_full + 4;                // [4] 256 epilogue-2 threads
  uint64_t* const wfull = st_empty + 4;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    for (uint32_t s = 0; s < p.slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 5); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc1_full[a], 1); mbar_init(&acc1_empty[a], 32 * E1W);
      mbar_init(&acc2_full[a], 1); mbar_init(&acc2_empty[a], 32 * E2W);
    }
    mbar_init(h_full, 32 * E1W); mbar_init(h_empty, 1);
    for (int a = 0; a < 4; ++a) { mbar_init(&st_full[a], 128); mbar_init(&st_empty[a], 32 * E2W); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < RC; i += R_THREADS) { bias1_s[i] = p.bias1[i]; bias2_s[i] = p.bias2[i]; }
  // h rows 128 / 129 are only ever read by the two dropped accumulator rows, but must stay finite
  for (int i = threadIdx.x; i < 2 * R_HPANEL / 4; i += R_THREADS) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_my = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, 2 * R_WBYTES);
#pragma unroll 1
      for (int blk = 0; blk < 6; ++blk) {            // (tap, K chunk) blocks of [64 rows][32 channels]
        tma_load_2d(&tmW1, wfull, w1s + blk * R_WBLK, (blk & 1) * 32, (blk >> 1) * RC);
        tma_load_2d(&tmW2, wfull, w2s + blk * R_WBLK, (blk & 1) * 32, (blk >> 1) * RC);
      }
    }
    __syncwarp();
    uint32_t s = 0, ph = 0;
    RTileIter it; it.init(p, blockIdx.x);
    const int nbox = p.halo ? 2 : 6;
    const uint32_t box_bytes = p.halo ? p.halo_rows * 128u : 128u * 128u;
    for (uint32_t i = 0; i < n_my; ++i) {
      const int p0 = it.p0(), b = it.b;
      it.next(p);

#pragma unroll 1
      for (int q = 0; q < nbox; ++q) {
        // per-tap order: tap 1 (k0, k1), tap 0 (k0, k1), tap 2 (k0, k1)
        const int tap = q < 2 ? 1 : (q < 4 ? 0 : 2);
        const int row = p.halo ? p0 - 1 - p.d : p0 - 1 + (tap - 1) * p.d;
        mbar_wait(&a_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], box_bytes);
          tma_load_4d(&tmA, &a_full[s], ring + (size_t)s * p.slot_bytes, (q & 1) * 32, row, 0, b);
        }
        __syncwarp();
        if (++s == p.slots) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(wfull, 0);
    tc_fence_after();
    uint32_t s = 0, ph = 0;
    const uint32_t dhi = desc_hi(64u /* 8 rows x 128 B >> 4 */, 2u /* SWIZZLE_128B */);
    const uint32_t w1a = smem_u32(w1s), w2a = smem_u32(w2s), ha = smem_u32(hs);
    for (uint32_t i = 0; i <= n_my; ++i) {
      if (i < n_my) {                                     // ---- conv1 of tile i
        mbar_wait(&acc1_empty[i & 1], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + R_ACC1 + (i & 1) * RC;
        if (p.halo) {
#pragma unroll 1
          for (int kc = 0; kc < 2; ++kc) {
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + (size_t)s * p.slot_bytes);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 3; ++tap) {
                const uint32_t a_lo = desc_lo(sa + (uint32_t)(tap * p.d) * 128u), b_lo = desc_lo(w1a + (uint32_t)(tap * 2 + kc) * R_WBLK);
                if (tap == 0 && kc == 0) tc_mma_lo<false, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
                else tc_mma_lo<true, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
                for (int k = 1; k < 4; ++k) tc_mma_lo<true, true>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
              }
              tc_commit(&a_empty[s]);
              if (kc == 1) tc_commit(&acc1_full[i & 1]);
            }
            __syncwarp();
            if (++s == p.slots) { s = 0; ph ^= 1; }
          }
        } else {
#pragma unroll 1
          for (int q = 0; q < 6; ++q) {
            const int tap = q < 2 ? 1 : (q < 4 ? 0 : 2), kc = q & 1;
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + (size_t)s * p.slot_bytes);
            if (elect_one()) {
              const uint32_t a_lo = desc_lo(sa), b_lo = desc_lo(w1a + (uint32_t)(tap * 2 + kc) * R_WBLK);
              if (q == 0) tc_mma_lo<false, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, true>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
              tc_commit(&a_empty[s]);
              if (q == 5) tc_commit(&acc1_full[i & 1]);
            }
            __syncwarp();
            if (++s == p.slots) { s = 0; ph ^= 1; }
          }
        }
      }
      if (i > 0) {                                        // ---- conv2 of tile i-1 (its h tile was produced meanwhile)
        const uint32_t j = i - 1;
        mbar_wait(&acc2_empty[j & 1], ((j >> 1) & 1) ^ 1);
        mbar_wait(h_full, j & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + R_ACC2 + (j & 1) * RC;
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              const uint32_t a_lo = desc_lo(ha + kc * R_HPANEL + tap * 128u), b_lo = desc_lo(w2a + (uint32_t)(tap * 2 + kc) * R_WBLK);
              if (tap == 0 && kc == 0) tc_mma_lo<false, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
              else tc_mma_lo<true, true>(d_tmem, a_lo, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) tc_mma_lo<true, true>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, dhi, p.idesc);
            }
          }
          tc_commit(h_empty);
          tc_commit(&acc2_full[j & 1]);
        }
        __syncwarp();
      }
    }
  } else if (warp < W_E2) {
    // ===================== epilogue 1: conv1 accumulator -> h tile (conv2's A operand) =====================
    const int sub = warp & 3;
    const int c_lo = E1W == 8 ? (warp - 2) >> 2 : 0, c_hi = E1W == 8 ? c_lo + 1 : 2;   // this warp's 32-column halves
    const int r = sub * 32 + lane;                        // accumulator row = h row; position q = p0 - 1 + r
    const uint32_t swz = (uint32_t)(r & 7);
    uint8_t* const hrow = hs + (size_t)r * 128;
    RTileIter it; it.init(p, blockIdx.x);
    for (uint32_t i = 0; i < n_my; ++i) {
      const int q = it.p0() - 1 + r;
      it.next(p);
      const bool inside = q >= 0 && q < p.L;
      mbar_wait(&acc1_full[i & 1], (i >> 1) & 1);
      mbar_wait(h_empty, (i & 1) ^ 1);                    // conv2 of tile i-1 has finished reading the h tile
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + R_ACC1 + (i & 1) * RC;
      // h = rna_tf32(lrelu(acc + b1)): lrelu as max(f, 0.01 f); the rounding as "+ half an ulp" on the bits -- the tensor
      // core drops the 13 low mantissa bits of a kind::tf32 operand, which completes cvt.rna (4 instructions per element)
      const bool all_in = __all_sync(0xffffffffu, inside);   // rows outside [0, L) are conv2's zero padding: edge tiles only
      const uint32_t half_ulp = inside ? 0x1000u : 0u;
      const float keep = inside ? 1.f : 0.f;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {                 // 32 channels = one whole 128-byte row of K chunk c
        uint32_t v[32];
        tc_ld32(t_row + c * 32, v);
        const float4* bp = reinterpret_cast<const float4*>(bias1_s + c * 32);
        uint8_t* const base = hrow + (size_t)c * R_HPANEL;
        if (all_in) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bp[j];
            float f0 = __uint_as_float(v[4 * j]) + b4.x, f1 = __uint_as_float(v[4 * j + 1]) + b4.y;
            float f2 = __uint_as_float(v[4 * j + 2]) + b4.z, f3 = __uint_as_float(v[4 * j + 3]) + b4.w;
            f0 = fmaxf(f0, f0 * 0.01f); f1 = fmaxf(f1, f1 * 0.01f);
            f2 = fmaxf(f2, f2 * 0.01f); f3 = fmaxf(f3, f3 * 0.01f);
            *reinterpret_cast<uint4*>(base + (((uint32_t)j ^ swz) << 4)) =
                make_uint4(__float_as_uint(f0) + 0x1000u, __float_as_uint(f1) + 0x1000u,
                           __float_as_uint(f2) + 0x1000u, __float_as_uint(f3) + 0x1000u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bp[j];
            float f0 = __uint_as_float(v[4 * j]) + b4.x, f1 = __uint_as_float(v[4 * j + 1]) + b4.y;
            float f2 = __uint_as_float(v[4 * j + 2]) + b4.z, f3 = __uint_as_float(v[4 * j + 3]) + b4.w;
            f0 = fmaxf(f0, f0 * 0.01f) * keep; f1 = fmaxf(f1, f1 * 0.01f) * keep;
            f2 = fmaxf(f2, f2 * 0.01f) * keep; f3 = fmaxf(f3, f3 * 0.01f) * keep;
            *reinterpret_cast<uint4*>(base + (((uint32_t)j ^ swz) << 4)) =
                make_uint4(__float_as_uint(f0) + half_ulp, __float_as_uint(f1) + half_ulp,
                           __float_as_uint(f2) + half_ulp, __float_as_uint(f3) + half_ulp);
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&acc1_empty[i & 1]);
      mbar_arrive(h_full);
    }
  } else if (warp < W_ST) {
    // ===================== epilogue 2: conv2 accumulator + b2 + residual (TMEM stash) -> S' / x' =====================
    const int ew = warp - W_E2, sub = warp & 3;
    const int c_lo = E2W == 8 ? ew >> 2 : 0, c_hi = E2W == 8 ? c_lo + 1 : 2;   // this warp's 32-column halves
    uint8_t* const stg = staging + ew * 4096;
    uint8_t* const ro = stg + lane * 128;
    const int r0 = sub * 32;
    const CUtensorMap* const mO = sub == 3 ? &tmO30 : &tmO;     // rows 126 / 127 of a tile belong to the next tile
    const uint32_t sw128 = (uint32_t)(lane & 7) << 4;
    RTileIter it; it.init(p, blockIdx.x);
    for (uint32_t i = 0; i < n_my; ++i) {
      const int p0 = it.p0(), b = it.b;
      it.next(p);
      mbar_wait(&acc2_full[i & 1], (i >> 1) & 1);
      mbar_wait(&st_full[i & 3], (i >> 2) & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(sub * 32) << 16);
      const uint32_t t_acc = lane_base + R_ACC2 + (i & 1) * RC, t_res = lane_base + R_STASH + (i & 3) * RC;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[32], s[32];
        tc_ld32(t_acc + c * 32, v);
        tc_ld32(t_res + c * 32, s);
        if (c == c_hi - 1) {                             // all TMEM reads are in registers: hand the columns back early
          tc_fence_before();
          mbar_arrive(&acc2_empty[i & 1]);
          mbar_arrive(&st_empty[i & 3]);
        }
        const float4* bp = reinterpret_cast<const float4*>(bias2_s + c * 32);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = bp[j];
          float f0 = __uint_as_float(v[4 * j]) + b4.x + stream_dec(__uint_as_float(s[4 * j]), p.enc_inv_slope);
          float f1 = __uint_as_float(v[4 * j + 1]) + b4.y + stream_dec(__uint_as_float(s[4 * j + 1]), p.enc_inv_slope);
          float f2 = __uint_as_float(v[4 * j + 2]) + b4.z + stream_dec(__uint_as_float(s[4 * j + 2]), p.enc_inv_slope);
          float f3 = __uint_as_float(v[4 * j + 3]) + b4.w + stream_dec(__uint_as_float(s[4 * j + 3]), p.enc_inv_slope);
          if (p.raw_enc) {
            f0 = stream_enc(f0, p.enc_slope); f1 = stream_enc(f1, p.enc_slope);
            f2 = stream_enc(f2, p.enc_slope); f3 = stream_enc(f3, p.enc_slope);
          }
          *reinterpret_cast<float4*>(ro + ((uint32_t)(j << 4) ^ sw128)) = make_float4(f0, f1, f2, f3);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(mO, stg, c * 32, p0 + r0, 0, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ===================== residual stash: the tile's own rows, shared memory -> TMEM =====================
    const int sub = warp & 3;
    const int r = sub * 32 + lane;                        // stash lane = conv2 accumulator row: position p0 + r
    // box row of position p0 + r: the halo box starts at p0 - 1 - d, the centre-tap box at p0 - 1 (its row 128 does not
    // exist; accumulator rows 126 / 127 are dropped anyway)
    const int rrow = p.halo ? r + 1 + p.d : (r + 1 < 128 ? r + 1 : 127);
    const uint32_t swz = (uint32_t)(rrow & 7);
    uint32_t s = 0, ph = 0;
    const int rest = p.halo ? 0 : 4;
    for (uint32_t i = 0; i < n_my; ++i) {
      mbar_wait(&st_empty[i & 3], ((i >> 2) & 1) ^ 1);
      tc_fence_after();
      const uint32_t t_res = tmem_base + ((uint32_t)(sub * 32) << 16) + R_STASH + (i & 3) * RC;
#pragma unroll 1
      for (int kc = 0; kc < 2; ++kc) {
        mbar_wait(&a_full[s], ph);
        const uint8_t* const row = ring + (size_t)s * p.slot_bytes + (size_t)rrow * 128;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(row + (((uint32_t)j ^ swz) << 4));
          v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
        }
        tc_st32(t_res + kc * 32, v);
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[s]);
        if (++s == p.slots) { s = 0; ph ^= 1; }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&st_full[i & 3]);
#pragma unroll 1
      for (int q = 0; q < rest; ++q) {                    // the other taps' slots: nothing to copy, only the arrival
        mbar_wait(&a_full[s], ph);
        if (lane == 0) mbar_arrive(&a_empty[s]);
        if (++s == p.slots) { s = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int resstack_pair3_tc(const vfx_pair_desc& d, cudaStream_t st) {
  if (d.C != RC || d.precision != VFX_PREC_TF32) return VFX_ERR_UNSUPPORTED;
  VFX_REQUIRE(d.a && d.x && d.w1 && d.w2 && d.b1 && d.b2, "resstack_pair3: null argument");
  VFX_REQUIRE(d.B > 0 && d.L > 0 && d.dilation >= 1, "resstack_pair3: empty problem");
  VFX_REQUIRE(d.write_raw, "resstack_pair3: the tf32 form writes the stream tensor");
  // the residual is taken from the operand boxes: one encoded tensor is both
  if (!d.stream_enc || d.a != (const void*)d.x) return VFX_ERR_UNSUPPORTED;
  VFX_REQUIRE(d.x_out && d.x_out != d.x, "resstack_pair3: needs an output buffer that does not alias the input (halo reads)");
  if (((uintptr_t)d.a & 15) || ((uintptr_t)d.w1 & 15) || ((uintptr_t)d.w2 & 15) || ((uintptr_t)d.x_out & 15) ||
      ((uintptr_t)d.b1 & 15) || ((uintptr_t)d.b2 & 15))
    return VFX_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("resstack_pair3: cuTensorMapEncodeTiled not available"); return VFX_ERR_CUDA; }

  Pair3Params p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.L = d.L; p.d = d.dilation;
  p.n_t = ceil_div(d.L, RTILE);
  const long long total = (long long)d.B * p.n_t;
  if (total >= (1LL << 31)) return VFX_ERR_UNSUPPORTED;
  p.total_tiles = (uint32_t)total;
  p.bias1 = d.b1; p.bias2 = d.b2;
  p.raw_enc = d.stream_enc_out ? 1u : 0u;
  p.enc_slope = 0.01f; p.enc_inv_slope = 100.0f;
  // c = F32, a = b = TF32 (2), K-major, N = 64, M = 128
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(RC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // Halo boxes (the tile's rows read once, taps as row-shifted views) while four of them fit; six aligned boxes per tile
  // beyond that.  VFX_PAIR3_HALO_MAX overrides the dilation limit (0 = never).
  static const int halo_max = getenv("VFX_PAIR3_HALO_MAX") ? atoi(getenv("VFX_PAIR3_HALO_MAX")) : 27;
  p.halo = (d.dilation <= halo_max && d.dilation <= 64) ? 1u : 0u;      // a TMA box is at most 256 rows
  p.halo_rows = 128u + 2u * (uint32_t)d.dilation;
  p.slot_bytes = p.halo ? (p.halo_rows * 128u + 1023u) / 1024u * 1024u : 128u * 128u;
  // Role split: 4 + 4 epilogue warps.  Measured alternatives (B200, B = 32, d = 3 / 243): 8 warps for epilogue 1: 1.54 / 2.15 ms,
  // 8 warps for epilogue 2 (32 KB of staging, one ring slot less): 1.60 / 2.21 ms, against 1.55 / 1.91 ms -- the kernel is
  // paced by shared-memory bandwidth (N = 64 MMAs fetch 6 KB of operands per 32 cycles), not by a single role.
  constexpr uint32_t e1w = 4u, e2w = 4u;
  const uint32_t fixed = 2u * R_WBYTES + 2u * R_HPANEL + e2w * 4096u + 2u * RC * 4u + 512u /*barriers*/;
  const uint32_t budget = 227u * 1024u;
  uint32_t slots = (budget - fixed) / p.slot_bytes;
  if (slots > (uint32_t)R_MAX_SLOTS) slots = R_MAX_SLOTS;
  if (slots < (p.halo ? 3u : 4u)) {
    if (!p.halo) return VFX_ERR_UNSUPPORTED;
    p.halo = 0; p.slot_bytes = 128u * 128u;               // box too tall for the ring: aligned boxes
    slots = (budget - fixed) / p.slot_bytes;
    if (slots > (uint32_t)R_MAX_SLOTS) slots = R_MAX_SLOTS;
  }
  p.slots = slots;
  const size_t smem_bytes = (size_t)fixed + (size_t)slots * p.slot_bytes;

  CUtensorMap tmA, tmW1, tmW2, tmO, tmO30;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  auto enc4 = [&](CUtensorMap* tm, const void* base, cuuint32_t box_rows) -> CUresult {
    cuuint64_t dims[4] = {(cuuint64_t)RC, (cuuint64_t)d.L, 1, (cuuint64_t)d.B};
    cuuint64_t strides[3] = {(cuuint64_t)RC * 4, (cuuint64_t)d.L * RC * 4, (cuuint64_t)d.L * RC * 4};
    cuuint32_t box[4] = {32, box_rows, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  auto encw = [&](CUtensorMap* tm, const void* base) -> CUresult {
    cuuint64_t dims[2] = {(cuuint64_t)RC, (cuuint64_t)3 * RC};
    cuuint64_t strides[1] = {(cuuint64_t)RC * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)RC};
    cuuint32_t es[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = enc4(&tmA, d.a, p.halo ? p.halo_rows : 128u);
  if (r == CUDA_SUCCESS) r = encw(&tmW1, d.w1);
  if (r == CUDA_SUCCESS) r = encw(&tmW2, d.w2);
  if (r == CUDA_SUCCESS) r = enc4(&tmO, d.x_out, 32);
  if (r == CUDA_SUCCESS) r = enc4(&tmO30, d.x_out, 30);
  if (r != CUDA_SUCCESS) { set_error("resstack_pair3: cuTensorMapEncodeTiled failed with %d", (int)r); return VFX_ERR_CUDA; }

  int dev = 0, num_sms = 0;
  VFX_CUDA_CHECK(cudaGetDevice(&dev));
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev]) num_sms = sms_of[dev];
  else {
    VFX_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    VFX_CUDA_CHECK(cudaFuncSetAttribute(resstack_pair3_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev < 64) sms_of[dev] = num_sms;
  }
  const int grid = (int)(p.total_tiles < (uint32_t)num_sms ? p.total_tiles : (uint32_t)num_sms);
  p.d_it = grid % p.n_t; p.d_b = grid / p.n_t;
  resstack_pair3_kernel<4, 4><<<grid, (int)(6u + e1w + e2w) * 32, smem_bytes, st>>>(tmA, tmW1, tmW2, tmO, tmO30, p);
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
