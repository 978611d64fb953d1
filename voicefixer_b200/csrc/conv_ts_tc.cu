// Dilated k = 3 convolution of width 128 in tf32 with the WEIGHTS RESIDENT IN TENSOR MEMORY (sm_100a, tcgen05 kind::tf32,
// A operand from TMEM): the convolutions of the C = 128 ResStack (voicefixer/vocoder/model/modules.py:550-576, 592-595) in the
// tf32 mode.
//
// 3 x [128][128] tf32 weights are 192 KB: conv_gemm_tc.cu cannot keep them in shared memory and re-streams them from the L2
// for every 128-row tile (256 KB of TMA traffic per tile, three quarters of it weights) -- 1.3-1.5 ms per launch where the
// HBM traffic alone would take 0.75 ms.  Here the GEMM is transposed,
//
//     D^T[out channel][position] = sum_tap  W_tap[out channel][in channel] . X[position + (tap-1) d][in channel]^T ,
//
// so that the weights are the M = 128 side: they are written ONCE per CTA into 384 TMEM columns (lane = out channel,
// column = tap * 128 + in channel) and every MMA takes its A operand from there (tcgen05.mma ... [d_tmem], [a_tmem], b_desc);
// the activations are the N side (B operand: K-major SWIZZLE_128B rows = positions, exactly the boxes the other kernels load)
// and the only thing that streams.  The remaining 128 TMEM columns hold two 64-position accumulators.
//
//   * Tile = 64 positions of one item (one accumulator); N = 64, K = 8 per MMA, 3 taps x 16 K steps = 48 MMAs per tile.
//   * Operand ring: one slot = one 128-byte K chunk (32 channels).  d <= 27: a halo box of 64 + 2d rows per chunk, taps are
//     row-shifted descriptor views; else three aligned 64-row boxes per chunk.
//   * Epilogue (8 warps: a pair per TMEM lane quarter = 32 out channels, one 32-position half of every tile each; thread = one
//     channel): the accumulator arrives transposed (a thread holds 32 positions of its channel), so results go through a
//     [32 positions][32 channels] staging block written one 4-byte column at a time (conflict-free: a warp writes 32
//     consecutive floats) and leave by TMA; the residual comes in the same way, one tile ahead (two blocks per warp).  All
//     residual loads precede the stores (they alias), bias is one scalar per thread, the output form is a template parameter.
//   MEASURED (B200, B = 32 x 147 882 positions, tools/bench_conv.py --only 128 --prec tf32 --kind pairenc): conv1 (halo boxes)
//   0.93 ms against 1.26 ms for conv_gemm_tc.cu, conv2 with the encoded residual stream 1.15 ms against 1.43 ms = 6.3 TB/s of
//   its 12 bytes per element (the HBM limit).  First version (4 epilogue warps, residual loads and result stores interleaved
//   element by element -- they alias, so every load waited for the previous store): 1.57 / 2.95 ms.
//   * Outputs: exactly one of out_raw (plain or encoded stream, vfx_conv_desc.raw_enc) and out_act (tf32 operand,
//     round-to-nearest); residual optional (plain or encoded).  Anything else stays with conv_gemm_tc.cu.
#include <stdlib.h>
#include <string.h>
#include "vfx_common.cuh"
#include "tc_ptx.cuh"

namespace vfx {

namespace {

constexpr int TC_ = 128;                        // channels (in = out)
constexpr int T_TILE = 64;                      // positions per tile
constexpr int T_THREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int T_MAX_SLOTS = 24;
constexpr int T_WCOLS = 3 * TC_;                // TMEM columns of the weights
constexpr int T_EPI_WARP_BYTES = 2 * 4096;      // per epilogue warp: two [32][32] fp32 blocks

struct TsParams {
  int B, L, d, n_t;
  uint32_t total_tiles;
  int d_b, d_it;
  uint32_t halo, halo_rows, slot_bytes, slots;
  const float* w;                 // [3][128][128] tf32-rounded fp32
  const float* bias;
  float act_param, enc_slope, enc_inv_slope;
  uint32_t idesc;
};

struct TsIter {
  int b, it;
  __device__ __forceinline__ void init(const TsParams& p, uint32_t tile) {
    b = (int)(tile / (uint32_t)p.n_t); it = (int)(tile % (uint32_t)p.n_t);
  }
  __device__ __forceinline__ void next(const TsParams& p) {
    it += p.d_it; const int c = it >= p.n_t; it -= c ? p.n_t : 0;
    b += p.d_b + c;
  }
  __device__ __forceinline__ int p0() const { return it * T_TILE; }
};

__device__ __forceinline__ void ts_st32(uint32_t taddr, const uint32_t* v) {
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
// D[tmem] (+)= A[tmem] * B[smem descriptor]
template <bool ACCUM>
__device__ __forceinline__ void ts_mma(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (ACCUM)
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
}

// ACT: activation of an operand output; RES: 0 none, 1 plain fp32 residual, 2 encoded stream; ENC_OUT: the raw output is the
// encoded stream; IS_ACT: the output is the tf32 operand act(result) instead of the raw result
template <int ACT, int RES, bool ENC_OUT, bool IS_ACT>
__global__ void __launch_bounds__(T_THREADS, 1)
conv_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ TsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // [operand ring | epilogue staging (8 warps x 2 blocks of 4 KB) | barriers]
  uint8_t* const ring = smem;
  uint8_t* const staging = ring + (size_t)p.slots * p.slot_bytes;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(staging + 8 * T_EPI_WARP_BYTES);
  uint64_t* const a_full = bars;
  uint64_t* const a_empty = a_full + T_MAX_SLOTS;
  uint64_t* const acc_full = a_empty + T_MAX_SLOTS;      // [2]
  uint64_t* const acc_empty = acc_full + 2;              // [2] 256 epilogue threads
  uint64_t* const w_ready = acc_empty + 2;               // 256 epilogue threads: the weights are in TMEM
  uint64_t* const res_full = w_ready + 1;                // [8 warps][2 blocks]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(res_full + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    for (uint32_t s = 0; s < p.slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 256); }
    mbar_init(w_ready, 256);
    for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_my = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nkc = TC_ / 32;                               // 4 K chunks of 32 channels

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t s = 0, ph = 0;
    TsIter it; it.init(p, blockIdx.x);
    const int nbox = p.halo ? nkc : 3 * nkc;
    const uint32_t box_bytes = p.halo ? p.halo_rows * 128u : (uint32_t)T_TILE * 128u;
    for (uint32_t i = 0; i < n_my; ++i) {
      const int p0 = it.p0(), b = it.b;
      it.next(p);
#pragma unroll 1
      for (int q = 0; q < nbox; ++q) {                    // aligned boxes: q = kc * 3 + tap
        const int kc = p.halo ? q : q / 3, tap = p.halo ? 0 : q % 3;
        const int row = p.halo ? p0 - p.d : p0 + (tap - 1) * p.d;
        mbar_wait(&a_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], box_bytes);
          tma_load_4d(&tmA, &a_full[s], ring + (size_t)s * p.slot_bytes, kc * 32, row, 0, b);
        }
        __syncwarp();
        if (++s == p.slots) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(w_ready, 0);
    tc_fence_after();
    uint32_t s = 0, ph = 0;
    const uint32_t dhi = desc_hi(64u /* 8 rows x 128 B >> 4 */, 2u /* SWIZZLE_128B */);
    for (uint32_t i = 0; i < n_my; ++i) {
      mbar_wait(&acc_empty[i & 1], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + T_WCOLS + (i & 1) * T_TILE;
#pragma unroll 1
      for (int kc = 0; kc < nkc; ++kc) {
        if (p.halo) {
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)s * p.slot_bytes);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
              const uint32_t b_lo = desc_lo(sa + (uint32_t)(tap * p.d) * 128u);
              const uint32_t a_t = tmem_base + (uint32_t)(tap * TC_ + kc * 32);
              if (tap == 0 && kc == 0) ts_mma<false>(d_tmem, a_t, b_lo, dhi, p.idesc);
              else ts_mma<true>(d_tmem, a_t, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) ts_mma<true>(d_tmem, a_t + 8 * k, b_lo + 2 * k, dhi, p.idesc);
            }
            tc_commit(&a_empty[s]);
            if (kc == nkc - 1) tc_commit(&acc_full[i & 1]);
          }
          __syncwarp();
          if (++s == p.slots) { s = 0; ph ^= 1; }
        } else {
#pragma unroll 1
          for (int tap = 0; tap < 3; ++tap) {
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + (size_t)s * p.slot_bytes);
            if (elect_one()) {
              const uint32_t b_lo = desc_lo(sa);
              const uint32_t a_t = tmem_base + (uint32_t)(tap * TC_ + kc * 32);
              if (tap == 0 && kc == 0) ts_mma<false>(d_tmem, a_t, b_lo, dhi, p.idesc);
              else ts_mma<true>(d_tmem, a_t, b_lo, dhi, p.idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k) ts_mma<true>(d_tmem, a_t + 8 * k, b_lo + 2 * k, dhi, p.idesc);
              tc_commit(&a_empty[s]);
              if (kc == nkc - 1 && tap == 2) tc_commit(&acc_full[i & 1]);
            }
            __syncwarp();
            if (++s == p.slots) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2-9): thread = out channel sub * 32 + lane, warp pair = the tile's two chunks ====
    const int ew = warp - 2, sub = warp & 3, c = ew >> 2;  // c: this warp's 32-position half of every tile
    const int ch = sub * 32 + lane;
    // ---- once: this thread's weight rows W[tap][ch][0..127] -> TMEM lane ch, columns tap * 128 + k (six blocks per warp)
    {
      const uint32_t t_w = tmem_base + ((uint32_t)(sub * 32) << 16);
#pragma unroll 1
      for (int blk = 6 * c; blk < 6 * c + 6; ++blk) {     // (tap, 32-channel chunk)
        const int tap = blk >> 2, kq = blk & 3;
        const uint4* src = reinterpret_cast<const uint4*>(p.w + ((size_t)tap * TC_ + ch) * TC_ + kq * 32);
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const uint4 q4 = src[j]; v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w; }
        ts_st32(t_w + (uint32_t)(tap * TC_ + kq * 32), v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(w_ready);
    }
    const float bias = p.bias ? p.bias[ch] : 0.f;
    uint8_t* const stg = staging + ew * T_EPI_WARP_BYTES;      // two [32 positions][32 channels] fp32 blocks
    uint64_t* const rfull = res_full + ew * 2;
    const int c0 = sub * 32;
    TsIter it; it.init(p, blockIdx.x);
    if (RES && n_my > 0 && lane == 0) {                         // residual block of the first tile
      mbar_expect_tx(&rfull[0], 4096);
      tma_load_4d(&tmR, &rfull[0], stg, c0, it.p0() + c * 32, 0, it.b);
    }
    for (uint32_t i = 0; i < n_my; ++i) {
      const int p0 = it.p0(), b = it.b;
      it.next(p);
      const uint32_t slot = i & 1;
      float* const blk = reinterpret_cast<float*>(stg + slot * 4096) + lane;       // blk[pos * 32] = (pos, this channel)
      mbar_wait(&acc_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tc_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + T_WCOLS + (i & 1) * T_TILE + c * 32, v);
      tc_fence_before();
      mbar_arrive(&acc_empty[i & 1]);
      float r[32];
      if (RES) {
        mbar_wait(&rfull[slot], (i >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = blk[j * 32];       // all loads first: the stores below may alias them
      } else {
        // this block was last stored from two tiles ago (one bulk group per tile)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float f = __uint_as_float(v[j]) + bias;
        if (RES == 2) f += stream_dec(r[j], p.enc_inv_slope);
        else if (RES == 1) f += r[j];
        if (IS_ACT) f = round_tf32(act_fast<ACT>(f, p.act_param));
        else if (ENC_OUT) f = stream_enc(f, p.enc_slope);
        v[j] = __float_as_uint(f);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) blk[j * 32] = __uint_as_float(v[j]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmO, stg + slot * 4096, c0, p0 + c * 32, 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (RES && i + 1 < n_my) {                                              // next tile's residual -> the other block,
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");       // whose store (tile i-1) has left it
          mbar_expect_tx(&rfull[slot ^ 1], 4096);
          tma_load_4d(&tmR, &rfull[slot ^ 1], stg + (slot ^ 1) * 4096, c0, it.p0() + c * 32, 0, it.b);
        }
      }
      __syncwarp();
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

// Returns VFX_ERR_UNSUPPORTED for every shape that is not "k = 3 dilated convolution, 128 -> 128 channels, contiguous
// [B][L][128] fp32 tensors, one output" (conv_gemm_tc.cu then handles it).
int conv_ts_tc(const vfx_conv_desc& d, cudaStream_t st) {
  static const bool off = getenv("VFX_NO_TS") != nullptr;
  if (off) return VFX_ERR_UNSUPPORTED;
  if (d.Cin != TC_ || d.N != TC_ || d.ntaps != 3 || d.H != 1 || d.Hq != 1 || d.Wq != d.W || d.OH != 1 || d.OW != d.W)
    return VFX_ERR_UNSUPPORTED;
  if (d.sh != 1 || d.sw != 1 || d.rh != 0 || d.rw != 0) return VFX_ERR_UNSUPPORTED;
  if (d.dh[0] || d.dh[1] || d.dh[2] || d.dw[1] != 0 || d.dw[2] <= 0 || d.dw[0] != -d.dw[2]) return VFX_ERR_UNSUPPORTED;
  for (int t = 0; t < 3; ++t) if (d.w_off[t] != (long long)t * TC_ * TC_) return VFX_ERR_UNSUPPORTED;
  const long long L = d.W;
  if (d.a_sW != TC_ || d.a_sB != L * TC_) return VFX_ERR_UNSUPPORTED;
  if ((d.out_raw != nullptr) == (d.out_act != nullptr)) return VFX_ERR_UNSUPPORTED;      // exactly one output
  if (d.act_scale || d.act_shift) return VFX_ERR_UNSUPPORTED;
  if (d.bias && d.bias_mod != TC_ && d.bias_mod != 0) return VFX_ERR_UNSUPPORTED;
  if (d.out_raw && (d.o_sW != TC_ || d.o_sB != L * TC_ || d.o_col != 0)) return VFX_ERR_UNSUPPORTED;
  if (d.out_act && (d.oa_sW != TC_ || d.oa_sB != L * TC_ || d.oa_col != 0)) return VFX_ERR_UNSUPPORTED;
  if (d.residual && (d.r_sW != TC_ || d.r_sB != L * TC_ || d.r_col != 0)) return VFX_ERR_UNSUPPORTED;
  const int act = d.out_act ? d.act : VFX_ACT_NONE;
  void* const out = d.out_raw ? (void*)d.out_raw : d.out_act;
  if (((uintptr_t)d.a & 15) || ((uintptr_t)d.w & 15) || ((uintptr_t)out & 15) || ((uintptr_t)d.residual & 15) ||
      ((uintptr_t)d.bias & 3))
    return VFX_ERR_UNSUPPORTED;
  if (out == d.a) return VFX_ERR_UNSUPPORTED;             // taps read neighbouring tiles' rows
  const bool res_enc = d.residual && d.res_enc, raw_enc = d.out_raw && d.raw_enc;
  if ((res_enc || raw_enc) && !(d.enc_slope > 0.f)) return VFX_ERR_UNSUPPORTED;
  // the output forms the tf32 vocoder uses (compile-time epilogues); everything else stays with the generic kernel
  const int res = !d.residual ? 0 : res_enc ? 2 : 1;
  int form = -1;
  if (d.out_act && act == VFX_ACT_LRELU && res == 0) form = 0;                  // conv1: lrelu operand out
  else if (d.out_raw && res == 2 && raw_enc) form = 1;                            // conv2: stream in, stream out
  else if (d.out_act && act == VFX_ACT_LRELU_XSINX && res == 2) form = 2;        // last conv2 of a stack: next up-sampler's operand
  else if (d.out_raw && res == 1 && !raw_enc) form = 3;                           // plain residual, plain result
  else if (d.out_raw && res == 2 && !raw_enc) form = 4;
  else if (d.out_raw && res == 0 && !raw_enc) form = 5;
  if (form < 0) return VFX_ERR_UNSUPPORTED;
  // conv1 with a dilation too large for a halo box would take three aligned boxes per K chunk: 1.39 ms against the generic
  // kernel's 1.27 ms (both L2 -> SM bound there) -- left to the generic kernel.  (conv2 always has dilation 1.)
  if (form == 0 && d.dw[2] > 27) return VFX_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return VFX_ERR_UNSUPPORTED;

  TsParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.L = (int)L; p.d = d.dw[2];
  p.n_t = ceil_div((int)L, T_TILE);
  const long long total = (long long)d.B * p.n_t;
  if (total >= (1LL << 31)) return VFX_ERR_UNSUPPORTED;
  p.total_tiles = (uint32_t)total;
  p.w = reinterpret_cast<const float*>(d.w); p.bias = d.bias;
  p.act_param = d.act_param; p.enc_slope = d.enc_slope; p.enc_inv_slope = d.enc_slope > 0.f ? 1.0f / d.enc_slope : 0.f;
  // c = F32, a = b = TF32, K-major, N = 64 positions, M = 128 out channels
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.halo = p.d <= 27 ? 1u : 0u;
  p.halo_rows = (uint32_t)T_TILE + 2u * (uint32_t)p.d;
  p.slot_bytes = p.halo ? (p.halo_rows * 128u + 1023u) / 1024u * 1024u : (uint32_t)T_TILE * 128u;
  const uint32_t fixed = 8u * T_EPI_WARP_BYTES + 1024u /*align*/ + 1024u /*barriers*/;
  uint32_t slots = (227u * 1024u - fixed) / p.slot_bytes;
  if (slots > (uint32_t)T_MAX_SLOTS) slots = T_MAX_SLOTS;
  p.slots = slots;
  const size_t smem_bytes = (size_t)fixed + (size_t)slots * p.slot_bytes;

  CUtensorMap tmA, tmR, tmO;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  auto enc4 = [&](CUtensorMap* tm, const void* base, cuuint32_t box_rows, CUtensorMapSwizzle swz) -> CUresult {
    cuuint64_t dims[4] = {(cuuint64_t)TC_, (cuuint64_t)L, 1, (cuuint64_t)d.B};
    cuuint64_t strides[3] = {(cuuint64_t)TC_ * 4, (cuuint64_t)L * TC_ * 4, (cuuint64_t)L * TC_ * 4};
    cuuint32_t box[4] = {32, box_rows, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = enc4(&tmA, d.a, p.halo ? p.halo_rows : (cuuint32_t)T_TILE, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r == CUDA_SUCCESS) r = enc4(&tmO, out, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
  tmR = tmO;
  if (r == CUDA_SUCCESS && d.residual) r = enc4(&tmR, d.residual, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_ts: cuTensorMapEncodeTiled failed with %d", (int)r); return VFX_ERR_CUDA; }

  int dev = 0, num_sms = 0;
  VFX_CUDA_CHECK(cudaGetDevice(&dev));
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev]) num_sms = sms_of[dev];
  else {
    VFX_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
#define VFX_TS_ATTR(...) VFX_CUDA_CHECK(cudaFuncSetAttribute(conv_ts_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    VFX_TS_ATTR(VFX_ACT_LRELU, 0, false, true); VFX_TS_ATTR(VFX_ACT_NONE, 2, true, false); VFX_TS_ATTR(VFX_ACT_LRELU_XSINX, 2, false, true);
    VFX_TS_ATTR(VFX_ACT_NONE, 1, false, false); VFX_TS_ATTR(VFX_ACT_NONE, 2, false, false); VFX_TS_ATTR(VFX_ACT_NONE, 0, false, false);
#undef VFX_TS_ATTR
    if (dev < 64) sms_of[dev] = num_sms;
  }
  const int grid = (int)(p.total_tiles < (uint32_t)num_sms ? p.total_tiles : (uint32_t)num_sms);
  p.d_it = grid % p.n_t; p.d_b = grid / p.n_t;
#define VFX_TS_LAUNCH(...) conv_ts_kernel<__VA_ARGS__><<<grid, T_THREADS, smem_bytes, st>>>(tmA, tmR, tmO, p)
  if (form == 0) VFX_TS_LAUNCH(VFX_ACT_LRELU, 0, false, true);
  else if (form == 1) VFX_TS_LAUNCH(VFX_ACT_NONE, 2, true, false);
  else if (form == 2) VFX_TS_LAUNCH(VFX_ACT_LRELU_XSINX, 2, false, true);
  else if (form == 3) VFX_TS_LAUNCH(VFX_ACT_NONE, 1, false, false);
  else if (form == 4) VFX_TS_LAUNCH(VFX_ACT_NONE, 2, false, false);
  else VFX_TS_LAUNCH(VFX_ACT_NONE, 0, false, false);
#undef VFX_TS_LAUNCH
  VFX_LAUNCH_CHECK();
  return VFX_OK;
}

}  // namespace vfx
