// vfx_b200 engine: weight registry, workspace planning and the launch sequence of the restore()
// hot path behind the C ABI of include/vfx_b200.h.  One engine per process/GPU; all device memory
// (weights, workspace, inputs, outputs) is owned by the caller; only two small constant tables
// (Hann window, FFT twiddles) are allocated here at create time.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "vfx_common.cuh"

namespace vfx {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

unsigned long long g_launches = 0;

struct Tensor { const void* p; size_t bytes; };

// per-tag CUDA-event timing of the launch sequence (option "profile")
struct ProfRec { std::string tag; cudaEvent_t a, b; double flops, bytes; };

}  // namespace vfx

enum { VFX_PART_ANALYSIS = 1, VFX_PART_VOCODER = 2 };   // analysis = front end + denoiser + UNet ("fe.", "dn.", "unet.")

struct vfx_engine {
  int device = 0;   // -1: planning-only engine (no CUDA calls; finalize and workspace queries only)
  int precision = VFX_PREC_FP32;
  bool finalized = false;
  int use_tc = 1;   // BF16 / TF32: tcgen05 kernel where the shape allows (0 = SIMT cross-check on the same operands)
  std::unordered_map<std::string, vfx::Tensor> tensors;
  float* d_window = nullptr;    // periodic Hann, 2048
  float2* d_tw = nullptr;       // exp(-2 pi i k / 2048), k < 1024
  std::vector<std::string> missing;
  int parts = 0;    // VFX_PART_* bits of the weight sets found complete by vfx_engine_finalize
  int tf32_stream = 1;  // TF32: vocoder residual streams are kept as ONE encoded fp32 tensor (operand + residual carrier)
  int fuse_pair = 1;  // BF16: ResStack pairs of width 64 run as ONE fused kernel (resstack_pair_tc.cu)
  int fuse_pair3 = 1; // TF32 width 64: one-SM fused pair with the residual stashed in tensor memory (resstack_pair3_tc.cu)
  int fuse_pair2 = 1; // two-CTA cluster pipeline (resstack_pair2_tc.cu): 1 = BF16 / FP16 width 128; 2 = also TF32 width 64 (no gain)
  int profile = 0;
  std::vector<vfx::ProfRec> prof;
  std::string prof_report;
  size_t esz() const { return vfx::prec_esz(precision); }
  bool tensor_core() const { return precision != VFX_PREC_FP32; }   // operands are a tensor-core format (bf16 / tf32)
};

namespace vfx {

namespace {

constexpr int UNET_C[7] = {2, 32, 64, 128, 256, 384, 384};       // channels per level (0 = input)
constexpr int VOC_CIN[4] = {1024, 512, 256, 128};
constexpr int VOC_COUT[4] = {512, 256, 128, 64};
constexpr int VOC_U[4] = {7, 7, 3, 3};

// ------------------------------------------------------------------ workspace bump allocator
struct Bump {
  char* base; size_t cap; size_t off = 0; size_t peak = 0; bool overflow = false;
  Bump(void* b, size_t c) : base((char*)b), cap(c) {}
  void* raw(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    size_t o = off;
    off += bytes;
    if (off > peak) peak = off;
    if (base == nullptr) return (void*)(uintptr_t)(256 + o);   // dry run: fake non-null pointer
    if (off > cap) { overflow = true; return base; }
    return base + o;
  }
  template <typename T> T* alloc(size_t n) { return (T*)raw(n * sizeof(T)); }
  size_t mark() const { return off; }
  void reset(size_t m) { off = m; }
};

struct Ctx {
  vfx_engine* e; cudaStream_t st; Bump* ws; bool dry; int mode; int rc = VFX_OK;
  bool train() const { return mode == VFX_MODE_TRAIN_BN; }
  int prec() const { return e->precision; }
};

struct ProfScope {
  Ctx& c; bool on;
  ProfScope(Ctx& c_, const char* tag, double flops, double bytes) : c(c_), on(false) {
    if (c.dry || !c.e->profile) return;
    ProfRec r; r.tag = tag; r.flops = flops; r.bytes = bytes;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, c.st);
    c.e->prof.push_back(r);
    on = true;
  }
  ~ProfScope() { if (on) cudaEventRecord(c.e->prof.back().b, c.st); }
};

double conv_flops(const vfx_conv_desc& d) {
  return 2.0 * d.B * d.Hq * d.Wq * (double)d.N * d.Cin * d.ntaps;
}
// algorithmic HBM bytes of one conv launch: the input tensor once (taps re-read from L2), fp32 residual in,
// fp32 raw out, operand out (weights are negligible and stay in L2)
double conv_bytes(const vfx_conv_desc& d, int prec) {
  const double esz = (double)prec_esz(prec);
  const double M = (double)d.B * d.Hq * d.Wq;
  return (double)d.B * d.H * d.W * d.Cin * esz + M * d.N * ((d.residual ? 4.0 : 0.0) + (d.out_raw ? 4.0 : 0.0) + (d.out_act ? esz : 0.0));
}

#define VFX_TRY(expr) do { int _r = (expr); if (_r != VFX_OK) return _r; } while (0)

const void* get(Ctx& c, const std::string& name, size_t bytes) {
  auto it = c.e->tensors.find(name);
  if (it == c.e->tensors.end()) {
    c.e->missing.push_back(name);
    set_error("weight tensor '%s' not registered", name.c_str());
    c.rc = VFX_ERR_MISSING_WEIGHT;
    return nullptr;
  }
  if (it->second.bytes != bytes) {
    set_error("weight tensor '%s' has %zu bytes, engine expects %zu", name.c_str(), it->second.bytes, bytes);
    c.rc = VFX_ERR_INVALID;
    return nullptr;
  }
  return it->second.p;
}
const float* getf(Ctx& c, const std::string& name, size_t n) { return (const float*)get(c, name, n * 4); }
const void* getw(Ctx& c, const std::string& name, size_t n, int prec) {
  return get(c, name, n * prec_esz(prec));
}

int run_conv(Ctx& c, int prec, const vfx_conv_desc& d, const char* tag = "conv") {
  if (c.dry) return VFX_OK;
  if (c.rc != VFX_OK) return c.rc;
  ProfScope ps(c, tag, conv_flops(d), conv_bytes(d, prec));
  if (prec != VFX_PREC_FP32 && c.e->use_tc) {
    int r = conv_gemm_tc(prec, d, c.st);
    if (r != VFX_ERR_UNSUPPORTED) return r;
  }
  return conv_gemm_simt(prec, d, c.st);
}

vfx_conv_desc conv_base(const void* a, int B, int H, int W, int Cin, const void* w, int N) {
  vfx_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.a = a; d.B = B; d.H = H; d.W = W; d.Cin = Cin;
  d.a_sW = Cin; d.a_sH = (long long)W * Cin; d.a_sB = (long long)H * W * Cin;
  d.w = w; d.N = N; d.Hq = H; d.Wq = W; d.sh = 1; d.sw = 1; d.OH = H; d.OW = W;
  d.bias_mod = N;
  return d;
}
void set_raw(vfx_conv_desc& d, float* p, long long ld, int col) {
  d.out_raw = p; d.o_sW = ld; d.o_sH = (long long)d.OW * ld; d.o_sB = (long long)d.OH * d.OW * ld; d.o_col = col;
}
void set_act(vfx_conv_desc& d, void* p, long long ld, int col, int act, float param) {
  d.out_act = p; d.oa_sW = ld; d.oa_sH = (long long)d.OW * ld; d.oa_sB = (long long)d.OH * d.OW * ld;
  d.oa_col = col; d.act = act; d.act_param = param;
}
void set_res(vfx_conv_desc& d, const float* p, long long ld, int col) {
  d.residual = p; d.r_sW = ld; d.r_sH = (long long)d.OW * ld; d.r_sB = (long long)d.OH * d.OW * ld; d.r_col = col;
}
void taps3x3(vfx_conv_desc& d, long long mat) {
  d.ntaps = 9;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      int t = kh * 3 + kw;
      d.dh[t] = kh - 1; d.dw[t] = kw - 1; d.w_off[t] = (long long)t * mat;
    }
}

// ------------------------------------------------------------------ BN handling
struct BnRef { const float* scale; const float* shift; int stat_sB; };

// Eval: the folded running-stat affine registered by the host ("<name>.scale/.shift").
// Train (mode 2): per-item biased batch statistics of `x`.
int bn_resolve(Ctx& c, const std::string& name, int bn_C, const float* x, long long x_sB, long long ldx,
               int B, long long P, int C, BnRef* out) {
  if (!c.train()) {
    out->scale = getf(c, name + ".scale", bn_C);
    out->shift = getf(c, name + ".shift", bn_C);
    out->stat_sB = 0;
    return c.rc;
  }
  const float* gamma = getf(c, name + ".gamma", bn_C);
  const float* beta = getf(c, name + ".beta", bn_C);
  float* sc = c.ws->alloc<float>((size_t)B * bn_C);
  float* sh = c.ws->alloc<float>((size_t)B * bn_C);
  double* acc = c.ws->alloc<double>((size_t)2 * B * bn_C);
  out->scale = sc; out->shift = sh; out->stat_sB = bn_C;
  if (c.dry || c.rc != VFX_OK) return c.rc;
  return bn_stats(x, x_sB, ldx, B, P, C, bn_C, gamma, beta, sc, sh, acc, c.st);
}

// operand = act(bn(x)); x raw fp32 [B][P][C] (pitch ldx) -> dense operand [B][P][C]
int bn_act_op(Ctx& c, int prec, const std::string& bn, int bn_C, const float* x, long long x_sB,
              long long ldx, int B, long long P, int C, int act, float slope, void* y) {
  BnRef r;
  VFX_TRY(bn_resolve(c, bn, bn_C, x, x_sB, ldx, B, P, C, &r));
  if (c.dry) return VFX_OK;
  ProfScope ps(c, "bn_act", 0.0, (double)B * P * C * (4 + prec_esz(prec)));
  return bn_act(prec, x, x_sB, ldx, B, P, C, r.scale, r.shift, bn_C, r.stat_sB, act, slope, y,
                (long long)P * C, C, c.st);
}

// ------------------------------------------------------------------ UNet
// ConvBlockRes.forward voicefixer/restorer/modules.py:68-76 on a channels-last raw tensor.
// in: raw fp32 [B][H][W][Cin] with pitch ld_in; out: raw fp32 with pitch ld_out (+col), may alias in.
struct BlockScratch { void* opA; float* rawH; void* opH; float* rawS; void* opX; };

// Eval-mode fusion hooks of a block (ignored in mode 2, which needs batch statistics between the ops):
//   opA_ready : s.opA already holds lrelu(bn1(in)) -- written by the previous conv's epilogue
//   next_bn   : BatchNorm whose eval affine + activation conv2's epilogue applies to produce the NEXT
//               consumer's operand in s.opA (next_slope: 0.01 LeakyReLU for ConvBlockRes.bn1, 0 = ReLU for
//               DecoderBlockRes.bn1)
struct BlockFuse { bool opA_ready = false; const char* next_bn = nullptr; float next_slope = 0.01f; };

int conv_block(Ctx& c, const std::string& p, const float* in, long long ld_in, int B, int H, int W,
               int Cin, int Cout, float* out, long long ld_out, const BlockScratch& s, const BlockFuse& fz = BlockFuse()) {
  const int prec = c.prec();
  const long long P = (long long)H * W;
  const bool fuse = !c.train();
  // bf16 / tf32 modes: a 2-channel input (first encoder block) is zero-padded to 32 operand channels so the block
  // runs on the tensor-core kernel; the host packs conv1 / shortcut weights as [..][Cout][32] accordingly.
  const int Cop = (prec != VFX_PREC_FP32 && Cin < 32) ? 32 : Cin;
  if (Cop != Cin && !c.dry) {
    VFX_CUDA_CHECK(cudaMemsetAsync(s.opA, 0, (size_t)B * P * Cop * prec_esz(prec), c.st));
    VFX_CUDA_CHECK(cudaMemsetAsync(s.opX, 0, (size_t)B * P * Cop * prec_esz(prec), c.st));
  }
  if (!(fuse && fz.opA_ready)) {
    BnRef r;
    VFX_TRY(bn_resolve(c, p + ".bn1", Cin, in, P * ld_in, ld_in, B, P, Cin, &r));
    if (!c.dry) {
      ProfScope ps(c, "bn_act", 0.0, (double)B * P * Cin * (4 + prec_esz(prec)));
      VFX_TRY(bn_act(prec, in, P * ld_in, ld_in, B, P, Cin, r.scale, r.shift, Cin, r.stat_sB, VFX_ACT_LRELU, 0.01f, s.opA,
                     P * Cop, Cop, c.st));
    }
  }
  char tg[64];
  if (fuse) {     // conv1 with bn2 folded in: h = lrelu(conv1f(a) + b) straight into the operand buffer
    vfx_conv_desc d = conv_base(s.opA, B, H, W, Cop, getw(c, p + ".conv1f.w", (size_t)9 * Cout * Cop, prec), Cout);
    taps3x3(d, (long long)Cout * Cop);
    d.bias = getf(c, p + ".conv1f.b", Cout);
    set_act(d, s.opH, Cout, 0, VFX_ACT_LRELU, 0.01f);
    snprintf(tg, sizeof(tg), "unet.conv3x3.W%d.%dto%d", W, Cin, Cout);
    VFX_TRY(run_conv(c, prec, d, c.e->profile > 1 ? tg : "unet.conv3x3"));
  } else {
    vfx_conv_desc d = conv_base(s.opA, B, H, W, Cop, getw(c, p + ".conv1.w", (size_t)9 * Cout * Cop, prec), Cout);
    taps3x3(d, (long long)Cout * Cop);
    set_raw(d, s.rawH, Cout, 0);
    snprintf(tg, sizeof(tg), "unet.conv3x3.W%d.%dto%d", W, Cin, Cout);
    VFX_TRY(run_conv(c, prec, d, c.e->profile > 1 ? tg : "unet.conv3x3"));
    VFX_TRY(bn_act_op(c, prec, p + ".bn2", Cout, s.rawH, P * Cout, Cout, B, P, Cout, VFX_ACT_LRELU, 0.01f, s.opH));
  }
  const float* res = in; long long ld_res = ld_in;
  if (Cin != Cout) {
    const void* xin = in;
    if (prec != VFX_PREC_FP32 || ld_in != Cin) {   // shortcut consumes raw x as a dense operand
      if (!c.dry) VFX_TRY(bn_act(prec, in, P * ld_in, ld_in, B, P, Cin, nullptr, nullptr, Cin, 0, VFX_ACT_NONE, 0.f,
                                 s.opX, P * Cop, Cop, c.st));
      xin = s.opX;
    }
    vfx_conv_desc d = conv_base(xin, B, H, W, Cop, getw(c, p + ".sc.w", (size_t)Cout * Cop, prec), Cout);
    d.ntaps = 1;
    d.bias = getf(c, p + ".sc.b", Cout);
    set_raw(d, s.rawS, Cout, 0);
    VFX_TRY(run_conv(c, prec, d, "unet.shortcut"));
    res = s.rawS; ld_res = Cout;
  }
  {
    vfx_conv_desc d = conv_base(s.opH, B, H, W, Cout, getw(c, p + ".conv2.w", (size_t)9 * Cout * Cout, prec), Cout);
    taps3x3(d, (long long)Cout * Cout);
    set_res(d, res, ld_res, 0);
    set_raw(d, out, ld_out, 0);
    if (fuse && fz.next_bn) {   // the consumer's eval BatchNorm + activation, fused: s.opA = act(bn(out))
      const std::string nb(fz.next_bn);
      d.act_scale = getf(c, nb + ".scale", Cout);
      d.act_shift = getf(c, nb + ".shift", Cout);
      set_act(d, s.opA, Cout, 0, VFX_ACT_LRELU, fz.next_slope);
    }
    snprintf(tg, sizeof(tg), "unet.conv3x3.W%d.%dto%d.res", W, Cout, Cout);
    VFX_TRY(run_conv(c, prec, d, c.e->profile > 1 ? tg : "unet.conv3x3"));
  }
  return c.rc;
}

// UNetResComplex_100Mb.forward voicefixer/restorer/model_kqq_bn.py:130-181.
// unet_in raw [B][Tp][127][2]; returns pointer to the last feature map raw [B][Tp][127][32].
int unet_forward(Ctx& c, const float* unet_in, int B, int Tp, float** feat_out) {
  const int prec = c.prec();
  int Hl[7], Wl[7];
  Hl[0] = Tp; Wl[0] = 127;                       // level l (1..6) works at Hl[l-1] x Wl[l-1]
  for (int l = 1; l < 7; ++l) { Hl[l] = Hl[l - 1] / 2; Wl[l] = Wl[l - 1] / 2; }
  const size_t esz = c.e->esz();
  const size_t P1 = (size_t)B * Tp * 127;
  BlockScratch s;
  s.opA = c.ws->raw(P1 * 64 * esz);
  s.rawH = c.ws->alloc<float>(P1 * 32);
  s.opH = c.ws->raw(P1 * 32 * esz);
  s.rawS = c.ws->alloc<float>(P1 * 32);
  s.opX = c.ws->raw(P1 * 64 * esz);
  float* cat[7]; float* pool[7];
  for (int l = 1; l <= 6; ++l) {
    const size_t P = (size_t)B * Hl[l - 1] * Wl[l - 1];
    cat[l] = c.ws->alloc<float>(P * 2 * UNET_C[l]);
    pool[l] = c.ws->alloc<float>((size_t)B * Hl[l] * Wl[l] * UNET_C[l]);
  }
  char name[96];
  // ---- encoder: level l output lives in the second half of the decoder's concat buffer
  const float* x = unet_in; int Cx = 2;
  for (int l = 1; l <= 6; ++l) {
    const int H = Hl[l - 1], W = Wl[l - 1], C = UNET_C[l];
    float* skip = cat[l] + C;                     // pitch 2C
    char nbn[4][96];
    for (int j = 1; j <= 4; ++j) {
      snprintf(name, sizeof(name), "unet.enc%d.b%d", l, j);
      BlockFuse fz;
      fz.opA_ready = j > 1;                               // produced by block j-1's conv2 epilogue
      if (j < 4) { snprintf(nbn[j], sizeof(nbn[j]), "unet.enc%d.b%d.bn1", l, j + 1); fz.next_bn = nbn[j]; }
      if (j == 1) VFX_TRY(conv_block(c, name, x, Cx, B, H, W, Cx, C, skip, 2 * C, s, fz));
      else VFX_TRY(conv_block(c, name, skip, 2 * C, B, H, W, C, C, skip, 2 * C, s, fz));
    }
    ProfScope ps(c, "unet.pool", 0.0, (double)B * H * W * C * 5.0);
    if (!c.dry) VFX_TRY(avgpool2x2(skip, (long long)H * W * 2 * C, (long long)W * 2 * C, 2 * C, B, H, W, C, pool[l], c.st));
    x = pool[l]; Cx = C;
  }
  // ---- centre block (in place on pool[6]): conv_block7
  {
    BlockFuse fz; fz.next_bn = "unet.dec1.bn1"; fz.next_slope = 0.0f;      // relu(bn1(x)) before the first ConvTranspose
    VFX_TRY(conv_block(c, "unet.center", pool[6], 384, B, Hl[6], Wl[6], 384, 384, pool[6], 384, s, fz));
  }
  float* xd = pool[6];
  int Cd = 384;
  const bool fuse = !c.train();
  // ---- decoder i = 1..6 at level l = 7 - i  (DecoderBlockRes.forward restorer/modules.py:149-157)
  for (int i = 1; i <= 6; ++i) {
    const int l = 7 - i;
    const int H = Hl[l], W = Wl[l];               // input grid
    const int OH = Hl[l - 1], OW = Wl[l - 1];     // output grid (= 2H, 2W+1)
    const int Cout = UNET_C[l];
    snprintf(name, sizeof(name), "unet.dec%d", i);
    const std::string p(name);
    const long long Pin = (long long)H * W;
    if (!fuse)   // eval: relu(bn1(x)) was written into s.opA by the producing conv2's epilogue
      VFX_TRY(bn_act_op(c, prec, p + ".bn1", Cd, xd, Pin * Cd, Cd, B, Pin, Cd, VFX_ACT_LRELU, 0.0f, s.opA));
    const void* wt = getw(c, p + ".up.w", (size_t)9 * Cout * Cd, prec);
    for (int rh = 0; rh < 2; ++rh)
      for (int rw = 0; rw < 2; ++rw) {
        vfx_conv_desc d = conv_base(s.opA, B, H, W, Cd, wt, Cout);
        d.Hq = H; d.Wq = W + 1; d.sh = 2; d.sw = 2; d.rh = rh; d.rw = rw; d.OH = OH; d.OW = OW;
        int n = 0;
        for (int kh = (rh ? 1 : 0); kh < 3; kh += 2)
          for (int kw = (rw ? 1 : 0); kw < 3; kw += 2) {
            d.dh[n] = kh == 2 ? -1 : 0; d.dw[n] = kw == 2 ? -1 : 0;
            d.w_off[n] = (long long)(kh * 3 + kw) * Cout * Cd;
            ++n;
          }
        d.ntaps = n;
        set_raw(d, cat[l], 2 * Cout, 0);
        VFX_TRY(run_conv(c, prec, d, "unet.convT"));
      }
    // conv_block2..5: first consumes the concat tensor (2C -> C, with shortcut)
    float* dl = pool[l];                          // dense [B][OH][OW][Cout]: reuse? sizes differ -> own buffer
    dl = c.ws->alloc<float>((size_t)B * OH * OW * Cout);
    char nbn[96];
    for (int j = 2; j <= 5; ++j) {
      snprintf(name, sizeof(name), "unet.dec%d.b%d", i, j);
      BlockFuse fz;
      fz.opA_ready = j > 2;
      if (j < 5) { snprintf(nbn, sizeof(nbn), "unet.dec%d.b%d.bn1", i, j + 1); fz.next_bn = nbn; }
      else if (i < 6) { snprintf(nbn, sizeof(nbn), "unet.dec%d.bn1", i + 1); fz.next_bn = nbn; fz.next_slope = 0.0f; }
      else { snprintf(nbn, sizeof(nbn), "unet.after.bn1"); fz.next_bn = nbn; }
      if (j == 2) VFX_TRY(conv_block(c, name, cat[l], 2 * Cout, B, OH, OW, 2 * Cout, Cout, dl, Cout, s, fz));
      else VFX_TRY(conv_block(c, name, dl, Cout, B, OH, OW, Cout, Cout, dl, Cout, s, fz));
    }
    xd = dl; Cd = Cout;
  }
  {
    BlockFuse fz; fz.opA_ready = true;                   // dec6.b5 produced lrelu(after.bn1(x))
    VFX_TRY(conv_block(c, "unet.after", xd, 32, B, Tp, 127, 32, 32, xd, 32, s, fz));
  }
  *feat_out = xd;
  return c.rc;
}

// ------------------------------------------------------------------ denoiser
// nn.Sequential voicefixer/restorer/model.py:69-99 (always fp32 SIMT: 1 % of the path's FLOPs,
// recurrent part is precision-sensitive).  mel [B][T][128] -> lin15 raw [B*T][128].
// Denoiser GEMMs: fp32 SIMT in fp32 mode; in bf16 mode the same tcgen05 kernel as the convolutions
// (bf16 operands, fp32 accumulate / bias / outputs).  The GRU recurrence itself is always fp32.
int linear(Ctx& c, const void* a, long long M, int K, const std::string& p, int N, float* out_raw,
           void* out_act, int act, const char* tag = "dn.linear") {
  const int prec = c.prec();
  vfx_conv_desc d = conv_base(a, 1, 1, (int)M, K, getw(c, p + ".w", (size_t)N * K, prec), N);
  d.ntaps = 1;
  d.bias = getf(c, p + ".b", N);
  if (out_raw) set_raw(d, out_raw, N, 0);
  if (out_act) set_act(d, out_act, N, 0, act, 0.f);
  return run_conv(c, prec, d, tag);
}

// fp32 tensor -> dense GEMM operand: the tensor itself in fp32 mode, a bf16 copy in bf16 mode
const void* dn_operand(Ctx& c, const float* x, long long M, int C, void* scratch, int* rc) {
  *rc = VFX_OK;
  if (c.prec() == VFX_PREC_FP32) return x;      // bf16: a bf16 copy; tf32: a copy rounded to tf32
  if (!c.dry) *rc = bn_act(c.prec(), x, M * C, C, 1, M, C, nullptr, nullptr, 1, 0, VFX_ACT_NONE, 0.f, scratch, M * C, C, c.st);
  return scratch;
}

int bn_gru(Ctx& c, const std::string& p, const float* x, int B, int T, void* op, float* gi,
           float* y0, float* y1) {
  const long long M = (long long)B * T;
  const int prec = c.prec();
  // BN2d(1) over the (T,512) plane of each item
  VFX_TRY(bn_act_op(c, prec, p + ".bn", 1, x, (long long)T * 512, 512, B, T, 512, VFX_ACT_NONE, 0.f, op));
  const void* in = op;
  float* outs[2] = {y0, y1};
  for (int layer = 0; layer < 2; ++layer) {
    char nm[64];
    snprintf(nm, sizeof(nm), "%s.l%d", p.c_str(), layer);
    const std::string q(nm);
    vfx_conv_desc d = conv_base(in, 1, 1, (int)M, 512, getw(c, q + ".wih", (size_t)1536 * 512, prec), 1536);
    d.ntaps = 1;
    d.bias = getf(c, q + ".bih", 1536);
    set_raw(d, gi, 1536, 0);
    const float* whh = getf(c, q + ".whh_t", (size_t)2 * 256 * 768);
    const float* bhh = getf(c, q + ".bhh", 2 * 768);
    VFX_TRY(run_conv(c, prec, d, "dn.gru_in"));
    if (!c.dry && c.rc == VFX_OK) {
      ProfScope ps(c, "dn.gru", 2.0 * M * 2 * 768 * 256, 0.0);
      VFX_TRY(gru_layer(gi, whh, bhh, B, T, outs[layer], c.st));
    }
    if (layer == 0) { int rc; in = dn_operand(c, y0, M, 512, op, &rc); VFX_TRY(rc); }
  }
  return c.rc;
}

int relu_inplace(Ctx& c, float* x, int B, int T, int C) {
  if (c.dry) return VFX_OK;
  return bn_act(VFX_PREC_FP32, x, (long long)T * C, C, B, T, C, nullptr, nullptr, 1, 0, VFX_ACT_LRELU, 0.f, x,
                (long long)T * C, C, c.st);
}

int denoiser_forward(Ctx& c, const float* mel, int B, int T, const uint8_t* drop, float* lin_out) {
  const long long M = (long long)B * T;
  const int prec = c.prec();
  void* a = c.ws->alloc<float>(M * 512);          // operand scratch (fp32-sized)
  float* b = c.ws->alloc<float>(M * 512);
  float* y0 = c.ws->alloc<float>(M * 512);
  float* gi = c.ws->alloc<float>(M * 1536);
  const std::string p = "dn.";
  VFX_TRY(bn_act_op(c, prec, p + "bn0", 1, mel, (long long)T * 128, 128, B, T, 128, VFX_ACT_NONE, 0.f, a));
  VFX_TRY(linear(c, a, M, 128, p + "lin1", 256, b, nullptr, 0));
  VFX_TRY(relu_inplace(c, b, B, T, 256));
  VFX_TRY(bn_act_op(c, prec, p + "bn3", 1, b, (long long)T * 256, 256, B, T, 256, VFX_ACT_NONE, 0.f, a));
  VFX_TRY(linear(c, a, M, 256, p + "lin4", 512, b, nullptr, 0));
  if (c.train() && drop && !c.dry) VFX_TRY(dropout_apply(b, drop, M * 512, c.st));
  VFX_TRY(relu_inplace(c, b, B, T, 512));
  VFX_TRY(bn_gru(c, p + "g7", b, B, T, a, gi, y0, b));       // result in b
  VFX_TRY(bn_gru(c, p + "g8", b, B, T, a, gi, y0, b));
  VFX_TRY(bn_act_op(c, prec, p + "bn9", 1, b, (long long)T * 512, 512, B, T, 512, VFX_ACT_LRELU, 0.f, a));
  VFX_TRY(linear(c, a, M, 512, p + "lin11", 512, b, nullptr, 0));
  if (c.train() && drop && !c.dry) VFX_TRY(dropout_apply(b, drop + M * 512, M * 512, c.st));
  VFX_TRY(bn_act_op(c, prec, p + "bn13", 1, b, (long long)T * 512, 512, B, T, 512, VFX_ACT_LRELU, 0.f, a));
  VFX_TRY(linear(c, a, M, 512, p + "lin15", 128, lin_out, nullptr, 0));
  return c.rc;
}

int analysis_forward(Ctx& c, const float* mel, int B, int T, const uint8_t* drop, float* mel_log_out) {
  const int Tp = (T + 63) / 64 * 64;
  // mode 2: the reference's train-mode BatchNorm raises "Expected more than 1 value per channel when training"
  // when the UNet centre is 1x1, i.e. for segments of at most 64 frames
  if (c.train() && !c.dry) VFX_REQUIRE(Tp > 64, "mode 2 needs more than 64 frames per segment (got T=%d): Expected more than 1 value per channel when training", T);
  const size_t m0 = c.ws->mark();
  float* xlog = c.ws->alloc<float>((size_t)B * T * 128);
  float* unet_in = c.ws->alloc<float>((size_t)B * Tp * 127 * 2);
  float* lin = c.ws->alloc<float>((size_t)B * T * 128);
  {
    const size_t m1 = c.ws->mark();
    VFX_TRY(denoiser_forward(c, mel, B, T, drop, lin));
    c.ws->reset(m1);
  }
  if (!c.dry) VFX_TRY(mask_log_pack(lin, mel, B, T, Tp, xlog, unet_in, c.st));
  float* feat = nullptr;
  VFX_TRY(unet_forward(c, unet_in, B, Tp, &feat));
  const float* hw = getf(c, "unet.head.w", 32);
  const float* hb = getf(c, "unet.head.b", 1);
  if (!c.dry && c.rc == VFX_OK) VFX_TRY(unet_head(feat, B, T, Tp, hw, hb, xlog, mel_log_out, c.st));
  c.ws->reset(m0);
  return c.rc;
}

// ------------------------------------------------------------------ vocoder
// vocoder Generator.forward voicefixer/vocoder/model/generator.py:127-145 on cond [B][Tc][128]
// (operand precision), then the fused post conv / tanh / trim.
int vocoder_generator(Ctx& c, const void* cond, int B, int Tc, float* wav_out, int trim_len, float scale) {
  const int prec = c.prec();
  const size_t esz = c.e->esz();
  const size_t m0 = c.ws->mark();
  const long long S = (long long)Tc * 441;
  // condnet buffers
  void* c0 = c.ws->raw((size_t)B * (Tc + 6) * 512 * esz);
  void* c1 = c.ws->raw((size_t)B * (Tc + 6) * 512 * esz);
  // full-rate buffers: X raw fp32, A0 / Hh / U operands
  const size_t maxel = (size_t)B * S * 64;
  float* X = c.ws->alloc<float>(maxel);
  void* A0 = c.ws->raw(maxel * esz);
  void* Hh = c.ws->raw(maxel * esz);
  void* U = c.ws->raw(maxel * esz);
  char name[96];
  // ---- condnet: 5 x (Conv1d k3 p1 + ELU), generator.py:33-54
  const void* in = cond; int Cin = 128;
  void* pp[2] = {c0, c1};
  for (int i = 0; i < 5; ++i) {
    snprintf(name, sizeof(name), "voc.cond%d", i);
    const std::string p(name);
    vfx_conv_desc d = conv_base(in, B, 1, Tc, Cin, getw(c, p + ".w", (size_t)3 * 512 * Cin, prec), 512);
    d.ntaps = 3;
    for (int k = 0; k < 3; ++k) { d.dw[k] = k - 1; d.w_off[k] = (long long)k * 512 * Cin; }
    d.bias = getf(c, p + ".b", 512);
    void* out = pp[i & 1];
    if (i == 4) {   // last ELU output goes to rows 3..Tc+2 of the reflect-padded buffer
      set_act(d, (char*)out + (size_t)3 * 512 * esz, 512, 0, VFX_ACT_ELU, 0.f);
      d.oa_sB = (long long)(Tc + 6) * 512;
    } else {
      set_act(d, out, 512, 0, VFX_ACT_ELU, 0.f);
    }
    VFX_TRY(run_conv(c, prec, d, "voc.condnet"));
    in = out; Cin = 512;
  }
  if (!c.dry) VFX_TRY(reflect_pad3((void*)in, B, Tc, 512, prec, c.st));
  // ---- ReflectionPad1d(3) + Conv1d(512,1024,k7) + LeakyReLU(0.2); the x + sin x of the first
  //      UpsampleNet (modules.py:504) is applied in the same epilogue.
  {
    vfx_conv_desc d = conv_base(in, B, 1, Tc + 6, 512, getw(c, "voc.pre.w", (size_t)7 * 1024 * 512, prec), 1024);
    d.ntaps = 7; d.Wq = Tc; d.OW = Tc;
    for (int k = 0; k < 7; ++k) { d.dw[k] = k; d.w_off[k] = (long long)k * 1024 * 512; }
    d.bias = getf(c, "voc.pre.b", 1024);
    set_act(d, U, 1024, 0, VFX_ACT_LRELU_XSINX, 0.2f);
    VFX_TRY(run_conv(c, prec, d, "voc.pre"));
  }
  // TF32 "encoded stream" (vfx_conv_desc.res_enc / raw_enc): X holds S = bits(lrelu(x)) + 0x1000, which conv1 reads as
  // its tf32 operand and conv2 decodes as the fp32 residual -- no separate activated copy (24 -> 20 bytes per element and pair)
  const bool enc = prec == VFX_PREC_TF32 && c.e->use_tc && c.e->tf32_stream;
  long long Lin = Tc;
  for (int j = 0; j < 4; ++j) {
    const int Ci = VOC_CIN[j], Co = VOC_COUT[j], u = VOC_U[j];
    const long long Lout = Lin * u;
    char up_tag[32], c1_tag[32], c2_tag[32];
    snprintf(up_tag, sizeof(up_tag), "voc.up%d", j);
    snprintf(c1_tag, sizeof(c1_tag), "voc.rs%d.c1", j);
    snprintf(c2_tag, sizeof(c2_tag), "voc.rs%d.c2", j);
    // ---- UpsampleNet: ConvTranspose1d(k=2u, s=u, p, op) as two phase-group GEMMs (modules.py:451-459)
    snprintf(name, sizeof(name), "voc.up%d", j);
    {
      const std::string p(name);
      const void* w = getw(c, p + ".w", (size_t)2 * u * Co * Ci, prec);
      const float* bias = getf(c, p + ".b", Co);
      const int pad = u / 2 + u % 2, nA = u - pad;
      const long long mat = (long long)Co * Ci;
      for (int grp = 0; grp < 2; ++grp) {
        const int nph = grp == 0 ? nA : u - nA;
        // Output described as the reshaped view [B][Lin][u*Co] (row q holds output samples q*u .. q*u+u-1):
        // the phase group is then a plain stride-1 GEMM writing columns [r0*Co, (r0+nph)*Co) of row q, which
        // lets the tensor-core kernel use its TMA-staged epilogue.
        vfx_conv_desc d = conv_base(U, B, 1, (int)Lin, Ci, w, nph * Co);
        d.ntaps = 2;
        int r0;
        if (grp == 0) { d.dw[0] = 0; d.w_off[0] = pad * mat; d.dw[1] = -1; d.w_off[1] = (pad + u) * mat; r0 = 0; }
        else          { d.dw[0] = 1; d.w_off[0] = 0;         d.dw[1] = 0;  d.w_off[1] = u * mat;         r0 = nA; }
        d.bias = bias; d.bias_mod = Co;
        set_raw(d, X, (long long)u * Co, r0 * Co);
        if (enc) { d.raw_enc = 1; d.enc_slope = 0.01f; }
        else set_act(d, A0, (long long)u * Co, r0 * Co, VFX_ACT_LRELU, 0.01f);
        VFX_TRY(run_conv(c, prec, d, up_tag));
      }
    }
    // ---- ResStack: 8 x (x + conv_k3_d1(lrelu(conv_k3_d3^i(lrelu(x))))), modules.py:550-576,592-595
    // bf16, width 64: each pair is ONE fused kernel (h stays on chip); the activated operand copy ping-pongs between
    // A0 and Hh because a pair reads it with a halo that neighbouring tiles would overwrite in place.
    const bool is16 = prec == VFX_PREC_BF16 || prec == VFX_PREC_FP16;          // 2-byte operand modes share every kernel
    const bool fuse_pairs = is16 && c.e->use_tc &&
                            ((c.e->fuse_pair && Co == 64) || (c.e->fuse_pair2 && Co == 128));
    // tf32, width 64: the encoded stream ping-pongs between X and the (otherwise unused) operand buffer; 8 pairs end in X
    const bool fuse_tf32 = enc && (c.e->fuse_pair3 || c.e->fuse_pair2 >= 2) && Co == 64 && j == 3;
    const int tf32_impl = c.e->fuse_pair3 ? 3 : 2;
    void* a_cur = A0; void* a_nxt = Hh;
    float* s_cur = X; float* s_nxt = (float*)A0;
    void* pair_scratch = nullptr;                       // two-CTA pipeline: the h tiles cross SMs through this (L2-resident) ring
    if ((fuse_tf32 && tf32_impl == 2) || (fuse_pairs && Co == 128)) pair_scratch = c.ws->raw(resstack_pair2_scratch_bytes());
    int dil = 1;
    for (int i = 0; i < 8; ++i, dil *= 3) {
      snprintf(name, sizeof(name), "voc.rs%d.l%d", j, i);
      const std::string p(name);
      if (fuse_tf32) {
        vfx_pair_desc pd;
        memset(&pd, 0, sizeof(pd));
        pd.a = s_cur; pd.x = s_cur; pd.x_out = s_nxt; pd.dilation = dil; pd.B = B; pd.L = (int)Lout; pd.C = Co;
        pd.w1 = getw(c, p + ".c1.w", (size_t)3 * Co * Co, prec); pd.b1 = getf(c, p + ".c1.b", Co);
        pd.w2 = getw(c, p + ".c2.w", (size_t)3 * Co * Co, prec); pd.b2 = getf(c, p + ".c2.b", Co);
        pd.write_raw = 1; pd.precision = VFX_PREC_TF32; pd.impl = tf32_impl;
        pd.scratch = pair_scratch; pd.scratch_bytes = resstack_pair2_scratch_bytes();
        pd.stream_enc = 1; pd.stream_enc_out = i < 7 ? 1 : 0;     // the last pair hands plain x' to the final convolution
        if (!c.dry && c.rc == VFX_OK) {
          char pt[48], ptd[64];
          snprintf(pt, sizeof(pt), "voc.rs%d.pair", j);
          snprintf(ptd, sizeof(ptd), "%s.d%d", pt, dil);
          const double els = (double)B * Lout * Co;
          ProfScope ps(c, c.e->profile > 1 ? ptd : pt, 2.0 * 2.0 * 3.0 * Co * els, 8.0 * els);
          VFX_TRY(tf32_impl == 3 ? resstack_pair3_tc(pd, c.st) : resstack_pair2_tc(pd, c.st));
        }
        float* t = s_cur; s_cur = s_nxt; s_nxt = t;
        continue;
      }
      if (fuse_pairs) {
        vfx_pair_desc pd;
        memset(&pd, 0, sizeof(pd));
        pd.precision = prec; pd.impl = Co == 64 ? 1 : 2;
        pd.scratch = pair_scratch; pd.scratch_bytes = resstack_pair2_scratch_bytes();
        pd.a = a_cur; pd.x = X; pd.dilation = dil; pd.B = B; pd.L = (int)Lout; pd.C = Co;
        pd.w1 = getw(c, p + ".c1.w", (size_t)3 * Co * Co, prec); pd.b1 = getf(c, p + ".c1.b", Co);
        pd.w2 = getw(c, p + ".c2.w", (size_t)3 * Co * Co, prec); pd.b2 = getf(c, p + ".c2.b", Co);
        pd.write_raw = (i < 7 || j == 3) ? 1 : 0;           // the next up-sampler consumes the operand only
        if (i < 7) { pd.out_act = a_nxt; pd.act = VFX_ACT_LRELU; pd.act_param = 0.01f; }
        else if (j < 3) { pd.out_act = U; pd.act = VFX_ACT_LRELU_XSINX; pd.act_param = 0.2f; }
        if (!c.dry && c.rc == VFX_OK) {
          char pt[48];
          snprintf(pt, sizeof(pt), "voc.rs%d.pair", j);
          char ptd[64];
          snprintf(ptd, sizeof(ptd), "%s.d%d", pt, dil);
          const double els = (double)B * Lout * Co;
          // algorithmic bytes of a pair: x in + x' out, fp32
          ProfScope ps(c, c.e->profile > 1 ? ptd : pt, 2.0 * 2.0 * 3.0 * Co * els, 8.0 * els);
          VFX_TRY(pd.impl == 1 ? resstack_pair_tc(pd, c.st) : resstack_pair2_tc(pd, c.st));
        }
        void* t = a_cur; a_cur = a_nxt; a_nxt = t;
        continue;
      }
      {
        vfx_conv_desc d = conv_base(enc ? (const void*)X : (const void*)A0, B, 1, (int)Lout, Co,
                                    getw(c, p + ".c1.w", (size_t)3 * Co * Co, prec), Co);
        d.ntaps = 3;
        for (int k = 0; k < 3; ++k) { d.dw[k] = (k - 1) * dil; d.w_off[k] = (long long)k * Co * Co; }
        d.bias = getf(c, p + ".c1.b", Co);
        set_act(d, Hh, Co, 0, VFX_ACT_LRELU, 0.01f);
        char c1d[48];
        snprintf(c1d, sizeof(c1d), "%s.d%d", c1_tag, dil);
        VFX_TRY(run_conv(c, prec, d, c.e->profile > 1 ? c1d : c1_tag));
      }
      {
        vfx_conv_desc d = conv_base(Hh, B, 1, (int)Lout, Co, getw(c, p + ".c2.w", (size_t)3 * Co * Co, prec), Co);
        d.ntaps = 3;
        for (int k = 0; k < 3; ++k) { d.dw[k] = k - 1; d.w_off[k] = (long long)k * Co * Co; }
        d.bias = getf(c, p + ".c2.b", Co);
        set_res(d, X, Co, 0);
        if (enc) {
          d.res_enc = 1; d.enc_slope = 0.01f;
          if (i < 7) { set_raw(d, X, Co, 0); d.raw_enc = 1; }               // S' in place
          else if (j == 3) set_raw(d, X, Co, 0);                            // last stack: plain x' for the final conv
          else set_act(d, U, Co, 0, VFX_ACT_LRELU_XSINX, 0.2f);             // next UpsampleNet's operand only
        } else {
          set_raw(d, X, Co, 0);
          if (i < 7) set_act(d, A0, Co, 0, VFX_ACT_LRELU, 0.01f);
          else if (j < 3) set_act(d, U, Co, 0, VFX_ACT_LRELU_XSINX, 0.2f);   // act + next UpsampleNet's x+sin x
        }
        VFX_TRY(run_conv(c, prec, d, c2_tag));
      }
    }
    Lin = Lout;
  }
  // ---- LeakyReLU(0.2) + ReflectionPad1d(3) + Conv1d(64,1,7) + Tanh + _trim_center
  int lo = 0; long long out_len = S;
  if (trim_len >= 0) {
    VFX_REQUIRE(trim_len <= S, "vocoder: trim_len %d exceeds generated length %lld", trim_len, S);
    const long long diff = S - trim_len;
    lo = (int)(diff / 2); out_len = trim_len;
    VFX_REQUIRE(diff == 0 || diff / 2 >= 1, "vocoder: _trim_center is degenerate for a length difference of 1");
  }
  const float* pw = getf(c, "voc.post.w", 7 * 64);
  const float* pb = getf(c, "voc.post.b", 1);
  ProfScope ps(c, "voc.post", 2.0 * B * out_len * 7 * 64, (double)B * S * 64 * 4);
  if (!c.dry && c.rc == VFX_OK) VFX_TRY(voc_post(X, B, (int)S, pw, pb, lo, (int)out_len, scale, wav_out, c.st));
  c.ws->reset(m0);
  return c.rc;
}

int vocoder_forward(Ctx& c, const float* mel, int B, int T, int input_is_log, float* wav_out, int trim_len,
                    float scale) {
  const int Tc = T + T % 2 + 4;
  const size_t m0 = c.ws->mark();
  void* cond = c.ws->raw((size_t)B * Tc * 128 * c.e->esz());
  const float* tab = getf(c, "voc.mel_tab", 129);
  if (!c.dry && c.rc == VFX_OK) VFX_TRY(voc_normalize(mel, B, T, Tc, input_is_log, tab, cond, c.prec(), c.st));
  VFX_TRY(vocoder_generator(c, cond, B, Tc, wav_out, trim_len, scale));
  c.ws->reset(m0);
  return c.rc;
}

int frontend_forward(Ctx& c, const float* wav, int B, int L, float* mel, float* sp) {
  const int T = 1 + L / 441;
  const float* fbT = getf(c, "fe.fbT", (size_t)128 * 1025);
  const int* fs = (const int*)get(c, "fe.fb_start", 128 * 4);
  const int* fl = (const int*)get(c, "fe.fb_len", 128 * 4);
  if (c.dry || c.rc != VFX_OK) return c.rc;
  ProfScope ps(c, "fe.stft_mel", (double)B * T * (5.0 * 1024 * 10 + 2.0 * 2050), (double)B * L * 4);
  return stft_mel(wav, B, L, T, c.e->d_window, c.e->d_tw, fbT, fs, fl, mel, sp, c.st);
}

int restore_forward(Ctx& c, const float* wav, int B, int L, const uint8_t* drop, float* wav_out) {
  const int T = 1 + L / 441;
  float* mel = c.ws->alloc<float>((size_t)B * T * 128);
  float* mel_log = c.ws->alloc<float>((size_t)B * T * 128);
  VFX_TRY(frontend_forward(c, wav, B, L, mel, nullptr));
  VFX_TRY(analysis_forward(c, mel, B, T, drop, mel_log));
  VFX_TRY(vocoder_forward(c, mel_log, B, T, 1, wav_out, L, 1.0f));
  return c.rc;
}

}  // namespace
}  // namespace vfx

// ====================================================================== C ABI
using namespace vfx;

extern "C" {

const char* vfx_last_error(void) { return vfx::g_err; }
int vfx_version(void) { return 100; }

int vfx_engine_create(vfx_engine** out, int device, int precision) {
  VFX_REQUIRE(out != nullptr, "engine_create: out is null");
  VFX_REQUIRE(precision == VFX_PREC_FP32 || precision == VFX_PREC_BF16 || precision == VFX_PREC_TF32 || precision == VFX_PREC_FP16,
              "engine_create: bad precision %d", precision);
  if (device == -1) {   // planning-only: weight-set validation and workspace sizing on a host without a GPU
    vfx_engine* e = new vfx_engine();
    e->device = -1; e->precision = precision;
    *out = e;
    return VFX_OK;
  }
  int ndev = 0;
  VFX_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  VFX_REQUIRE(device >= 0 && device < ndev, "engine_create: device %d not present (%d devices)", device, ndev);
  VFX_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  VFX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("vfx_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return VFX_ERR_UNSUPPORTED;
  }
  vfx_engine* e = new vfx_engine();
  e->device = device; e->precision = precision;
  if (getenv("VFX_FUSE_PAIR3")) e->fuse_pair3 = atoi(getenv("VFX_FUSE_PAIR3"));
  if (getenv("VFX_FUSE_PAIR2")) e->fuse_pair2 = atoi(getenv("VFX_FUSE_PAIR2"));      // A/B knob for bench runs
  std::vector<float> win(2048);
  std::vector<float2> tw(1024);
  for (int n = 0; n < 2048; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 2048.0));
  for (int k = 0; k < 1024; ++k) {
    tw[k].x = (float)cos(-2.0 * M_PI * k / 2048.0);
    tw[k].y = (float)sin(-2.0 * M_PI * k / 2048.0);
  }
  VFX_CUDA_CHECK(cudaMalloc(&e->d_window, 2048 * sizeof(float)));
  VFX_CUDA_CHECK(cudaMalloc(&e->d_tw, 1024 * sizeof(float2)));
  VFX_CUDA_CHECK(cudaMemcpy(e->d_window, win.data(), 2048 * sizeof(float), cudaMemcpyHostToDevice));
  VFX_CUDA_CHECK(cudaMemcpy(e->d_tw, tw.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
  *out = e;
  return VFX_OK;
}

int vfx_engine_destroy(vfx_engine* e) {
  if (!e) return VFX_OK;
  if (e->device >= 0) {
    cudaFree(e->d_window);
    cudaFree(e->d_tw);
  }
  delete e;
  return VFX_OK;
}

unsigned long long vfx_launch_count(void) { return vfx::g_launches; }

int vfx_profile_report(vfx_engine* e, char* buf, size_t cap) {
  VFX_REQUIRE(e && buf && cap > 0, "profile_report: bad arguments");
  VFX_REQUIRE(e->device >= 0, "profile_report: planning-only engine");
  VFX_CUDA_CHECK(cudaSetDevice(e->device));
  VFX_CUDA_CHECK(cudaDeviceSynchronize());
  struct Agg { double ms = 0, flops = 0, bytes = 0; long n = 0; };
  std::vector<std::pair<std::string, Agg>> agg;
  for (auto& r : e->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    size_t i = 0;
    for (; i < agg.size(); ++i) if (agg[i].first == r.tag) break;
    if (i == agg.size()) agg.push_back({r.tag, Agg()});
    agg[i].second.ms += ms; agg[i].second.flops += r.flops; agg[i].second.bytes += r.bytes; agg[i].second.n++;
  }
  e->prof.clear();
  std::string out;
  char line[256];
  for (auto& a : agg) {
    snprintf(line, sizeof(line), "%s %ld %.6f %.6e %.6e\n", a.first.c_str(), a.second.n, a.second.ms, a.second.flops,
             a.second.bytes);
    out += line;
  }
  snprintf(buf, cap, "%s", out.c_str());
  return VFX_OK;
}

int vfx_engine_set_option(vfx_engine* e, const char* key, int value) {
  VFX_REQUIRE(e && key, "set_option: null argument");
  if (!strcmp(key, "use_tc")) { e->use_tc = value; return VFX_OK; }
  if (!strcmp(key, "profile")) { e->profile = value; return VFX_OK; }
  if (!strcmp(key, "fuse_pair")) { e->fuse_pair = value; return VFX_OK; }
  if (!strcmp(key, "fuse_pair2")) { e->fuse_pair2 = value; return VFX_OK; }
  if (!strcmp(key, "fuse_pair3")) { e->fuse_pair3 = value; return VFX_OK; }
  if (!strcmp(key, "tf32_stream")) { e->tf32_stream = value; return VFX_OK; }
  set_error("set_option: unknown key '%s'", key);
  return VFX_ERR_INVALID;
}

int vfx_engine_set_tensor(vfx_engine* e, const char* name, const void* dev_ptr, size_t bytes) {
  VFX_REQUIRE(e && name && dev_ptr, "set_tensor: null argument");
  e->tensors[name] = vfx::Tensor{dev_ptr, bytes};
  e->finalized = false;
  return VFX_OK;
}

static int dry_run(vfx_engine* e, int B, int T, int L, size_t* bytes) {
  Bump ws(nullptr, 0);
  Ctx c{e, nullptr, &ws, true, VFX_MODE_TRAIN_BN};
  int r;
  if (L > 0) r = restore_forward(c, nullptr, B, L, nullptr, nullptr);
  else {
    float* mel = ws.alloc<float>((size_t)B * T * 128);
    const bool voc_only = e->finalized && !(e->parts & VFX_PART_ANALYSIS);
    r = voc_only ? VFX_OK : analysis_forward(c, mel, B, T, nullptr, mel);
    if (r == VFX_OK) r = vocoder_forward(c, mel, B, T, 1, nullptr, -1, 1.f);
  }
  *bytes = ws.peak + 4096;
  return r;
}

int vfx_engine_finalize(vfx_engine* e) {
  VFX_REQUIRE(e, "finalize: null engine");
  e->missing.clear();
  size_t bytes = 0;
  int r = dry_run(e, 1, 64, 0, &bytes);
  {   // also the front-end tables
    Bump ws(nullptr, 0);
    Ctx c{e, nullptr, &ws, true, VFX_MODE_EVAL};
    int r2 = frontend_forward(c, nullptr, 1, 44100, nullptr, nullptr);
    if (r == VFX_OK) r = r2;
  }
  if (!e->missing.empty()) {
    // The reference's stand-alone Vocoder class (voicefixer/vocoder/base.py:10-40) loads the synthesis checkpoint only.
    // An engine that holds nothing but "voc." tensors, all of them present, is complete for vfx_vocoder /
    // vfx_vocoder_cond; the analysis-side entry points refuse it.
    bool voc_only = !e->tensors.empty();
    for (auto& kv : e->tensors) if (kv.first.rfind("voc.", 0) != 0) voc_only = false;
    if (voc_only) {
      std::vector<std::string> analysis_missing;
      analysis_missing.swap(e->missing);
      Bump ws(nullptr, 0);
      Ctx c{e, nullptr, &ws, true, VFX_MODE_EVAL};
      float* mel = ws.alloc<float>((size_t)64 * 128);
      const int rv = vocoder_forward(c, mel, 1, 64, 1, nullptr, -1, 1.f);
      if (rv == VFX_OK && c.rc == VFX_OK && e->missing.empty()) {
        e->parts = VFX_PART_VOCODER;
        e->finalized = true;
        return VFX_OK;
      }
      if (e->missing.empty()) return rv != VFX_OK ? rv : c.rc;     // e.g. a tensor of the wrong size: message already set
    }
    std::string s = "missing weight tensors:";
    for (size_t i = 0; i < e->missing.size() && i < 12; ++i) s += " " + e->missing[i];
    if (e->missing.size() > 12) s += " ...";
    set_error("%s (%zu total)", s.c_str(), e->missing.size());
    return VFX_ERR_MISSING_WEIGHT;
  }
  if (r != VFX_OK) return r;
  e->parts = VFX_PART_ANALYSIS | VFX_PART_VOCODER;
  e->finalized = true;
  return VFX_OK;
}

size_t vfx_workspace_bytes(const vfx_engine* e, int B, int L) {
  if (!e || B <= 0 || L <= 1024) return 0;
  if (e->finalized && !(e->parts & VFX_PART_ANALYSIS)) return 0;   // restore() needs the analysis weights
  size_t bytes = 0;
  dry_run(const_cast<vfx_engine*>(e), B, 0, L, &bytes);
  return bytes;
}

size_t vfx_workspace_bytes_frames(const vfx_engine* e, int B, int T) {
  if (!e || B <= 0 || T <= 0) return 0;
  size_t bytes = 0;
  dry_run(const_cast<vfx_engine*>(e), B, T, 0, &bytes);
  return bytes;
}

#define VFX_ENTER(e)                                                                     \
  VFX_REQUIRE((e) != nullptr, "null engine");                                            \
  VFX_REQUIRE((e)->finalized, "engine not finalized (call vfx_engine_finalize)");        \
  VFX_REQUIRE((e)->device >= 0, "planning-only engine (device -1) cannot launch work");  \
  VFX_CUDA_CHECK(cudaSetDevice((e)->device))

#define VFX_NEED_ANALYSIS(e)                                                              \
  VFX_REQUIRE((e)->parts & VFX_PART_ANALYSIS,                                            \
              "this engine holds the vocoder weights only (stand-alone Vocoder); the analysis module is not loaded")

// Pre-flight: the same forward pass on a null allocator gives the exact workspace need of THIS call; a too-small
// workspace is refused before the first launch (a bump allocator that overflows mid-sequence would alias buffers).
#define VFX_PREFLIGHT(mode_, call_)                                                                       \
  do {                                                                                                    \
    Bump dry_(nullptr, 0);                                                                                \
    Ctx c{e, nullptr, &dry_, true, (mode_)};                                                              \
    int r_ = (call_);                                                                                     \
    if (r_ != VFX_OK) return r_;                                                                          \
    if (dry_.peak > workspace_bytes) {                                                                    \
      set_error("workspace too small: %zu bytes given, this call needs %zu", (size_t)workspace_bytes, dry_.peak); \
      return VFX_ERR_WORKSPACE;                                                                           \
    }                                                                                                     \
  } while (0)

#define VFX_FINISH(c, ws)                                                                \
  if ((ws).overflow) { set_error("workspace too small: need %zu bytes", (ws).peak); return VFX_ERR_WORKSPACE; } \
  return (c).rc

int vfx_frontend_mel(vfx_engine* e, const float* wav, int B, int L, float* mel, float* sp_out, void* stream) {
  VFX_ENTER(e);
  VFX_NEED_ANALYSIS(e);
  VFX_REQUIRE(wav && mel && B > 0, "frontend: bad arguments");
  Bump ws(nullptr, 0);
  Ctx c{e, (cudaStream_t)stream, &ws, false, VFX_MODE_EVAL};
  return frontend_forward(c, wav, B, L, mel, sp_out);
}

int vfx_analysis(vfx_engine* e, const float* mel, int B, int T, int mode, const uint8_t* drop_masks,
                 float* mel_log_out, void* workspace, size_t workspace_bytes, void* stream) {
  VFX_ENTER(e);
  VFX_NEED_ANALYSIS(e);
  VFX_REQUIRE(mel && mel_log_out && B > 0 && T > 0 && workspace, "analysis: bad arguments");
  VFX_REQUIRE(mode == VFX_MODE_EVAL || mode == VFX_MODE_TRAIN_BN, "analysis: bad mode %d", mode);
  VFX_PREFLIGHT(mode, analysis_forward(c, mel, B, T, nullptr, mel_log_out));
  Bump ws(workspace, workspace_bytes);
  Ctx c{e, (cudaStream_t)stream, &ws, false, mode};
  int r = analysis_forward(c, mel, B, T, drop_masks, mel_log_out);
  if (r != VFX_OK) return r;
  VFX_FINISH(c, ws);
}

int vfx_vocoder(vfx_engine* e, const float* mel, int B, int T, int input_is_log, float* wav_out, int trim_len,
                float scale, void* workspace, size_t workspace_bytes, void* stream) {
  VFX_ENTER(e);
  VFX_REQUIRE(mel && wav_out && B > 0 && T > 0 && workspace, "vocoder: bad arguments");
  VFX_PREFLIGHT(VFX_MODE_EVAL, vocoder_forward(c, mel, B, T, input_is_log, wav_out, trim_len, scale));
  Bump ws(workspace, workspace_bytes);
  Ctx c{e, (cudaStream_t)stream, &ws, false, VFX_MODE_EVAL};
  int r = vocoder_forward(c, mel, B, T, input_is_log, wav_out, trim_len, scale);
  if (r != VFX_OK) return r;
  VFX_FINISH(c, ws);
}

int vfx_vocoder_cond(vfx_engine* e, const float* cond, int B, int Tc, float* wav_out, int trim_len, float scale,
                     void* workspace, size_t workspace_bytes, void* stream) {
  VFX_ENTER(e);
  VFX_REQUIRE(cond && wav_out && B > 0 && Tc >= 4 && workspace, "vocoder_cond: bad arguments");
  VFX_PREFLIGHT(VFX_MODE_EVAL, (dry_.raw((size_t)B * Tc * 128 * e->esz()), vocoder_generator(c, nullptr, B, Tc, wav_out, trim_len, scale)));
  Bump ws(workspace, workspace_bytes);
  Ctx c{e, (cudaStream_t)stream, &ws, false, VFX_MODE_EVAL};
  void* op = ws.raw((size_t)B * Tc * 128 * e->esz());
  if (ws.overflow) { set_error("workspace too small"); return VFX_ERR_WORKSPACE; }
  int r = cast_rows(cond, (long long)B * Tc * 128, op, e->precision, c.st);
  if (r != VFX_OK) return r;
  r = vocoder_generator(c, op, B, Tc, wav_out, trim_len, scale);
  if (r != VFX_OK) return r;
  VFX_FINISH(c, ws);
}

int vfx_restore(vfx_engine* e, const float* wav, int B, int L, int mode, const uint8_t* drop_masks,
                float* wav_out, void* workspace, size_t workspace_bytes, void* stream) {
  VFX_ENTER(e);
  VFX_NEED_ANALYSIS(e);
  VFX_REQUIRE(wav && wav_out && B > 0 && workspace, "restore: bad arguments");
  VFX_REQUIRE(mode == VFX_MODE_EVAL || mode == VFX_MODE_TRAIN_BN, "restore: bad mode %d", mode);
  VFX_REQUIRE(L > 1024, "restore: L=%d must exceed 1024 samples", L);
  VFX_PREFLIGHT(mode, restore_forward(c, wav, B, L, nullptr, wav_out));
  Bump ws(workspace, workspace_bytes);
  Ctx c{e, (cudaStream_t)stream, &ws, false, mode};
  int r = restore_forward(c, wav, B, L, drop_masks, wav_out);
  if (r != VFX_OK) return r;
  VFX_FINISH(c, ws);
}

int vfx_hf_cut(vfx_engine* e, const float* wav, int B, int L, float ratio, float* wav_out, int* cut_bins,
               void* workspace, size_t workspace_bytes, void* stream) {
  VFX_ENTER(e);
  return hf_cut(wav, B, L, ratio, e->d_window, e->d_tw, wav_out, cut_bins, workspace, workspace_bytes,
                (cudaStream_t)stream);
}

int vfx_conv_gemm(int precision, int impl, const vfx_conv_desc* d, void* stream) {
  VFX_REQUIRE(d, "conv_gemm: null descriptor");
  if (impl == 1) {
    VFX_REQUIRE(precision == VFX_PREC_BF16 || precision == VFX_PREC_TF32 || precision == VFX_PREC_FP16,
                "conv_gemm: the tcgen05 implementation takes bf16, fp16 or tf32 operands");
    return conv_gemm_tc(precision, *d, (cudaStream_t)stream);
  }
  return conv_gemm_simt(precision, *d, (cudaStream_t)stream);
}

size_t vfx_resstack_pair_scratch_bytes(void) { return resstack_pair2_scratch_bytes(); }

int vfx_resstack_pair(const vfx_pair_desc* d, void* stream) {
  VFX_REQUIRE(d, "resstack_pair: null descriptor");
  const bool tf32 = d->precision == VFX_PREC_TF32;
  const bool one_cta = d->impl == 1 || (d->impl == 0 && !tf32 && d->C == 64 && !d->x_out);
  int r;
  if (tf32 && (d->impl == 3 || d->impl == 0)) {            // one SM, residual stashed in TMEM; the two-CTA form on request
    r = resstack_pair3_tc(*d, (cudaStream_t)stream);
    if (r == VFX_ERR_UNSUPPORTED && d->impl == 0) r = resstack_pair2_tc(*d, (cudaStream_t)stream);
  } else if (d->impl == 3) r = VFX_ERR_UNSUPPORTED;
  else
    r = one_cta ? ((tf32 || d->x_out) ? VFX_ERR_UNSUPPORTED : resstack_pair_tc(*d, (cudaStream_t)stream))
                : resstack_pair2_tc(*d, (cudaStream_t)stream);
  if (r == VFX_ERR_UNSUPPORTED)
    set_error("resstack_pair: unsupported shape (bf16: C = 64 or 128; tf32: C = 64 with stream_enc; 16-byte aligned tensors)");
  return r;
}

int vfx_gru_layer(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out, void* stream) {
  return gru_layer(gi, whh_t, bhh, B, T, out, (cudaStream_t)stream);
}

}  // extern "C"
