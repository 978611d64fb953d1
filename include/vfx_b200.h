/* vfx_b200.h — C ABI of the B200-native VoiceFixer restore() hot path.
 *
 * The reference (haoheliu/voicefixer, pure Python/PyTorch) has no FFI; its boundary for this path
 * is five nn.Module.forward seams.  Each entry point below replaces one of them and
 * cites it (paths relative to the reference checkout).  Conventions:
 *   - extern "C", plain pointers and sizes, no torch types; all data pointers are DEVICE pointers
 *     unless the name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 (VFX_OK) or a negative vfx_status; vfx_last_error() gives the text.
 *   - no allocation inside compute calls: the caller passes a workspace obtained from
 *     vfx_workspace_bytes().  Weights are registered once by name (device pointers owned by caller).
 *   - tensors are fp32 unless stated; activations inside the engine are channels-last.
 */
#ifndef VFX_B200_H
#define VFX_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vfx_engine vfx_engine;

enum vfx_status {
  VFX_OK = 0,
  VFX_ERR_INVALID = -1,     /* bad argument / shape (mirrors the reference's AssertionError/ValueError) */
  VFX_ERR_MISSING_WEIGHT = -2,
  VFX_ERR_WORKSPACE = -3,   /* workspace too small */
  VFX_ERR_CUDA = -4,
  VFX_ERR_UNSUPPORTED = -5
};

/* Arithmetic of the convolution GEMMs.
 *   FP32: SIMT fp32 FMA (validation path, bit-for-bit deterministic).
 *   BF16: tcgen05 tensor-core MMA (kind::f16), bf16 operands and weights, fp32 accumulation in TMEM.
 *   TF32: tcgen05 tensor-core MMA (kind::tf32), fp32 storage with operands and weights rounded to tf32
 *         (round-to-nearest) by their producers, fp32 accumulation -- the arithmetic of the reference's own
 *         CUDA path (torch.backends.cudnn.allow_tf32 defaults to True; the reference's test/test.py:27-35
 *         compares that path with its CPU path).  Weights are registered in the FP32 layout.
 *   FP16: tcgen05 kind::f16 on fp16 operands and weights, fp32 accumulation: the SAME 10-bit mantissa as tf32 (so the same
 *         waveform parity) in 2-byte storage, i.e. the BF16 mode's memory traffic, kernels and speed -- at fp16's exponent
 *         range (|x| < 65504; smaller than 6e-5 loses relative precision).  Everything outside the GEMM operands stays fp32. */
enum vfx_precision { VFX_PREC_FP32 = 0, VFX_PREC_BF16 = 1, VFX_PREC_TF32 = 2, VFX_PREC_FP16 = 3 };

/* restore() modes, voicefixer/base.py:110-115. mode 1's pre-filter is vfx_hf_cut(). */
enum vfx_mode { VFX_MODE_EVAL = 0, VFX_MODE_TRAIN_BN = 2 };

const char* vfx_last_error(void);
int vfx_version(void);

/* ---- engine lifetime ------------------------------------------------------------------ */
/* device >= 0: an engine on that sm_100a device.  device == -1: a planning-only engine that makes no
 * CUDA call -- vfx_engine_set_tensor / vfx_engine_finalize validate a weight set (names and byte sizes)
 * and vfx_workspace_bytes* size the workspace on a host without a GPU; every launching entry point
 * refuses it with VFX_ERR_INVALID. */
int vfx_engine_create(vfx_engine** out, int device, int precision);
int vfx_engine_destroy(vfx_engine* e);

/* Register one weight tensor (device pointer stays owned by the caller; `bytes` is checked
 * against what the engine expects when it is first used).  Names and layouts: DESIGN.md §3.
 * Replaces torch.load + load_state_dict: voicefixer/base.py:15-30, voicefixer/vocoder/base.py:24-32. */
int vfx_engine_set_tensor(vfx_engine* e, const char* name, const void* dev_ptr, size_t bytes);
/* Engine options: "use_tc" (BF16 / TF32; 1 = tcgen05 kernel [default], 0 = SIMT cross-check on the same operands),
 * "profile" (0/1/2, see vfx_profile_report; 2 = one tag per shape / dilation),
 * "fuse_pair" (BF16: width-64 ResStack pairs as one fused kernel, default 1),
 * "tf32_stream" (TF32: vocoder residual streams as one encoded tensor, vfx_conv_desc.res_enc / raw_enc, default 1),
 * "fuse_pair3" (TF32: width-64 ResStack pairs as one fused kernel, residual stashed in tensor memory, default 1),
 * "fuse_pair2" (two-CTA cluster pipeline: 1 [default] = BF16 / FP16 width-128 ResStack pairs, 2 = also TF32 width 64 when
 *  "fuse_pair3" is 0 [no gain there], 0 = off). */
int vfx_engine_set_option(vfx_engine* e, const char* key, int value);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long vfx_launch_count(void);
/* With option "profile" = 1 every launch group is bracketed by CUDA events on the call's stream;
 * this synchronises, writes one line per tag "tag count total_ms flops bytes" and clears the log. */
int vfx_profile_report(vfx_engine* e, char* buf, size_t cap);
/* Resolve every name the engine needs; returns VFX_ERR_MISSING_WEIGHT and lists them otherwise.
 * Two weight sets are complete: analysis + vocoder (VoiceFixer, voicefixer/base.py:11-30), or the
 * vocoder alone -- nothing but "voc." tensors registered, all present (the reference's stand-alone
 * Vocoder class, voicefixer/vocoder/base.py:10-40).  On a vocoder-only engine vfx_frontend_mel,
 * vfx_analysis and vfx_restore return VFX_ERR_INVALID and vfx_workspace_bytes returns 0. */
int vfx_engine_finalize(vfx_engine* e);

/* Bytes of workspace needed by vfx_restore()/vfx_analysis()/vfx_vocoder() for a batch of B
 * segments of L samples each (T = 1 + L/441 frames). */
size_t vfx_workspace_bytes(const vfx_engine* e, int B, int L);
size_t vfx_workspace_bytes_frames(const vfx_engine* e, int B, int T);

/* ---- the five seams -------------------------------------------------------------------- */

/* wav[B][L] -> mel[B][T][128], T = 1 + L/441 (requires L > 1024: reflect pad).
 * Replaces FDomainHelper.wav_to_spectrogram_phase (voicefixer/tools/modules/fDomainHelper.py:88-110,
 * torchlibrosa STFT n_fft 2048 / hop 441 / periodic Hann / reflect) + MelScale.forward
 * (voicefixer/tools/mel_scale.py:63-77) as called from VoiceFixer._pre (voicefixer/base.py:78-85).
 * sp_out (optional, may be NULL): magnitude spectrogram [B][T][1025]. */
int vfx_frontend_mel(vfx_engine* e, const float* wav, int B, int L, float* mel, float* sp_out,
                     void* stream);

/* mel[B][T][128] (linear) -> mel_log[B][T][128] (log10 domain).
 * Replaces restorer Generator.forward (voicefixer/restorer/model.py:103-120): denoiser
 * (:69-99), to_log (tools/pytorch_util.py:18-22), UNetResComplex_100Mb.forward
 * (restorer/model_kqq_bn.py:130-181).  mode: vfx_mode.  drop_masks (mode 2 only, may be NULL =
 * no dropout): two keep-masks uint8 [2][B][T][512] for the Dropout(0.5)s at model.py:75,94. */
int vfx_analysis(vfx_engine* e, const float* mel, int B, int T, int mode,
                 const uint8_t* drop_masks, float* mel_log_out, void* workspace,
                 size_t workspace_bytes, void* stream);

/* mel[B][T][128] -> wav[B][out_len].
 * Replaces Vocoder.forward (voicefixer/vocoder/base.py:42-56: de-weight, dB, normalise, tr_pre)
 * + vocoder Generator.forward (voicefixer/vocoder/model/generator.py:127-145).
 * input_is_log != 0: input is the analysis stage's log10 mel and from_log
 * (tools/pytorch_util.py:25-27) is applied first.  The generator emits S = (T + T%2 + 4)*441
 * samples; trim_len < 0: write all S (out_len = S); else _trim_center (voicefixer/base.py:63-76)
 * to trim_len samples (out_len = trim_len).  scale multiplies the output (oracle(): 2^15). */
int vfx_vocoder(vfx_engine* e, const float* mel, int B, int T, int input_is_log, float* wav_out,
                int trim_len, float scale, void* workspace, size_t workspace_bytes, void* stream);

/* Vocoder Generator alone on normalised conditions cond[B][Tc][128] (channels-last view of the
 * reference's (B,128,Tc)) -> wav[B][Tc*441]*scale.  Used by Vocoder.oracle (vocoder/base.py:74-77). */
int vfx_vocoder_cond(vfx_engine* e, const float* cond, int B, int Tc, float* wav_out, int trim_len,
                     float scale, void* workspace, size_t workspace_bytes, void* stream);

/* Whole restore_inmem loop body (voicefixer/base.py:120-135) for B equal-length segments:
 * wav[B][L] -> wav_out[B][L].  Equivalent to frontend -> analysis -> vocoder(trim_len = L). */
int vfx_restore(vfx_engine* e, const float* wav, int B, int L, int mode, const uint8_t* drop_masks,
                float* wav_out, void* workspace, size_t workspace_bytes, void* stream);

/* mode 1 pre-filter, VoiceFixer.remove_higher_frequency (voicefixer/base.py:87-104) with the
 * librosa 0.10.1 stft/istft defaults (n_fft 2048, hop 512, Hann, center, zero pad):
 * wav[B][L] -> wav_out[B][512*(L/512)].  Returns the per-item cut bin in cut_bins[B] (optional). */
int vfx_hf_cut(vfx_engine* e, const float* wav, int B, int L, float ratio, float* wav_out,
               int* cut_bins, void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-kernel test entry points (tests/ call the kernels through these) --------------- */

/* Generic channels-last "shifted-window" convolution GEMM; the one kernel family behind every
 * Conv1d/Conv2d/ConvTranspose/Linear on the path.  See DESIGN.md §4 for the field semantics. */
typedef struct vfx_conv_desc {
  const void* a;            /* operand activations [B][H][W][Cin], element type per `precision` */
  int B, H, W, Cin;
  long long a_sB, a_sH, a_sW;   /* element strides (channels contiguous) */
  const void* w;            /* weights [*][N][Cin] K-major, element type per `precision` */
  int ntaps;
  int dh[9], dw[9];         /* input offset of each tap relative to the output grid position */
  long long w_off[9];       /* element offset of each tap's [N][Cin] matrix inside w */
  int Hq, Wq;               /* output grid (GEMM M = B*Hq*Wq) */
  int N;                    /* output channels of this GEMM */
  int sh, rh, sw, rw;       /* output coordinate = (qh*sh+rh, qw*sw+rw) (transposed-conv phases) */
  int OH, OW;               /* outputs outside [0,OH)x[0,OW) are dropped */
  float* out_raw;           /* fp32 result (+bias +residual), may be NULL */
  long long o_sB, o_sH, o_sW; int o_col;
  void* out_act;            /* act(result) as next operand (element type per precision), may be NULL */
  long long oa_sB, oa_sH, oa_sW; int oa_col;
  const float* bias; int bias_mod;          /* bias[n % bias_mod], NULL = none */
  const float* residual;    /* fp32, added before activation, NULL = none */
  long long r_sB, r_sH, r_sW; int r_col;
  int act;                  /* vfx_act applied to out_act */
  float act_param;          /* leaky slope */
  const float* act_scale;   /* optional per-channel affine before the activation of out_act:          */
  const float* act_shift;   /* out_act = act(result * act_scale[n] + act_shift[n]) (fused eval-mode BN) */
  /* TF32 "encoded stream" (vocoder ResStacks in VFX_PREC_TF32): ONE fp32 tensor S serves both as the next convolution's
   * tf32 operand and as the lossless carrier of the fp32 residual stream:  S = bits(lrelu(x, enc_slope)) + 0x1000.
   * kind::tf32 ignores the low 13 mantissa bits of an operand, so reading S as an operand yields exactly
   * round-to-nearest-tf32(lrelu(x)); subtracting 0x1000 from the bits and inverting the (bijective) leaky ReLU gives x back.
   * res_enc: `residual` holds S, it is decoded before the add.  raw_enc: `out_raw` receives S of the result. */
  int res_enc, raw_enc;
  float enc_slope;
} vfx_conv_desc;

enum vfx_act { VFX_ACT_NONE = 0, VFX_ACT_LRELU = 1, VFX_ACT_ELU = 2,
               VFX_ACT_LRELU_XSINX = 3 /* v = lrelu(x, p); v + sin v */, VFX_ACT_SIGMOID = 4 };

/* impl: 0 = SIMT, 1 = tcgen05 (BF16, TF32 or FP16). */
int vfx_conv_gemm(int precision, int impl, const vfx_conv_desc* d, void* stream);

/* Fused ResStack pair on tcgen05:
 *   x' = x + conv2_{k3,d=1}( lrelu_0.01( conv1_{k3,dilation}( a ) + b1 ) ) + b2
 * ResStack.forward voicefixer/vocoder/model/modules.py:592-595 (layers :550-576).  The intermediate never leaves the chip.
 * Three implementations (impl: 0 = pick, 1 = one CTA per tile [16-bit operands], 2 = two-CTA cluster pipeline: conv1 on one SM,
 * conv2 on its neighbour, the intermediate crossing through an L2-resident scratch ring, 3 = one CTA per tile in tf32: the
 * residual rides in the operand boxes and is stashed in tensor memory, 8 bytes of HBM traffic per element):
 *   VFX_PREC_BF16 / VFX_PREC_FP16, C = 64 (impl 1), C = 128 (impl 2): a = lrelu_0.01(x) in the operand format [B][L][C]; x fp32 is read as the residual
 *     and, if write_raw, overwritten with x' (or written to x_out); out_act (optional, bf16, must not alias a) receives act(x').
 *   VFX_PREC_TF32, C = 64 (impl 3, or 2), stream_enc = 1: a == x == the encoded stream S (vfx_conv_desc.raw_enc), fp32 [B][L][C];
 *     the result goes to x_out (required, aliasing neither input) as S' (stream_enc_out = 1) or as plain x' (0).
 * w1 / w2: [3][C][C] in the operand format, tap-major.  Other shapes return VFX_ERR_UNSUPPORTED. */
typedef struct vfx_pair_desc {
  const void* a;
  float* x;
  const void* w1; const float* b1; int dilation;
  const void* w2; const float* b2;
  int B, L, C;
  int write_raw;
  void* out_act; int act; float act_param;
  int precision;            /* vfx_precision: VFX_PREC_BF16 (0 is read as BF16 for compatibility) or VFX_PREC_TF32 */
  int impl;
  float* x_out;             /* NULL = in place over x (bf16 only) */
  int stream_enc, stream_enc_out;
  void* scratch;            /* impl 2: device scratch of vfx_resstack_pair_scratch_bytes() bytes, 128-byte aligned (the */
  size_t scratch_bytes;     /* intermediate tile crosses from one SM to its neighbour through it; stays in L2)          */
} vfx_pair_desc;
size_t vfx_resstack_pair_scratch_bytes(void);
int vfx_resstack_pair(const vfx_pair_desc* d, void* stream);

/* One direction-pair GRU layer recurrence: gi[B][T][2][768] (x W_ih^T + b_ih, fwd|bwd),
 * whh_t[2][256][768] (W_hh transposed), bhh[2][768] -> out[B][T][512] (fwd 256 | bwd 256).
 * torch.nn.GRU gate order r,z,n (voicefixer/restorer/model.py:35-42,61). */
int vfx_gru_layer(const float* gi, const float* whh_t, const float* bhh, int B, int T, float* out,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VFX_B200_H */
