// tcgen05 implementation of the conv GEMM (placeholder until the TMA/TMEM kernel lands).
#include "vfx_common.cuh"
namespace vfx {
int conv_gemm_tc(const vfx_conv_desc& d, cudaStream_t st) { (void)d; (void)st; return VFX_ERR_UNSUPPORTED; }
}
