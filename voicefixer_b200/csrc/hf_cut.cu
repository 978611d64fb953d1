// mode-1 pre-filter (placeholder).
#include "vfx_common.cuh"
namespace vfx {
size_t hf_cut_workspace(int B, int L) { (void)B; (void)L; return 0; }
int hf_cut(const float*, int, int, float, const float*, const float2*, float*, int*, void*, size_t, cudaStream_t) {
  set_error("vfx_hf_cut: not implemented yet");
  return VFX_ERR_UNSUPPORTED;
}
}
