// bf16 tcgen05 path of the box-classifier head (placeholder until the tensor-core kernels land).
#include "c2d_common.cuh"
#include "c2d_head_plan.h"

using namespace c2d;

extern "C" {
int c2d_has_tensor_core_head(void) { return 0; }
int c2d_head_mixed5_fwd_bf16(const void*, int, const float*, const HeadPlan&, char*, const float*, float, float*,
                             cudaStream_t) {
  set_error("head: bf16 tcgen05 path not built yet");
  return C2D_ERR_UNSUPPORTED;
}
int c2d_head_mixed5_bwd_bf16(const void*, int, const float*, const HeadPlan&, char*, const float*, float,
                             const float*, float*, void*, cudaStream_t) {
  set_error("head: bf16 tcgen05 path not built yet");
  return C2D_ERR_UNSUPPORTED;
}
}
