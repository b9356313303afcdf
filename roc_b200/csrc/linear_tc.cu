// linear_tc.cu — tcgen05/TMEM tensor-core GEMMs for the Linear op (3xTF32).
// Placeholder until the tensor-core kernels land: every entry reports
// ROC_ERR_UNSUPPORTED so the dispatcher in linear.cu takes the exact-fp32 SIMT path.
#include "common.cuh"
namespace roc {
int tc_linear_fwd(int64_t, int, int, const float*, int64_t, const float*, float*, int64_t, int, const uint64_t*,
                  uint64_t, cudaStream_t) { return ROC_ERR_UNSUPPORTED; }
size_t tc_dw_workspace_bytes(int64_t, int, int) { return 0; }
int tc_linear_dw(int64_t, int, int, const float*, int64_t, const float*, int64_t, float*, float*, size_t,
                 cudaStream_t) { return ROC_ERR_UNSUPPORTED; }
}  // namespace roc
