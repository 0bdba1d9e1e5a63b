// tcgen05 (5th-gen tensor core) engine -- placeholder until the kernels land: reports every
// shape as unsupported so FRCNN_ENGINE_AUTO resolves to the fp32 CUDA-core engine.
#include "common.cuh"

namespace frcnn {

bool tc_fwd_supported(int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_fwd_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
int tc_conv2d_fwd(const float *, const float *, const float *, const float *, const float *, float *,
                  int, int, int, int, int, int, int, int, int, int, void *, size_t, cudaStream_t)
{ return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: not built"); }
bool tc_dgrad_supported(int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_dgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
int tc_conv2d_dgrad(const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, void *, size_t, cudaStream_t)
{ return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: not built"); }
bool tc_wgrad_supported(int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_wgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
int tc_conv2d_wgrad(const float *, const float *, float *, int, int, int, int, int, int, int, int, int, void *, size_t, cudaStream_t)
{ return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: not built"); }

}  // namespace frcnn
