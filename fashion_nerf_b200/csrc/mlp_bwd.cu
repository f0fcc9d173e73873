#include "common.cuh"
namespace fnerf {
int64_t mlp_bwd_workspace_bytes(int64_t, int64_t) { return 256; }
int launch_mlp_bwd_fp32(const MlpArgs&, const float*, float*, void*, int64_t, cudaStream_t) {
  return set_error(FNERF_ERR_ARG, "mlp_bwd: not built");
}
}
