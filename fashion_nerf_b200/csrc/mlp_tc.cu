// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace fnerf {
int launch_mlp_tc(const MlpArgs&, cudaStream_t) { return set_error(FNERF_ERR_ARG, "mlp_tc: not built"); }
}
