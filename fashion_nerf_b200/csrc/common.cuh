// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include "../../include/fnerf.h"
#include "layout.h"

namespace fnerf {

// thread-local error text behind fnerf_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

// process-wide count of kernels this library has launched (fnerf_launch_count(); bench.py reports it as gpu_launches)
void note_launches(int n);

// called after every kernel launch (n = kernels launched since the previous call)
inline int check_launch(const char* what, int n = 1) {
  note_launches(n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

#define FN_REQUIRE(cond, code, ...) \
  do { if (!(cond)) return ::fnerf::set_error((code), __VA_ARGS__); } while (0)
#define FN_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)

// One-time, per-device initialisation (cudaFuncSetAttribute, occupancy queries, device properties) guarded by
// std::call_once, as SURVEY.md 8(b) prescribes: the entry points are re-entrant and may be called from several
// host threads on different streams.  The callable's cudaError_t is kept and returned on every later call.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
  std::once_flag flag[kMaxDevices];
  cudaError_t err[kMaxDevices];
};
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}
template <class F>
inline cudaError_t device_once(DeviceOnce& o, F&& f) {
  const int dev = current_device();
  std::call_once(o.flag[dev], [&] { o.err[dev] = f(); });
  return o.err[dev];
}
// kernels that need more than 48 KB of dynamic shared memory opt in once per device
template <class K>
inline cudaError_t opt_in_smem_once(DeviceOnce& o, K kernel, size_t bytes) {
  return device_once(o, [&] { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); });
}

inline int num_sms() {
  static DeviceOnce once;
  static int sms[kMaxDevices];
  const int dev = current_device();
  device_once(once, [&] {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
    return cudaSuccess;
  });
  return sms[dev];
}

// kernel launchers implemented in the individual .cu files (host side, return 0 / cudaError_t)
int launch_ray_setup(const float* rays_d, float* viewdirs, float* dnorm, int64_t R, cudaStream_t s);
int launch_stratified(const float* near, const float* far, const float* t_vals, const float* u,
                      float* z, int64_t R, int64_t N, int lindisp, cudaStream_t s);
int launch_importance(const float* z_c, const float* w_c, const float* u, int64_t u_stride,
                      float* z_samples, float* z_f, int32_t* bin_idx, float* z_std, int64_t R,
                      int64_t Nc, int64_t Nf, cudaStream_t s);
int launch_debug_fdiv(int64_t n, uint64_t seed, unsigned long long* mismatches, cudaStream_t s);
int launch_posenc(const float* x, float* out, int64_t M, int L, cudaStream_t s);
int launch_composite_fwd(const float* raw, const float* z, const float* dnorm, const float* noise,
                         float* rgb, float* depth, float* acc, float* disp, float* weights,
                         int64_t R, int64_t S, int white, cudaStream_t s);
int launch_composite_bwd(const float* raw, const float* z, const float* dnorm, const float* noise, const float* g_rgb,
                         const float* g_depth, const float* g_acc, float* g_raw, int64_t R,
                         int64_t S, int white, cudaStream_t s);
int launch_pack(const float* flat, void* packed, int cond, cudaStream_t s);
int launch_unpack(const void* packed, float* flat, int cond, cudaStream_t s);
int launch_allreduce_adam(const float* const* peer_grads, int world, int64_t offset, float* p, float* m, float* v, int64_t n, float lr,
                          float b1, float b2, float eps, int64_t t, float grad_scale, cudaStream_t s);
int launch_multimem_allreduce(float* mc, int rank, int world, int64_t n, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int64_t t,
                float grad_scale, cudaStream_t s);
int launch_cond_project(const void* packed, const float* cond, float* proj, int64_t C, cudaStream_t s);

struct MlpArgs {
  const void* packed; int cond;
  const float* rays_o; const float* rays_d; const float* viewdirs; const float* z;
  const float* cond_proj; const int32_t* cond_index; int64_t C;
  float* raw; int64_t R, S;
};
// outputs / extra inputs of the network query with fused compositing (mlp_tc.cu)
struct CompositeOut {
  const float* dnorm; const float* noise;
  float* rgb; float* depth; float* acc; float* disp; float* weights;
  int white;
};
int mlp_tc_composite_group(int64_t S);
int launch_mlp_tc_composite(const MlpArgs& a, const CompositeOut& c, cudaStream_t s);
int launch_mlp_fp32(const MlpArgs& a, cudaStream_t s);
int launch_mlp_tc(const MlpArgs& a, cudaStream_t s);
int launch_mlp_tc_tape(const MlpArgs& a, uint8_t* tape, uint32_t* mask_tape, cudaStream_t s);   // training forward: also writes the tape

int64_t mlp_bwd_workspace_bytes(int64_t R, int64_t S);
int64_t mlp_bwd_tc_workspace_bytes(int64_t R, int64_t S);
int64_t mlp_tape_bytes(int64_t M);                       // forward tape of one bf16 network query
int64_t mlp_bwd_from_tape_workspace_bytes(int64_t M);
int launch_mlp_fwd_tape(const MlpArgs& a, void* tape, cudaStream_t s);
int launch_mlp_bwd_from_tape(const void* packed, int cond, const float* g_raw, const void* tape, const float* cond_rows,
                             const int32_t* cond_index, int64_t C, int64_t S, float* flat_grad, void* ws, int64_t M,
                             cudaStream_t s);
int launch_mlp_bwd_tc(const MlpArgs& a, const float* g_raw, float* flat_grad, void* ws, int64_t ws_bytes, cudaStream_t s);
int launch_mlp_bwd_fp32(const MlpArgs& a, const float* g_raw, float* flat_grad, void* ws,
                        int64_t ws_bytes, cudaStream_t s);

}  // namespace fnerf
