// C ABI of libfnerf.so (include/fnerf.h): argument validation, error text, and the render_rays
// orchestration (A.9).  Everything here only enqueues work on the caller's stream.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace fnerf {

static std::atomic<int64_t> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct RenderWorkspace {
  int64_t viewdirs, dnorm, z_c, raw_c, weights_c, z_samples, z_f, raw_f, depth0, total;
};
static RenderWorkspace render_layout(int64_t R, int64_t Nc, int64_t Nf) {
  RenderWorkspace w;
  int64_t off = 0;
  auto take = [&](int64_t floats) { int64_t o = off; off = align_up(off + floats * 4, 256); return o; };
  w.viewdirs = take(R * 3);
  w.dnorm = take(R);
  w.z_c = take(R * Nc);
  w.raw_c = take(R * Nc * 4);
  w.weights_c = take(R * Nc);
  w.z_samples = take(R * (Nf > 0 ? Nf : 1));
  w.z_f = take(R * (Nc + Nf));
  w.raw_f = take(R * (Nc + Nf) * 4);
  w.depth0 = take(R);
  w.total = off;
  return w;
}

}  // namespace fnerf

using namespace fnerf;

extern "C" {

int fnerf_abi_version(void) { return FNERF_ABI_VERSION; }
int64_t fnerf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* fnerf_last_error(void) { return error_buffer(); }

int64_t fnerf_param_count(int cond) { return flat_count(cond ? 1 : 0); }
int64_t fnerf_packed_bytes(int cond) { return (packed_bytes(cond ? 1 : 0) + 255) / 256 * 256; }

int fnerf_pack_weights(const float* flat, void* packed, int cond, fnerf_stream_t stream) {
  FN_REQUIRE(flat && packed, FNERF_ERR_NULL, "pack_weights: null pointer");
  FN_REQUIRE(FN_ALIGNED16(packed), FNERF_ERR_ALIGN, "pack_weights: packed must be 16-byte aligned");
  return launch_pack(flat, packed, cond ? 1 : 0, (cudaStream_t)stream);
}

int fnerf_unpack_weights(const void* packed, float* flat, int cond, fnerf_stream_t stream) {
  FN_REQUIRE(flat && packed, FNERF_ERR_NULL, "unpack_weights: null pointer");
  return launch_unpack(packed, flat, cond ? 1 : 0, (cudaStream_t)stream);
}

int fnerf_ray_setup(const float* rays_d, float* viewdirs, float* dnorm, int64_t R, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0, FNERF_ERR_SIZE, "ray_setup: R < 0");
  if (R == 0) return 0;
  FN_REQUIRE(rays_d && viewdirs && dnorm, FNERF_ERR_NULL, "ray_setup: null pointer");
  return launch_ray_setup(rays_d, viewdirs, dnorm, R, (cudaStream_t)stream);
}

int fnerf_stratified(const float* near, const float* far, const float* t_vals, const float* u_strat,
                     float* z, int64_t R, int64_t N, int lindisp, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0 && N >= 1 && N <= (1 << 20), FNERF_ERR_SIZE, "stratified: bad R=%lld N=%lld", (long long)R, (long long)N);
  if (R == 0) return 0;
  FN_REQUIRE(near && far && t_vals && z, FNERF_ERR_NULL, "stratified: null pointer");
  FN_REQUIRE(R * N < ((int64_t)1 << 40), FNERF_ERR_SIZE, "stratified: R*N too large");
  return launch_stratified(near, far, t_vals, u_strat, z, R, N, lindisp, (cudaStream_t)stream);
}

int fnerf_debug_fdiv_mismatches(int64_t n, uint64_t seed, unsigned long long* mismatches, fnerf_stream_t stream) {
  FN_REQUIRE(n >= 0, FNERF_ERR_SIZE, "debug_fdiv: n < 0");
  FN_REQUIRE(mismatches, FNERF_ERR_NULL, "debug_fdiv: null pointer");
  return launch_debug_fdiv(n, seed, mismatches, (cudaStream_t)stream);
}

int fnerf_importance(const float* z_c, const float* weights_c, const float* u, int64_t u_row_stride,
                     float* z_samples, float* z_f, int32_t* bin_idx, float* z_std, int64_t R,
                     int64_t Nc, int64_t Nf, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0, FNERF_ERR_SIZE, "importance: R < 0");
  FN_REQUIRE(Nc >= 3 && Nc <= 4096, FNERF_ERR_SIZE, "importance: Nc=%lld outside [3,4096]", (long long)Nc);
  FN_REQUIRE(Nf >= 1 && Nf <= 8192, FNERF_ERR_SIZE, "importance: Nf=%lld outside [1,8192]", (long long)Nf);
  FN_REQUIRE(u_row_stride == 0 || u_row_stride >= Nf, FNERF_ERR_ARG, "importance: u_row_stride must be 0 or >= Nf");
  if (R == 0) return 0;
  FN_REQUIRE(z_c && weights_c && u && z_samples && z_f, FNERF_ERR_NULL, "importance: null pointer");
  return launch_importance(z_c, weights_c, u, u_row_stride, z_samples, z_f, bin_idx, z_std, R, Nc, Nf,
                           (cudaStream_t)stream);
}

int fnerf_posenc(const float* x, float* out, int64_t M, int L, fnerf_stream_t stream) {
  FN_REQUIRE(M >= 0 && L >= 0 && L <= 24, FNERF_ERR_SIZE, "posenc: bad M=%lld L=%d", (long long)M, L);
  if (M == 0) return 0;
  FN_REQUIRE(x && out, FNERF_ERR_NULL, "posenc: null pointer");
  return launch_posenc(x, out, M, L, (cudaStream_t)stream);
}

int fnerf_cond_project(const void* packed, const float* cond, float* proj, int64_t C, fnerf_stream_t stream) {
  FN_REQUIRE(C >= 0, FNERF_ERR_SIZE, "cond_project: C < 0");
  if (C == 0) return 0;
  FN_REQUIRE(packed && cond && proj, FNERF_ERR_NULL, "cond_project: null pointer");
  return launch_cond_project(packed, cond, proj, C, (cudaStream_t)stream);
}

static int validate_mlp(const char* who, int precision, const void* packed, int cond, const float* rays_o,
                        const float* rays_d, const float* viewdirs, const float* z, const float* cond_proj,
                        const int32_t* cond_index, int64_t C, const float* out, int64_t R, int64_t S) {
  FN_REQUIRE(precision == FNERF_PRECISION_FP32 || precision == FNERF_PRECISION_BF16, FNERF_ERR_ARG,
             "%s: unknown precision %d", who, precision);
  FN_REQUIRE(R >= 0 && S >= 1 && S <= (1 << 20), FNERF_ERR_SIZE, "%s: bad R=%lld S=%lld", who, (long long)R, (long long)S);
  if (R == 0) return 0;
  FN_REQUIRE(packed && rays_o && rays_d && viewdirs && z && out, FNERF_ERR_NULL, "%s: null pointer", who);
  FN_REQUIRE(FN_ALIGNED16(packed) && FN_ALIGNED16(out), FNERF_ERR_ALIGN, "%s: packed/raw must be 16-byte aligned", who);
  if (cond) {
    FN_REQUIRE(cond_proj != nullptr, FNERF_ERR_NULL, "%s: cond=1 needs cond_proj", who);
    FN_REQUIRE(C >= 1, FNERF_ERR_SIZE, "%s: cond=1 needs C >= 1", who);
    FN_REQUIRE(cond_index != nullptr || C == 1 || C == R, FNERF_ERR_ARG,
               "%s: cond_index is required unless C == 1 or C == R", who);
  }
  return 0;
}

int fnerf_mlp_fwd(int precision, const void* packed, int cond, const float* rays_o, const float* rays_d,
                  const float* viewdirs, const float* z, const float* cond_proj, const int32_t* cond_index,
                  int64_t C, float* raw, int64_t R, int64_t S, fnerf_stream_t stream) {
  int rc = validate_mlp("mlp_fwd", precision, packed, cond, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, raw, R, S);
  if (rc != 0 || R == 0) return rc;
  MlpArgs a{packed, cond ? 1 : 0, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, raw, R, S};
  return precision == FNERF_PRECISION_BF16 ? launch_mlp_tc(a, (cudaStream_t)stream)
                                           : launch_mlp_fp32(a, (cudaStream_t)stream);
}

// the path fnerf_mlp_bwd takes for (precision, cond): bf16 tape + tcgen05 backward, or the fp32 SGEMM chain
static bool bwd_uses_tc(int precision, int cond) { return precision == FNERF_PRECISION_BF16 && !cond; }

int64_t fnerf_mlp_bwd_workspace_bytes(int precision, int cond, int64_t R, int64_t S) {
  if (R < 0 || S < 1) return -1;
  return bwd_uses_tc(precision, cond) ? mlp_bwd_tc_workspace_bytes(R, S) : mlp_bwd_workspace_bytes(R, S);
}

int fnerf_mlp_bwd(int precision, const void* packed, int cond, const float* rays_o, const float* rays_d,
                  const float* viewdirs, const float* z, const float* cond_proj, const int32_t* cond_index,
                  int64_t C, const float* g_raw, float* flat_grad, void* workspace, int64_t workspace_bytes,
                  int64_t R, int64_t S, fnerf_stream_t stream) {
  int rc = validate_mlp("mlp_bwd", precision, packed, cond, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, g_raw, R, S);
  if (rc != 0 || R == 0) return rc;
  FN_REQUIRE(flat_grad && workspace, FNERF_ERR_NULL, "mlp_bwd: null pointer");
  FN_REQUIRE(workspace_bytes >= fnerf_mlp_bwd_workspace_bytes(precision, cond, R, S), FNERF_ERR_WORKSPACE, "mlp_bwd: workspace too small");
  FN_REQUIRE(FN_ALIGNED16(workspace) && FN_ALIGNED16(g_raw), FNERF_ERR_ALIGN, "mlp_bwd: workspace / g_raw must be 16-byte aligned");
  MlpArgs a{packed, cond ? 1 : 0, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, nullptr, R, S};
  // bf16: forward with tape + tcgen05 dgrad chain + tcgen05 wgrad (unconditioned networks); the conditioned
  // variant and FNERF_PRECISION_FP32 take the fp32 SGEMM chain
  if (bwd_uses_tc(precision, cond))
    return launch_mlp_bwd_tc(a, g_raw, flat_grad, workspace, workspace_bytes, (cudaStream_t)stream);
  return launch_mlp_bwd_fp32(a, g_raw, flat_grad, workspace, workspace_bytes, (cudaStream_t)stream);
}

int fnerf_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, int64_t step, float grad_scale, fnerf_stream_t stream) {
  FN_REQUIRE(n >= 0 && step >= 1, FNERF_ERR_SIZE, "adam_step: bad n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return 0;
  FN_REQUIRE(params && grad && exp_avg && exp_avg_sq, FNERF_ERR_NULL, "adam_step: null pointer");
  return launch_adam(params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

int fnerf_allreduce_adam_step(const float* const* peer_grads, int world, int64_t offset, float* params, float* exp_avg,
                              float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                              float grad_scale, fnerf_stream_t stream) {
  FN_REQUIRE(n >= 0 && step >= 1 && offset >= 0, FNERF_ERR_SIZE, "allreduce_adam_step: bad n=%lld step=%lld", (long long)n, (long long)step);
  FN_REQUIRE(world >= 1 && world <= 64, FNERF_ERR_SIZE, "allreduce_adam_step: bad world=%d", world);
  if (n == 0) return 0;
  FN_REQUIRE(peer_grads && params && exp_avg && exp_avg_sq, FNERF_ERR_NULL, "allreduce_adam_step: null pointer");
  return launch_allreduce_adam(peer_grads, world, offset, params, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale,
                               (cudaStream_t)stream);
}

int fnerf_multimem_allreduce(float* multicast_ptr, int rank, int world, int64_t n, fnerf_stream_t stream) {
  FN_REQUIRE(n >= 0 && (n & 3) == 0, FNERF_ERR_SIZE, "multimem_allreduce: n=%lld must be a multiple of 4", (long long)n);
  FN_REQUIRE(world >= 1 && rank >= 0 && rank < world, FNERF_ERR_SIZE, "multimem_allreduce: bad rank %d / world %d", rank, world);
  if (n == 0) return 0;
  FN_REQUIRE(multicast_ptr != nullptr, FNERF_ERR_NULL, "multimem_allreduce: null multicast pointer");
  FN_REQUIRE(FN_ALIGNED16(multicast_ptr), FNERF_ERR_ALIGN, "multimem_allreduce: pointer must be 16-byte aligned");
  return launch_multimem_allreduce(multicast_ptr, rank, world, n, (cudaStream_t)stream);
}

int fnerf_mlp_fwd_composite_supported(int64_t S) { return mlp_tc_composite_group(S) > 0 ? 1 : 0; }

int fnerf_mlp_fwd_composite(const void* packed, int cond, const float* rays_o, const float* rays_d, const float* viewdirs,
                            const float* dnorm, const float* z, const float* cond_proj, const int32_t* cond_index, int64_t C,
                            const float* raw_noise, float* raw, float* rgb, float* depth, float* acc, float* disp, float* weights,
                            int64_t R, int64_t S, int white_bkgd, fnerf_stream_t stream) {
  // `raw` is nullable here: validate with a stand-in for the output pointer
  int rc = validate_mlp("mlp_fwd_composite", FNERF_PRECISION_BF16, packed, cond, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C,
                        raw ? raw : reinterpret_cast<const float*>(packed), R, S);
  if (rc != 0 || R == 0) return rc;
  FN_REQUIRE(dnorm && rgb && depth && acc && disp, FNERF_ERR_NULL, "mlp_fwd_composite: null pointer");
  FN_REQUIRE(mlp_tc_composite_group(S) > 0, FNERF_ERR_SIZE, "mlp_fwd_composite: S=%lld is not served by the fused epilogue "
             "(see fnerf_mlp_fwd_composite_supported)", (long long)S);
  MlpArgs a{packed, cond ? 1 : 0, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, raw, R, S};
  CompositeOut c{dnorm, raw_noise, rgb, depth, acc, disp, weights, white_bkgd};
  return launch_mlp_tc_composite(a, c, (cudaStream_t)stream);
}

int64_t fnerf_mlp_tape_bytes(int64_t R, int64_t S) { return mlp_tape_bytes(R * S); }
int64_t fnerf_mlp_bwd_tape_workspace_bytes(int64_t R, int64_t S) { return mlp_bwd_from_tape_workspace_bytes(R * S); }

int fnerf_mlp_fwd_tape(const void* packed, int cond, const float* rays_o, const float* rays_d, const float* viewdirs,
                       const float* z, const float* cond_proj, const int32_t* cond_index, int64_t C, float* raw,
                       void* tape, int64_t tape_bytes, int64_t R, int64_t S, fnerf_stream_t stream) {
  int rc = validate_mlp("mlp_fwd_tape", FNERF_PRECISION_BF16, packed, cond, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, raw, R, S);
  if (rc != 0 || R == 0) return rc;
  FN_REQUIRE(tape != nullptr, FNERF_ERR_NULL, "mlp_fwd_tape: null tape");
  FN_REQUIRE(FN_ALIGNED16(tape), FNERF_ERR_ALIGN, "mlp_fwd_tape: tape must be 16-byte aligned");
  FN_REQUIRE(tape_bytes >= mlp_tape_bytes(R * S), FNERF_ERR_WORKSPACE, "mlp_fwd_tape: tape too small");
  MlpArgs a{packed, cond ? 1 : 0, rays_o, rays_d, viewdirs, z, cond_proj, cond_index, C, raw, R, S};
  return launch_mlp_fwd_tape(a, tape, (cudaStream_t)stream);
}

int fnerf_mlp_bwd_tape(const void* packed, int cond, const float* g_raw, const void* tape, int64_t tape_bytes,
                       const float* cond_rows, const int32_t* cond_index, int64_t C, float* flat_grad, void* workspace,
                       int64_t workspace_bytes, int64_t R, int64_t S, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0 && S >= 1, FNERF_ERR_SIZE, "mlp_bwd_tape: bad R=%lld S=%lld", (long long)R, (long long)S);
  if (R == 0) return 0;
  FN_REQUIRE(packed && g_raw && tape && flat_grad && workspace, FNERF_ERR_NULL, "mlp_bwd_tape: null pointer");
  if (cond) {
    FN_REQUIRE(cond_rows != nullptr && C >= 1, FNERF_ERR_NULL, "mlp_bwd_tape: conditioned network needs the raw codes cond_rows[C,256]");
    FN_REQUIRE(FN_ALIGNED16(cond_rows), FNERF_ERR_ALIGN, "mlp_bwd_tape: cond_rows must be 16-byte aligned");
    FN_REQUIRE(cond_index != nullptr || C == 1 || C == R, FNERF_ERR_SIZE, "mlp_bwd_tape: C must be 1 or R without cond_index");
  }
  FN_REQUIRE(FN_ALIGNED16(packed) && FN_ALIGNED16(g_raw) && FN_ALIGNED16(tape) && FN_ALIGNED16(workspace), FNERF_ERR_ALIGN,
             "mlp_bwd_tape: packed / g_raw / tape / workspace must be 16-byte aligned");
  FN_REQUIRE(tape_bytes >= mlp_tape_bytes(R * S), FNERF_ERR_WORKSPACE, "mlp_bwd_tape: tape too small");
  FN_REQUIRE(workspace_bytes >= mlp_bwd_from_tape_workspace_bytes(R * S), FNERF_ERR_WORKSPACE, "mlp_bwd_tape: workspace too small");
  return launch_mlp_bwd_from_tape(packed, cond ? 1 : 0, g_raw, tape, cond_rows, cond_index, C, S, flat_grad, workspace, R * S,
                                  (cudaStream_t)stream);
}

int fnerf_composite_fwd(const float* raw, const float* z, const float* dnorm, const float* raw_noise,
                        float* rgb, float* depth, float* acc, float* disp, float* weights, int64_t R,
                        int64_t S, int white_bkgd, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0 && S >= 1 && S <= (1 << 20), FNERF_ERR_SIZE, "composite_fwd: bad R=%lld S=%lld", (long long)R, (long long)S);
  if (R == 0) return 0;
  FN_REQUIRE(raw && z && dnorm && rgb && depth && acc && disp, FNERF_ERR_NULL, "composite_fwd: null pointer");
  FN_REQUIRE(FN_ALIGNED16(raw), FNERF_ERR_ALIGN, "composite_fwd: raw must be 16-byte aligned");
  return launch_composite_fwd(raw, z, dnorm, raw_noise, rgb, depth, acc, disp, weights, R, S, white_bkgd,
                              (cudaStream_t)stream);
}

int fnerf_composite_bwd(const float* raw, const float* z, const float* dnorm, const float* raw_noise, const float* g_rgb,
                        const float* g_depth, const float* g_acc, float* g_raw, int64_t R, int64_t S,
                        int white_bkgd, fnerf_stream_t stream) {
  FN_REQUIRE(R >= 0 && S >= 1 && S <= 4096, FNERF_ERR_SIZE, "composite_bwd: bad R=%lld S=%lld", (long long)R, (long long)S);
  if (R == 0) return 0;
  FN_REQUIRE(raw && z && dnorm && g_rgb && g_raw, FNERF_ERR_NULL, "composite_bwd: null pointer");
  FN_REQUIRE(FN_ALIGNED16(raw) && FN_ALIGNED16(g_raw), FNERF_ERR_ALIGN, "composite_bwd: raw/g_raw must be 16-byte aligned");
  return launch_composite_bwd(raw, z, dnorm, raw_noise, g_rgb, g_depth, g_acc, g_raw, R, S, white_bkgd, (cudaStream_t)stream);
}

int64_t fnerf_render_rays_workspace_bytes(int64_t R, int64_t Nc, int64_t Nf) {
  if (R < 0 || Nc < 1 || Nf < 0) return -1;
  return render_layout(R, Nc, Nf).total;
}

int fnerf_render_rays(const fnerf_render_args* a, fnerf_stream_t stream) {
  FN_REQUIRE(a != nullptr, FNERF_ERR_NULL, "render_rays: null args");
  const int64_t R = a->R, Nc = a->Nc, Nf = a->Nf;
  FN_REQUIRE(R >= 0 && Nc >= 1 && Nf >= 0, FNERF_ERR_SIZE, "render_rays: bad sizes");
  FN_REQUIRE(Nf == 0 || Nc >= 3, FNERF_ERR_SIZE, "render_rays: importance sampling needs Nc >= 3");
  if (R == 0) return 0;
  FN_REQUIRE(a->packed_coarse && a->rays_o && a->rays_d && a->near && a->far && a->t_vals, FNERF_ERR_NULL,
             "render_rays: null input");
  FN_REQUIRE(a->rgb && a->disp && a->acc && a->depth && a->rgb0 && a->disp0 && a->acc0 && a->z_std, FNERF_ERR_NULL,
             "render_rays: null output");
  FN_REQUIRE(Nf == 0 || (a->packed_fine && a->u_fine), FNERF_ERR_NULL, "render_rays: fine pass needs packed_fine and u_fine");
  FN_REQUIRE(a->workspace != nullptr, FNERF_ERR_NULL, "render_rays: null workspace");
  FN_REQUIRE((!a->tape_coarse && !a->tape_fine) || a->precision == FNERF_PRECISION_BF16, FNERF_ERR_ARG,
             "render_rays: training tapes need the bf16 path");
  FN_REQUIRE(FN_ALIGNED16(a->workspace), FNERF_ERR_ALIGN, "render_rays: workspace must be 16-byte aligned");
  const RenderWorkspace L = render_layout(R, Nc, Nf);
  FN_REQUIRE(a->workspace_bytes >= L.total, FNERF_ERR_WORKSPACE, "render_rays: workspace %lld < %lld",
             (long long)a->workspace_bytes, (long long)L.total);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  auto f = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  float* viewdirs = f(L.viewdirs);
  float* dnorm = f(L.dnorm);
  float* z_c = a->z_c ? a->z_c : f(L.z_c);
  float* raw_c = a->raw_c ? a->raw_c : f(L.raw_c);
  float* weights_c = a->weights_c ? a->weights_c : f(L.weights_c);
  float* z_samples = f(L.z_samples);
  float* z_f = a->z_f ? a->z_f : f(L.z_f);
  float* raw_f = a->raw_f ? a->raw_f : f(L.raw_f);
  float* depth0 = a->depth0 ? a->depth0 : f(L.depth0);
  FN_REQUIRE(FN_ALIGNED16(raw_c) && FN_ALIGNED16(raw_f), FNERF_ERR_ALIGN, "render_rays: raw taps must be 16-byte aligned");

  int rc;
  auto record = [&](void* ev) -> int {            // optional profiling events: a failed record is an error, not a silent gap
    if (ev == nullptr) return 0;
    cudaError_t e = cudaEventRecord((cudaEvent_t)ev, s);
    return e == cudaSuccess ? 0 : set_error((int)e, "render_rays: cudaEventRecord: %s", cudaGetErrorString(e));
  };
  auto d2d = [&](float* dst, const float* src, int64_t floats) -> int {
    cudaError_t e = cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? 0 : set_error((int)e, "render_rays: cudaMemcpyAsync: %s", cudaGetErrorString(e));
  };
  if ((rc = fnerf_ray_setup(a->rays_d, viewdirs, dnorm, R, stream))) return rc;
  if ((rc = fnerf_stratified(a->near, a->far, a->t_vals, a->u_strat, z_c, R, Nc, a->lindisp, stream))) return rc;
  // fused path: compositing inside the network-query kernel, raw[R,S,4] written only if the caller asks
  // for the tap.  bf16 inference only (the training tape keeps the separate kernels: the backward needs raw anyway).
  const bool fuse_c = a->fuse_composite && a->precision == FNERF_PRECISION_BF16 && !a->tape_coarse && mlp_tc_composite_group(Nc) > 0;
  const bool fuse_f = Nf > 0 && a->fuse_composite && a->precision == FNERF_PRECISION_BF16 && !a->tape_fine && mlp_tc_composite_group(Nc + Nf) > 0;
  if ((rc = record(a->ev_coarse_start))) return rc;
  if (fuse_c) {
    float* rgb_c = Nf == 0 ? a->rgb : a->rgb0;
    float* dep_c = Nf == 0 ? a->depth : depth0;
    float* acc_c = Nf == 0 ? a->acc : a->acc0;
    float* dis_c = Nf == 0 ? a->disp : a->disp0;
    if ((rc = fnerf_mlp_fwd_composite(a->packed_coarse, a->cond, a->rays_o, a->rays_d, viewdirs, dnorm, z_c, a->cond_proj_coarse,
                                      a->cond_index, a->C, a->raw_noise_coarse, a->raw_c, rgb_c, dep_c, acc_c, dis_c,
                                      (Nf > 0 || a->weights_c) ? weights_c : nullptr, R, Nc, a->white_bkgd, stream))) return rc;
  } else if (a->tape_coarse) {
    if ((rc = fnerf_mlp_fwd_tape(a->packed_coarse, a->cond, a->rays_o, a->rays_d, viewdirs, z_c, a->cond_proj_coarse,
                                 a->cond_index, a->C, raw_c, a->tape_coarse, a->tape_coarse_bytes, R, Nc, stream))) return rc;
  } else if ((rc = fnerf_mlp_fwd(a->precision, a->packed_coarse, a->cond, a->rays_o, a->rays_d, viewdirs, z_c,
                                 a->cond_proj_coarse, a->cond_index, a->C, raw_c, R, Nc, stream))) return rc;
  if ((rc = record(a->ev_coarse_stop))) return rc;
  if (Nf == 0) {
    if (!fuse_c && (rc = fnerf_composite_fwd(raw_c, z_c, dnorm, a->raw_noise_coarse, a->rgb, a->depth, a->acc, a->disp,
                                             a->weights_c ? weights_c : nullptr, R, Nc, a->white_bkgd, stream))) return rc;
    if ((rc = d2d(a->rgb0, a->rgb, R * 3)) || (rc = d2d(a->disp0, a->disp, R)) || (rc = d2d(a->acc0, a->acc, R))) return rc;
    if (a->depth0 && (rc = d2d(a->depth0, a->depth, R))) return rc;
    if (cudaError_t e = cudaMemsetAsync(a->z_std, 0, R * sizeof(float), s))
      return set_error((int)e, "render_rays: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return check_launch("render_rays");
  }
  if (!fuse_c && (rc = fnerf_composite_fwd(raw_c, z_c, dnorm, a->raw_noise_coarse, a->rgb0, depth0, a->acc0, a->disp0, weights_c, R, Nc,
                                           a->white_bkgd, stream))) return rc;
  if ((rc = fnerf_importance(z_c, weights_c, a->u_fine, a->u_fine_row_stride, z_samples, z_f, nullptr, a->z_std,
                             R, Nc, Nf, stream))) return rc;
  if ((rc = record(a->ev_fine_start))) return rc;
  if (fuse_f) {
    if ((rc = fnerf_mlp_fwd_composite(a->packed_fine, a->cond, a->rays_o, a->rays_d, viewdirs, dnorm, z_f, a->cond_proj_fine,
                                      a->cond_index, a->C, a->raw_noise_fine, a->raw_f, a->rgb, a->depth, a->acc, a->disp, a->weights_f,
                                      R, Nc + Nf, a->white_bkgd, stream))) return rc;
  } else if (a->tape_fine) {
    if ((rc = fnerf_mlp_fwd_tape(a->packed_fine, a->cond, a->rays_o, a->rays_d, viewdirs, z_f, a->cond_proj_fine,
                                 a->cond_index, a->C, raw_f, a->tape_fine, a->tape_fine_bytes, R, Nc + Nf, stream))) return rc;
  } else if ((rc = fnerf_mlp_fwd(a->precision, a->packed_fine, a->cond, a->rays_o, a->rays_d, viewdirs, z_f,
                                 a->cond_proj_fine, a->cond_index, a->C, raw_f, R, Nc + Nf, stream))) return rc;
  if ((rc = record(a->ev_fine_stop))) return rc;
  if (!fuse_f && (rc = fnerf_composite_fwd(raw_f, z_f, dnorm, a->raw_noise_fine, a->rgb, a->depth, a->acc, a->disp, a->weights_f, R,
                                           Nc + Nf, a->white_bkgd, stream))) return rc;
  return 0;
}

}  // extern "C"
