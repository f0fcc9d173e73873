// bf16 tensor-core weight gradients (A.4 backward, wgrad) for sm_100a.
//
//   dW[n][k] += sum_m dZ[m][n] * X[m][k]          (m = samples; bf16 operands, fp32 accumulation)
//
// Both operands come from the training tape (layout.h): per 128-sample tile, 16 KB K-block images of
// 128 rows (samples) x 64 columns (features) with the 128-byte swizzle -- exactly what the forward /
// dgrad kernels keep in shared memory.  For this GEMM the reduction runs over SAMPLES, i.e. over the rows
// of those images, so they are fed to tcgen05.mma as MN-MAJOR operands (the feature dimension is the
// contiguous one): no transposed copy of any activation or gradient is ever made.
//   A = dZ image(s): M = 128 output features = two 64-column K-blocks (LBO apart), K = samples (rows)
//   B = X  image(s): N = 64 * x_kb input features,                    K = samples (rows)
//   D = one 128 x N fp32 accumulator in TMEM per block of 128 output features, accumulated over every
//       tile the CTA owns, then flushed once with fp32 atomicAdd into the flat gradient buffer.
// One K=16 MMA consumes 16 samples = two 8-row swizzle atoms (SBO = 1024 B); a pipeline stage is half a
// tile (64 samples = 8 KB of every K-block image, one bulk copy each).
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kWgStages = 3;
constexpr int kWgThreads = 192;                    // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr uint32_t kHalfImg = 8192;                // 64 rows of a K-block image
constexpr uint32_t kWgStageBytes = 8 * kHalfImg;   // up to 4 dZ + 4 X half-images
constexpr uint32_t kWgOffBar = kWgStages * kWgStageBytes;
constexpr uint32_t kWgSmem = kWgOffBar + (2 * kWgStages + 1) * 8 + 16 + 1024;

constexpr int kWgMaxJobs = 16;
struct WgradJob {
  const uint8_t* dz; int64_t dz_tile_stride; int dz_slot0; int n_kb;   // dZ images: tile t, K-block kb at dz + t*stride + (slot0+kb)*16K
  const uint8_t* x;  int64_t x_tile_stride;  int x_slot0;  int x_kb;   // X images
  float* dw; int64_t ld_n, ld_k;                                       // dW[n][k] at dw[n*ld_n + (k - k0)*ld_k]
  int k0, n_valid;                                                     // columns k0 <= k < n_valid are written (63 / 27 / 64*x_kb)
  float* bias;                                                         // nullable: bias[n] += sum_m dZ[m][n] (column sums of the A operand)
  int cta0, ncta;                                                      // the CTAs [cta0, cta0 + ncta) share this job's tiles (set by the launcher)
};
// One launch runs every weight-gradient product of a network backward side by side: each job owns a slice
// of the grid proportional to the bytes it streams (the kernel is HBM-bound), so a CTA accumulates ONE
// product over ntiles / ncta tiles and flushes one accumulator -- 12x fewer atomics than one launch per
// product over the whole grid, and no per-product launch tail.
struct WgradBatch {
  WgradJob jobs[kWgMaxJobs];
  int njobs;
  int64_t ntiles;
};

__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {   // A and B both MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const WgradBatch B) {
  int job = 0;
  for (int j = 1; j < B.njobs; ++j)
    if ((int)blockIdx.x >= B.jobs[j].cta0) job = j;
  const WgradJob& P = B.jobs[job];
  const int64_t tile0 = (int)blockIdx.x - P.cta0, tile_step = P.ncta;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kWgOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kWgStages + s); };
  const uint32_t bar_done = bar0 + 8u * (2 * kWgStages);
  const uint32_t tmem_slot = bar0 + 8u * (2 * kWgStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = P.n_kb / 2;                  // blocks of 128 output features
  const int N = 64 * P.x_kb;                        // accumulator width
  const uint32_t stage_bytes = (uint32_t)(P.n_kb + P.x_kb) * kHalfImg;

  if (threadIdx.x == 0) {
    // a stage is released by the MMA commit and, when bias sums are wanted, by the four epilogue warps that read it
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), P.bias ? 5 : 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kWgOffBar + 8 * (2 * kWgStages + 1));

  int64_t my_tiles = 0;
  for (int64_t t = tile0; t < B.ntiles; t += tile_step) ++my_tiles;
  const int64_t nhalf = 2 * my_tiles;               // pipeline items: half tiles

  if (warp == 0) {
    if (lane == 0) {
      int64_t it = 0;
      for (int64_t t = tile0; t < B.ntiles; t += tile_step) {
        for (int h = 0; h < 2; ++h, ++it) {
          const uint32_t s = (uint32_t)(it % kWgStages);
          mbar_wait(bar_empty(s), (uint32_t)((it / kWgStages) & 1) ^ 1u);
          mbar_expect_tx(bar_full(s), stage_bytes);
          const uint32_t dst = base + s * kWgStageBytes;
          for (int kb = 0; kb < P.n_kb; ++kb)
            bulk_g2s(dst + kb * kHalfImg, P.dz + t * P.dz_tile_stride + (int64_t)(P.dz_slot0 + kb) * 16384 + h * kHalfImg,
                     kHalfImg, bar_full(s));
          for (int kb = 0; kb < P.x_kb; ++kb)
            bulk_g2s(dst + (4 + kb) * kHalfImg, P.x + t * P.x_tile_stride + (int64_t)(P.x_slot0 + kb) * 16384 + h * kHalfImg,
                     kHalfImg, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_mn(128, N);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    for (int64_t it = 0; it < nhalf; ++it) {
      const uint32_t s = (uint32_t)(it % kWgStages);
      mbar_wait(bar_full(s), (uint32_t)((it / kWgStages) & 1));
      tc_fence_after();
      const uint32_t st = base + s * kWgStageBytes;
      if (elect_one()) {
        for (int mb = 0; mb < n_blocks; ++mb) {
          const uint64_t a0 = umma_desc_mn_sw128(st + (uint32_t)(2 * mb) * kHalfImg, kHalfImg);
          const uint64_t b0 = umma_desc_mn_sw128(st + 4 * kHalfImg, kHalfImg);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)              // 16 samples = 2 atoms = 2048 B per K step
            umma_bf16(tmem_u + (uint32_t)mb * 256u, a0 + (uint64_t)(ks * (2048 >> 4)), b0 + (uint64_t)(ks * (2048 >> 4)), idesc,
                      (it == 0 && ks == 0) ? 0u : 1u);
        }
        umma_commit(bar_empty(s));
        if (it == nhalf - 1) umma_commit(bar_done);
      }
      __syncwarp();
    }
  } else if (nhalf > 0) {
    const uint32_t q = (uint32_t)warp & 3u;
    if (P.bias != nullptr) {
      // bias gradients while the MMAs stream: thread e sums the column pair (kb = e / 32, columns 2p, 2p+1) of
      // the dZ half-images of every stage; a warp reads one 128-byte image row per instruction (conflict-free)
      const int e = (int)threadIdx.x - 64, kb = e >> 5;
      const uint32_t p = (uint32_t)e & 31u;
      const bool active = kb < P.n_kb;
      float b0 = 0.0f, b1 = 0.0f;
      for (int64_t it = 0; it < nhalf; ++it) {
        const uint32_t s = (uint32_t)(it % kWgStages);
        mbar_wait(bar_full(s), (uint32_t)((it / kWgStages) & 1));
        if (active) {
          const uint8_t* img = base_ptr + s * kWgStageBytes + (uint32_t)kb * kHalfImg + (p & 3u) * 4u;
#pragma unroll 8
          for (uint32_t r = 0; r < 64; ++r) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(img + r * 128u + (((p >> 2) ^ (r & 7u)) << 4));
            b0 += __uint_as_float(w << 16);
            b1 += __uint_as_float(w & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
      }
      if (active) {
        atomicAdd(P.bias + kb * 64 + 2 * (int)p, b0);
        atomicAdd(P.bias + kb * 64 + 2 * (int)p + 1, b1);
      }
    }
    // epilogue: TMEM lane = output feature (row of dW), columns = input features
    const bool vec = P.ld_k == 1 && P.k0 == 0 && (P.ld_n & 3) == 0 && (P.n_valid & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(P.dw) & 15u) == 0;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int mb = 0; mb < n_blocks; ++mb) {
      const int n = mb * 128 + (int)(q * 32u) + lane;
      float* out = P.dw + (int64_t)n * P.ld_n;
      for (int c0 = 0; c0 < N && c0 < P.n_valid; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((q * 32u) << 16) + (uint32_t)mb * 256u + (uint32_t)c0, v);
        tmem_ld_wait();
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < P.n_valid)
              red_add_v4(out + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j >= P.k0 && c0 + j < P.n_valid) atomicAdd(out + (int64_t)(c0 + j - P.k0) * P.ld_k, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Assigns grid slices (cost = K-block images streamed per tile) and launches.
int launch_wgrad_tc(WgradBatch& B, cudaStream_t s) {
  if (B.ntiles == 0 || B.njobs == 0) return 0;
  static DeviceOnce once;
  if (cudaError_t e = opt_in_smem_once(once, k_wgrad_tc, kWgSmem)) return set_error((int)e, "wgrad_tc attr: %s", cudaGetErrorString(e));
  const int sms = num_sms();
  if (B.njobs > kWgMaxJobs || B.njobs > sms) return set_error(FNERF_ERR_ARG, "wgrad_tc: too many jobs");
  int total_cost = 0, used = 0;
  for (int j = 0; j < B.njobs; ++j) total_cost += B.jobs[j].n_kb + B.jobs[j].x_kb;
  for (int j = 0; j < B.njobs; ++j) {
    int n = (B.jobs[j].n_kb + B.jobs[j].x_kb) * sms / total_cost;
    B.jobs[j].ncta = n < 1 ? 1 : n;
    used += B.jobs[j].ncta;
  }
  for (int cost = 8; used < sms && cost > 0; --cost)          // leftover CTAs go to the heaviest jobs first
    for (int j = 0; j < B.njobs && used < sms; ++j)
      if (B.jobs[j].n_kb + B.jobs[j].x_kb == cost) { ++B.jobs[j].ncta; ++used; }
  while (used > sms)                                          // (only if rounding up the tiny jobs overshot)
    for (int j = 0; j < B.njobs && used > sms; ++j)
      if (B.jobs[j].ncta > 1) { --B.jobs[j].ncta; --used; }
  int cta = 0;
  for (int j = 0; j < B.njobs; ++j) {
    if ((int64_t)B.jobs[j].ncta > B.ntiles) B.jobs[j].ncta = (int)B.ntiles;
    B.jobs[j].cta0 = cta;
    cta += B.jobs[j].ncta;
  }
  k_wgrad_tc<<<(unsigned)cta, kWgThreads, kWgSmem, s>>>(B);
  return check_launch("wgrad_tc");
}

int launch_mlp_dgrad_tc(const void* packed, int cond, const float* g_raw, const uint32_t* mask_tape, uint8_t* bwd_tape,
                        float* flat_grad, int64_t M, cudaStream_t s);
int launch_mlp_tc_tape(const MlpArgs& a, uint8_t* tape, uint32_t* mask_tape, cudaStream_t s);

// forward tape of one network query: K-block images of every tile, then the ReLU bitmask words of every tile
int64_t mlp_tape_bytes(int64_t M) {
  const int64_t ntiles = (M + 127) / 128;
  return ntiles * ((int64_t)kTapeFwdSlots * 16384 + kMaskTileBytes);
}
int64_t mlp_bwd_pipe_workspace_bytes();
bool mlp_bwd_pipe_supported();
int launch_mlp_bwd_pipe(const void* packed, const float* g_raw, const void* tape, float* flat_grad, void* ws, int64_t M, cudaStream_t s);

// Which backward runs from a forward tape: the layer-pipelined fused dgrad + wgrad kernel (mlp_bwd_pipe.cu) for
// unconditioned networks; the tile-major dgrad chain + grouped wgrad launch below for conditioned ones (their extra
// code-block product needs a 256 x 256 accumulator no pipeline role has room for).  FNERF_BWD_PIPE=0 forces the latter.
static bool use_bwd_pipe(int cond) {
  static std::once_flag flag;
  static int enabled = 1;
  std::call_once(flag, [] { const char* e = getenv("FNERF_BWD_PIPE"); if (e) enabled = atoi(e) != 0; });
  return enabled && !cond && mlp_bwd_pipe_supported();
}

// backward tape + (conditioned networks) the per-sample garment codes as 4 K-block images per tile; or the pipeline's
// hand-off rings
int64_t mlp_bwd_from_tape_workspace_bytes(int64_t M) {
  const int64_t tile_major = (M + 127) / 128 * (int64_t)(kTapeBwdSlots + 4) * 16384, pipe = mlp_bwd_pipe_workspace_bytes();
  return tile_major > pipe ? tile_major : pipe;
}

// A.8 backward: the code block of W5 sees, for every sample, the code of its ray.  Written once per backward as
// bf16 K-block images (same layout as the activation tape) so that dW5[:, 63:319] is one more wgrad product.
__global__ void __launch_bounds__(256) k_tape_codes(const float* __restrict__ cond_rows, const int32_t* __restrict__ cond_index,
                                                    int64_t C, int64_t M, int S, uint8_t* __restrict__ code_tape) {
  const int64_t tile = blockIdx.x;
  uint8_t* timg = code_tape + (size_t)tile * 4 * 16384;
  for (int e = threadIdx.x; e < 128 * 32; e += blockDim.x) {       // (row, 16-byte chunk of the 256 columns)
    const uint32_t r = (uint32_t)e >> 5, ch = (uint32_t)e & 31u;    // ch: K-block ch / 8, chunk ch % 8
    int64_t g = tile * 128 + r;
    if (g >= M) g = M - 1;
    const int64_t ray = g / S;
    const int64_t crow = cond_row(cond_index, C, ray);
    const float4* src = reinterpret_cast<const float4*>(cond_rows + crow * kCond + ch * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    *reinterpret_cast<uint4*>(timg + (size_t)(ch >> 3) * 16384 + r * 128u + (((ch & 7u) ^ (r & 7u)) << 4)) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}
// recomputing backward: forward tape | backward tape | raw scratch, for one chunk of whole rays (the tapes cost ~11 KB
// per sample, so the recompute runs over chunks of at most kBwdTcChunk samples and the workspace stays bounded)
constexpr int64_t kBwdTcChunk = 1 << 18;
static int64_t bwd_tc_chunk_rays(int64_t R, int64_t S) {
  int64_t rays = kBwdTcChunk / S;
  if (rays < 1) rays = 1;
  return rays < R ? rays : R;
}
static int64_t bwd_tc_chunk_bytes(int64_t M) {
  return (mlp_tape_bytes(M) + 1023) / 1024 * 1024 + mlp_bwd_from_tape_workspace_bytes(M) + (M + 127) / 128 * 128 * 16 + 4096;
}
int64_t mlp_bwd_tc_workspace_bytes(int64_t R, int64_t S) { return bwd_tc_chunk_bytes(bwd_tc_chunk_rays(R, S) * S); }

int launch_mlp_fwd_tape(const MlpArgs& a, void* tape, cudaStream_t s) {
  const int64_t ntiles = (a.R * a.S + 127) / 128;
  uint8_t* t = reinterpret_cast<uint8_t*>(tape);
  return launch_mlp_tc_tape(a, t, reinterpret_cast<uint32_t*>(t + ntiles * (int64_t)kTapeFwdSlots * 16384), s);
}

// bf16 tensor-core backward of one network query from its forward tape: dgrad chain, then every
// weight-gradient product (which also sums the bias gradients) and the two head products in one grouped
// launch.  flat_grad += dL/dparams.  Conditioned networks (A.8) add the product with the per-sample codes.
int launch_mlp_bwd_from_tape(const void* packed, int cond, const float* g_raw, const void* tape, const float* cond_rows,
                             const int32_t* cond_index, int64_t C, int64_t S, float* flat_grad, void* ws, int64_t M,
                             cudaStream_t s) {
  if (M == 0) return 0;
  if (use_bwd_pipe(cond)) return launch_mlp_bwd_pipe(packed, g_raw, tape, flat_grad, ws, M, s);
  const int64_t ntiles = (M + 127) / 128;
  const uint8_t* fwd_tape = reinterpret_cast<const uint8_t*>(tape);
  const uint32_t* mask_tape = reinterpret_cast<const uint32_t*>(fwd_tape + ntiles * (int64_t)kTapeFwdSlots * 16384);
  uint8_t* bwd_tape = reinterpret_cast<uint8_t*>(ws);
  uint8_t* code_tape = bwd_tape + ntiles * (int64_t)kTapeBwdSlots * 16384;
  int rc;
  if (cond) {
    k_tape_codes<<<(unsigned)ntiles, 256, 0, s>>>(cond_rows, cond_index, C, M, (int)S, code_tape);
    if ((rc = check_launch("tape_codes"))) return rc;
  }
  if ((rc = launch_mlp_dgrad_tc(packed, cond, g_raw, mask_tape, bwd_tape, flat_grad, M, s))) return rc;

  const int64_t fstride = (int64_t)kTapeFwdSlots * 16384, bstride = (int64_t)kTapeBwdSlots * 16384;
  WgradBatch B;
  B.njobs = 0; B.ntiles = ntiles;
  auto wg = [&](const uint8_t* dz, int64_t dz_stride, int dz_slot, int n_kb, const uint8_t* x, int64_t x_stride, int x_slot,
                int x_kb, float* dw, int64_t ld_n, int64_t ld_k, int k0, int n_valid, float* bias) {
    WgradJob& P = B.jobs[B.njobs++];
    P.dz = dz; P.dz_tile_stride = dz_stride; P.dz_slot0 = dz_slot; P.n_kb = n_kb;
    P.x = x; P.x_tile_stride = x_stride; P.x_slot0 = x_slot; P.x_kb = x_kb;
    P.dw = dw; P.ld_n = ld_n; P.ld_k = ld_k; P.k0 = k0; P.n_valid = n_valid; P.bias = bias; P.cta0 = 0; P.ncta = 1;
    return 0;
  };
  // dZ (backward tape) x forward activation
  auto wz = [&](int dz_slot, int n_kb, int x_slot, int x_kb, float* dw, int64_t ld, int n_valid, float* bias) {
    return wg(bwd_tape, bstride, dz_slot, n_kb, fwd_tape, fstride, x_slot, x_kb, dw, ld, 1, 0, n_valid, bias);
  };
  auto gw = [&](int l) { return flat_grad + flat_weight_offset(l, cond); };
  auto gb = [&](int l) { return flat_grad + flat_bias_offset(l, cond); };
  auto zslot = [&](int l) { return kTapeBwdSlotZ + 4 * (7 - l); };
  const int in5 = kPE + (cond ? kCond : 0) + kW;
  if ((rc = wz(zslot(0), 4, kTapeSlotPe, 1, gw(0), kPE, kPE, gb(0)))) return rc;
  for (int l = 1; l <= 7; ++l) {
    if (l == 5) {
      if ((rc = wz(zslot(5), 4, kTapeSlotPe, 1, gw(5), in5, kPE, nullptr))) return rc;
      if (cond && (rc = wg(bwd_tape, bstride, zslot(5), 4, code_tape, 4 * 16384, 0, 4, gw(5) + kPE, in5, 1, 0, kCond, nullptr))) return rc;
      if ((rc = wz(zslot(5), 4, kTapeSlotH + 4 * 4, 4, gw(5) + (in5 - kW), in5, kW, gb(5)))) return rc;
    } else {
      if ((rc = wz(zslot(l), 4, kTapeSlotH + 4 * (l - 1), 4, gw(l), kW, kW, gb(l)))) return rc;
    }
  }
  if ((rc = wz(kTapeBwdSlotFeat, 4, kTapeSlotH + 28, 4, gw(9), kW, kW, gb(9)))) return rc;
  if ((rc = wz(kTapeBwdSlotZv, 2, kTapeSlotFeat, 4, gw(10), kW + kPED, kW, gb(10)))) return rc;
  if ((rc = wz(kTapeBwdSlotZv, 2, kTapeSlotPed, 1, gw(10) + kW, kW + kPED, kPED, nullptr))) return rc;
  // heads: the "dZ" operand is a forward activation, the "X" operand the g_raw image (columns rgb, sigma)
  //   rgb_linear.weight[c][n]  += sum_m HV[m][n] g_rgb[m][c]      alpha_linear.weight[0][n] += sum_m H7[m][n] g_sigma[m]
  if ((rc = wg(fwd_tape, fstride, kTapeSlotHv, 2, bwd_tape, bstride, kTapeBwdSlotG, 1, gw(11), 1, kWV, 0, 3, nullptr))) return rc;
  if ((rc = wg(fwd_tape, fstride, kTapeSlotH + 28, 4, bwd_tape, bstride, kTapeBwdSlotG, 1, gw(8), 1, 0, 3, 4, nullptr))) return rc;
  return launch_wgrad_tc(B, s);
}

// Recomputing variant (unconditioned networks): per chunk of rays, forward with tape into the workspace, then the
// backward above; flat_grad accumulates over the chunks.
int launch_mlp_bwd_tc(const MlpArgs& a, const float* g_raw, float* flat_grad, void* ws, int64_t ws_bytes, cudaStream_t s) {
  if (a.R * a.S == 0) return 0;
  if (ws_bytes < mlp_bwd_tc_workspace_bytes(a.R, a.S)) return set_error(FNERF_ERR_WORKSPACE, "mlp_bwd_tc: workspace too small");
  const int64_t chunk = bwd_tc_chunk_rays(a.R, a.S);
  for (int64_t r0 = 0; r0 < a.R; r0 += chunk) {
    const int64_t rays = a.R - r0 < chunk ? a.R - r0 : chunk, M = rays * a.S;
    uint8_t* tape = reinterpret_cast<uint8_t*>(ws);
    uint8_t* bwd_ws = tape + (mlp_tape_bytes(M) + 1023) / 1024 * 1024;
    MlpArgs fa = a;
    fa.rays_o = a.rays_o + 3 * r0; fa.rays_d = a.rays_d + 3 * r0; fa.viewdirs = a.viewdirs + 3 * r0; fa.z = a.z + r0 * a.S;
    fa.R = rays;
    fa.raw = reinterpret_cast<float*>(bwd_ws + mlp_bwd_from_tape_workspace_bytes(M));
    int rc;
    if ((rc = launch_mlp_fwd_tape(fa, tape, s))) return rc;
    if ((rc = launch_mlp_bwd_from_tape(a.packed, 0, g_raw + r0 * a.S * 4, tape, nullptr, nullptr, 0, a.S, flat_grad, bwd_ws, M, s))) return rc;
  }
  return 0;
}

}  // namespace fnerf

// ---- debug entry (tests only): dw[n_kb*64, ld] += dZ^T X from two image buffers ---------------------
extern "C" int fnerf_debug_wgrad_tc(const void* dz_img, int n_kb, const void* x_img, int x_kb, float* dw, int64_t ld,
                                    int n_valid, int64_t ntiles, fnerf_stream_t stream) {
  fnerf::WgradBatch B;
  B.njobs = 1; B.ntiles = ntiles;
  fnerf::WgradJob& P = B.jobs[0];
  P.dz = reinterpret_cast<const uint8_t*>(dz_img); P.dz_tile_stride = (int64_t)n_kb * 16384; P.dz_slot0 = 0; P.n_kb = n_kb;
  P.x = reinterpret_cast<const uint8_t*>(x_img); P.x_tile_stride = (int64_t)x_kb * 16384; P.x_slot0 = 0; P.x_kb = x_kb;
  P.dw = dw; P.ld_n = ld; P.ld_k = 1; P.k0 = 0; P.n_valid = n_valid; P.bias = nullptr; P.cta0 = 0; P.ncta = 1;
  return fnerf::launch_wgrad_tc(B, (cudaStream_t)stream);
}
