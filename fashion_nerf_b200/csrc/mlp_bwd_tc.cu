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

struct WgradParams {
  const uint8_t* dz; int64_t dz_tile_stride; int dz_slot0; int n_kb;   // dZ images: tile t, K-block kb at dz + t*stride + (slot0+kb)*16K
  const uint8_t* x;  int64_t x_tile_stride;  int x_slot0;  int x_kb;   // X images
  float* dw; int64_t ld;                                               // dW[n][k] at dw[n*ld + k]
  int n_valid;                                                         // k < n_valid columns are written (63 / 27 / 64*x_kb)
  int64_t ntiles;
};

// MN-major shared-memory matrix descriptor, 128-byte swizzle: 64 contiguous MN elements per 128-byte row,
// rows = K; LBO = byte distance between 64-element MN groups, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {   // A and B both MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const WgradParams P) {
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
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kWgOffBar + 8 * (2 * kWgStages + 1));

  int64_t my_tiles = 0;
  for (int64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) ++my_tiles;
  const int64_t nhalf = 2 * my_tiles;               // pipeline items: half tiles

  if (warp == 0) {
    if (lane == 0) {
      int64_t it = 0;
      for (int64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
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
    // epilogue: TMEM lane = output feature (row of dW), columns = input features
    const uint32_t q = (uint32_t)warp & 3u;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int mb = 0; mb < n_blocks; ++mb) {
      const int n = mb * 128 + (int)(q * 32u) + lane;
      float* out = P.dw + (int64_t)n * P.ld;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((q * 32u) << 16) + (uint32_t)mb * 256u + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < P.n_valid) atomicAdd(out + c0 + j, __uint_as_float(v[j]));
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

int launch_wgrad_tc(const WgradParams& P, cudaStream_t s) {
  if (P.ntiles == 0) return 0;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    if (e != cudaSuccess) return set_error((int)e, "wgrad_tc attr: %s", cudaGetErrorString(e));
    attr_done[dev] = true;
  }
  int64_t blocks = num_sms();
  if (blocks > P.ntiles) blocks = P.ntiles;
  k_wgrad_tc<<<(unsigned)blocks, kWgThreads, kWgSmem, s>>>(P);
  return check_launch("wgrad_tc");
}

// ---- bias and head gradients straight from the tape ---------------------------------------------------
// One CTA per tile, 256 threads = columns.  Reads every dZ image once (HBM-bound) and adds the column
// sums into the bias gradients; the two tiny heads (sigma: 256->1 on H7, rgb: 128->3 on HV) are fp32 dot
// products against g_raw.
__device__ __forceinline__ float tape_elem(const uint8_t* img, int r, int col) {   // image of 128 x 64, column col < 64
  const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)col & 7u) * 2u;
  const unsigned short b = *reinterpret_cast<const unsigned short*>(img + off);
  return __uint_as_float((uint32_t)b << 16);
}

__global__ void __launch_bounds__(256) k_bias_heads_from_tape(const uint8_t* __restrict__ fwd_tape, const uint8_t* __restrict__ bwd_tape,
                                                              const float4* __restrict__ g_raw, float* __restrict__ flat_grad,
                                                              int cond, int64_t M, int64_t ntiles) {
  __shared__ float4 s_g[128];
  const int n = threadIdx.x;
  float bsum[10];                                     // dZ0..dZ7, dFEAT, dZv
#pragma unroll
  for (int i = 0; i < 10; ++i) bsum[i] = 0.0f;
  float wa = 0.0f, wr0 = 0.0f, wr1 = 0.0f, wr2 = 0.0f, gs = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (n < 128) {
      const int64_t g = tile * 128 + n;
      s_g[n] = g < M ? g_raw[g] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const uint8_t* ft = fwd_tape + (size_t)tile * kTapeFwdSlots * 16384;
    const uint8_t* bt = bwd_tape + (size_t)tile * kTapeBwdSlots * 16384;
    const int kb = n >> 6, col = n & 63;
#pragma unroll 1
    for (int l = 0; l < 8; ++l) {
      const uint8_t* img = bt + (size_t)(kTapeBwdSlotZ + 4 * (7 - l) + kb) * 16384;
      float a = 0.0f;
      for (int r = 0; r < 128; ++r) a += tape_elem(img, r, col);
      bsum[l] += a;
    }
    {
      const uint8_t* img = bt + (size_t)(kTapeBwdSlotFeat + kb) * 16384;
      const uint8_t* h7 = ft + (size_t)(kTapeSlotH + 28 + kb) * 16384;
      float a = 0.0f, w = 0.0f;
      for (int r = 0; r < 128; ++r) { a += tape_elem(img, r, col); w = fmaf(s_g[r].w, tape_elem(h7, r, col), w); }
      bsum[8] += a; wa += w;
    }
    if (n < 128) {
      const uint8_t* img = bt + (size_t)(kTapeBwdSlotZv + kb) * 16384;
      const uint8_t* hv = ft + (size_t)(kTapeSlotHv + kb) * 16384;
      float a = 0.0f;
      for (int r = 0; r < 128; ++r) {
        a += tape_elem(img, r, col);
        const float h = tape_elem(hv, r, col);
        const float4 gq = s_g[r];
        wr0 = fmaf(gq.x, h, wr0); wr1 = fmaf(gq.y, h, wr1); wr2 = fmaf(gq.z, h, wr2);
      }
      bsum[9] += a;
    }
    if (n == 0) for (int r = 0; r < 128; ++r) { gs += s_g[r].w; g0 += s_g[r].x; g1 += s_g[r].y; g2 += s_g[r].z; }
  }
  for (int l = 0; l < 8; ++l) atomicAdd(flat_grad + flat_bias_offset(l, cond) + n, bsum[l]);
  atomicAdd(flat_grad + flat_bias_offset(9, cond) + n, bsum[8]);
  atomicAdd(flat_grad + flat_weight_offset(8, cond) + n, wa);
  if (n < 128) {
    atomicAdd(flat_grad + flat_bias_offset(10, cond) + n, bsum[9]);
    atomicAdd(flat_grad + flat_weight_offset(11, cond) + n, wr0);
    atomicAdd(flat_grad + flat_weight_offset(11, cond) + kWV + n, wr1);
    atomicAdd(flat_grad + flat_weight_offset(11, cond) + 2 * kWV + n, wr2);
  }
  if (n == 0) {
    atomicAdd(flat_grad + flat_bias_offset(8, cond), gs);
    atomicAdd(flat_grad + flat_bias_offset(11, cond), g0);
    atomicAdd(flat_grad + flat_bias_offset(11, cond) + 1, g1);
    atomicAdd(flat_grad + flat_bias_offset(11, cond) + 2, g2);
  }
}

int launch_mlp_dgrad_tc(const void* packed, int cond, const float* g_raw, const uint8_t* fwd_tape, uint8_t* bwd_tape,
                        int64_t M, cudaStream_t s);

int64_t mlp_bwd_tc_workspace_bytes(int64_t M) {
  const int64_t ntiles = (M + 127) / 128;
  return ntiles * (int64_t)(kTapeFwdSlots + kTapeBwdSlots) * 16384 + ntiles * 128 * 16 + 4096;
}

// Full bf16 tensor-core backward of one network query: forward with tape, dgrad chain, wgrad GEMMs, bias /
// head reductions.  flat_grad += dL/dparams.  (Unconditioned networks; the conditioned variant uses the
// fp32 path.)
int launch_mlp_bwd_tc(const MlpArgs& a, const float* g_raw, float* flat_grad, void* ws, int64_t ws_bytes, cudaStream_t s) {
  const int64_t M = a.R * a.S;
  if (M == 0) return 0;
  const int64_t ntiles = (M + 127) / 128;
  if (ws_bytes < mlp_bwd_tc_workspace_bytes(M)) return set_error(FNERF_ERR_WORKSPACE, "mlp_bwd_tc: workspace too small");
  uint8_t* fwd_tape = reinterpret_cast<uint8_t*>(ws);
  uint8_t* bwd_tape = fwd_tape + ntiles * (int64_t)kTapeFwdSlots * 16384;
  float* raw_scratch = reinterpret_cast<float*>(bwd_tape + ntiles * (int64_t)kTapeBwdSlots * 16384);
  const int cond = 0;
  MlpArgs fa = a;
  fa.raw = raw_scratch;
  int rc;
  if ((rc = launch_mlp_tc_save(fa, fwd_tape, s))) return rc;
  if ((rc = launch_mlp_dgrad_tc(a.packed, cond, g_raw, fwd_tape, bwd_tape, M, s))) return rc;

  const int64_t fstride = (int64_t)kTapeFwdSlots * 16384, bstride = (int64_t)kTapeBwdSlots * 16384;
  auto wg = [&](int dz_slot, int n_kb, int x_slot, int x_kb, float* dw, int64_t ld, int n_valid) {
    WgradParams P;
    P.dz = bwd_tape; P.dz_tile_stride = bstride; P.dz_slot0 = dz_slot; P.n_kb = n_kb;
    P.x = fwd_tape; P.x_tile_stride = fstride; P.x_slot0 = x_slot; P.x_kb = x_kb;
    P.dw = dw; P.ld = ld; P.n_valid = n_valid; P.ntiles = ntiles;
    return launch_wgrad_tc(P, s);
  };
  auto gw = [&](int l) { return flat_grad + flat_weight_offset(l, cond); };
  auto zslot = [&](int l) { return kTapeBwdSlotZ + 4 * (7 - l); };
  const int in5 = kPE + kW;
  if ((rc = wg(zslot(0), 4, kTapeSlotPe, 1, gw(0), kPE, kPE))) return rc;
  for (int l = 1; l <= 7; ++l) {
    if (l == 5) {
      if ((rc = wg(zslot(5), 4, kTapeSlotPe, 1, gw(5), in5, kPE))) return rc;
      if ((rc = wg(zslot(5), 4, kTapeSlotH + 4 * 4, 4, gw(5) + kPE, in5, kW))) return rc;
    } else {
      if ((rc = wg(zslot(l), 4, kTapeSlotH + 4 * (l - 1), 4, gw(l), kW, kW))) return rc;
    }
  }
  if ((rc = wg(kTapeBwdSlotFeat, 4, kTapeSlotH + 28, 4, gw(9), kW, kW))) return rc;
  if ((rc = wg(kTapeBwdSlotZv, 2, kTapeSlotFeat, 4, gw(10), kW + kPED, kW))) return rc;
  if ((rc = wg(kTapeBwdSlotZv, 2, kTapeSlotPed, 1, gw(10) + kW, kW + kPED, kPED))) return rc;
  int64_t blocks = ntiles < 4 * (int64_t)num_sms() ? ntiles : 4 * (int64_t)num_sms();
  k_bias_heads_from_tape<<<(unsigned)blocks, 256, 0, s>>>(fwd_tape, bwd_tape, reinterpret_cast<const float4*>(g_raw), flat_grad,
                                                         cond, M, ntiles);
  return check_launch("mlp_bwd_tc");
}

}  // namespace fnerf

// ---- debug entry (tests only): dw[n_kb*64, ld] += dZ^T X from two image buffers ---------------------
extern "C" int fnerf_debug_wgrad_tc(const void* dz_img, int n_kb, const void* x_img, int x_kb, float* dw, int64_t ld,
                                    int n_valid, int64_t ntiles, void* stream) {
  fnerf::WgradParams P;
  P.dz = reinterpret_cast<const uint8_t*>(dz_img); P.dz_tile_stride = (int64_t)n_kb * 16384; P.dz_slot0 = 0; P.n_kb = n_kb;
  P.x = reinterpret_cast<const uint8_t*>(x_img); P.x_tile_stride = (int64_t)x_kb * 16384; P.x_slot0 = 0; P.x_kb = x_kb;
  P.dw = dw; P.ld = ld; P.n_valid = n_valid; P.ntiles = ntiles;
  return fnerf::launch_wgrad_tc(P, (cudaStream_t)stream);
}
