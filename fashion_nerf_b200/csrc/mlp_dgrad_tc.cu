// bf16 tensor-core data-gradient chain of the NeRF MLP (A.4 backward, dgrad) for sm_100a.
//
// Walks the network top-down for a 128-sample tile, exactly like the forward kernel walks it bottom-up:
//   dZv   = (g_rgb . W_rgb) (.) [HV > 0]                              CUDA cores (K = 3)
//   dFEAT = dZv . Wv[:, :256]                                         tcgen05, K = 128
//   dZ7   = (dFEAT . W_f + g_sigma (x) w_alpha) (.) [H7 > 0]          tcgen05, K = 256
//   dZ_{l-1} = (dZ_l . W_l[:, trunk cols]) (.) [H_{l-1} > 0], l = 7..1
// A = the current gradient tile in shared memory (K-major over the layer's OUTPUT features, 128B swizzle),
// B = chunks of the TRANSPOSED weights (packed section E, [256 input features x 64 output features]),
// D = one of two 128x256 fp32 accumulators in TMEM.  The ReLU masks come from the forward tape (the saved
// post-activation images), every dZ is written to the backward tape as a K-block image for the wgrad
// kernel (mlp_bwd_tc.cu), and to shared memory as the next step's A operand.  Gradients are rounded to
// bf16 between layers (bf16 in / fp32 accumulate); inputs (encodings, codes) get no gradient (A.4).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kDgTileM = 128;
constexpr int kDgStages = 3;
constexpr int kDgThreads = 576;
constexpr int kDgWorkerThreads = 512;
constexpr uint32_t kDgKB = 16384;                                  // one K-block image
constexpr uint32_t kDgOffAct = 0;                                  // 4 K-blocks: current gradient tile
constexpr uint32_t kDgOffW = 4 * kDgKB;                            // weight stages, 32 KB each
constexpr uint32_t kDgOffHeads = kDgOffW + kDgStages * 32768;      // fp32 w_alpha[256] + pad | w_rgb[3][128]
constexpr int kDgHeadFloats = kAuxFloats - kAuxWAlpha;
constexpr uint32_t kDgOffBar = kDgOffHeads + kDgHeadFloats * 4;
constexpr uint32_t kDgNumBars = 2 * kDgStages + 4 + 2;
constexpr uint32_t kDgSmem = kDgOffBar + kDgNumBars * 8 + 16 + 1024;
static_assert(kDgOffBar % 8 == 0, "barrier alignment");
constexpr int kDgSteps = 9;                                         // MMA steps per tile

struct DgradParams {
  const uint8_t* packed;        // blob: section E (transposed chunks) + aux
  int cond;
  const float4* g_raw;          // [M] (rgb_raw grads, sigma_raw grad)
  const uint8_t* fwd_tape;      // kTapeFwdSlots images per tile
  uint8_t* bwd_tape;            // kTapeBwdSlots images per tile
  int64_t M; int64_t ntiles;
};

__device__ __forceinline__ void dg_worker_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kDgWorkerThreads) : "memory"); }

__global__ void __launch_bounds__(kDgThreads, 1) k_mlp_dgrad_tc(const DgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kDgOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kDgStages + s); };
  auto bar_act = [&](int kb) { return bar0 + 8u * (2 * kDgStages + kb); };
  auto bar_acc = [&](int a) { return bar0 + 8u * (2 * kDgStages + 4 + a); };
  const uint32_t tmem_slot = bar0 + 8u * kDgNumBars;
  float* heads_s = reinterpret_cast<float*>(base_ptr + kDgOffHeads);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    const float* aux_g = reinterpret_cast<const float*>(P.packed + kSecBOffset) + kAuxWAlpha;
    for (int i = threadIdx.x; i < kDgHeadFloats; i += kDgThreads) heads_s[i] = aux_g[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar_act(kb), 8);     // 2 worker groups x 4 warps per K-block
    mbar_init(bar_acc(0), 1);
    mbar_init(bar_acc(1), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kDgOffBar + 8 * kDgNumBars);
  const uint8_t* wsrc = P.packed + sec_e_offset(P.cond);

  if (warp == 0) {
    // ================================ weight producer ==============================================
    if (lane == 0) {
      uint32_t wc = 0;
      for (int64_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        for (int c = 0; c < kNumChunksT; ++c, ++wc) {
          const uint32_t s = wc % kDgStages;
          mbar_wait(bar_empty(s), ((wc / kDgStages) & 1u) ^ 1u);
          mbar_expect_tx(bar_full(s), 32768);
          bulk_g2s(base + kDgOffW + s * 32768, wsrc + (size_t)c * 32768, 32768, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform, elected lane issues) ===============
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_act = umma_desc_sw128(base + kDgOffAct);
    const uint64_t desc_w = umma_desc_sw128(base + kDgOffW);
    uint32_t wc = 0;
    uint32_t act_phase[4] = {0u, 0u, 0u, 0u};
    for (int64_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
#pragma unroll 1
      for (int t = 0; t < kDgSteps; ++t) {
        const int nkb = t == 0 ? 2 : 4;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, ++wc) {
          const uint32_t s = wc % kDgStages;
          mbar_wait(bar_act(kb), act_phase[kb] & 1u);
          ++act_phase[kb];
          mbar_wait(bar_full(s), (wc / kDgStages) & 1u);
          tc_fence_after();
          const uint64_t a_desc = desc_act + (uint64_t)((uint32_t)kb * (kDgKB >> 4));
          const uint64_t b_desc = desc_w + (uint64_t)(s * (32768u >> 4));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_u + (uint32_t)(t & 1) * 256u, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                        (kb == 0 && ks == 0) ? 0u : 1u);
            umma_commit(bar_empty(s));
            if (kb == nkb - 1) umma_commit(bar_acc(t & 1));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ workers ======================================================
    const uint32_t q = (uint32_t)warp & 3u;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;
    const uint32_t row = q * 32u + (uint32_t)lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    const uint32_t act_row = base + kDgOffAct + row * 128u;
    const float* walpha_s = heads_s;
    const float* wrgb_s = heads_s + (kAuxWRgb - kAuxWAlpha);
    uint32_t acc_cnt[2] = {0u, 0u};
    for (int64_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int64_t g = tile * kDgTileM + row;
      const float4 gr = g < P.M ? __ldg(P.g_raw + g) : make_float4(0.f, 0.f, 0.f, 0.f);   // padding rows carry no gradient
      const uint8_t* ftape = P.fwd_tape + (size_t)tile * kTapeFwdSlots * kDgKB + row * 128u;
      uint8_t* btape = P.bwd_tape + (size_t)tile * kTapeBwdSlots * kDgKB + row * 128u;
      // ---- step b0: dZv = (g_rgb . W_rgb) (.) [HV > 0], 32 of the 128 columns per group -------------
      {
        const uint32_t kb = grp >> 1, ch0 = (grp & 1u) * 4u;
        const uint8_t* hv = ftape + (size_t)(kTapeSlotHv + (int)kb) * kDgKB;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t c16 = ch0 + (uint32_t)c;
          const uint4 mk = __ldg(reinterpret_cast<const uint4*>(hv + ((c16 ^ (row & 7u)) << 4)));
          const uint32_t m[4] = {mk.x, mk.y, mk.z, mk.w};
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = (int)grp * 32 + c * 8 + 2 * j;
            float lo = gr.x * wrgb_s[col] + gr.y * wrgb_s[kWV + col] + gr.z * wrgb_s[2 * kWV + col];
            float hi = gr.x * wrgb_s[col + 1] + gr.y * wrgb_s[kWV + col + 1] + gr.z * wrgb_s[2 * kWV + col + 1];
            if ((m[j] & 0xFFFFu) == 0u) lo = 0.0f;
            if ((m[j] >> 16) == 0u) hi = 0.0f;
            pk[j] = pack_bf16(lo, hi);
          }
          st_shared_v4(act_row + kb * kDgKB + ((c16 ^ (row & 7u)) << 4), pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(btape + (size_t)(kTapeBwdSlotZv + (int)kb) * kDgKB + ((c16 ^ (row & 7u)) << 4)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_act(kb));
      }
      // ---- MMA steps: t = 0 views -> dFEAT, t = 1 feature -> dZ7, t = 2..8 layers 7..1 -> dZ6..dZ0 -----
#pragma unroll 1
      for (int t = 0; t < kDgSteps; ++t) {
        const int a = t & 1;
        const int out_slot = t == 0 ? kTapeBwdSlotFeat : kTapeBwdSlotZ + 4 * (t - 1);
        const int mask_slot = kTapeSlotH + 4 * (8 - t);            // H_{8-t}: only used for t >= 1
        // masks of both rounds, fetched while the MMAs run
        uint4 mk[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t unit = grp + 4u * (uint32_t)r;
          const uint32_t kb = unit >> 1, ch0 = (unit & 1u) * 4u;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t c16 = ch0 + (uint32_t)c;
            mk[r][c] = t == 0 ? make_uint4(~0u, ~0u, ~0u, ~0u)
                              : __ldg(reinterpret_cast<const uint4*>(ftape + (size_t)(mask_slot + (int)kb) * kDgKB + ((c16 ^ (row & 7u)) << 4)));
          }
        }
        mbar_wait(bar_acc(a), acc_cnt[a] & 1u);
        ++acc_cnt[a];
        tc_fence_after();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t unit = grp + 4u * (uint32_t)r;
          const uint32_t kb = unit >> 1, ch0 = (unit & 1u) * 4u, col0 = unit * 32u;
          uint32_t v[32];
          tmem_ld32(tmem_row + (uint32_t)a * 256u + col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t m[4] = {mk[r][c].x, mk[r][c].y, mk[r][c].z, mk[r][c].w};
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float lo = __uint_as_float(v[c * 8 + 2 * j]), hi = __uint_as_float(v[c * 8 + 2 * j + 1]);
              if (t == 1) {                                          // sigma head: + g_sigma * w_alpha[k]
                lo = fmaf(gr.w, walpha_s[col0 + c * 8 + 2 * j], lo);
                hi = fmaf(gr.w, walpha_s[col0 + c * 8 + 2 * j + 1], hi);
              }
              if ((m[j] & 0xFFFFu) == 0u) lo = 0.0f;
              if ((m[j] >> 16) == 0u) hi = 0.0f;
              pk[j] = pack_bf16(lo, hi);
            }
            const uint32_t c16 = ch0 + (uint32_t)c;
            if (t < kDgSteps - 1) st_shared_v4(act_row + kb * kDgKB + ((c16 ^ (row & 7u)) << 4), pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(btape + (size_t)(out_slot + (int)kb) * kDgKB + ((c16 ^ (row & 7u)) << 4)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (t < kDgSteps - 1 && lane == 0) mbar_arrive(bar_act(kb));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_mlp_dgrad_tc(const void* packed, int cond, const float* g_raw, const uint8_t* fwd_tape, uint8_t* bwd_tape,
                        int64_t M, cudaStream_t s) {
  if (M == 0) return 0;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_dgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDgSmem);
    if (e != cudaSuccess) return set_error((int)e, "mlp_dgrad_tc attr: %s", cudaGetErrorString(e));
    attr_done[dev] = true;
  }
  DgradParams P;
  P.packed = reinterpret_cast<const uint8_t*>(packed); P.cond = cond;
  P.g_raw = reinterpret_cast<const float4*>(g_raw);
  P.fwd_tape = fwd_tape; P.bwd_tape = bwd_tape;
  P.M = M; P.ntiles = (M + kDgTileM - 1) / kDgTileM;
  int64_t blocks = num_sms();
  if (blocks > P.ntiles) blocks = P.ntiles;
  k_mlp_dgrad_tc<<<(unsigned)blocks, kDgThreads, kDgSmem, s>>>(P);
  return check_launch("mlp_dgrad_tc");
}

}  // namespace fnerf
