// bf16 tensor-core data-gradient chain of the NeRF MLP (A.4 backward, dgrad) for sm_100a.
//
// Walks the network top-down for a 128-sample tile, exactly like the forward kernel walks it bottom-up:
//   dZv   = (g_rgb . W_rgb) (.) [HV > 0]                              CUDA cores (K = 3)
//   dFEAT = dZv . Wv[:, :256]                                         tcgen05, K = 128
//   dZ7   = (dFEAT . W_f + g_sigma (x) w_alpha) (.) [H7 > 0]          tcgen05, K = 256
//   dZ_{l-1} = (dZ_l . W_l[:, trunk cols]) (.) [H_{l-1} > 0], l = 7..1
// A = the current gradient tile in shared memory (K-major over the layer's OUTPUT features, 128B swizzle),
// B = chunks of the TRANSPOSED weights (packed section E, [256 input features x 64 output features]),
// D = one of two 128x256 fp32 accumulators in TMEM.  The ReLU masks are the forward kernel's bitmask tape
// (one u32 per row and 32-column unit, coalesced); every dZ tile is written to shared memory as the next
// step's A operand (two buffers, step parity) and streamed from there to the backward tape by 16 KB bulk
// stores, as K-block images for the wgrad kernel (mlp_bwd_tc.cu).  Gradients are rounded to bf16 between
// layers (bf16 in / fp32 accumulate); inputs (encodings, codes) get no gradient (A.4).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kDgTileM = 128;
constexpr int kDgStages = 3;
constexpr int kDgThreads = 576;
constexpr int kDgWorkerThreads = 512;
constexpr uint32_t kDgKB = 16384;                                  // one K-block image
constexpr uint32_t kDgOffAct = 0;                                  // 2 buffers x 4 K-blocks: gradient tiles (step parity)
constexpr uint32_t kDgOffW = 8 * kDgKB;                            // weight stages, 32 KB each
constexpr uint32_t kDgOffBar = kDgOffW + kDgStages * 32768;
constexpr uint32_t kDgNumBars = 2 * kDgStages + 8 + 2;
constexpr uint32_t kDgSmem = kDgOffBar + kDgNumBars * 8 + 16 + 1024;
static_assert(kDgSmem <= 227 * 1024, "shared memory budget");
constexpr int kDgSteps = 9;                                         // MMA steps per tile

struct DgradParams {
  const uint8_t* packed;        // blob: section E (transposed chunks) + aux
  int cond;
  const float4* g_raw;          // [M] (rgb_raw grads, sigma_raw grad)
  const uint32_t* mask_tape;    // kMaskUnits x 128 u32 per tile (layout.h)
  uint8_t* bwd_tape;            // kTapeBwdSlots images per tile
  float* flat_grad;             // head bias gradients (sums of g_raw) are added here
  int64_t M; int64_t ntiles;
};

// tape slot of the gradient tile that is the A operand of MMA step t (t = 9: dZ0, consumed by wgrad only)
__device__ __forceinline__ int dg_slot_of_step(int t) { return t == 0 ? kTapeBwdSlotZv : (t == 1 ? kTapeBwdSlotFeat : kTapeBwdSlotZ + 4 * (t - 2)); }

__global__ void __launch_bounds__(kDgThreads, 1) k_mlp_dgrad_tc(const DgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kDgOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kDgStages + s); };
  auto bar_act = [&](int buf, int kb) { return bar0 + 8u * (2 * kDgStages + 4 * buf + kb); };
  auto bar_acc = [&](int a) { return bar0 + 8u * (2 * kDgStages + 8 + a); };
  const uint32_t tmem_slot = bar0 + 8u * kDgNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int i = 0; i < 8; ++i) mbar_init(bar_act(i >> 2, i & 3), 8);   // 2 worker groups x 4 warps per K-block
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
    // ================================ MMA issuer + tape writer ======================================
    // Warp-uniform loop, one elected lane issues the MMAs.  Lane 0 also streams every finished gradient
    // image (the A operand it is about to consume) to the backward tape with one 16 KB bulk store; the
    // two activation buffers alternate by step, and before an accumulator is handed to the epilogue that
    // will overwrite a buffer, the stores still reading that buffer are drained (bulk wait_group.read).
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t desc_act = umma_desc_sw128(base + kDgOffAct);
    const uint64_t desc_w = umma_desc_sw128(base + kDgOffW);
    uint32_t wc = 0, tile_cnt = 0;
    const uint64_t pol_stream = l2_policy_evict_first();
    for (int64_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++tile_cnt) {
      uint8_t* btile = P.bwd_tape + (size_t)tile * kTapeBwdSlots * kDgKB;
#pragma unroll 1
      for (int t = 0; t <= kDgSteps; ++t) {
        const int buf = t & 1;
        const int nkb = t == 0 ? 2 : 4;
        const int slot0 = dg_slot_of_step(t);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          // phases per tile: buffer 0 K-blocks 0,1: b0 + 4 epilogues; K-blocks 2,3: 4 epilogues; buffer 1: 5 epilogues
          const uint32_t ph = buf ? 5u * tile_cnt + (uint32_t)(t >> 1) : (kb < 2 ? 5u * tile_cnt + (uint32_t)(t >> 1) : 4u * tile_cnt + (uint32_t)(t >> 1) - 1u);
          mbar_wait(bar_act(buf, kb), ph & 1u);
          if (lane == 0) {
            // evict-first: the tape is read next by wgrad, gigabytes later; keeping it out of L2's way is worth 2-3 % of a step
            bulk_s2g_hint(btile + (size_t)(slot0 + kb) * kDgKB, base + kDgOffAct + (uint32_t)(4 * buf + kb) * kDgKB, kDgKB, pol_stream);
            bulk_commit();
          }
          if (t < kDgSteps) {
            const uint32_t s = wc % kDgStages;
            mbar_wait(bar_full(s), (wc / kDgStages) & 1u);
            tc_fence_after();
            const uint64_t a_desc = desc_act + (uint64_t)((uint32_t)(4 * buf + kb) * (kDgKB >> 4));
            const uint64_t b_desc = desc_w + (uint64_t)(s * (32768u >> 4));
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmem_u + (uint32_t)buf * 256u, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                          (kb == 0 && ks == 0) ? 0u : 1u);
              umma_commit(bar_empty(s));
            }
            ++wc;
          }
          __syncwarp();
        }
        if (t < kDgSteps) {
          // epilogue t writes buffer (t+1)&1; the groups issued since that buffer's last stores: 2 (t = 0), 4 (else);
          // after step 8 the workers go straight to the next tile's b0 (buffer 0), so drain everything there
          if (lane == 0) {
            if (t == 0) bulk_wait_read<2>();
            else if (t == kDgSteps - 1) bulk_wait_read<0>();
            else bulk_wait_read<4>();
          }
          __syncwarp();
          if (elect_one()) umma_commit(bar_acc(buf));
          __syncwarp();
        }
      }
    }
    if (lane == 0) bulk_wait_all<0>();
  } else {
    // ================================ workers ======================================================
    const uint32_t q = (uint32_t)warp & 3u;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;
    const uint32_t row = q * 32u + (uint32_t)lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    const uint32_t act_row = base + kDgOffAct + row * 128u;
    const float* aux = reinterpret_cast<const float*>(P.packed + kSecBOffset);
    const float* walpha_g = aux + kAuxWAlpha;
    const float* wrgb_g = aux + kAuxWRgb;
    uint32_t acc_cnt[2] = {0u, 0u};
    float gsum0 = 0.0f, gsum1 = 0.0f, gsum2 = 0.0f, gsum3 = 0.0f;
    for (int64_t tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int64_t g = tile * kDgTileM + row;
      const float4 gr = g < P.M ? __ldg(P.g_raw + g) : make_float4(0.f, 0.f, 0.f, 0.f);   // padding rows carry no gradient
      const uint32_t* mtile = P.mask_tape + (size_t)tile * (kMaskUnits * 128) + row;
      if (grp == 0) {
        // head bias gradients + the g_raw image (bf16, chunk 0 of the row) the head wgrads multiply with
        gsum0 += gr.x; gsum1 += gr.y; gsum2 += gr.z; gsum3 += gr.w;
        uint8_t* grow = P.bwd_tape + ((size_t)tile * kTapeBwdSlots + kTapeBwdSlotG) * kDgKB + row * 128u;
        *reinterpret_cast<uint4*>(grow + ((row & 7u) << 4)) = make_uint4(pack_bf16(gr.x, gr.y), pack_bf16(gr.z, gr.w), 0u, 0u);
      }
      // ---- step b0: dZv = (g_rgb . W_rgb) (.) [HV > 0], 32 of the 128 columns per group -> buffer 0 ---
      {
        const uint32_t kb = grp >> 1, ch0 = (grp & 1u) * 4u;
        const uint32_t mb = __ldg(mtile + (kMaskUnitHv + (int)grp) * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = c * 4 + j, col = (int)grp * 32 + 2 * i;
            float lo = gr.x * __ldg(wrgb_g + col) + gr.y * __ldg(wrgb_g + kWV + col) + gr.z * __ldg(wrgb_g + 2 * kWV + col);
            float hi = gr.x * __ldg(wrgb_g + col + 1) + gr.y * __ldg(wrgb_g + kWV + col + 1) + gr.z * __ldg(wrgb_g + 2 * kWV + col + 1);
            if (!(mb & (1u << i))) lo = 0.0f;
            if (!(mb & (1u << (16 + i)))) hi = 0.0f;
            pk[j] = pack_bf16(lo, hi);
          }
          const uint32_t c16 = ch0 + (uint32_t)c;
          st_shared_v4(act_row + kb * kDgKB + ((c16 ^ (row & 7u)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_act(0, (int)kb));
      }
      // ---- MMA steps: t = 0 views -> dFEAT, t = 1 feature -> dZ7, t = 2..8 layers 7..1 -> dZ6..dZ0 -----
#pragma unroll 1
      for (int t = 0; t < kDgSteps; ++t) {
        const int a = t & 1, obuf = (t + 1) & 1;
        // ReLU bitmasks of H_{8-t} for this thread's two units, fetched while the MMAs run (dFEAT has none)
        uint32_t mk[2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
          mk[r] = t == 0 ? 0xFFFFFFFFu : __ldg(mtile + ((8 - t) * 8 + (int)grp + 4 * r) * 128);
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
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = c * 4 + j;
              float lo = __uint_as_float(v[2 * i]), hi = __uint_as_float(v[2 * i + 1]);
              if (t == 1) {                                          // sigma head: + g_sigma * w_alpha[k]
                lo = fmaf(gr.w, __ldg(walpha_g + col0 + 2 * i), lo);
                hi = fmaf(gr.w, __ldg(walpha_g + col0 + 2 * i + 1), hi);
              }
              if (!(mk[r] & (1u << i))) lo = 0.0f;
              if (!(mk[r] & (1u << (16 + i)))) hi = 0.0f;
              pk[j] = pack_bf16(lo, hi);
            }
            const uint32_t c16 = ch0 + (uint32_t)c;
            st_shared_v4(act_row + (uint32_t)(4 * obuf + (int)kb) * kDgKB + ((c16 ^ (row & 7u)) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_act(obuf, (int)kb));
        }
      }
    }
    if (grp == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gsum0 += __shfl_xor_sync(0xffffffffu, gsum0, o); gsum1 += __shfl_xor_sync(0xffffffffu, gsum1, o);
        gsum2 += __shfl_xor_sync(0xffffffffu, gsum2, o); gsum3 += __shfl_xor_sync(0xffffffffu, gsum3, o);
      }
      if (lane == 0) {
        atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond), gsum0);
        atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond) + 1, gsum1);
        atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond) + 2, gsum2);
        atomicAdd(P.flat_grad + flat_bias_offset(8, P.cond), gsum3);
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

int launch_mlp_dgrad_tc(const void* packed, int cond, const float* g_raw, const uint32_t* mask_tape, uint8_t* bwd_tape,
                        float* flat_grad, int64_t M, cudaStream_t s) {
  if (M == 0) return 0;
  static DeviceOnce once;
  if (cudaError_t e = opt_in_smem_once(once, k_mlp_dgrad_tc, kDgSmem)) return set_error((int)e, "mlp_dgrad_tc attr: %s", cudaGetErrorString(e));
  DgradParams P;
  P.packed = reinterpret_cast<const uint8_t*>(packed); P.cond = cond;
  P.g_raw = reinterpret_cast<const float4*>(g_raw);
  P.mask_tape = mask_tape; P.bwd_tape = bwd_tape; P.flat_grad = flat_grad;
  P.M = M; P.ntiles = (M + kDgTileM - 1) / kDgTileM;
  int64_t blocks = num_sms();
  if (blocks > P.ntiles) blocks = P.ntiles;
  k_mlp_dgrad_tc<<<(unsigned)blocks, kDgThreads, kDgSmem, s>>>(P);
  return check_launch("mlp_dgrad_tc");
}

}  // namespace fnerf
