// Layer-pipelined tensor-core backward of the NeRF MLP (A.4 backward: dgrad + wgrad in ONE kernel) for sm_100a.
//
// Why: the tile-major backward (mlp_dgrad_tc.cu + mlp_bwd_tc.cu) writes every gradient image dZ_l to an HBM tape and
// reads it back for the weight gradients (5 GB + 12 GB per 4096-ray step) and re-streams all transposed weights for
// every 128-sample tile.  Here every CTA is pinned to ONE layer (one 128-column half of it) for the whole launch:
//   * its transposed weights W_l^T[half] (64 KB) are loaded ONCE and stay in shared memory;
//   * its weight-gradient accumulator dW_l[:, half] (256 x 128 fp32) stays in TMEM for the whole launch and is flushed
//     once with red.global.add;
//   * per 128-sample tile it loads the incoming gradient tile dZ_l (4 K-block images, written by the CTAs of layer
//     l+1) and its half of the forward activation H_{l-1} from the tape, runs
//         dgrad:  dZ_{l-1}[:, half] = (dZ_l . W_l[:, half]) (.) [H_{l-1} > 0]     M = 128 samples, N = 128, K = 256
//         wgrad:  dW_l[:, half]    += dZ_l^T . H_{l-1}[:, half]                  M = 2 x 128 outputs, N = 128, K = 128 samples
//     on the SAME shared-memory images (K-major for dgrad, MN-major for wgrad), and hands dZ_{l-1}[:, half] to the CTAs
//     of layer l-1 through a small ring in global memory that lives in L2 (4 tiles x 64 KB per hand-off).
// dZ never reaches HBM; the only HBM streams are the forward tape (read once) and g_raw.
//
// Grid: kPipeRoles role-halves x kPipeLanes lanes.  Lane g owns the tiles t = g (mod kPipeLanes); the CTAs (role, g) of
// one lane form a linear pipeline  V -> F -> L7 -> ... -> L1 -> L0  (backward order), each boundary a ring with
// release/acquire counters in global memory (`ready`: images written, `done`: consumer halves that have the tile in
// shared memory).  All CTAs must be co-resident (cooperative launch, grid <= #SMs, 1 CTA / SM).
//   V0, V1   view branch: dZv = (g_rgb . W_rgb) (.) [HV > 0] on CUDA cores, dFEAT[:, half] = dZv . Wv[:, half],
//            dWv[:, half] += dZv^T . FEAT[:, half];  V1 also dWv[:, 256:283] (direction encoding);  V0 the rgb head;
//            both the alpha head for their half of H7 (B operand = a bf16 image of g_raw built in shared memory)
//   F0, F1   feature layer: dZ7[:, half] = (dFEAT . Wf[:, half] + g_sigma (x) w_alpha[half]) (.) [H7 > 0]
//   Ll_0/1   trunk layers l = 7..1 (layer 5: the trunk block of W5; L5_0 also dW5[:, 0:63] against the xyz encoding)
//   L0       dW0 += dZ0^T . PE   (no dgrad)
// Warps (448 threads): 0 loader (bulk TMA + ring flags), 1 MMA issuer, 2..9 epilogue (TMEM -> mask -> bf16 -> staging ->
// bulk store into the ring), 10..13 bias sums (column sums of the incoming gradient images, in registers).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kPipeLanes = 6;
constexpr int kPipeRoles = 24;                 // V0a V0b V1a V1b F0 F1 L7_0 L7_1 ... L1_0 L1_1 Z0a Z0b Z5a Z5b
constexpr int kPipeRings = 9;                  // V->F, F->L7, L7->L6, ..., L1->L0
#ifndef FNERF_PIPE_DEPTH
#define FNERF_PIPE_DEPTH 8
#endif
constexpr int kPipeDepth = FNERF_PIPE_DEPTH;   // tiles per ring (covers the store -> flag -> poll -> load round trip)
constexpr int kPipeDepthLast = FNERF_PIPE_DEPTH;
constexpr int kRingZ5 = 3;                     // ring L6 -> L5 (dZ5): third consumer = the L0 role (dW5[:, 0:63])
__host__ __device__ constexpr int ring_depth(int r) { return r == kPipeRings - 1 ? kPipeDepthLast : kPipeDepth; }
__host__ __device__ constexpr int64_t ring_first_tile(int r) { return (int64_t)r * kPipeLanes * kPipeDepth; }   // rings 0..r-1 are kPipeDepth deep
constexpr int64_t kRingTiles = ring_first_tile(kPipeRings - 1) + (int64_t)kPipeLanes * kPipeDepthLast;
constexpr int kPipeThreads = 512;              // warps: 0 loader, 1 MMA, 2..9 epilogue, 10..13 bias sums, 14..15 ring stores
constexpr uint32_t kImg = 16384, kPairB = 32768, kTileB = 65536;
constexpr int kMaxPairs = 7;
// shared memory (role dependent, see the map in the kernel):
//   trunk: W^T half 64 KB | staging 32 KB | 4 pair slots 128 KB
//   view:  W^T half 32 KB | staging 32 KB | dZv pair 32 KB | G image 16 KB | W_rgb tile 16 KB | 3 pair slots 96 KB
//   L0:    4 pair slots
constexpr uint32_t kPOffWt = 0;
constexpr uint32_t kPOffBar = 229376;
constexpr uint32_t kPNumBars = 2 * kMaxPairs + 16;
constexpr uint32_t kPOffWal = kPOffBar + 256;      // F roles: w_alpha[half] (128 fp32) for the rank-1 sigma term; L0 role: its job queue
static_assert(kPNumBars * 8 + 16 <= 256, "barrier block");   // 30 barriers + the TMEM slot
constexpr uint32_t kPipeSmem = kPOffWal + 512 + 1024;
static_assert(kPipeSmem <= 227 * 1024, "shared memory budget");
constexpr int kPipeStatSlots = 12;             // per role: cycles the warps spent waiting (debug, see fnerf_debug_pipe_stats)

enum { ROLE_V = 0, ROLE_T = 1, ROLE_Z = 2 };

struct PipeProduct {            // one weight-gradient product accumulated in TMEM and flushed at the end
  float* dw;                    // element (lane n of M-block mb, column c) -> dw[(mb*128 + n) * ld_n + (c - k0) * ld_k]
  int64_t ld_n, ld_k;
  int tmem_col, n_mb, ncols, k0, n_valid;
};
struct PipeRole {
  int kind, half;
  int t0, tstep;                // the role runs the lane's tiles t0, t0 + tstep, ... (view roles: two CTAs alternate)
  int in_ring, out_ring;        // -1: none
  int wt_chunk0, wt_nchunks;    // section E chunks of W^T (rows [128*half, +128) of each)
  int x_slot;                   // forward-tape slot of the activation pair (X operand of the main wgrad product)
  int mask_unit0;               // first ReLU-mask unit of the output columns (-1: none, dFEAT)
  int rank1;                    // F: add g_sigma (x) w_alpha[half] before the mask
  int e_slot;                   // forward-tape slot of an extra single image (PE / PED), -1 none
  int p_slot[2];                // forward-tape slots of extra pairs (HV, H7 half), -1 none
  float* bias;                  // column sums of the incoming gradient images {2*half, 2*half+1} (Z: all four)
  PipeProduct prod[3];
  int nprod;
};
struct PipeParams {
  PipeRole roles[kPipeRoles];
  const uint8_t* packed; int cond;
  const float4* g_raw;
  const uint8_t* fwd_tape; const uint32_t* mask_tape;
  uint8_t* ring;                // [kPipeRings][kPipeLanes][kPipeDepth] tiles of 64 KB
  uint32_t* flags;              // ready[kPipeRings][kPipeLanes][2 halves] then done[...], 32 words per (ring, lane)
  float* flat_grad;
  unsigned long long* stats;    // nullable: [kPipeRoles][kPipeStatSlots] summed wait cycles
  int64_t M, ntiles;
};

// Polls use RELAXED gpu-scope loads: an acquire load drags a CCTL.IVALL (whole-L1 invalidate) along on every poll, and
// what a successful poll guards is read by bulk TMA from L2, never through L1.  One 16-byte load fetches all counters.
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool cnt_ge(uint32_t v, uint32_t target) { return (int32_t)(v - target) >= 0; }
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Every spin on a counter written by another CTA of this launch is bounded: a peer that never arrives (a bug, or CTAs
// that are not co-resident) aborts the launch with a trap after 4 s instead of hanging the GPU.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__host__ __device__ constexpr uint32_t pipe_idesc_mn(int M, int N) {   // A and B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// mbarrier wait that also accounts the cycles spent (debug statistics)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc) {
  if (mbar_test_wait(bar, parity)) return;
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

__global__ void __launch_bounds__(kPipeThreads, 1) k_mlp_bwd_pipe(const __grid_constant__ PipeParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int role_id = (int)blockIdx.x / kPipeLanes, lane_g = (int)blockIdx.x % kPipeLanes;
  const PipeRole& Rl = P.roles[role_id];
  const int kind = Rl.kind, half = Rl.half;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_lane = P.ntiles > lane_g ? (P.ntiles - lane_g + kPipeLanes - 1) / kPipeLanes : 0;   // tiles of this lane
  const int64_t t0 = Rl.t0, tstep = Rl.tstep;                                 // ... of which this CTA runs i = t0, t0 + tstep, ...
  const bool has_work = t0 < n_lane;

  // ---- shared-memory map --------------------------------------------------------------------------
  const uint32_t wt_bytes = (uint32_t)Rl.wt_nchunks * kImg;                       // 64 KB (trunk), 32 KB (view), 0 (L0)
  const uint32_t off_stage = kind == ROLE_Z ? 0u : wt_bytes;
  const uint32_t off_zv = off_stage + kPairB;                                     // V only: dZv pair, G image, W_rgb tile
  const uint32_t off_g = off_zv + kPairB;
  const uint32_t off_wrgb = off_g + kImg;
  const uint32_t off_ring = kind == ROLE_V ? off_wrgb + kImg : (kind == ROLE_Z ? 0u : off_stage + kPairB);
  const int npairs = kind == ROLE_V ? 3 : (kind == ROLE_Z ? 7 : 4);   // L0 role: all of shared memory is its load ring
  const uint32_t bar0 = base + kPOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kMaxPairs + s); };
  const uint32_t bx = bar0 + 8u * (2 * kMaxPairs);
  const uint32_t bar_dg_full = bx, bar_dg_empty = bx + 8u;
  const uint32_t bar_zv_full = bx + 16u, bar_zv_empty = bx + 24u;
  const uint32_t bar_g_full = bx + 32u, bar_g_empty = bx + 40u;
  const uint32_t bar_done = bx + 48u, bar_wt = bx + 56u;
  const uint32_t bar_zacc_full = bx + 64u;
  auto bar_img_full = [&](uint32_t j) { return bx + 72u + 8u * j; };
  auto bar_img_empty = [&](uint32_t j) { return bx + 88u + 8u * j; };
  // trunk roles keep TWO dgrad accumulators (TMEM columns 0 and 384, tile parity) so the next tile's first MMAs need not
  // wait for this tile's epilogue to have drained the accumulator; buffer 1 has its own barrier pair
  const uint32_t bar_dg_full1 = bx + 104u, bar_dg_empty1 = bx + 112u;
  const uint32_t tmem_slot = bar0 + 8u * kPNumBars;
  auto pair_addr = [&](int s) { return base + off_ring + (uint32_t)s * kPairB; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxPairs; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1 + 4); }   // MMA commit + 4 bias warps
    mbar_init(bar_dg_full, 1);  mbar_init(bar_dg_empty, 8);
    mbar_init(bar_dg_full1, 1); mbar_init(bar_dg_empty1, 8);
    mbar_init(bar_zv_full, 8);  mbar_init(bar_zv_empty, 1 + 4);
    mbar_init(bar_g_full, 8);   mbar_init(bar_g_empty, 1);
    mbar_init(bar_done, 1);     mbar_init(bar_wt, 1);
    mbar_init(bar_zacc_full, 1);
    for (uint32_t j = 0; j < 2; ++j) { mbar_init(bar_img_full(j), 4); mbar_init(bar_img_empty(j), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (Rl.rank1) {
    const float* wa = reinterpret_cast<const float*>(P.packed + kSecBOffset) + kAuxWAlpha + half * 128;
    for (int n = threadIdx.x; n < 128; n += kPipeThreads) reinterpret_cast<float*>(base_ptr + kPOffWal)[n] = wa[n];
  }
  if (kind == ROLE_V) {
    // constant B operand of the dZv product (K-major, K = 16): row n = HV unit n, columns (k): 0..2 = bf16(w_c[n]), 3 = 0
    // (the sigma column of the G image), 4..6 = bf16(w_c[n]) again (they meet the low parts of g), 7..9 = the low parts
    // of w_c[n] (they meet the high parts of g): g.w to ~16 mantissa bits with bf16 operands
    const float* wrgb = reinterpret_cast<const float*>(P.packed + kSecBOffset) + kAuxWRgb;
    for (int n = threadIdx.x; n < kWV; n += kPipeThreads) {
      float w[3];
      uint32_t hi[3], lo[3];
      for (int c = 0; c < 3; ++c) {
        w[c] = wrgb[c * kWV + n];
        hi[c] = pack_bf16(w[c], 0.0f) & 0xFFFFu;
        lo[c] = pack_bf16(w[c] - __uint_as_float(hi[c] << 16), 0.0f) & 0xFFFFu;
      }
      const uint32_t r = (uint32_t)n;
      // chunk 0: k = 0..7 = (h0, h1, h2, 0, h0, h1, h2, l0)   chunk 1: k = 8..15 = (l1, l2, 0, ...)
      st_shared_v4(base + off_wrgb + r * 128u + ((r & 7u) << 4), hi[0] | (hi[1] << 16), hi[2], hi[0] | (hi[1] << 16), hi[2] | (lo[0] << 16));
      st_shared_v4(base + off_wrgb + r * 128u + (((r & 7u) ^ 1u) << 4), lo[1] | (lo[2] << 16), 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kPOffBar + 8 * kPNumBars);

  uint32_t* ready = P.flags;
  uint32_t* done = P.flags + kPipeRings * kPipeLanes * 32;
  auto flag_idx = [&](int ring) { return (ring * kPipeLanes + lane_g) * 32; };
  // `ready` counters: one per store thread of a ring = (tile parity for ring 0 whose producers alternate, half, image),
  // word 4 * parity + 2 * half + image, counting the tiles that thread has published IN ORDER.  The four counters of a
  // parity sit in one 16-byte word: a consumer polls with ONE load (each poll is an L2 round trip on the loader's critical
  // path) and remembers what it saw, so a producer that runs ahead is not polled again.
  uint32_t seen_ready[2] = {0u, 0u};
  auto is_ready = [&](int ring, int64_t i) {
    const int np = ring == 0 ? 2 : 1, par = (int)(i % np);
    const uint32_t target = (uint32_t)(i / np + 1);
    if (cnt_ge(seen_ready[par], target)) return true;
    const uint4 v = ld_relaxed_gpu_v4(ready + flag_idx(ring) + 4 * par);
    uint32_t m = v.x;
    if (!cnt_ge(v.y, m)) m = v.y;
    if (!cnt_ge(v.z, m)) m = v.z;
    if (!cnt_ge(v.w, m)) m = v.w;
    seen_ready[par] = m;                       // the slowest of the four publishers
    return cnt_ge(m, target);
  };
  auto wait_ready = [&](int ring, int64_t i) {
    if (is_ready(ring, i)) return;
    const uint64_t tstart = global_timer_ns();
    while (!is_ready(ring, i)) {
      __nanosleep(32);
      if (global_timer_ns() - tstart > 4000000000ull) __trap();
    }
  };
  auto ring_tile = [&](int ring, int64_t i) {
    const int d = ring_depth(ring);
    return P.ring + ((size_t)ring_first_tile(ring) + (size_t)lane_g * d + (size_t)(i % d)) * kTileB;
  };
  // job sequence of this CTA: its k-th job is tile t0 + k * tstep of the lane
  auto next_job = [&](int64_t it, int64_t& i, bool& s5) -> int {
    i = t0 + it * tstep; s5 = false;
    return i < n_lane ? 1 : 0;
  };
  const int64_t fstride = (int64_t)kTapeFwdSlots * kImg;
  // pairs a tile consumes, in this order (loader, MMA warp and bias warps walk the same sequence of slots):
  //   T: A (dZ images 0,1)  C (X pair)  B (dZ images 2,3)  [E (xyz encoding image)]
  //   V: C (FEAT half pair)  P0 (HV pair | PED image)  P1 (H7 half pair)
  //   Z: A  B  E
  const int n_seq = kind == ROLE_T ? (Rl.e_slot >= 0 ? 4 : 3) : 3;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;            // debug statistics: cycles spent waiting

  if (warp == 0) {
    // ================================ loader ========================================================
    if (lane == 0 && has_work) {
      if (wt_bytes) {                                   // stationary W^T half: rows [128*half, +128) of every chunk
        mbar_expect_tx(bar_wt, wt_bytes);
        const uint8_t* wsrc = P.packed + sec_e_offset(P.cond);
        for (int c = 0; c < Rl.wt_nchunks; ++c)
          bulk_g2s(base + kPOffWt + (uint32_t)c * kImg, wsrc + (size_t)(Rl.wt_chunk0 + c) * kChunkTBytes + (size_t)half * kImg, kImg, bar_wt);
      }
      uint32_t pc = 0;
      auto load = [&](const void* src, uint32_t bytes) {
        const uint32_t s = pc % npairs;
        mbar_wait_t(bar_empty(s), ((pc / npairs) & 1u) ^ 1u, w1);
        mbar_expect_tx(bar_full(s), bytes);
        bulk_g2s(pair_addr(s), src, bytes, bar_full(s));
        ++pc;
      };
      for (int64_t it = 0;; ++it) {
        int64_t i; bool s5;
        const int jr = next_job(it, i, s5);
        if (jr == 0) break;
        if (jr == 2) continue;
        const int64_t tile = lane_g + i * kPipeLanes;
        const uint8_t* ft = P.fwd_tape + tile * fstride;
        const uint8_t* src = nullptr;
        if (kind != ROLE_V) {
          // both halves of the producing layer have published their two images of tile i
          const int in_ring = Rl.in_ring;
          const long long c0 = clock64();
          wait_ready(in_ring, i);
          w0 += clock64() - c0;
          fence_proxy_async_global();
          src = ring_tile(in_ring, i);
        }
        if (kind == ROLE_T) {
          load(src, kPairB);
          load(ft + (size_t)Rl.x_slot * kImg, kPairB);
          load(src + kPairB, kPairB);
          if (Rl.e_slot >= 0) load(ft + (size_t)Rl.e_slot * kImg, kImg);
        } else if (kind == ROLE_Z) {
          load(src, kPairB);
          load(src + kPairB, kPairB);
          load(ft + (size_t)Rl.e_slot * kImg, kImg);
        } else {
          load(ft + (size_t)Rl.x_slot * kImg, kPairB);
          if (Rl.p_slot[0] >= 0) load(ft + (size_t)Rl.p_slot[0] * kImg, kPairB);
          else load(ft + (size_t)Rl.e_slot * kImg, kImg);
          load(ft + (size_t)Rl.p_slot[1] * kImg, kPairB);
        }
      }
      if (P.stats) { atomicAdd(P.stats + role_id * kPipeStatSlots + 0, (unsigned long long)w0); atomicAdd(P.stats + role_id * kPipeStatSlots + 1, (unsigned long long)w1); }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ====================================================
    if (has_work) {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc_dg = umma_idesc_bf16(128, 128);
      uint32_t pc = 0;
      auto slot_of = [&](uint32_t& s, uint32_t& ph) { s = pc % npairs; ph = (pc / npairs) & 1u; ++pc; };
      auto kdesc = [&](uint32_t addr) { return umma_desc_sw128(addr); };
      auto mndesc = [&](uint32_t addr) { return umma_desc_mn_sw128(addr, kImg); };
      // one wgrad product block: D[tmem_col ..] (+)= A(MN-major pair / image)^T-view . B(MN-major), K = 128 samples
      auto wgrad = [&](uint32_t dcol, uint32_t a_addr, uint32_t b_addr, int ncols, uint32_t first) {
        const uint32_t idesc = pipe_idesc_mn(128, ncols);
        const uint64_t a0 = mndesc(a_addr), b0 = mndesc(b_addr);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tm + dcol, a0 + (uint64_t)(ks * 128), b0 + (uint64_t)(ks * 128), idesc, (ks == 0) ? first : 1u);
      };
      if (wt_bytes) mbar_wait(bar_wt, 0);
      constexpr uint32_t idesc_wg = pipe_idesc_mn(128, 128);
      const uint64_t kd_ring = kdesc(base + off_ring), md_ring = mndesc(base + off_ring), kd_w = kdesc(base + kPOffWt);
      int64_t kk = 0;
      for (int64_t it = 0;; ++it) {
        int64_t i; bool s5;
        const int jr = next_job(it, i, s5);
        if (jr == 0) break;
        if (jr == 2) continue;
        const uint32_t first = kk == 0 ? 0u : 1u;        // accumulate flag of the launch-long wgrad accumulators
        const uint32_t par = (uint32_t)(kk & 1);
        const bool last = i + tstep >= n_lane;
        ++kk;
        if (kind == ROLE_T) {
          uint32_t sA, pA, sC, pCc, sB, pB;
          slot_of(sA, pA); slot_of(sC, pCc); slot_of(sB, pB);
          // dgrad over K-blocks 0,1 | wgrad M-block 0 | dgrad over K-blocks 2,3 | wgrad M-block 1: pair A is released at
          // half time, so the loader's prefetch distance (4 pair slots = 1 1/3 tiles) covers the L2 latency.  All 32 MMAs
          // are N = 128 (64 tensor-pipe cycles each): the issuing thread must not spend more than that per MMA, so every
          // descriptor below is one 64-bit add on a base computed before the tile loop and every loop is unrolled.
          const uint64_t aA = kd_ring + (uint64_t)(sA * (kPairB >> 4)), aB = kd_ring + (uint64_t)(sB * (kPairB >> 4));
          const uint64_t mA = md_ring + (uint64_t)(sA * (kPairB >> 4)), mB = md_ring + (uint64_t)(sB * (kPairB >> 4));
          const uint64_t mC = md_ring + (uint64_t)(sC * (kPairB >> 4));
          const uint32_t dbuf = (uint32_t)((kk - 1) & 1), dpar = (uint32_t)(((kk - 1) >> 1) & 1);   // kk was incremented above
          const uint32_t tmd = tm + dbuf * 384u;
          mbar_wait_t(dbuf ? bar_dg_empty1 : bar_dg_empty, dpar ^ 1u, w1);
          mbar_wait_t(bar_full(sA), pA, w0);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmd, aA + (uint64_t)(kb * (kImg >> 4) + 2 * ks), kd_w + (uint64_t)(kb * (kImg >> 4) + 2 * ks), idesc_dg, (kb | ks) ? 1u : 0u);
          }
          __syncwarp();
          mbar_wait_t(bar_full(sC), pCc, w0);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16(tm + 128u, mA + (uint64_t)(ks * 128), mC + (uint64_t)(ks * 128), idesc_wg, ks == 0 ? first : 1u);
            umma_commit(bar_empty(sA));
          }
          __syncwarp();
          mbar_wait_t(bar_full(sB), pB, w0);
          tc_fence_after();
          if (lane == 0) red_relaxed_gpu_add(done + flag_idx(Rl.in_ring) + half, 1u);    // this half holds tile i in shared memory
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmd, aB + (uint64_t)(kb * (kImg >> 4) + 2 * ks), kd_w + (uint64_t)((kb + 2) * (kImg >> 4) + 2 * ks), idesc_dg, 1u);
            umma_commit(dbuf ? bar_dg_full1 : bar_dg_full);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16(tm + 256u, mB + (uint64_t)(ks * 128), mC + (uint64_t)(ks * 128), idesc_wg, ks == 0 ? first : 1u);
            umma_commit(bar_empty(sC));
            umma_commit(bar_empty(sB));
            if (last) umma_commit(bar_done);
          }
          __syncwarp();
        } else if (kind == ROLE_Z) {
          uint32_t sA, pA, sB, pB, sE, pE;
          slot_of(sA, pA); slot_of(sB, pB); slot_of(sE, pE);
          mbar_wait_t(bar_full(sA), pA, w0);
          mbar_wait_t(bar_full(sB), pB, w0);
          if (lane == 0) red_relaxed_gpu_add(done + flag_idx(Rl.in_ring) + 2 + (int)t0, 1u);
          mbar_wait_t(bar_full(sE), pE, w0);
          tc_fence_after();
          if (elect_one()) {
            const PipeProduct& pr = Rl.prod[0];
            wgrad((uint32_t)pr.tmem_col, pair_addr(sA), pair_addr(sE), pr.ncols, first);
            wgrad((uint32_t)(pr.tmem_col + pr.ncols), pair_addr(sB), pair_addr(sE), pr.ncols, first);
            umma_commit(bar_empty(sA)); umma_commit(bar_empty(sB)); umma_commit(bar_empty(sE));
            if (last) umma_commit(bar_done);
          }
          __syncwarp();
        } else {
          uint32_t sC, pCc, sP0, pP0, sP1, pP1;
          slot_of(sC, pCc); slot_of(sP0, pP0); slot_of(sP1, pP1);
          // ---- dZv accumulator = G . W_rgb^T (one K = 16 MMA), then the workers mask it into the dZv images ---------
          mbar_wait_t(bar_g_full, par, w1);
          tc_fence_after();
          if (elect_one()) {
            umma_bf16(tm + 320u, kdesc(base + off_g), kdesc(base + off_wrgb), idesc_dg, 0u);
            umma_commit(bar_zacc_full);
          }
          __syncwarp();
          // ---- dgrad: dFEAT[:, half] = dZv . Wv[:, half] ----------------------------------------------------
          mbar_wait_t(bar_dg_empty, par ^ 1u, w1);
          mbar_wait_t(bar_zv_full, par, w1);
          tc_fence_after();
          if (elect_one()) {
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t a = kdesc(base + off_zv + (uint32_t)kb * kImg), b = kdesc(base + kPOffWt + (uint32_t)kb * kImg);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tm, a + (uint64_t)(2 * ks), b + (uint64_t)(2 * ks), idesc_dg, (kb | ks) ? 1u : 0u);
            }
            umma_commit(bar_dg_full);
          }
          __syncwarp();
          mbar_wait_t(bar_full(sC), pCc, w0);
          tc_fence_after();
          if (elect_one()) {
            wgrad((uint32_t)Rl.prod[0].tmem_col, base + off_zv, pair_addr(sC), Rl.prod[0].ncols, first);
            umma_commit(bar_empty(sC));
          }
          __syncwarp();
          mbar_wait_t(bar_full(sP0), pP0, w0);
          mbar_wait_t(bar_full(sP1), pP1, w0);
          tc_fence_after();
          if (elect_one()) {
            // prod[1]: V0: rgb head, A = HV pair, B = G image;  V1: dWv[:, 256:283], A = dZv, B = PED image
            if (Rl.p_slot[0] >= 0) wgrad((uint32_t)Rl.prod[1].tmem_col, pair_addr(sP0), base + off_g, Rl.prod[1].ncols, first);
            else wgrad((uint32_t)Rl.prod[1].tmem_col, base + off_zv, pair_addr(sP0), Rl.prod[1].ncols, first);
            // prod[2]: alpha head for this half of H7: A = H7 pair, B = G image
            wgrad((uint32_t)Rl.prod[2].tmem_col, pair_addr(sP1), base + off_g, Rl.prod[2].ncols, first);
            umma_commit(bar_empty(sP0)); umma_commit(bar_empty(sP1));
            umma_commit(bar_zv_empty); umma_commit(bar_g_empty);
            if (last) umma_commit(bar_done);
          }
          __syncwarp();
        }
      }
      if (P.stats && lane == 0) { atomicAdd(P.stats + role_id * kPipeStatSlots + 2, (unsigned long long)w0); atomicAdd(P.stats + role_id * kPipeStatSlots + 3, (unsigned long long)w1); }
    }
  } else if (warp < 10) {
    // ================================ epilogue warps ================================================
    const uint32_t q = (uint32_t)warp & 3u;                   // TMEM lane quadrant of this warp
    const uint32_t j = (uint32_t)(warp - 2) >> 2;             // output image of the half (64 columns)
    const uint32_t row = q * 32u + (uint32_t)lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    float gs0 = 0.f, gs1 = 0.f, gs2 = 0.f, gs3 = 0.f;         // V0: sums of g_raw (head bias gradients)
    // accumulator columns -> (+ rank-1 term) -> mask -> bf16 -> one swizzled 64-column image row
    auto emit = [&](const uint32_t (&v)[32], uint32_t mk, uint32_t dst_row, int u, float gsig, const float* wal) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int ii = c * 4 + jj;
          float lo = __uint_as_float(v[2 * ii]), hi = __uint_as_float(v[2 * ii + 1]);
          if (wal != nullptr) {                        // shared memory, same address in every lane: broadcast reads
            lo = fmaf(gsig, wal[2 * ii], lo);
            hi = fmaf(gsig, wal[2 * ii + 1], hi);
          }
          if (!(mk & (1u << ii))) lo = 0.0f;
          if (!(mk & (1u << (16 + ii)))) hi = 0.0f;
          pk[jj] = pack_bf16(lo, hi);
        }
        const uint32_t c16 = (uint32_t)(u * 4 + c);
        st_shared_v4(dst_row + ((c16 ^ (row & 7u)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
    };
    if (kind != ROLE_Z) {
      const uint32_t stage_row = base + off_stage + j * kImg + row * 128u;
      // per-tile global inputs (g_raw row, ReLU-mask words) are fetched ONE TILE AHEAD: on the tile's own critical path
      // their DRAM latency (~1 us) was the longest link of the view roles' chain
      struct TileIn { float4 gr; uint32_t mk0, mk1, mhv0, mhv1; };
      auto fetch = [&](int64_t i) {
        TileIn t;
        t.gr = make_float4(0.f, 0.f, 0.f, 0.f);
        t.mk0 = t.mk1 = 0xFFFFFFFFu; t.mhv0 = t.mhv1 = 0u;
        if (i >= n_lane) return t;
        const int64_t tile = lane_g + i * kPipeLanes;
        const int64_t g = tile * 128 + row;
        const uint32_t* mtile = P.mask_tape + (size_t)tile * (kMaskUnits * 128) + row;
        if ((kind == ROLE_V || Rl.rank1) && g < P.M) t.gr = __ldg(P.g_raw + g);
        if (Rl.mask_unit0 >= 0) {
          t.mk0 = __ldg(mtile + (Rl.mask_unit0 + (int)j * 2) * 128);
          t.mk1 = __ldg(mtile + (Rl.mask_unit0 + (int)j * 2 + 1) * 128);
        }
        if (kind == ROLE_V) {
          t.mhv0 = __ldg(mtile + (kMaskUnitHv + (int)j * 2) * 128);
          t.mhv1 = __ldg(mtile + (kMaskUnitHv + (int)j * 2 + 1) * 128);
        }
        return t;
      };
      TileIn nxt = fetch(t0);
      int64_t kk = 0;
      for (int64_t i = t0; i < n_lane; i += tstep, ++kk) {
        const uint32_t par = (uint32_t)(kk & 1);
        const TileIn cur = nxt;
        nxt = fetch(i + tstep);
        const float4 gr = cur.gr;
        uint32_t v0[32], v1[32];
        if (kind == ROLE_V) {
          // ---- the bf16 image of g_raw: A operand of the dZv product, B operand of the two head products ----------
          const uint32_t mhv0 = cur.mhv0, mhv1 = cur.mhv1;
          mbar_wait_t(bar_g_empty, par ^ 1u, w1);
          if (j == 0) {
            // k = 0..2: high parts of g_rgb, 3: g_sigma, 4..6: low parts of g_rgb, 7..9: high parts again (see the W_rgb tile)
            const uint32_t h01 = pack_bf16(gr.x, gr.y), h2s = pack_bf16(gr.z, gr.w);
            const float hx = __uint_as_float(h01 << 16), hy = __uint_as_float(h01 & 0xFFFF0000u), hz = __uint_as_float(h2s << 16);
            const uint32_t l01 = pack_bf16(gr.x - hx, gr.y - hy), l2 = pack_bf16(gr.z - hz, 0.0f) & 0xFFFFu;
            st_shared_v4(base + off_g + row * 128u + ((row & 7u) << 4), h01, h2s, l01, l2 | (h01 << 16));
            st_shared_v4(base + off_g + row * 128u + (((row & 7u) ^ 1u) << 4), (h01 >> 16) | (h2s << 16), 0u, 0u, 0u);
            fence_proxy_async_smem();
            if (half == 0) { gs0 += gr.x; gs1 += gr.y; gs2 += gr.z; gs3 += gr.w; }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_g_full);                 // all 8 warps: also "my reads of the dZv accumulator are done"
          // ---- dZv image j = accumulator (.) [HV > 0] -----------------------------------------------------------
          mbar_wait_t(bar_zacc_full, par, w0);
          tc_fence_after();
          tmem_ld32(tmem_row + 320u + j * 64u, v0);
          tmem_ld32(tmem_row + 320u + j * 64u + 32u, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_wait_t(bar_zv_empty, par ^ 1u, w1);
          const uint32_t zv_row = base + off_zv + j * kImg + row * 128u;
          emit(v0, mhv0, zv_row, 0, 0.0f, nullptr);
          emit(v1, mhv1, zv_row, 1, 0.0f, nullptr);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_zv_full);
        }
        // ---- dgrad epilogue: accumulator -> (+ rank-1 sigma term) -> ReLU mask -> bf16 -> staging image j -----
        const uint32_t mk0 = cur.mk0, mk1 = cur.mk1;
        const uint32_t dbuf = kind == ROLE_T ? (uint32_t)(kk & 1) : 0u;
        const uint32_t dpar = kind == ROLE_T ? (uint32_t)((kk >> 1) & 1) : par;
        mbar_wait_t(dbuf ? bar_dg_full1 : bar_dg_full, dpar, w0);
        tc_fence_after();
        tmem_ld32(tmem_row + dbuf * 384u + j * 64u, v0);
        tmem_ld32(tmem_row + dbuf * 384u + j * 64u + 32u, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dbuf ? bar_dg_empty1 : bar_dg_empty);   // this accumulator may be overwritten again
        mbar_wait_t(bar_img_empty(j), par ^ 1u, w2);            // the store of the previous tile has read the staging image
        const float* wal = Rl.rank1 ? reinterpret_cast<const float*>(base_ptr + kPOffWal) + (int)j * 64 : nullptr;
        emit(v0, mk0, stage_row, 0, gr.w, wal);
        emit(v1, mk1, stage_row, 1, gr.w, wal ? wal + 32 : nullptr);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_img_full(j));
      }
    }
    // ---- flush the launch-long weight-gradient accumulators -------------------------------------------
    if (has_work) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      for (int p = 0; p < Rl.nprod; ++p) {
        const PipeProduct& pr = Rl.prod[p];
        const bool vec = pr.ld_k == 1 && pr.k0 == 0 && (pr.ld_n & 3) == 0 && (pr.n_valid & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(pr.dw) & 15u) == 0;
        for (int mb = 0; mb < pr.n_mb; ++mb) {
          float* out = pr.dw + (int64_t)(mb * 128 + (int)row) * pr.ld_n;
          // the two warps of a lane quadrant split the 32-column groups
          for (int c0 = (int)j * 32; c0 < pr.ncols && c0 < pr.n_valid; c0 += 64) {
            uint32_t v[32];
            tmem_ld32(tmem_row + (uint32_t)(pr.tmem_col + mb * pr.ncols + c0), v);
            tmem_ld_wait();
            if (vec) {
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                if (c0 + c < pr.n_valid)
                  red_add_v4(out + c0 + c, __uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (c0 + c >= pr.k0 && c0 + c < pr.n_valid) atomicAdd(out + (int64_t)(c0 + c - pr.k0) * pr.ld_k, __uint_as_float(v[c]));
            }
          }
        }
      }
      tc_fence_before();
      if (kind == ROLE_V && half == 0 && j == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs0 += __shfl_xor_sync(0xffffffffu, gs0, o); gs1 += __shfl_xor_sync(0xffffffffu, gs1, o);
          gs2 += __shfl_xor_sync(0xffffffffu, gs2, o); gs3 += __shfl_xor_sync(0xffffffffu, gs3, o);
        }
        if (lane == 0) {
          atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond), gs0);
          atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond) + 1, gs1);
          atomicAdd(P.flat_grad + flat_bias_offset(11, P.cond) + 2, gs2);
          atomicAdd(P.flat_grad + flat_bias_offset(8, P.cond), gs3);
        }
      }
      if (P.stats && lane == 0) {
        atomicAdd(P.stats + role_id * kPipeStatSlots + 4, (unsigned long long)w0); atomicAdd(P.stats + role_id * kPipeStatSlots + 5, (unsigned long long)w1);
        atomicAdd(P.stats + role_id * kPipeStatSlots + 6, (unsigned long long)w2);
      }
    }
  } else if (warp < 14) {
    // ================================ bias warps ====================================================
    // column sums of the incoming gradient images (bias gradients), accumulated in registers over the launch: thread e
    // owns the column pair (2p, 2p+1) of one image over a block of rows
    const int e = (int)threadIdx.x - 320;                 // 0..127
    const uint32_t p = (uint32_t)e & 31u;
    const int sel = e >> 5;                               // warp of the four
    float b0 = 0.0f, b1 = 0.0f, c0 = 0.0f, c1 = 0.0f;
    auto colsum = [&](const uint8_t* img, uint32_t r0, uint32_t nrows, float& s0, float& s1) {
      const uint8_t* src = img + (p & 3u) * 4u;
#pragma unroll 8
      for (uint32_t r = r0; r < r0 + nrows; ++r) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + r * 128u + (((p >> 2) ^ (r & 7u)) << 4));
        s0 += __uint_as_float(w << 16);
        s1 += __uint_as_float(w & 0xFFFF0000u);
      }
    };
    uint32_t pc = 0;
    int64_t kk = 0;
    for (int64_t it = 0;; ++it) {
      int64_t i; bool s5;
      const int jr = next_job(it, i, s5);
      if (jr == 0) break;
      if (jr == 2) continue;
      for (int k = 0; k < n_seq; ++k, ++pc) {
        const uint32_t s = pc % npairs;
        mbar_wait(bar_full(s), (pc / npairs) & 1u);
        const uint8_t* pair = base_ptr + off_ring + s * kPairB;
        if (kind == ROLE_T && k == (half ? 2 : 0)) {
          // images {2 half, 2 half + 1} of dZ (pair A or B): warp sel -> image sel / 2, rows half sel % 2
          colsum(pair + (uint32_t)(sel >> 1) * kImg, (uint32_t)(sel & 1) * 64u, 64u, b0, b1);
        } else if (kind == ROLE_Z && k < 2 && Rl.bias != nullptr) {
          // all four images: pair k, warp sel -> image sel / 2 of the pair, rows half sel % 2
          if (k == 0) colsum(pair + (uint32_t)(sel >> 1) * kImg, (uint32_t)(sel & 1) * 64u, 64u, b0, b1);
          else colsum(pair + (uint32_t)(sel >> 1) * kImg, (uint32_t)(sel & 1) * 64u, 64u, c0, c1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
      }
      if (kind == ROLE_V) {
        mbar_wait(bar_zv_full, (uint32_t)(kk & 1));
        // dZv image `half`: 64 columns = 32 pairs; warp sel sums rows [32 sel, +32)
        colsum(base_ptr + off_zv + (uint32_t)half * kImg, (uint32_t)sel * 32u, 32u, b0, b1);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_zv_empty);
      }
      ++kk;
    }
    if (has_work && Rl.bias != nullptr) {
      if (kind == ROLE_T) {
        float* dst = Rl.bias + half * 128 + (sel >> 1) * 64 + 2 * (int)p;
        atomicAdd(dst, b0); atomicAdd(dst + 1, b1);
      } else if (kind == ROLE_Z) {
        float* dst = Rl.bias + (sel >> 1) * 64 + 2 * (int)p;
        atomicAdd(dst, b0); atomicAdd(dst + 1, b1);
        atomicAdd(dst + 128, c0); atomicAdd(dst + 129, c1);
      } else {
        float* dst = Rl.bias + half * 64 + 2 * (int)p;
        atomicAdd(dst, b0); atomicAdd(dst + 1, b1);
      }
    }
  } else {
    // ================================ ring stores ===================================================
    // warp 14 + j hands image 2*half + j of every tile to the next layer: bulk store from the staging image, then the
    // `ready` counter.  The staging image is released as soon as the store has READ it; the counter is published when the
    // store has COMPLETED, which is only awaited once the next tile's store is in flight (or at once, when no next image
    // is waiting), so the global write latency is off the per-tile critical path.
    const uint32_t j = (uint32_t)warp - 14u;
    if (lane == 0 && kind != ROLE_Z && has_work) {
      int64_t published = 0;                                   // in units of this CTA's tiles
      uint32_t* rdy = ready + flag_idx(Rl.out_ring) + 4 * (int)t0 + 2 * half + (int)j;
      // credits: the `done` counters of the ring's consumers (words 0, 1: the halves of the next layer, in tile order;
      // words 2, 3: on the dZ5 / dZ0 rings the alternating Z CTAs, CTA p counting the tiles of parity p), cached
      uint4 seen_done = make_uint4(0u, 0u, 0u, 0u);
      const bool has_halves = Rl.out_ring != kPipeRings - 1, has_z = Rl.out_ring == kRingZ5 || Rl.out_ring == kPipeRings - 1;
      auto has_credit = [&](int64_t old_tile) {               // every consumer has `old_tile` in its shared memory
        const uint32_t th = (uint32_t)(old_tile + 1), tz = (uint32_t)(old_tile / 2 + 1);
        const uint32_t zc = (old_tile & 1) ? seen_done.w : seen_done.z;
        return (!has_halves || (cnt_ge(seen_done.x, th) && cnt_ge(seen_done.y, th))) && (!has_z || cnt_ge(zc, tz));
      };
      auto publish_upto = [&](int64_t n) {                     // tiles [published, n) have completed their stores
        if (n > published) {
          // the stores of these tiles have completed (wait_group); the proxy fence orders the async-proxy writes before
          // this thread's release, which publishes them at gpu scope
          fence_proxy_async_global();
#ifdef FNERF_PIPE_RELAXED_PUBLISH              // experiment: ~1500 cycles cheaper per tile, but the publication then leans on the proxy fence alone
          red_relaxed_gpu_add(rdy, (uint32_t)(n - published));
#else
          red_release_gpu_add(rdy, (uint32_t)(n - published));
#endif
          published = n;
        }
      };
      int64_t kk = 0;
      for (int64_t i = t0; i < n_lane; i += tstep, ++kk) {
        const uint32_t par = (uint32_t)(kk & 1);
        // Publish the previous tile BEFORE the next store is issued: the release fence of the publication waits for every
        // memory operation this thread has in flight, a freshly issued 16 KB store included (that cost 1650 cycles per
        // tile).  Normally this thread is early and the wait for the store's completion falls into its idle time.
        {
          long long c1 = clock64();
          bulk_wait_all<0>();
          long long c2 = clock64();
          w3 += c2 - c1;
          publish_upto(kk);
          w4 += clock64() - c2;
          mbar_wait_t(bar_img_full(j), par, w0);
        }
        const int depth = ring_depth(Rl.out_ring);
        if (i >= depth && !has_credit(i - depth)) {            // the ring slot still holds a tile some consumer has not loaded
          const long long c0 = clock64();
          const uint64_t tstart = global_timer_ns();
          for (;;) {
            seen_done = ld_relaxed_gpu_v4(done + flag_idx(Rl.out_ring));
            if (has_credit(i - depth)) break;
            __nanosleep(32);
            if (global_timer_ns() - tstart > 4000000000ull) __trap();
          }
          w1 += clock64() - c0;
        }
        bulk_s2g(ring_tile(Rl.out_ring, i) + (size_t)(2 * half + (int)j) * kImg, base + off_stage + j * kImg, kImg);
        bulk_commit();
        const long long c1 = clock64();
        bulk_wait_read<0>();
        w2 += clock64() - c1;
        mbar_arrive(bar_img_empty(j));
      }
      bulk_wait_all<0>();
      publish_upto(kk);
      if (P.stats && j == 0) {
        atomicAdd(P.stats + role_id * kPipeStatSlots + 7, (unsigned long long)w0); atomicAdd(P.stats + role_id * kPipeStatSlots + 8, (unsigned long long)w1);
        atomicAdd(P.stats + role_id * kPipeStatSlots + 9, (unsigned long long)w2); atomicAdd(P.stats + role_id * kPipeStatSlots + 10, (unsigned long long)w3);
        atomicAdd(P.stats + role_id * kPipeStatSlots + 11, (unsigned long long)w4);
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

static unsigned long long* g_pipe_stats = nullptr;     // debug: device buffer [kPipeRoles][kPipeStatSlots], see fnerf_debug_pipe_stats

// the pipeline needs every one of its CTAs resident at the same time, one per SM
bool mlp_bwd_pipe_supported() { return num_sms() >= kPipeRoles * kPipeLanes; }

int64_t mlp_bwd_pipe_workspace_bytes() {
  return kRingTiles * kTileB + 2 * (int64_t)kPipeRings * kPipeLanes * 32 * 4 + 1024;
}

// flat_grad += dL/dparams of an UNCONDITIONED network from its forward tape and g_raw[M,4]
int launch_mlp_bwd_pipe(const void* packed, const float* g_raw, const void* tape, float* flat_grad, void* ws, int64_t M, cudaStream_t s) {
  if (M == 0) return 0;
  const int cond = 0;
  const int64_t ntiles = (M + 127) / 128;
  static DeviceOnce once;
  if (cudaError_t e = opt_in_smem_once(once, k_mlp_bwd_pipe, kPipeSmem)) return set_error((int)e, "mlp_bwd_pipe attr: %s", cudaGetErrorString(e));
  if (num_sms() < kPipeRoles * kPipeLanes) return set_error(FNERF_ERR_ARG, "mlp_bwd_pipe: needs %d co-resident CTAs", kPipeRoles * kPipeLanes);
  PipeParams P = {};
  P.packed = reinterpret_cast<const uint8_t*>(packed); P.cond = cond;
  P.g_raw = reinterpret_cast<const float4*>(g_raw);
  P.fwd_tape = reinterpret_cast<const uint8_t*>(tape);
  P.mask_tape = reinterpret_cast<const uint32_t*>(P.fwd_tape + ntiles * (int64_t)kTapeFwdSlots * kImg);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  P.ring = w8;
  P.flags = reinterpret_cast<uint32_t*>(w8 + kRingTiles * kTileB);
  P.flat_grad = flat_grad; P.M = M; P.ntiles = ntiles;
  P.stats = g_pipe_stats;
  auto gw = [&](int l) { return flat_grad + flat_weight_offset(l, cond); };
  auto gb = [&](int l) { return flat_grad + flat_bias_offset(l, cond); };
  const int in5 = kPE + kW;
  int r = 0;
  auto prod = [](float* dw, int64_t ld_n, int64_t ld_k, int tmem_col, int n_mb, int ncols, int k0, int n_valid) {
    PipeProduct p; p.dw = dw; p.ld_n = ld_n; p.ld_k = ld_k; p.tmem_col = tmem_col; p.n_mb = n_mb; p.ncols = ncols; p.k0 = k0; p.n_valid = n_valid;
    return p;
  };
  // V0a V0b V1a V1b: per half two CTAs that alternate over the lane's tiles (the view roles' per-tile chain -- g_raw image,
  // dZv product, mask, dgrad, epilogue -- is serial and about twice as long as a trunk role's tile)
  for (int hp = 0; hp < 4; ++hp) {
    const int h = hp >> 1;
    PipeRole& R = P.roles[r++];
    R.t0 = hp & 1; R.tstep = 2;
    R.kind = ROLE_V; R.half = h; R.in_ring = -1; R.out_ring = 0; R.wt_chunk0 = 0; R.wt_nchunks = 2;
    R.x_slot = kTapeSlotFeat + 2 * h; R.mask_unit0 = -1; R.rank1 = 0;
    R.bias = gb(10);
    R.prod[0] = prod(gw(10) + 128 * h, kW + kPED, 1, 128, 1, 128, 0, 128);
    if (h == 0) {
      R.e_slot = -1; R.p_slot[0] = kTapeSlotHv; R.p_slot[1] = kTapeSlotH + 28;
      R.prod[1] = prod(gw(11), 1, kWV, 256, 1, 16, 0, 3);                        // rgb_linear.weight[c][n]
    } else {
      R.e_slot = kTapeSlotPed; R.p_slot[0] = -1; R.p_slot[1] = kTapeSlotH + 30;
      R.prod[1] = prod(gw(10) + kW, kW + kPED, 1, 256, 1, 32, 0, kPED);          // views weight, direction columns
    }
    R.prod[2] = prod(gw(8) + 128 * h, 1, 0, 288, 1, 16, 3, 4);                    // alpha_linear.weight[128 h + n] (column 3 = sigma)
    R.nprod = 3;
  }
  // F0 / F1, L7 .. L1
  for (int st = 1; st <= 8; ++st) {
    const int layer = st == 1 ? 9 : 9 - st;                                      // flat layer id: 9 = feature, then 7..1
    for (int h = 0; h < 2; ++h) {
      PipeRole& R = P.roles[r++];
      R.t0 = 0; R.tstep = 1;
      R.kind = ROLE_T; R.half = h; R.in_ring = st - 1; R.out_ring = st;
      R.wt_chunk0 = st == 1 ? 2 : 6 + 4 * (st - 2); R.wt_nchunks = 4;
      const int hsrc = st == 1 ? 7 : layer - 1;                                  // forward activation H_hsrc is the X operand and the mask
      R.x_slot = kTapeSlotH + 4 * hsrc + 2 * h;
      R.mask_unit0 = hsrc * 8 + 4 * h;
      R.rank1 = st == 1;
      R.e_slot = -1; R.p_slot[0] = R.p_slot[1] = -1;
      R.bias = gb(layer);
      const int ld = layer == 5 ? in5 : kW;
      R.prod[0] = prod(gw(layer) + (layer == 5 ? kPE : 0) + 128 * h, ld, 1, 128, 2, 128, 0, 128);
      R.nprod = 1;

    }
  }
  // Z0a Z0b Z5a Z5b: the two products against the xyz encoding, dW0 += dZ0^T . PE (+ bias 0) and dW5[:, 0:63] += dZ5^T . PE,
  // each by two CTAs that alternate over the lane's tiles (a tile costs 80 KB of loads for ~600 cycles of MMAs: latency-bound)
  for (int zp = 0; zp < 4; ++zp) {
    const bool z5 = zp >= 2;
    PipeRole& R = P.roles[r++];
    R.t0 = zp & 1; R.tstep = 2;
    R.kind = ROLE_Z; R.half = 0; R.in_ring = z5 ? kRingZ5 : kPipeRings - 1; R.out_ring = -1; R.wt_chunk0 = 0; R.wt_nchunks = 0; R.x_slot = 0;
    R.mask_unit0 = -1; R.rank1 = 0; R.e_slot = kTapeSlotPe; R.p_slot[0] = R.p_slot[1] = -1;
    R.bias = z5 ? nullptr : gb(0);
    R.prod[0] = z5 ? prod(gw(5), in5, 1, 0, 2, 64, 0, kPE) : prod(gw(0), kPE, 1, 0, 2, 64, 0, kPE);
    R.nprod = 1;
  }
  cudaError_t e = cudaMemsetAsync(P.flags, 0, 2 * (size_t)kPipeRings * kPipeLanes * 32 * 4, s);
  if (e != cudaSuccess) return set_error((int)e, "mlp_bwd_pipe memset: %s", cudaGetErrorString(e));
  void* args[] = {(void*)&P};
  e = cudaLaunchCooperativeKernel((const void*)k_mlp_bwd_pipe, dim3(kPipeRoles * kPipeLanes), dim3(kPipeThreads), args, kPipeSmem, s);
  if (e != cudaSuccess) return set_error((int)e, "mlp_bwd_pipe launch: %s", cudaGetErrorString(e));
  return check_launch("mlp_bwd_pipe");
}

}  // namespace fnerf

// ---- debug entry (tools only): every following pipelined backward adds, per role, the cycles its warps spent waiting
// into stats[role * 8 + slot] (slots: loader ring flags / free slots, MMA operands / accumulators, epilogue accumulators /
// buffers / staging, store warp images).  NULL switches the accounting off.
extern "C" int fnerf_debug_pipe_stats(unsigned long long* stats) {
  fnerf::g_pipe_stats = stats;
  return fnerf::kPipeRoles * fnerf::kPipeStatSlots;
}
