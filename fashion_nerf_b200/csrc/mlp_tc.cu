// Fused bf16 tensor-core network query (A.3 + A.4 + A.8) for sm_100a: tcgen05.mma with TMEM
// accumulators, weights streamed by bulk TMA, activations resident in shared memory.
//
// One persistent CTA per SM walks 128-sample tiles.  Warp roles (576 threads):
//   warp 0      weight producer: one lane streams the 48 pre-swizzled weight chunks of the network
//               (layout.h section A, consumption order) through a ring of 32 KB stages with
//               cp.async.bulk + mbarrier complete_tx;
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=256|128, K=16) for every layer,
//               A = activation / encoding tile in smem (K-major, 128B swizzle), B = weight stage,
//               D = one of two 128x256 fp32 accumulators in TMEM (layer parity);
//   warps 2..17 workers, four groups of four warps (a warp's TMEM lane quadrant is warp%4, so each
//               group covers the 128 rows): group 0 / 1 build the xyz / direction encoding tiles;
//               all groups run the per-layer epilogue in 16-column pieces: tcgen05.ld -> ReLU -> bf16
//               -> swizzled st.shared as the next layer's A operand.  In round kb (0..3) group g converts
//               columns 16g..16g+15 of K-block kb, so all 16 warps finish K-block 0 first.
//   warps 18,19 training forward only (kSave, 640 threads): tape writers -- copy every finished 16 KB activation /
//               encoding image from shared memory to the HBM tape, so that no epilogue warp ever issues a global store
//               ahead of a barrier arrival (the arrival's release waits for the thread's outstanding stores).
// kComp (inference): alpha compositing (A.5) runs in the finishing group, deferred into the next tile (comp_phase).
// The epilogue hands the activation tile over in 64-column K-blocks (one mbarrier each, 16 arrivals),
// so the next layer's MMAs on K-block 0 start while K-blocks 1..3 are still being converted; the
// two TMEM accumulators make that overlap legal.
//
// Biases ride in the GEMM: the direction-encoding tile carries 1.0 in its two spare columns (27, 28)
// and every layer ends with one extra K=16 MMA of that tile's K-step 1 against a BIAS chunk holding
// bf16(b) and bf16(b - bf16(b)) in those columns (zeros elsewhere), so the fp32 accumulator already
// contains the bias (to ~16 mantissa bits) and the epilogue is a single F2FP.RELU per two elements.
// History (profiles/r1_microbench.md): 4 epilogue warps -> 43 % tensor-pipe (one latency-bound warp
// per scheduler); 16 warps with an fp32 bias add -> 48-52 % (FADD + F2FP issue bound, ~2250 cycles
// per layer against 2048 MMA cycles).
//
// Encoded samples and activations never touch HBM: per sample the kernel reads 4 B (z) and writes
// 16 B (raw).  sigma (256->1) and rgb (128->3) are fp32 dot products inside the epilogues of layer 7
// and of the view layer.
#include <stdlib.h>
#include "common.cuh"
#include "composite_math.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kDefaultMlpCluster = 1;     // render kernel as single CTAs (1) or CTA pairs (2, see k_mlp_tc); FNERF_MLP_CLUSTER overrides
constexpr int kStages = 3;
constexpr int kTcThreads = 576;
constexpr int kWorkerThreads = 512;
#ifdef EXP_TAPE_NOAUX
constexpr bool kTapeAux = false;          // experiment: no mask / encoding / view-layer images on the tape (timing only)
#else
constexpr bool kTapeAux = true;
#endif
constexpr int kTapeWarps = 2;             // training forward only: warps that copy finished activation images to the tape
constexpr uint32_t kKBlockBytes = kTileM * 128;                 // 16 KB: 128 rows x 64 bf16
constexpr uint32_t kOffAct = 0;                                 // 4 K-blocks
constexpr uint32_t kOffPe = 4 * kKBlockBytes;                   // xyz encoding (63 -> 64), two buffers (tile parity)
constexpr uint32_t kOffPed = kOffPe + 2 * kKBlockBytes;         // direction encoding (27) + ones (cols 27, 28) for the view layer's DIR chunk
constexpr uint32_t kOffOnes = kOffPed + kKBlockBytes;           // constant K = 16 A tile of the BIAS MMAs: 128 x 16, MN-major, ones in K rows 11, 12
constexpr uint32_t kOnesBytes = 128 * 16 * 2;
constexpr uint32_t kOffW = kOffOnes + kOnesBytes;               // weight stages
constexpr uint32_t kOffHeads = kOffW + kStages * kBigChunkBytes;  // fp32 head weights (aux from kAuxWAlpha on)
constexpr int kHeadFloats = kAuxFloats - kAuxWAlpha;
constexpr uint32_t kOffBar = kOffHeads + kHeadFloats * 4;
constexpr uint32_t kNumBars = 3 * (2 * kStages) + 4 + 1 + 2;   // room for the pair mode: 2*kStages x (full, empty, peer-weights) + act-ready, encodings, acc-full
constexpr uint32_t kOffComp = kOffBar + kNumBars * 8 + 16;   // fused compositing (kComp): block products + tile carry, running sums, block operands
constexpr uint32_t kCompFloats = 32 + 5 * 32 + 4 * 6 * 32;
constexpr uint32_t kTcSmemBytes = kOffComp + kCompFloats * 4 + 1024;  // + alignment slack
static_assert(kOffBar % 8 == 0, "barrier alignment");
static_assert(kTcSmemBytes <= 227 * 1024, "shared memory budget");

// Optional timeline tracer (compile with -DFNERF_TRACE, see tools/trace_tc.py): CTA 0 stamps clock64()
// at pipeline events of its 3rd tile into a global buffer.  Compiled out of the product library.
#ifdef FNERF_TRACE
__device__ long long* g_trace_buf = nullptr;
#define FN_TRACE(cond, slot) do { if (blockIdx.x == 0 && (cond) && g_trace_buf) g_trace_buf[(slot)] = clock64(); } while (0)
#else
#define FN_TRACE(cond, slot) do { } while (0)
#endif

// The MMA schedule of one tile, derived from the packed chunk order (layout.h chunk_desc).
struct MmaChunk {
  int a_sel;       // 0 = activation K-block kb, 1 = xyz tile, 2 = direction tile, 3 = constant ones tile (MN-major)
  int kb;          // activation K-block (also the act-ready barrier to honour when `gated`)
  int kstep0;      // first K=16 step inside the A tile / weight chunk
  int ksteps;      // number of K=16 MMAs
  int n128;        // 1: N = 128 (view layer), 0: N = 256
  int acc;         // TMEM accumulator 0 / 1
  int fresh;       // first MMA overwrites the accumulator
  int gated;       // wait for the epilogue's act-ready[kb]
  int commit_acc;  // last chunk of a layer: commit to acc-full[acc]
  int bias;        // B = compact MN-major BIAS chunk (one K-step at the start of the stage)
};
__host__ __device__ constexpr int layer_step(int flat_layer) { return flat_layer < 8 ? flat_layer : (flat_layer == 9 ? 8 : 9); }
__host__ __device__ constexpr MmaChunk mma_chunk(int c) {
  const ChunkDesc d = chunk_desc(c);
  const int acc = layer_step(d.layer) & 1;
  const bool first = c == 0 || chunk_desc(c - 1).layer != d.layer;
  const bool last = c == kNumChunks - 1 || chunk_desc(c + 1).layer != d.layer;
  if (d.kind == CHUNK_TRUNK) return {0, d.kb, 0, 4, d.layer == 10, acc, first, 1, last, 0};
  if (d.kind == CHUNK_XYZ) return {1, 0, 0, 4, 0, acc, first, 0, last, 0};
  if (d.kind == CHUNK_BIAS) return {3, 0, 0, 1, 0, acc, first, 0, last, 1};  // A: the constant ones tile
  return {2, 0, 0, 2, 1, acc, first, 0, last, 0};                           // CHUNK_DIR: cols 0..31
}

struct TcParams {
  const uint8_t* packed;
  const float* rays_o; const float* rays_d; const float* viewdirs; const float* z;
  const float* cond_proj; const int32_t* cond_index; int64_t C;
  float4* raw;
  int64_t M; int S; int64_t ntiles; int cond;
  uint8_t* tape;        // kSave only: per tile kTapeFwdSlots K-block images (layout.h), the shared-memory bytes verbatim
  uint32_t* mask_tape;  // kSave only: per tile kMaskUnits x 128 ReLU bitmask words (layout.h)
  // kComp only: alpha compositing (A.5) fused into the last epilogue.  Tiles are walked in GROUPS of `gt` consecutive
  // tiles that hold whole rays (gt * 128 = rays_per_group * S), so transmittance and the ray sums never leave the CTA.
  const float* dnorm; const float* noise;          // [R]; [R,S] nullable
  float* rgb_map; float* depth_map; float* acc_map; float* disp_map;   // [R,3], [R], [R], [R]
  float* weights;                                   // [R,S] nullable
  int gt; int rays_per_group; int white;
};

// two non-negative bf16 in one word -> bit 0 = (low half != 0), bit 16 = (high half != 0): adding 0x7FFF to a
// non-negative bf16 (<= 0x7F80) sets bit 15 exactly when it is non-zero and never carries into the other half
__device__ __forceinline__ uint32_t relu_bits(uint32_t packed) { return ((packed + 0x7FFF7FFFu) >> 15) & 0x00010001u; }

// epilogue of one 16-column piece (two 16-byte chunks of a K-block row): TMEM -> (+rowbias) -> [ReLU] -> bf16 -> swizzled
// smem.  The layer bias is already in the accumulator.  `chunk0` is the index (0, 2, 4, 6) of the piece's first 16-byte
// chunk inside its 128-byte K-block row.  Returns the ReLU bits of the piece (training tape): bit j = column 2j > 0,
// bit 16 + j = column 2j + 1 > 0, j = 0..7.
template <bool kRelu, bool kSigma, bool kCond>
__device__ __forceinline__ uint32_t epilogue_piece(uint32_t taddr, const float* __restrict__ walpha_s, const float* __restrict__ rowbias,
                                                   uint32_t act_row_addr, uint32_t chunk0, uint32_t row, float& sigma) {
#ifdef EXP_NOEPI
  return 0u;   // experiment: how fast is the kernel when the epilogue costs nothing (results are garbage)
#endif
  uint32_t v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
#ifdef EXP_LDONLY
  if (v[0] != 0x12345678u) return 0u;   // experiment: TMEM read-out only
#endif
  uint32_t mask = 0u;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[c * 8 + j]);
    const int col = c * 8;
    if (kCond) {
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(rowbias + col));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(rowbias + col + 4));
      x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
      x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
    }
    if (kSigma) {
      const float4 w0 = *reinterpret_cast<const float4*>(walpha_s + col);
      const float4 w1 = *reinterpret_cast<const float4*>(walpha_s + col + 4);
      sigma = fmaf(fmaxf(x[0], 0.f), w0.x, sigma); sigma = fmaf(fmaxf(x[1], 0.f), w0.y, sigma);
      sigma = fmaf(fmaxf(x[2], 0.f), w0.z, sigma); sigma = fmaf(fmaxf(x[3], 0.f), w0.w, sigma);
      sigma = fmaf(fmaxf(x[4], 0.f), w1.x, sigma); sigma = fmaf(fmaxf(x[5], 0.f), w1.y, sigma);
      sigma = fmaf(fmaxf(x[6], 0.f), w1.z, sigma); sigma = fmaf(fmaxf(x[7], 0.f), w1.w, sigma);
    }
    uint32_t p0, p1, p2, p3;
    if (kRelu) {
      p0 = pack_bf16_relu(x[0], x[1]); p1 = pack_bf16_relu(x[2], x[3]);
      p2 = pack_bf16_relu(x[4], x[5]); p3 = pack_bf16_relu(x[6], x[7]);
    } else {
      p0 = pack_bf16(x[0], x[1]); p1 = pack_bf16(x[2], x[3]);
      p2 = pack_bf16(x[4], x[5]); p3 = pack_bf16(x[6], x[7]);
    }
    const uint32_t c16 = chunk0 + (uint32_t)c;
    st_shared_v4(act_row_addr + ((c16 ^ (row & 7u)) << 4), p0, p1, p2, p3);
    if (kRelu) mask |= relu_bits(p0) << (4 * c) | relu_bits(p1) << (4 * c + 1) | relu_bits(p2) << (4 * c + 2) | relu_bits(p3) << (4 * c + 3);
  }
  return mask;
}

__device__ __forceinline__ void worker_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory"); }

// kCl == 2: CTA pair (thread-block cluster of 2, tcgen05 cta_group::2).  The pair runs two 128-sample tiles in
// lock step as ONE M = 256 MMA stream issued by the leader (cluster rank 0): every CTA keeps its own tile (A operand,
// TMEM accumulators, epilogue) but streams only ITS HALF of every weight chunk (N/2 rows of B); the tensor cores of the
// two SMs exchange the halves.  Per SM that halves the bytes written into shared memory by the weight stream and the
// bytes of B read back by the MMAs -- the two costs that A/B builds showed to limit the single-CTA kernel (streaming
// half of each chunk: +6-8 %, none: +12 %).  Handshakes: both CTAs' epilogue / encoding warps arrive on the LEADER's
// act-ready and encoding barriers (mbarrier.arrive.release.cluster on mapa addresses); the peer's otherwise idle MMA warp
// relays "my half of stage s has landed" to the leader; tcgen05.commit multicasts stage-empty and accumulator-full to
// both CTAs.  Both CTAs run the same number of tiles (a tile past the end is computed on clamped inputs, not stored).
template <bool kSave, int kCl, bool kComp = false>
__global__ void __launch_bounds__(kSave ? kTcThreads + 32 * kTapeWarps : kTcThreads, 1) k_mlp_tc(const TcParams P) {
  static_assert(!kSave || kCl == 1, "the training forward runs as single CTAs");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  // pair mode streams half chunks: twice the stages of half the size cover the longer refill round trip
  // (leader commit -> peer's empty barrier -> load -> peer's relay -> leader)
  constexpr int kSt = kCl > 1 ? 2 * kStages : kStages;
  constexpr uint32_t kStageBytes = kCl > 1 ? kBigChunkBytes / 2 : kBigChunkBytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kSt + s); };
  auto bar_act = [&](int kb) { return bar0 + 8u * (2 * kSt + kb); };
  const uint32_t bar_pe = bar0 + 8u * (2 * kSt + 4);
  auto bar_acc = [&](int a) { return bar0 + 8u * (2 * kSt + 5 + a); };
  auto bar_wpeer = [&](int s) { return bar0 + 8u * (2 * kSt + 7 + s); };   // pair: the peer's half of stage s has landed
  [[maybe_unused]] auto bar_taped = [&](int kb) { return bar0 + 8u * (3 * kSt + 7 + kb); };   // kSave: K-block kb's image has been read for the tape
  [[maybe_unused]] const uint32_t bar_hv = bar0 + 8u * (3 * kSt + 11);                        // kSave: the view layer's output images are in K-blocks 0, 1
  const uint32_t tmem_slot = bar0 + 8u * kNumBars;
  float* heads_s = reinterpret_cast<float*>(base_ptr + kOffHeads);   // index with (kAux* - kAuxWAlpha)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ---------------------------------------------------------------------------
  {
    const float* aux_g = reinterpret_cast<const float*>(P.packed + kSecBOffset) + kAuxWAlpha;
    for (int i = threadIdx.x; i < kHeadFloats; i += blockDim.x) heads_s[i] = aux_g[i];
  }
  // constant A operand of the BIAS MMAs: ones in K rows 11 and 12 (= columns 27, 28 of a K-step-1 view), zeros elsewhere
  for (int e = threadIdx.x; e < 128 * 16; e += blockDim.x) {
    const uint32_t m = (uint32_t)e >> 4, k = (uint32_t)e & 15u;
    *reinterpret_cast<unsigned short*>(base_ptr + kOffOnes + bias_chunk_offset(m, k)) =
        (k == (uint32_t)(kBiasColHi - 16) || k == (uint32_t)(kBiasColLo - 16)) ? (unsigned short)0x3F80 : (unsigned short)0;
  }
  if (kComp && threadIdx.x < 32 + 5 * 32) reinterpret_cast<float*>(base_ptr + kOffComp)[threadIdx.x] = threadIdx.x == 4 ? 1.0f : 0.0f;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSt; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); mbar_init(bar_wpeer(s), 1); }
    // one arrive per warp: every K-block is handed over by all 16 worker warps (16-column pieces; x CTAs)
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar_act(kb), 16 * kCl);
    mbar_init(bar_pe, 128 * kCl);                                     // xyz-encoding group (x CTAs)
    mbar_init(bar_acc(0), 1);
    mbar_init(bar_acc(1), 1);
    if (kSave) {
      for (int kb = 0; kb < 4; ++kb) mbar_init(bar_taped(kb), 1);
      mbar_init(bar_hv, kWorkerThreads / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCl > 1) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (kCl > 1) cluster_sync();               // no peer may arrive on barriers that are not initialised yet
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBar + 8 * kNumBars);

  const int64_t first_tile = blockIdx.x;
  const int64_t tile_stride = gridDim.x;
  // every CTA runs n_iter tiles (identical across a cluster); without clusters that is exactly the tiles it owns
  int64_t n_iter = kCl > 1 ? (P.ntiles + tile_stride - 1) / tile_stride
                           : (P.ntiles > first_tile ? (P.ntiles - first_tile + tile_stride - 1) / tile_stride : 0);
  // kComp: the CTA owns whole groups of gt consecutive tiles (groups blockIdx.x, + gridDim.x, ...); only the last group of
  // the launch can be short
  const int64_t gt = kComp ? P.gt : 1;
  if (kComp) {
    const int64_t ngroups = (P.ntiles + gt - 1) / gt;
    const int64_t mine = ngroups > first_tile ? (ngroups - first_tile + tile_stride - 1) / tile_stride : 0;
    n_iter = mine * gt;
    if (mine > 0 && first_tile + (mine - 1) * tile_stride == ngroups - 1) n_iter -= ngroups * gt - P.ntiles;
  }
  auto tile_of = [&](int64_t it) -> int64_t {
    if (!kComp) return first_tile + it * tile_stride;
    return (first_tile + (it / gt) * tile_stride) * gt + it % gt;
  };
  [[maybe_unused]] const uint32_t cta_rank = kCl > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (uint16_t)((1u << kCl) - 1u);
  // pair mode: act-ready / encoding arrivals of BOTH CTAs go to the leader's barriers
  [[maybe_unused]] const uint32_t lead_bar_act0 = kCl > 1 ? mapa_shared(bar_act(0), 0) : 0u;
  [[maybe_unused]] const uint32_t lead_bar_pe = kCl > 1 ? mapa_shared(bar_pe, 0) : 0u;

  if (warp == 0) {
    // ================================ weight producer ==============================================
    if (lane == 0) {
      uint32_t wc = 0;
      for (int64_t it = 0; it < n_iter; ++it) {
        [[maybe_unused]] const int64_t tile = tile_of(it);
        uint32_t coff = 0;                       // chunk_offset(c), accumulated
        for (int c = 0; c < kNumChunks; ++c, ++wc) {
          const uint32_t s = wc % kSt;
          const uint32_t bytes = (uint32_t)chunk_bytes(c);
          const uint8_t* csrc = P.packed + coff;
          coff += bytes;
          mbar_wait(bar_empty(s), ((wc / kSt) & 1u) ^ 1u);
#ifdef EXP_NOTMA
          // experiment (tools/ab_tc.py): weights are streamed for the CTA's first tile only, afterwards the stages keep
          // those (real, but stale) chunks -- bounds what removing the L2 -> SM weight traffic could buy
          if (tile != first_tile) { mbar_arrive(bar_full(s)); continue; }
#endif
#ifdef EXP_HALFW
          // experiment: only the first half of every chunk is streamed (the MMAs still read the whole, half-stale stage)
          mbar_expect_tx(bar_full(s), bytes / 2);
          bulk_g2s(base + kOffW + s * kStageBytes, csrc, bytes / 2, bar_full(s));
#else
          if (kCl > 1) {                          // this CTA's N/2 rows of the chunk, at the start of the stage
            const uint32_t part = bytes / 2;
            mbar_expect_tx(bar_full(s), part);
            bulk_g2s(base + kOffW + s * kStageBytes, csrc + cta_rank * part, part, bar_full(s));
          } else {
            mbar_expect_tx(bar_full(s), bytes);
            bulk_g2s(base + kOffW + s * kStageBytes, csrc, bytes, bar_full(s));
          }
#endif
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================================
    // The issuing lane must stay ahead of the tensor pipe: one 64-wide K chunk is 4 MMAs = 512 pipe
    // cycles (N=256).  A timeline trace of the first version showed ~850 issue-side cycles per chunk
    // (two serial mbarrier probes at ~150 cycles each, descriptor rebuilds, ~100 cycles per MMA), i.e.
    // the issuer, not the pipe, paced the kernel.  Hence: the chunk schedule is a compile-time table
    // (fully unrolled), descriptors advance by 64-bit adds, and the barriers of chunk c+1 are probed
    // (non-blocking test_wait) before chunk c's MMAs are issued so the probe latency hides behind them.
    // The loop runs warp-uniformly on all 32 lanes and only the MMA / commit instructions are
    // predicated on one elected lane: with a divergent `if (lane == 0)` around the loop ptxas cannot
    // prove the descriptors uniform and wraps EVERY tcgen05.mma in an elect/R2UR waterfall loop.
    if (kCl > 1 && cta_rank != 0) {
      // peer of a CTA pair: no MMAs to issue; tell the leader when this CTA's half of a weight stage has landed
      const uint32_t lead_wpeer0 = mapa_shared(bar_wpeer(0), 0);
      uint32_t wc = 0;
      for (int64_t it = 0; it < n_iter; ++it) {
        for (int c = 0; c < kNumChunks; ++c, ++wc) {
          const uint32_t s = wc % kSt;
          mbar_wait(bar_full(s), (wc / kSt) & 1u);
          if (lane == 0) mbar_arrive_cluster(lead_wpeer0 + 8u * s);
          __syncwarp();
        }
      }
    } else {
      constexpr uint32_t idesc256 = umma_idesc_bf16(128 * kCl, 256);
      constexpr uint32_t idesc128 = umma_idesc_bf16(128 * kCl, 128);
      constexpr uint32_t idesc256b = umma_idesc_bf16(128 * kCl, 256) | (1u << 15) | (1u << 16);   // BIAS MMAs: A (ones) and B MN-major
      uint32_t wc = 0, act_cnt = 0, tile_cnt = 0;
      [[maybe_unused]] uint32_t tslot = 0;   // tracer: slots [0,256): (before waits, operands ready, issued) per chunk
      const uint64_t desc_act = umma_desc_sw128(base + kOffAct);
      const uint64_t desc_pe = umma_desc_sw128(base + kOffPe);
      const uint64_t desc_ped = umma_desc_sw128(base + kOffPed);
      const uint64_t desc_w = umma_desc_sw128(base + kOffW);
      const uint64_t desc_wb = umma_desc_mn_sw128(base + kOffW, 2048);
      const uint64_t desc_ones = umma_desc_mn_sw128(base + kOffOnes, 2048);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform
      const uint32_t acc_addr[2] = {tmem_u, tmem_u + 256};
      for (int64_t it = 0; it < n_iter; ++it, ++tile_cnt) {
        mbar_wait(bar_pe, tile_cnt & 1u);
        bool w_ready = __all_sync(0xffffffffu, mbar_test_wait(bar_full(wc % kSt), (wc / kSt) & 1u) &&
                                                   (kCl == 1 || mbar_test_wait(bar_wpeer(wc % kSt), (wc / kSt) & 1u)));
        bool a_ready = true;
#pragma unroll
        for (int c = 0; c < kNumChunks; ++c) {
          const MmaChunk op = mma_chunk(c);
          const uint32_t s = wc % kSt;
          FN_TRACE(tile_cnt == 2 && lane == 0, tslot++);
          if (op.gated && !a_ready) mbar_wait(bar_act(op.kb), act_cnt & 1u);
          if (!w_ready) {
            mbar_wait(bar_full(s), (wc / kSt) & 1u);
            if (kCl > 1) mbar_wait(bar_wpeer(s), (wc / kSt) & 1u);
          }
          tc_fence_after();
          FN_TRACE(tile_cnt == 2 && lane == 0, tslot++);
          const uint64_t a_desc = op.a_sel == 3 ? desc_ones
                                  : (op.a_sel == 0 ? desc_act + (uint64_t)(op.kb * (kKBlockBytes >> 4))
                                                   : (op.a_sel == 1 ? desc_pe + (uint64_t)((tile_cnt & 1u) * (kKBlockBytes >> 4)) : desc_ped)) +
                                        (uint64_t)(2 * op.kstep0);
          // BIAS chunk: MN-major K = 16 tile at the start of the stage (4 groups of 64 outputs, 2 KB apart; per CTA of a pair: 2)
          const uint64_t b_desc = op.bias ? desc_wb + (uint64_t)(s * (kStageBytes >> 4))
                                          : desc_w + (uint64_t)(s * (kStageBytes >> 4)) + (uint64_t)(2 * op.kstep0);
          const uint32_t idesc = op.bias ? idesc256b : (op.n128 ? idesc128 : idesc256);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < op.ksteps; ++ks)
#ifdef EXP_NOMMA
              if (P.M < 0)
#endif
              if (kCl > 1) umma_bf16_pair(acc_addr[op.acc], a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                                          (op.fresh && ks == 0) ? 0u : 1u);
              else umma_bf16(acc_addr[op.acc], a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                             (op.fresh && ks == 0) ? 0u : 1u);
            if (kCl > 1) {
              umma_commit_pair(bar_empty(s), kClMask);
              if (op.commit_acc) umma_commit_pair(bar_acc(op.acc), kClMask);
            } else {
              umma_commit(bar_empty(s));
              if (op.commit_acc) umma_commit(bar_acc(op.acc));
            }
          }
          __syncwarp();
          // probe the next chunk's barriers AFTER this chunk's MMAs are in the queue: the probes (two test_wait round trips and
          // two votes) cost ~100 cycles that used to sit between an act-ready wake-up and the first MMA of the chunk
          const uint32_t act_next = act_cnt + ((op.gated && op.kb == 3) ? 1u : 0u);
          w_ready = __all_sync(0xffffffffu, mbar_test_wait(bar_full((wc + 1) % kSt), ((wc + 1) / kSt) & 1u) &&
                                                (kCl == 1 || mbar_test_wait(bar_wpeer((wc + 1) % kSt), ((wc + 1) / kSt) & 1u)));
          if (c + 1 < kNumChunks) {
            const MmaChunk nx = mma_chunk(c + 1);
            a_ready = nx.gated ? __all_sync(0xffffffffu, mbar_test_wait(bar_act(nx.kb), act_next & 1u)) : true;
          } else {
            a_ready = true;
          }
          FN_TRACE(tile_cnt == 2 && lane == 0, tslot++);
          act_cnt = act_next;
          ++wc;
        }
      }
    }
  } else if (kSave && warp >= 2 + kWorkerThreads / 32) {
    // ================================ tape writers (training forward) ==============================
    // Every finished activation K-block image (16 KB, the shared-memory bytes verbatim) goes to the tape.  Global stores
    // stall their issuer (an SM drains ~45 B/clk into L2), so they must not be issued by the epilogue warps, whose work
    // is on the layer-to-layer critical path: two extra warps wait for act-ready like the MMA issuer does, copy the image
    // linearly (512 contiguous bytes per instruction) and tell the epilogue that the K-block may be overwritten.
    const uint32_t sw = (uint32_t)warp - (2 + kWorkerThreads / 32);      // 0: K-blocks 0, 2, xyz tile;  1: K-blocks 1, 3, direction tile
    uint32_t ph = 0;                                                      // act-ready phases consumed: one per step
    // 16 KB image: shared memory -> tape slot, 512 contiguous bytes per instruction
    auto copy_image = [&](uint32_t src_img, uint8_t* dst_img) {
#ifdef EXP_TAPE_BULK
      // experiment: one bulk-TMA store per image (the writers fenced their stores for the async proxy); returns once the
      // engine has READ the image
      if (lane == 0) {
        bulk_s2g(dst_img, src_img, kKBlockBytes);
        bulk_commit();
        bulk_wait_read<0>();
      }
      __syncwarp();
      return;
#endif
#ifdef EXP_TAPE_NOLD
      if (P.M >= 0) { __syncwarp(); return; }     // experiment: the hand-shake alone, no shared-memory read-out, no stores
#endif
      const uint32_t src = src_img + ((uint32_t)lane << 4);
      uint8_t* dstg = dst_img + ((uint32_t)lane << 4);
#if defined(EXP_TAPE_LDONLY)
      // experiment: read the image out of shared memory, store nothing (one guard at the end keeps the loads alive)
      uint32_t acc_x = 0u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint4 t4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t4[j] = ld_shared_v4(src + (uint32_t)(b * 8 + j) * 512u);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_x ^= t4[j].x ^ t4[j].y ^ t4[j].z ^ t4[j].w;
      }
      if (acc_x == 0x12345678u && P.M < 0) *reinterpret_cast<uint32_t*>(dstg) = acc_x;
#elif defined(EXP_TAPE_STONLY)
      // experiment: store the image bytes' worth of a register pattern, read nothing from shared memory
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dstg + (b * 8 + j) * 512), "r"(src), "r"(src), "r"(src), "r"(src) : "memory");
#else
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint4 t4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t4[j] = ld_shared_v4(src + (uint32_t)(b * 8 + j) * 512u);
#pragma unroll
        for (int j = 0; j < 8; ++j)     // streaming (evict-first) stores: the tape is next read by the backward, gigabytes later
#ifdef EXP_TAPE_NOSTG
          if (t4[j].x == 0x12345678u && P.M < 0)
#endif
          asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dstg + (b * 8 + j) * 512), "r"(t4[j].x), "r"(t4[j].y), "r"(t4[j].z), "r"(t4[j].w) : "memory");
      }
#endif
      __syncwarp();                      // every lane's loads have returned (their stores consumed them)
    };
    // A waiter must observe a barrier phase before the NEXT phase of that barrier can complete (parity waits cannot tell
    // phases two apart).  For act-ready[kb] and the view-layer barrier that holds because the next write of the K-block
    // waits for this warp's `taped` arrival; the encoding tiles have no such hand-shake of their own, so warp 0 copies
    // them inside hand-shakes that do order them: the direction tile together with K-block 0 of step 0 (group 1 builds it
    // before its first arrival on act-ready[0]), and the NEXT tile's xyz-ready phase is observed before the last `taped`
    // arrival of this tile (the phase after it needs group 0 to get past the next tile's first epilogue).
    if (sw == 0 && kTapeAux && n_iter > 0) mbar_wait(bar_pe, 0u);
    for (int64_t it = 0; it < n_iter; ++it) {
      uint8_t* tape_tile = P.tape + (size_t)tile_of(it) * kTapeFwdSlots * kKBlockBytes;
      if (sw == 0 && kTapeAux)      // xyz tile (built a tile ahead; its buffer is rewritten two tiles later)
        copy_image(base + kOffPe + ((uint32_t)it & 1u) * kKBlockBytes, tape_tile + (size_t)kTapeSlotPe * kKBlockBytes);
#pragma unroll 1
      for (int step = 0; step < 9; ++step, ++ph) {
#pragma unroll 1
        for (uint32_t r = 0; r < 2; ++r) {
          const uint32_t kb = sw + 2u * r;
          mbar_wait(bar_act(kb), ph & 1u);
          copy_image(base + kOffAct + kb * kKBlockBytes, tape_tile + (size_t)(kTapeSlotH + 4 * step + (int)kb) * kKBlockBytes);
          if (step == 0 && kb == 0 && kTapeAux) copy_image(base + kOffPed, tape_tile + (size_t)kTapeSlotPed * kKBlockBytes);
          if (lane == 0) mbar_arrive(bar_taped(kb));
        }
      }
      // view-layer output: the epilogue parks its two images in K-blocks 0, 1 (free once that layer's MMAs are done)
      mbar_wait(bar_hv, (uint32_t)it & 1u);
      copy_image(base + kOffAct + sw * kKBlockBytes, tape_tile + (size_t)(kTapeSlotHv + (int)sw) * kKBlockBytes);
      if (sw == 0 && kTapeAux && it + 1 < n_iter) mbar_wait(bar_pe, (uint32_t)(it + 1) & 1u);
      if (lane == 0) mbar_arrive(bar_taped(sw));
    }
  } else {
    // ================================ workers: encodings + epilogues ===============================
    const uint32_t q = (uint32_t)warp & 3u;                 // TMEM lane quadrant this warp may read
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;         // 0..3
    const uint32_t row = q * 32u + (uint32_t)lane;
    const uint32_t tmem_row = tmem_base + ((q * 32u) << 16);
    uint32_t acc_cnt[2] = {0u, 0u};
    [[maybe_unused]] uint32_t wtile = 0, wslot = 256 + grp * 64;   // tracer: 64 slots per group from 256
    const uint32_t act_row = base + kOffAct + row * 128u;
    uint8_t* ped_row_ptr = base_ptr + kOffPed + row * 128u;
    // xyz encoding of this thread's row of tile `tile_n` into xyz buffer `buf` (group 0): sincos once per coordinate,
    // then the double-angle recurrence per octave; arrives on the encoding barrier
    auto compute_pe = [&](int64_t tile_n, uint32_t buf) {
      const int64_t gn = tile_n * kTileM + row;
      const int64_t gcn = gn < P.M ? gn : P.M - 1;
      const int64_t rayn = gcn / P.S;
      const float zv = P.z[gcn];
      float f[64];
      f[63] = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p = __fadd_rn(P.rays_o[3 * rayn + c], __fmul_rn(P.rays_d[3 * rayn + c], zv));
        f[c] = p;
        float sn, cs;
        sincosf(p, &sn, &cs);
#pragma unroll
        for (int k = 0; k < kLX; ++k) {
          f[3 + 6 * k + c] = sn;
          f[3 + 6 * k + 3 + c] = cs;
          const float s2 = 2.0f * sn * cs;
          cs = 1.0f - 2.0f * sn * sn;
          sn = s2;
        }
      }
      const uint32_t pe_row = base + kOffPe + buf * kKBlockBytes + row * 128u;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint4 pk = make_uint4(pack_bf16(f[c16 * 8 + 0], f[c16 * 8 + 1]), pack_bf16(f[c16 * 8 + 2], f[c16 * 8 + 3]),
                                    pack_bf16(f[c16 * 8 + 4], f[c16 * 8 + 5]), pack_bf16(f[c16 * 8 + 6], f[c16 * 8 + 7]));
        st_shared_v4(pe_row + (((uint32_t)c16 ^ (row & 7u)) << 4), pk.x, pk.y, pk.z, pk.w);
      }
      fence_proxy_async_smem();
      if (kCl > 1) mbar_arrive_cluster(lead_bar_pe);
      else mbar_arrive(bar_pe);
    };
    // ---- alpha compositing (A.5) of the PREVIOUS tile's 128 samples (kComp; finishing group only) --------------------
    // S % 32 == 0, so a warp holds one 32-sample block of ONE ray: exactly the unit the stand-alone kernel (composite.cu)
    // walks.  Same per-sample arithmetic (composite_math.cuh), same order -- the transmittance entering a block is the
    // sequential product of the blocks before it, every lane adds its samples block after block, one warp reduction per
    // ray -- so the maps and weights carry the stand-alone kernel's bits whatever the alignment of rays to tiles.  A ray
    // that spans tiles hands over through shared memory.  The work is cut into three short phases that run after the
    // group's first three layer epilogues of the next tile, where the group would otherwise wait for an accumulator:
    // done in one piece at the end of its own tile it delays the next tile's first layers (+3 % per launch, measured).
    [[maybe_unused]] auto comp_phase = [&](int ph) {
      float* cs = reinterpret_cast<float*>(base_ptr + kOffComp);
      float* s_P = cs;                       // [4]  product of (1 - alpha + 1e-10) over each warp's block
      float* s_carry = cs + 4;               // [2]  transmittance entering the tile (ray continued from the one before); its successor
      float* s_run = cs + 32;                // [5][32] per-lane sums of that ray so far
      float* s_ops = cs + 32 + 5 * 32;       // [4][6][32] the tile's blocks: alpha -> w, sigmoid(r,g,b), z, inclusive scan per lane
      asm volatile("bar.sync 4, 128;" ::: "memory");
      const int S = P.S;
      const int64_t t0 = reinterpret_cast<const long long*>(cs + 8)[0] * (int64_t)kTileM;
      const int k0 = reinterpret_cast<const int*>(cs + 6)[0];           // sample index in its ray of the tile's first row
      const int64_t pray = reinterpret_cast<const long long*>(cs + 10)[q];
      auto blk_k = [&](uint32_t w) { return (int)((uint32_t)(k0 + 32 * (int)w) % (uint32_t)S); };
      const int ck = blk_k(q);
      const bool bvalid = t0 + 32 * q < P.M;                            // M % 32 == 0: a block is valid as a whole
      const bool btail = ck + 32 == S;
      const int64_t pg = t0 + row;
      float* o = s_ops + q * (6 * 32) + lane;
      if (ph == 1) {
        const int64_t pgc = bvalid ? pg : P.M - 1;
        const float zv = P.z[pgc];
        float z_up = __shfl_down_sync(0xffffffffu, zv, 1);
        if (lane == 31 && !btail) z_up = P.z[pgc + 1 < P.M ? pgc + 1 : pgc];
        float dist = (lane == 31 && btail) ? 1e10f : (z_up - zv);
        dist *= P.dnorm[pray];
        const float4 fin = *reinterpret_cast<const float4*>(ped_row_ptr + ((4u ^ (row & 7u)) << 4));
        float sgm = fin.w;
        if (P.noise != nullptr) sgm += P.noise[pgc];
        const float alpha = comp_alpha(sgm, dist, bvalid);
        const float pin = comp_scan(alpha, bvalid, (int)lane);
        if (lane == 31) s_P[q] = pin;
        o[0] = alpha; o[32] = sigmoidf_(fin.x); o[64] = sigmoidf_(fin.y); o[96] = sigmoidf_(fin.z); o[128] = zv; o[160] = pin;
      } else if (ph == 2) {
        // transmittance entering this block: the chain restarts at a ray head, else continues the previous block's
        float carry = (k0 == 0) ? 1.0f : s_carry[0];
        for (uint32_t w2 = 0; w2 < q; ++w2) {
          carry *= s_P[w2];
          if (blk_k(w2 + 1) == 0) carry = 1.0f;
        }
        const float pin = o[160];
        const float w = comp_weight(o[0], pin, carry, (int)lane);
        if (P.weights != nullptr && bvalid) P.weights[pg] = w;
        o[0] = w;
        if (q == 3 && lane == 31) s_carry[1] = carry * pin;             // becomes s_carry[0] once every warp has read the old one
      } else {
        // the warp holding a ray's last block in this tile adds up the ray's blocks of this tile, in order, on top of
        // what earlier tiles left; it finishes the ray or leaves the sums for the next tile
        if (bvalid && (btail || q == 3)) {
          int first = (int)q - ck / 32;                                 // block of this tile where the ray starts (< 0: earlier tile)
          CompSums a = {0.f, 0.f, 0.f, 0.f, 0.f};
          if (first < 0) { first = 0; a.r = s_run[lane]; a.g = s_run[32 + lane]; a.b = s_run[64 + lane]; a.d = s_run[96 + lane]; a.w = s_run[128 + lane]; }
          for (int w2 = first; w2 <= (int)q; ++w2) {
            const float* ow = s_ops + w2 * (6 * 32) + lane;
            comp_accum(a, ow[0], ow[32], ow[64], ow[96], ow[128]);
          }
          if (btail) comp_finish(a, (int)lane, pray, P.white, P.rgb_map, P.depth_map, P.acc_map, P.disp_map);
          else { s_run[lane] = a.r; s_run[32 + lane] = a.g; s_run[64 + lane] = a.b; s_run[96 + lane] = a.d; s_run[128 + lane] = a.w; }
        }
        if (q == 3 && lane == 31) s_carry[0] = s_carry[1];
      }
    };
    for (int64_t it = 0; it < n_iter; ++it) {
      const int64_t tile = tile_of(it);     // >= ntiles only in a cluster's padding tiles: g >= M, nothing stored
      const int64_t g = tile * kTileM + row;
      const int64_t gc = g < P.M ? g : P.M - 1;
      const int64_t ray = gc / P.S;
      [[maybe_unused]] uint32_t* mask_row = kSave ? P.mask_tape + (size_t)tile * (kMaskUnits * 128) + row : nullptr;
      [[maybe_unused]] uint8_t* tape_tile = kSave ? P.tape + (size_t)tile * kTapeFwdSlots * kKBlockBytes : nullptr;
      // ---- encodings (A.3).  The xyz tile of tile t+1 is built by group 0 while tile t runs (after its step-0
      // epilogue, in time the workers would otherwise spend waiting for the next accumulator) into the other xyz
      // buffer, so the first MMAs of a tile never wait for sincos.  The direction tile is needed by the view layer
      // only (the BIAS MMAs multiply the constant ones tile), so group 1 builds it at the start of the tile, off the
      // critical path; its stores are published by the fence below and ordered by the group's act-ready arrivals.
      if (grp == 0 && it == 0) compute_pe(tile, 0u);
      if (grp == 1) {
        float d[32];
#pragma unroll
        for (int i = kPED; i < 32; ++i) d[i] = 0.0f;
        d[kBiasColHi] = 1.0f;               // the two "ones" columns the DIR chunk's bias rows multiply
        d[kBiasColLo] = 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float p = P.viewdirs[3 * ray + c];
          d[c] = p;
          float sn, cs;
          sincosf(p, &sn, &cs);
#pragma unroll
          for (int k = 0; k < kLD; ++k) {
            d[3 + 6 * k + c] = sn;
            d[3 + 6 * k + 3 + c] = cs;
            const float s2 = 2.0f * sn * cs;
            cs = 1.0f - 2.0f * sn * sn;
            sn = s2;
          }
        }
        const uint32_t ped_row = base + kOffPed + row * 128u;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          // columns 32..63 of the tile are never multiplied by the forward, and wgrad writes out only the first 27
          // columns of its product with the taped image, so the upper half of the image stays undefined
          const uint4 pk = make_uint4(pack_bf16(d[c16 * 8 + 0], d[c16 * 8 + 1]), pack_bf16(d[c16 * 8 + 2], d[c16 * 8 + 3]),
                                      pack_bf16(d[c16 * 8 + 4], d[c16 * 8 + 5]), pack_bf16(d[c16 * 8 + 6], d[c16 * 8 + 7]));
          st_shared_v4(ped_row + (((uint32_t)c16 ^ (row & 7u)) << 4), pk.x, pk.y, pk.z, pk.w);
        }
        fence_proxy_async_smem();
      }

      const float* rowbias = nullptr;
      if (P.cond) {
        const int64_t crow = cond_row(P.cond_index, P.C, ray);
        rowbias = P.cond_proj + crow * kW;
      }
      float sigma = 0.0f;                                   // this thread's share of the sigma head
      // ---- trunk + feature epilogues ---------------------------------------------------------------
#pragma unroll 1
      for (int step = 0; step < 9; ++step) {
        const int a = step & 1;
        FN_TRACE(wtile == 2 && row == 0, wslot++);
        mbar_wait(bar_acc(a), acc_cnt[a] & 1u);
        FN_TRACE(wtile == 2 && row == 0, wslot++);
        ++acc_cnt[a];
        tc_fence_after();
        const uint32_t tacc = tmem_row + (uint32_t)a * 256u;
        // K-block kb of the next layer's A operand is produced in round kb by ALL 16 warps (group g: its columns
        // 16g..16g+15): the first K-block is complete after a quarter of the layer's epilogue instead of half of it, and
        // the issuer's wait for it is the longest stall of a layer (timeline: ~1000 of ~3250 cycles).
        [[maybe_unused]] uint32_t mask_a = 0u, mask_b = 0u;   // training tape: ReLU bits of the pieces of K-blocks 0, 1 / 2, 3
#pragma unroll 1
        for (uint32_t kb = 0; kb < 4; ++kb) {
          const uint32_t col0 = kb * 64u + grp * 16u;
          const uint32_t dst = act_row + kb * kKBlockBytes;
          // training forward: the tape writers have read this K-block's previous image (K-blocks 0 / 1 carry ten images
          // per tile -- nine layers and the view layer's output --, K-blocks 2 / 3 nine)
          if (kSave) mbar_wait(bar_taped(kb), (((kb < 2 ? 10u : 9u) * (uint32_t)it + (uint32_t)step) & 1u) ^ 1u);
          uint32_t mbits;
          if (step == 8)
            mbits = epilogue_piece<false, false, false>(tacc + col0, nullptr, nullptr, dst, grp * 2u, row, sigma);
          else if (step == 7)
            mbits = epilogue_piece<true, true, false>(tacc + col0, heads_s + col0, nullptr, dst, grp * 2u, row, sigma);
          else if (step == 5 && P.cond)
            mbits = epilogue_piece<true, false, true>(tacc + col0, nullptr, rowbias + col0, dst, grp * 2u, row, sigma);
          else
            mbits = epilogue_piece<true, false, false>(tacc + col0, nullptr, nullptr, dst, grp * 2u, row, sigma);
          // every lane publishes its own stores to the async proxy; one lane per warp then arrives
          // (512 per-thread arrives on two barriers cost ~10 % of the epilogue in SYNCS throttling)
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kCl > 1) mbar_arrive_cluster(lead_bar_act0 + 8u * kb);
            else mbar_arrive(bar_act(kb));
          }
          if (kSave) { if (kb < 2) mask_a |= mbits << (8u * kb); else mask_b |= mbits << (8u * (kb - 2u)); }
          FN_TRACE(wtile == 2 && row == 0 && (kb & 1u), wslot++);
        }
        if (kSave && kTapeAux && step < 8) {
          // ReLU bits of piece (kb, grp) = bytes h and 2 + h (h = grp & 1) of the mask word of 32-column unit 2 kb + grp / 2.
          // All eight bytes are stored after the step's LAST act-ready arrival: an arrival's release (and the proxy fence)
          // wait for the thread's outstanding global stores -- an L2 round trip on the layer-to-layer critical path
          uint8_t* mw = reinterpret_cast<uint8_t*>(mask_row + (step * 8 + (int)(grp >> 1)) * 128) + (grp & 1u);
          mw[0] = (uint8_t)mask_a;                     mw[2] = (uint8_t)(mask_a >> 16);          // K-block 0: unit 0 / 1
          mw[2 * 512] = (uint8_t)(mask_a >> 8);        mw[2 * 512 + 2] = (uint8_t)(mask_a >> 24); // K-block 1: unit 2 / 3
          mw[4 * 512] = (uint8_t)mask_b;               mw[4 * 512 + 2] = (uint8_t)(mask_b >> 16);
          mw[6 * 512] = (uint8_t)(mask_b >> 8);        mw[6 * 512 + 2] = (uint8_t)(mask_b >> 24);
        }
        // next tile's xyz encodings, in the shadow of layer 1's MMAs (the buffer's last reader, layer 5 of the tile
        // before this one, completed long ago)
        if (step == 0 && grp == 0 && it + 1 < n_iter) compute_pe(tile_of(it + 1), (uint32_t)((it + 1) & 1));
        if (kComp && grp == 3 && step < 3 && it > 0) comp_phase(step + 1);
      }
      // ---- view layer epilogue + rgb head: each group reduces 32 of the 128 columns ---------------
      {
        mbar_wait(bar_acc(1), acc_cnt[1] & 1u);
        ++acc_cnt[1];
        tc_fence_after();
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
        uint32_t v[32];
        tmem_ld32(tmem_row + 256u + grp * 32u, v);
        tmem_ld_wait();
        const float* wrgb = heads_s + (kAuxWRgb - kAuxWAlpha);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = (int)grp * 32 + j;
          const float h = fmaxf(__uint_as_float(v[j]), 0.0f);         // view-layer bias is in the accumulator
          c0 = fmaf(h, wrgb[col], c0);
          c1 = fmaf(h, wrgb[kWV + col], c1);
          c2 = fmaf(h, wrgb[2 * kWV + col], c2);
        }
        [[maybe_unused]] uint32_t hv_mask = 0u;
        if (kSave) {
          // view-layer output for the tape: 2 images of 64 columns, parked in K-blocks 0 / 1 (their last readers, this
          // layer's MMAs, are complete) in the layout of every other activation image; the tape writers copy them out
          const uint32_t kb = grp >> 1;
          mbar_wait(bar_taped(kb), ((10u * (uint32_t)it + 9u) & 1u) ^ 1u);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 pk = make_uint4(pack_bf16_relu(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1])),
                                        pack_bf16_relu(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3])),
                                        pack_bf16_relu(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5])),
                                        pack_bf16_relu(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7])));
            hv_mask |= relu_bits(pk.x) << (4 * c) | relu_bits(pk.y) << (4 * c + 1) | relu_bits(pk.z) << (4 * c + 2) | relu_bits(pk.w) << (4 * c + 3);
            const uint32_t c16 = (grp & 1u) * 4u + (uint32_t)c;
            st_shared_v4(act_row + kb * kKBlockBytes + ((c16 ^ (row & 7u)) << 4), pk.x, pk.y, pk.z, pk.w);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_hv);
        }
        tc_fence_before();
        // The partials of three groups park in logical chunks 5..7 of this row of the direction tile (its unused upper
        // half; the view layer's MMAs, its last async-proxy readers this tile, are complete); the fourth group sums them.
        // That group is 0 in the plain kernel and 3 when compositing is fused in (group 0 already builds the next tile's
        // xyz encodings); chunk 4 + grp with the roles of 0 and kFin swapped.
        constexpr uint32_t kFin = kComp ? 3u : 0u;
        const uint32_t park = grp == 0 ? kFin : grp;          // group 0 parks in the finishing group's chunk when it is not the finisher
        if (grp != kFin) *reinterpret_cast<float4*>(ped_row_ptr + (((4u + park) ^ (row & 7u)) << 4)) = make_float4(c0, c1, c2, sigma);
        worker_bar_sync();
        if (kSave && kTapeAux) mask_row[(kMaskUnitHv + (int)grp) * 128] = hv_mask;
        if (grp == kFin) {
          // chunks 5, 6 hold groups 1, 2; chunk 7 holds group 3 (plain kernel) or group 0 (kComp).  The sum is always
          // (g0 + g1) + (g2 + g3): fp32 addition commutes, so both kernels produce the same bits
          const float4 p1 = *reinterpret_cast<const float4*>(ped_row_ptr + ((5u ^ (row & 7u)) << 4));
          const float4 p2 = *reinterpret_cast<const float4*>(ped_row_ptr + ((6u ^ (row & 7u)) << 4));
          float4 p3 = *reinterpret_cast<const float4*>(ped_row_ptr + ((7u ^ (row & 7u)) << 4));
          if (kComp) {                                        // own partial takes group 3's place, group 0's takes "own"
            const float4 own = make_float4(c0, c1, c2, sigma);
            c0 = p3.x; c1 = p3.y; c2 = p3.z; sigma = p3.w;
            p3 = own;
          }
          const float* brgb = heads_s + (kAuxBRgb - kAuxWAlpha);
          c0 = brgb[0] + ((c0 + p1.x) + (p2.x + p3.x));
          c1 = brgb[1] + ((c1 + p1.y) + (p2.y + p3.y));
          c2 = brgb[2] + ((c2 + p1.z) + (p2.z + p3.z));
          const float sg = heads_s[kAuxBAlpha - kAuxWAlpha] + ((sigma + p1.w) + (p2.w + p3.w));
          // streaming store: raw is gigabytes per frame, read once by compositing
          if (g < P.M && (!kComp || P.raw != nullptr))
            asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.raw + g), "f"(c0), "f"(c1), "f"(c2), "f"(sg) : "memory");
          if (kComp) {
            // hand the tile to the compositing phases, which run inside the NEXT tile (comp_phase): final raw values in the
            // free chunk 4 of this row of the direction tile, the tile's coordinates next to the running sums
            *reinterpret_cast<float4*>(ped_row_ptr + ((4u ^ (row & 7u)) << 4)) = make_float4(c0, c1, c2, sg);
            float* cs = reinterpret_cast<float*>(base_ptr + kOffComp);
            if (lane == 0) reinterpret_cast<long long*>(cs + 10)[q] = ray;
            if (row == 0) { reinterpret_cast<long long*>(cs + 8)[0] = tile; reinterpret_cast<int*>(cs + 6)[0] = (int)(gc - ray * (int64_t)P.S); }
          }
        }
      }
      ++wtile;
    }
    if (kComp && grp == 3 && n_iter > 0) { comp_phase(1); comp_phase(2); comp_phase(3); }     // the CTA's last tile
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (kCl > 1) cluster_sync();               // the peer may still arrive on this CTA's barriers / its MMAs read this CTA's tiles
  if (warp == 1) {
    tc_fence_after();
    if (kCl > 1) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

#ifdef FNERF_TRACE
extern "C" int fnerf_debug_set_trace(long long* buf) {
  return (int)cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf));
}
#endif

// tape == nullptr: render path.  tape != nullptr: training forward, every intermediate activation is also
// streamed to the tape as K-block images plus one ReLU bitmask word per row and 32-column unit
// (mlp_dgrad_tc.cu / mlp_bwd_tc.cu consume them).
template <int kCl>
static int launch_mlp_tc_cluster(const TcParams& P, cudaStream_t s) {
  static DeviceOnce once;
  static int max_clusters[kMaxDevices];              // < 0 = clusters unavailable
  const int dev = current_device();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kTcSmemBytes; cfg.stream = s; cfg.attrs = attr; cfg.numAttrs = 1;
  device_once(once, [&] {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc<false, kCl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    int n = 0;
    cfg.gridDim = dim3((unsigned)(num_sms() / kCl * kCl));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, k_mlp_tc<false, kCl>, &cfg);
    max_clusters[dev] = (e == cudaSuccess && n > 0) ? n : -1;
    (void)cudaGetLastError();
    return cudaSuccess;
  });
  if (max_clusters[dev] < 0) return -1;              // caller falls back to the single-CTA kernel
  int64_t clusters = max_clusters[dev];
  const int64_t want = (P.ntiles + kCl - 1) / kCl;
  if (clusters > want) clusters = want;
  cfg.gridDim = dim3((unsigned)(clusters * kCl));
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_mlp_tc<false, kCl>, P);
  if (e != cudaSuccess) return set_error((int)e, "mlp_tc cluster launch: %s", cudaGetErrorString(e));
  return check_launch("mlp_tc");
}

// render kernel as single CTAs (1) or CTA pairs (2): FNERF_MLP_CLUSTER overrides the default
static int mlp_cluster_size() {
  static std::once_flag flag;
  static int v = kDefaultMlpCluster;
  std::call_once(flag, [] {
    const char* e = getenv("FNERF_MLP_CLUSTER");
    v = e ? atoi(e) : kDefaultMlpCluster;
    if (v != 1 && v != 2) v = 1;
  });
  return v;
}

int launch_mlp_tc_tape(const MlpArgs& a, uint8_t* tape, uint32_t* mask_tape, cudaStream_t s) {
  const int64_t M = a.R * a.S;
  if (M == 0) return 0;
  static DeviceOnce once[2];
  const int sv = tape != nullptr ? 1 : 0;
  if (cudaError_t e = sv ? opt_in_smem_once(once[1], k_mlp_tc<true, 1>, kTcSmemBytes) : opt_in_smem_once(once[0], k_mlp_tc<false, 1>, kTcSmemBytes))
    return set_error((int)e, "mlp_tc attr: %s", cudaGetErrorString(e));
  TcParams P;
  P.packed = reinterpret_cast<const uint8_t*>(a.packed);
  P.rays_o = a.rays_o; P.rays_d = a.rays_d; P.viewdirs = a.viewdirs; P.z = a.z;
  P.cond_proj = a.cond_proj; P.cond_index = a.cond_index; P.C = a.C;
  P.raw = reinterpret_cast<float4*>(a.raw);
  P.M = M; P.S = (int)a.S; P.ntiles = (M + kTileM - 1) / kTileM; P.cond = a.cond;
  P.tape = tape; P.mask_tape = mask_tape;
  if (!sv && P.ntiles >= 2 * (int64_t)num_sms()) {     // enough tiles per CTA for the shared weight stream to pay
    const int cl = mlp_cluster_size();
    int rc = -1;
    if (cl == 2) rc = launch_mlp_tc_cluster<2>(P, s);
    if (rc >= 0) return rc;
  }
  int64_t blocks = num_sms();
  if (blocks > P.ntiles) blocks = P.ntiles;
  if (sv) k_mlp_tc<true, 1><<<(unsigned)blocks, kTcThreads + 32 * kTapeWarps, kTcSmemBytes, s>>>(P);
  else k_mlp_tc<false, 1><<<(unsigned)blocks, kTcThreads, kTcSmemBytes, s>>>(P);
  return check_launch("mlp_tc");
}

int launch_mlp_tc(const MlpArgs& a, cudaStream_t s) { return launch_mlp_tc_tape(a, nullptr, nullptr, s); }

// Tiles per whole-ray group for a sample count S (gt * 128 = rays * S), or 0 when the fused compositing epilogue cannot
// serve S: a warp must hold 32 samples of one ray (S % 32 == 0), and a group is capped at 16 tiles so that the static
// round-robin of groups over the CTAs stays balanced.
int mlp_tc_composite_group(int64_t S) {
  if (S < 1) return 0;
  int64_t a = S, b = kTileM;
  while (b) { const int64_t t = a % b; a = b; b = t; }          // a = gcd(S, 128)
  const int64_t gt = S / a;
  return (S % 32 == 0 && gt <= 16) ? (int)gt : 0;
}

// Network query with alpha compositing (A.5) fused into the last epilogue: raw[R,S,4] is written only when a.raw != NULL.
int launch_mlp_tc_composite(const MlpArgs& a, const CompositeOut& c, cudaStream_t s) {
  const int64_t M = a.R * a.S;
  if (M == 0) return 0;
  const int gt = mlp_tc_composite_group(a.S);
  if (gt == 0) return set_error(FNERF_ERR_SIZE, "mlp_fwd_composite: S=%lld has no whole-ray tile group", (long long)a.S);
  static DeviceOnce once;
  if (cudaError_t e = opt_in_smem_once(once, k_mlp_tc<false, 1, true>, kTcSmemBytes)) return set_error((int)e, "mlp_tc attr: %s", cudaGetErrorString(e));
  TcParams P = {};
  P.packed = reinterpret_cast<const uint8_t*>(a.packed);
  P.rays_o = a.rays_o; P.rays_d = a.rays_d; P.viewdirs = a.viewdirs; P.z = a.z;
  P.cond_proj = a.cond_proj; P.cond_index = a.cond_index; P.C = a.C;
  P.raw = reinterpret_cast<float4*>(a.raw);
  P.M = M; P.S = (int)a.S; P.ntiles = (M + kTileM - 1) / kTileM; P.cond = a.cond;
  P.dnorm = c.dnorm; P.noise = c.noise; P.rgb_map = c.rgb; P.depth_map = c.depth; P.acc_map = c.acc; P.disp_map = c.disp;
  P.weights = c.weights; P.gt = gt; P.rays_per_group = (int)((int64_t)gt * kTileM / a.S); P.white = c.white;
  const int64_t ngroups = (P.ntiles + gt - 1) / gt;
  int64_t blocks = num_sms();
  if (blocks > ngroups) blocks = ngroups;
  k_mlp_tc<false, 1, true><<<(unsigned)blocks, kTcThreads, kTcSmemBytes, s>>>(P);
  return check_launch("mlp_tc_composite");
}

}  // namespace fnerf
