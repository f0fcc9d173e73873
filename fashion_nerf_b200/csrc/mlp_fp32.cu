// fp32 SIMT network query (A.3 + A.4 + A.8): the "fp32 CUDA path" and the correctness anchor of
// the tensor-core kernel.  One CTA walks 64-sample tiles; activations stay in shared memory in
// K-major form ([feature][sample], row stride 68 floats so the 8-row float4 reads are broadcasts
// and the column writes are bank-conflict free).  Each of the 256 compute threads owns an 8 (samples) x N/32 (features)
// block; every output accumulates over k in order, so the result does not depend on the blocking.
// Weights (packed section C, K-major: 32 consecutive k rows are one contiguous block) are streamed by a ninth warp with
// bulk TMA through two 32 KB stages of shared memory (full / empty mbarriers, the 77 chunks of a tile in consumption
// order), so no compute warp ever waits on a global load: with one CTA of 8 compute warps per SM (163 KB of activations)
// the first version, which read the weights with __ldg, sat on the long scoreboard (ncu: 1.2 stalled warps per issue, FMA
// pipe 50 %).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fnerf {
using namespace ptx;

constexpr int kTM = 64;      // samples per tile
constexpr int kLds = 68;      // row stride (floats) of the K-major activation buffers
constexpr int kThreads = 256;                  // compute threads; one more warp streams the weights
constexpr int kWRows = 32;                     // k rows per weight chunk
constexpr int kWStageFloats = kWRows * kW;     // 32 KB
constexpr size_t kFp32ActFloats = (size_t)(2 * kW + kPE + kPED) * kLds + 4 * kTM;
constexpr size_t kFp32Smem = (kFp32ActFloats + 2 * kWStageFloats) * sizeof(float) + 64;   // + 4 mbarriers
static_assert(kFp32Smem + 1024 <= 227 * 1024, "shared memory budget (512 B static for the row-bias pointers)");

// feature owned by lane tx as its j-th output column: groups of 4 consecutive columns, 128 apart, so that the weights of
// one k are N/128 16-byte shared-memory loads per lane
__device__ __forceinline__ int feat_of(int tx, int j) { return (j >> 2) * 128 + 4 * tx + (j & 3); }

// the weight ring as the compute warps see it: stage s of chunk c is c & 1, its phase (c >> 1) & 1
struct WRing {
  const float* stage0; uint32_t full0, empty0; uint32_t c;     // c: chunks consumed so far (identical in all compute threads)
};

__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

template <int N, bool RELU>
__device__ __forceinline__ void dense(WRing& ring, const float* __restrict__ inA, int KA, const float* __restrict__ inB, int KB,
                                      const float* __restrict__ bias, const float* const* rowbias,
                                      float* __restrict__ out) {
  constexpr int NJ = N / 32, NG = N / 128;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8][NJ];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float4 b = *reinterpret_cast<const float4*>(bias + g * 128 + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][4 * g] = b.x; acc[i][4 * g + 1] = b.y; acc[i][4 * g + 2] = b.z; acc[i][4 * g + 3] = b.w; }
  }
  if (rowbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* rb = rowbias[ty * 8 + i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] += rb[feat_of(tx, j)];
    }
  }
  auto segment = [&](const float* in, int K) {
    for (int k0 = 0; k0 < K; k0 += kWRows) {
      const uint32_t st = ring.c & 1u;
      mbar_wait(ring.full0 + 8u * st, (ring.c >> 1) & 1u);
      const float* wst = ring.stage0 + st * kWStageFloats + 4 * tx;
      const int rows = K - k0 < kWRows ? K - k0 : kWRows;
#pragma unroll 4
      for (int kk = 0; kk < rows; ++kk) {
        const float* ink = in + (k0 + kk) * kLds + ty * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(ink);
        const float4 a1 = *reinterpret_cast<const float4*>(ink + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float wv[NJ];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 t = *reinterpret_cast<const float4*>(wst + kk * N + g * 128);
          wv[4 * g] = t.x; wv[4 * g + 1] = t.y; wv[4 * g + 2] = t.z; wv[4 * g + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], wv[j], acc[i][j]);
      }
      __syncwarp();
      if (tx == 0) mbar_arrive(ring.empty0 + 8u * st);        // this warp is done with the stage
      ++ring.c;
    }
  };
  segment(inA, KA);
  if (KB > 0) segment(inB, KB);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = RELU ? fmaxf(acc[i][j], 0.0f) : acc[i][j];
    float* o = out + feat_of(tx, j) * kLds + ty * 8;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__global__ void __launch_bounds__(kThreads + 32, 1)
k_mlp_fp32(const uint8_t* __restrict__ packed, int cond, const float* __restrict__ rays_o,
           const float* __restrict__ rays_d, const float* __restrict__ viewdirs,
           const float* __restrict__ z, const float* __restrict__ cond_proj,
           const int32_t* __restrict__ cond_index, int64_t C, float4* __restrict__ raw,
           int64_t R, int S) {
  extern __shared__ __align__(16) float smem_f[];
  float* H0 = smem_f;
  float* H1 = H0 + kW * kLds;
  float* PE = H1 + kW * kLds;
  float* PED = PE + kPE * kLds;
  float* scratch = PED + kPED * kLds;           // [4][64]
  float* wstage = scratch + 4 * kTM;            // 2 x kWStageFloats
  const uint32_t bar0 = smem_u32(wstage + 2 * kWStageFloats);   // full[2], empty[2]
  __shared__ const float* s_rowbias[kTM];

  const float* aux = reinterpret_cast<const float*>(packed + kSecBOffset);
  const float* secC = reinterpret_cast<const float*>(packed + kSecCOffset);
  const float* wt[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) wt[j] = secC + simt_offset_floats(j, cond);
  const int hoff = kPE + (cond ? kCond : 0);

  const int64_t M = R * (int64_t)S;
  const int64_t ntiles = (M + kTM - 1) / kTM;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);                       // full: the producer's expect_tx arrival
    mbar_init(bar0 + 16, kThreads / 32); mbar_init(bar0 + 24, kThreads / 32);   // empty: one arrival per compute warp
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x >= kThreads) {
    // ================================ weight producer ================================================
    // the segments of one tile in the order the dense() calls below consume them: (weights, K rows, N columns)
    if (threadIdx.x == kThreads) {
      uint32_t c = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int sgm = 0; sgm < 12; ++sgm) {
          const float* w; int K, N = kW;
          if (sgm == 0) { w = wt[0]; K = kPE; }
          else if (sgm <= 4) { w = wt[sgm]; K = kW; }
          else if (sgm == 5) { w = wt[5]; K = kPE; }
          else if (sgm == 6) { w = wt[5] + (size_t)hoff * kW; K = kW; }
          else if (sgm <= 8) { w = wt[sgm - 1]; K = kW; }
          else if (sgm == 9) { w = wt[8]; K = kW; }
          else if (sgm == 10) { w = wt[9]; K = kW; N = kWV; }
          else { w = wt[9] + (size_t)kW * kWV; K = kPED; N = kWV; }
          for (int k0 = 0; k0 < K; k0 += kWRows, ++c) {
            const uint32_t st = c & 1u;
            const int rows = K - k0 < kWRows ? K - k0 : kWRows;
            const uint32_t bytes = (uint32_t)(rows * N) * 4u;
            mbar_wait(bar0 + 16 + 8u * st, ((c >> 1) & 1u) ^ 1u);
            mbar_expect_tx(bar0 + 8u * st, bytes);
            bulk_g2s(smem_u32(wstage + st * kWStageFloats), w + (size_t)k0 * N, bytes, bar0 + 8u * st);
          }
        }
      }
    }
    return;
  }
  WRing ring = {wstage, bar0, bar0 + 16, 0u};
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t g0 = tile * kTM;
    // ---- positional encodings (A.3), K-major into PE / PED -----------------------------------
    for (int e = threadIdx.x; e < kTM * 33; e += kThreads) {
      const int m = e & (kTM - 1), q = e / kTM;
      int64_t g = g0 + m; if (g >= M) g = M - 1;
      const int64_t ray = g / S;
      const int c = q < 3 ? q : (q - 3) % 3;
      // pts = o + d*z with the oracle's two roundings
      const float p = __fadd_rn(rays_o[3 * ray + c], __fmul_rn(rays_d[3 * ray + c], z[g]));
      if (q < 3) PE[q * kLds + m] = p;
      else {
        const int k = (q - 3) / 3;
        float sn, cs;
        sincosf(__fmul_rn(p, (float)(1 << k)), &sn, &cs);
        PE[(3 + 6 * k + c) * kLds + m] = sn;
        PE[(3 + 6 * k + 3 + c) * kLds + m] = cs;
      }
    }
    for (int e = threadIdx.x; e < kTM * 15; e += kThreads) {
      const int m = e & (kTM - 1), q = e / kTM;
      int64_t g = g0 + m; if (g >= M) g = M - 1;
      const int64_t ray = g / S;
      const int c = q < 3 ? q : (q - 3) % 3;
      const float p = viewdirs[3 * ray + c];
      if (q < 3) PED[q * kLds + m] = p;
      else {
        const int k = (q - 3) / 3;
        float sn, cs;
        sincosf(__fmul_rn(p, (float)(1 << k)), &sn, &cs);
        PED[(3 + 6 * k + c) * kLds + m] = sn;
        PED[(3 + 6 * k + 3 + c) * kLds + m] = cs;
      }
    }
    if (cond && threadIdx.x < kTM) {
      int64_t g = g0 + threadIdx.x; if (g >= M) g = M - 1;
      const int64_t ray = g / S;
      const int64_t row = cond_row(cond_index, C, ray);
      s_rowbias[threadIdx.x] = cond_proj + row * kW;
    }
    compute_bar_sync();

    // ---- trunk (A.4) ---------------------------------------------------------------------------
    dense<kW, true>(ring, PE, kPE, nullptr, 0, aux + kAuxBiasPts, nullptr, H0);
    compute_bar_sync();
    dense<kW, true>(ring, H0, kW, nullptr, 0, aux + kAuxBiasPts + 256, nullptr, H1);
    compute_bar_sync();
    dense<kW, true>(ring, H1, kW, nullptr, 0, aux + kAuxBiasPts + 512, nullptr, H0);
    compute_bar_sync();
    dense<kW, true>(ring, H0, kW, nullptr, 0, aux + kAuxBiasPts + 768, nullptr, H1);
    compute_bar_sync();
    dense<kW, true>(ring, H1, kW, nullptr, 0, aux + kAuxBiasPts + 1024, nullptr, H0);
    compute_bar_sync();
    // layer 5: cat([pe, (cond), h]) -- the cond block enters as the hoisted per-ray projection
    dense<kW, true>(ring, PE, kPE, H0, kW, aux + kAuxBiasPts + 1280, cond ? s_rowbias : nullptr, H1);
    compute_bar_sync();
    dense<kW, true>(ring, H1, kW, nullptr, 0, aux + kAuxBiasPts + 1536, nullptr, H0);
    compute_bar_sync();
    dense<kW, true>(ring, H0, kW, nullptr, 0, aux + kAuxBiasPts + 1792, nullptr, H1);
    compute_bar_sync();
    // ---- heads -----------------------------------------------------------------------------------
    {  // sigma partials: 4 threads per sample, 64 features each
      const int m = threadIdx.x & (kTM - 1), part = threadIdx.x / kTM;
      float acc = 0.0f;
      for (int k = part * 64; k < part * 64 + 64; ++k) acc = fmaf(aux[kAuxWAlpha + k], H1[k * kLds + m], acc);
      scratch[part * kTM + m] = acc;
    }
    dense<kW, false>(ring, H1, kW, nullptr, 0, aux + kAuxBiasFeat, nullptr, H0);
    compute_bar_sync();
    dense<kWV, true>(ring, H0, kW, PED, kPED, aux + kAuxBiasViews, nullptr, H1);
    compute_bar_sync();
    if (threadIdx.x < kTM) {
      const int m = threadIdx.x;
      const int64_t g = g0 + m;
      float c0 = aux[kAuxBRgb], c1 = aux[kAuxBRgb + 1], c2 = aux[kAuxBRgb + 2];
      for (int k = 0; k < kWV; ++k) {
        const float h = H1[k * kLds + m];
        c0 = fmaf(aux[kAuxWRgb + k], h, c0);
        c1 = fmaf(aux[kAuxWRgb + kWV + k], h, c1);
        c2 = fmaf(aux[kAuxWRgb + 2 * kWV + k], h, c2);
      }
      const float sg = aux[kAuxBAlpha] + ((scratch[m] + scratch[kTM + m]) + (scratch[2 * kTM + m] + scratch[3 * kTM + m]));
      if (g < M) raw[g] = make_float4(c0, c1, c2, sg);
    }
    compute_bar_sync();
  }
}

int launch_mlp_fp32(const MlpArgs& a, cudaStream_t s) {
  const int64_t M = a.R * a.S;
  if (M == 0) return 0;
  static DeviceOnce once;
  if (cudaError_t e = opt_in_smem_once(once, k_mlp_fp32, kFp32Smem)) return set_error((int)e, "mlp_fp32 attr: %s", cudaGetErrorString(e));
  const int64_t ntiles = (M + kTM - 1) / kTM;
  int64_t blocks = num_sms();
  if (blocks > ntiles) blocks = ntiles;
  k_mlp_fp32<<<(unsigned)blocks, kThreads + 32, kFp32Smem, s>>>(
      reinterpret_cast<const uint8_t*>(a.packed), a.cond, a.rays_o, a.rays_d, a.viewdirs, a.z,
      a.cond_proj, a.cond_index, a.C, reinterpret_cast<float4*>(a.raw), a.R, (int)a.S);
  return check_launch("mlp_fp32");
}

}  // namespace fnerf
