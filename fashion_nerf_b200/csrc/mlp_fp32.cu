// fp32 SIMT network query (A.3 + A.4 + A.8): the "fp32 CUDA path" and the correctness anchor of
// the tensor-core kernel.  One CTA walks 64-sample tiles; activations stay in shared memory in
// K-major form ([feature][sample], row stride 68 floats so the 8-row float4 reads are broadcasts
// and the column writes are bank-conflict free); weights are read K-major (packed section C) from
// L1/L2 with 128-byte coalesced rows.  Each thread owns an 8 (samples) x N/32 (features) block.
#include "common.cuh"

namespace fnerf {

constexpr int kTM = 64;      // samples per tile
constexpr int kLds = 68;      // row stride (floats) of the K-major activation buffers
constexpr int kThreads = 256;
constexpr size_t kFp32Smem = (size_t)(2 * kW + kPE + kPED) * kLds * sizeof(float) + 4 * kTM * sizeof(float);

template <int N, bool RELU>
__device__ __forceinline__ void dense(const float* __restrict__ inA, int KA, const float* __restrict__ wA,
                                      const float* __restrict__ inB, int KB, const float* __restrict__ wB,
                                      const float* __restrict__ bias, const float* const* rowbias,
                                      float* __restrict__ out) {
  constexpr int NJ = N / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float b = bias[tx + 32 * j];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] = b;
  }
  if (rowbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* rb = rowbias[ty * 8 + i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] += rb[tx + 32 * j];
    }
  }
  auto segment = [&](const float* in, int K, const float* w) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(in + k * kLds + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(in + k * kLds + ty * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float wv[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) wv[j] = __ldg(w + (size_t)k * N + tx + 32 * j);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], wv[j], acc[i][j]);
    }
  };
  segment(inA, KA, wA);
  if (KB > 0) segment(inB, KB, wB);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = RELU ? fmaxf(acc[i][j], 0.0f) : acc[i][j];
    float* o = out + (tx + 32 * j) * kLds + ty * 8;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
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
  __shared__ const float* s_rowbias[kTM];

  const float* aux = reinterpret_cast<const float*>(packed + kSecBOffset);
  const float* secC = reinterpret_cast<const float*>(packed + kSecCOffset);
  const float* wt[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) wt[j] = secC + simt_offset_floats(j, cond);
  const int hoff = kPE + (cond ? kCond : 0);

  const int64_t M = R * (int64_t)S;
  const int64_t ntiles = (M + kTM - 1) / kTM;
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
    __syncthreads();

    // ---- trunk (A.4) ---------------------------------------------------------------------------
    dense<kW, true>(PE, kPE, wt[0], nullptr, 0, nullptr, aux + kAuxBiasPts, nullptr, H0);
    __syncthreads();
    dense<kW, true>(H0, kW, wt[1], nullptr, 0, nullptr, aux + kAuxBiasPts + 256, nullptr, H1);
    __syncthreads();
    dense<kW, true>(H1, kW, wt[2], nullptr, 0, nullptr, aux + kAuxBiasPts + 512, nullptr, H0);
    __syncthreads();
    dense<kW, true>(H0, kW, wt[3], nullptr, 0, nullptr, aux + kAuxBiasPts + 768, nullptr, H1);
    __syncthreads();
    dense<kW, true>(H1, kW, wt[4], nullptr, 0, nullptr, aux + kAuxBiasPts + 1024, nullptr, H0);
    __syncthreads();
    // layer 5: cat([pe, (cond), h]) -- the cond block enters as the hoisted per-ray projection
    dense<kW, true>(PE, kPE, wt[5], H0, kW, wt[5] + (size_t)hoff * kW, aux + kAuxBiasPts + 1280,
                    cond ? s_rowbias : nullptr, H1);
    __syncthreads();
    dense<kW, true>(H1, kW, wt[6], nullptr, 0, nullptr, aux + kAuxBiasPts + 1536, nullptr, H0);
    __syncthreads();
    dense<kW, true>(H0, kW, wt[7], nullptr, 0, nullptr, aux + kAuxBiasPts + 1792, nullptr, H1);
    __syncthreads();
    // ---- heads -----------------------------------------------------------------------------------
    {  // sigma partials: 4 threads per sample, 64 features each
      const int m = threadIdx.x & (kTM - 1), part = threadIdx.x / kTM;
      float acc = 0.0f;
      for (int k = part * 64; k < part * 64 + 64; ++k) acc = fmaf(aux[kAuxWAlpha + k], H1[k * kLds + m], acc);
      scratch[part * kTM + m] = acc;
    }
    dense<kW, false>(H1, kW, wt[8], nullptr, 0, nullptr, aux + kAuxBiasFeat, nullptr, H0);
    __syncthreads();
    dense<kWV, true>(H0, kW, wt[9], PED, kPED, wt[9] + (size_t)kW * kWV, aux + kAuxBiasViews, nullptr, H1);
    __syncthreads();
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
    __syncthreads();
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
  k_mlp_fp32<<<(unsigned)blocks, kThreads, kFp32Smem, s>>>(
      reinterpret_cast<const uint8_t*>(a.packed), a.cond, a.rays_o, a.rays_d, a.viewdirs, a.z,
      a.cond_proj, a.cond_index, a.C, reinterpret_cast<float4*>(a.raw), a.R, (int)a.S);
  return check_launch("mlp_fp32");
}

}  // namespace fnerf
