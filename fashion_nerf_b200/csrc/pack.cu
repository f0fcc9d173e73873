// Weight packer / unpacker (flat fp32 nn.Linear layout <-> packed device blob, see layout.h),
// standalone positional encoding (A.3) and the hoisted conditioning projection (A.8).
#include <cuda_bf16.h>
#include "common.cuh"

namespace fnerf {

// The four sections of the blob are written by ONE launch (k_pack_all): a re-pack follows every optimiser step, and four
// launches per network were 2 % of a training step.  Each section keeps its own block decomposition (bx, by).
// section A, blocks (chunk, quarter of the chunk's elements): 48 CTAs alone left most of the chip idle for 36 us per re-pack
__device__ __forceinline__ void pack_bf16_block(const float* __restrict__ flat, uint8_t* __restrict__ packed, int cond, int bx, int by) {
  const int c = bx;
  const ChunkDesc cd = chunk_desc(c);
  const LayerDim d = layer_dim(cd.layer, cond);
  const float* W = flat + flat_weight_offset(cd.layer, cond);
  const float* bvec = flat + flat_bias_offset(cd.layer, cond);
  uint8_t* dst = packed + chunk_offset(c);
  if (cd.kind == CHUNK_BIAS) {                            // MN-major K = 16 tile: K rows 11 / 12 = bias hi / lo
    for (int e = by * 1024 + threadIdx.x; e < (by + 1) * 1024; e += blockDim.x) {
      const int n = e >> 4, k = e & 15;
      const float b = bvec[n];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      const float v = k == kBiasColHi - 16 ? hi : (k == kBiasColLo - 16 ? b - hi : 0.0f);
      *reinterpret_cast<__nv_bfloat16*>(dst + bias_chunk_offset((uint32_t)n, (uint32_t)k)) = __float2bfloat16_rn(v);
    }
    return;
  }
  const int rows = c >= kFirstViewChunk ? 128 : 256;
  for (int e = by * rows * 16 + threadIdx.x; e < (by + 1) * rows * 16; e += blockDim.x) {
    const int r = e >> 6, k = e & 63;                     // output feature r, K column k of the chunk
    float v = 0.0f;
    if (cd.kind == CHUNK_TRUNK) {
      const int base = (cd.layer == 5) ? kPE + (cond ? kCond : 0) : 0;
      v = W[(int64_t)r * d.in + base + cd.kb * 64 + k];
    } else if (cd.kind == CHUNK_XYZ) {
      if (k < kPE) v = W[(int64_t)r * d.in + k];
    } else {                                              // DIR: view-layer weights of the direction encoding + bias hi / lo
      if (k < kPED) v = W[(int64_t)r * d.in + kW + k];
      const float b = bvec[r];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      if (k == kBiasColHi) v = hi;
      if (k == kBiasColLo) v = b - hi;
    }
    *reinterpret_cast<__nv_bfloat16*>(dst + sw128_offset(r, k)) = __float2bfloat16_rn(v);
  }
}

// Section E: transposed chunks for dgrad: chunk row k = input feature, column j = output feature 64*kb + j
__device__ __forceinline__ void pack_bf16_T_block(const float* __restrict__ flat, uint8_t* __restrict__ secE, int cond, int bx, int by) {
  const int c = bx;
  const ChunkTDesc cd = chunk_t_desc(c);
  const LayerDim d = layer_dim(cd.layer, cond);
  const float* W = flat + flat_weight_offset(cd.layer, cond);
  const int base = (cd.layer == 5) ? kPE + (cond ? kCond : 0) : 0;
  for (int e = by * 4096 + threadIdx.x; e < (by + 1) * 4096; e += blockDim.x) {   // by = quarter of the chunk
    const int k = e >> 6, j = e & 63;
    const int n = cd.kb * 64 + j;
    const float v = W[(int64_t)n * d.in + base + k];
    *reinterpret_cast<__nv_bfloat16*>(secE + (size_t)c * kChunkTBytes + sw128_offset((uint32_t)k, (uint32_t)j)) = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ void pack_aux_block(const float* __restrict__ flat, float* __restrict__ aux, int cond, int bx) {
  int i = bx * blockDim.x + threadIdx.x;
  if (i >= kAuxFloats) return;
  float v = 0.0f;
  if (i < kAuxBiasFeat) v = flat[flat_bias_offset(i >> 8, cond) + (i & 255)];
  else if (i < kAuxBiasViews) v = flat[flat_bias_offset(9, cond) + (i - kAuxBiasFeat)];
  else if (i < kAuxWAlpha) v = flat[flat_bias_offset(10, cond) + (i - kAuxBiasViews)];
  else if (i < kAuxBAlpha) v = flat[flat_weight_offset(8, cond) + (i - kAuxWAlpha)];
  else if (i == kAuxBAlpha) v = flat[flat_bias_offset(8, cond)];
  else if (i >= kAuxWRgb && i < kAuxBRgb) v = flat[flat_weight_offset(11, cond) + (i - kAuxWRgb)];
  else if (i >= kAuxBRgb && i < kAuxBRgb + 3) v = flat[flat_bias_offset(11, cond) + (i - kAuxBRgb)];
  aux[i] = v;
}

// Section C: layer j stored K-major: wt[k][n] = W[n][k]
__device__ __forceinline__ void pack_simt_block(const float* __restrict__ flat, float* __restrict__ secC, int cond, int bx, int by) {
  const int j = by;                              // one row of blocks per fp32 layer
  const int l = simt_layer_id(j);
  const LayerDim d = layer_dim(l, cond);
  int64_t e = (int64_t)bx * blockDim.x + threadIdx.x;
  if (e >= (int64_t)d.out * d.in) return;
  const int k = (int)(e / d.out), n = (int)(e % d.out);
  secC[simt_offset_floats(j, cond) + e] = flat[flat_weight_offset(l, cond) + (int64_t)n * d.in + k];
}

__global__ void k_unpack(const uint8_t* __restrict__ packed, float* __restrict__ flat, int cond) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= flat_count(cond)) return;
  // locate layer
  int l = 0;
  while (l + 1 < kNumLayers && e >= flat_weight_offset(l + 1, cond)) ++l;
  const LayerDim d = layer_dim(l, cond);
  const int64_t within = e - flat_weight_offset(l, cond);
  const float* aux = reinterpret_cast<const float*>(packed + kSecBOffset);
  const float* secC = reinterpret_cast<const float*>(packed + kSecCOffset);
  float v;
  if (within >= (int64_t)d.out * d.in) {            // bias
    const int n = (int)(within - (int64_t)d.out * d.in);
    if (l < 8) v = aux[kAuxBiasPts + l * 256 + n];
    else if (l == 8) v = aux[kAuxBAlpha];
    else if (l == 9) v = aux[kAuxBiasFeat + n];
    else if (l == 10) v = aux[kAuxBiasViews + n];
    else v = aux[kAuxBRgb + n];
  } else {
    const int n = (int)(within / d.in), k = (int)(within % d.in);
    if (l == 8) v = aux[kAuxWAlpha + k];
    else if (l == 11) v = aux[kAuxWRgb + n * kWV + k];
    else {
      const int j = l < 8 ? l : (l == 9 ? 8 : 9);
      v = secC[simt_offset_floats(j, cond) + (int64_t)k * d.out + n];
    }
  }
  flat[e] = v;
}

constexpr int kPackBlocksA = kNumChunks * 4, kPackBlocksAux = (kAuxFloats + 255) / 256, kPackBlocksT = kNumChunksT * 4;
__global__ void __launch_bounds__(256) k_pack_all(const float* __restrict__ flat, uint8_t* __restrict__ p, int cond, int simt_bx) {
  int b = (int)blockIdx.x;
  if (b < kPackBlocksA) return pack_bf16_block(flat, p, cond, b >> 2, b & 3);
  b -= kPackBlocksA;
  if (b < kPackBlocksT) return pack_bf16_T_block(flat, p + sec_e_offset(cond), cond, b >> 2, b & 3);
  b -= kPackBlocksT;
  if (b < kPackBlocksAux) return pack_aux_block(flat, reinterpret_cast<float*>(p + kSecBOffset), cond, b);
  b -= kPackBlocksAux;
  pack_simt_block(flat, reinterpret_cast<float*>(p + kSecCOffset), cond, b % simt_bx, b / simt_bx);
}

int launch_pack(const float* flat, void* packed, int cond, cudaStream_t s) {
  uint8_t* p = reinterpret_cast<uint8_t*>(packed);
  int64_t nmax = 0;
  for (int j = 0; j < 10; ++j) {
    const LayerDim d = layer_dim(simt_layer_id(j), cond);
    if ((int64_t)d.out * d.in > nmax) nmax = (int64_t)d.out * d.in;
  }
  const int simt_bx = (int)((nmax + 255) / 256);
  k_pack_all<<<kPackBlocksA + kPackBlocksT + kPackBlocksAux + simt_bx * 10, 256, 0, s>>>(flat, p, cond, simt_bx);
  return check_launch("pack_weights");
}

int launch_unpack(const void* packed, float* flat, int cond, cudaStream_t s) {
  const int64_t n = flat_count(cond);
  k_unpack<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(packed), flat, cond);
  return check_launch("unpack_weights");
}

// ------------------------------------------------------------------------------------------ A.10
// Adam on a flat fp32 buffer (torch semantics: m, v updated first; step = lr / (1 - b1^t); denom = sqrt(v / (1 - b2^t)) + eps)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       int64_t n, float b1, float b2, float one_minus_b1, float one_minus_b2, float step_size, float inv_bc2,
                       float eps, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = fmaf(one_minus_b1, gi, m[i] * b1);
  const float vi = fmaf(one_minus_b2 * gi, gi, v[i] * b2);
  m[i] = mi;
  v[i] = vi;
  p[i] -= step_size * mi / (sqrtf(vi * inv_bc2) + eps);
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int64_t t,
                float grad_scale, cudaStream_t s) {
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, g, m, v, n, b1, b2, 1.0f - b1, 1.0f - b2, (float)(lr / bc1), (float)(1.0 / bc2),
                                                       eps, grad_scale);
  return check_launch("adam_step");
}

// Gradient all-reduce fused with Adam over peer memory (NVLink / NVSwitch P2P loads): every rank reads element i of all
// `world` ranks' gradient buffers (same rank order everywhere -> bit-identical sums -> bit-identical replicas), scales,
// and applies Adam to its own parameters.  The reduced gradient is never written anywhere.
__global__ void __launch_bounds__(256) k_allreduce_adam(const float* const* __restrict__ peer_grads, int world, int64_t offset,
                                                        float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        float b1, float b2, float one_minus_b1, float one_minus_b2, float step_size,
                                                        float inv_bc2, float eps, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = 0.0f;
  for (int r = 0; r < world; ++r) g += __ldcv(peer_grads[r] + offset + i);     // volatile-cached: peers just wrote these
  const float gi = g * grad_scale;
  const float mi = fmaf(one_minus_b1, gi, m[i] * b1);
  const float vi = fmaf(one_minus_b2 * gi, gi, v[i] * b2);
  m[i] = mi;
  v[i] = vi;
  p[i] -= step_size * mi / (sqrtf(vi * inv_bc2) + eps);
}

int launch_allreduce_adam(const float* const* peer_grads, int world, int64_t offset, float* p, float* m, float* v, int64_t n, float lr,
                          float b1, float b2, float eps, int64_t t, float grad_scale, cudaStream_t s) {
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
  k_allreduce_adam<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(peer_grads, world, offset, p, m, v, n, b1, b2, 1.0f - b1, 1.0f - b2,
                                                                 (float)(lr / bc1), (float)(1.0 / bc2), eps, grad_scale);
  return check_launch("allreduce_adam_step");
}

// In-place sum of the ranks' gradient buffers through an NVSwitch multicast mapping (NVLS): rank r owns every world-th
// 16-byte group, pulls the switch-reduced sum of that group (multimem.ld_reduce: the switch adds the `world` copies) and
// multicast-stores it back to all ranks.  One reducer per element -> every rank ends up with identical bits.
__global__ void __launch_bounds__(256) k_multimem_allreduce(float* __restrict__ mc, int rank, int world, int64_t n4) {
  // this rank's groups: i = j * world + rank; four of them in flight per thread (the switch round trip is long)
  const int64_t mine = (n4 - rank + world - 1) / world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < mine; j0 += 4 * stride) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = j0 + u * stride;
      if (j < mine)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]) : "l"(mc + 4 * (j * world + rank)) : "memory");
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = j0 + u * stride;
      if (j < mine)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(mc + 4 * (j * world + rank)), "f"(v[u][0]), "f"(v[u][1]), "f"(v[u][2]), "f"(v[u][3]) : "memory");
    }
  }
}

int launch_multimem_allreduce(float* mc, int rank, int world, int64_t n, cudaStream_t s) {
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 / world / 4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_multimem_allreduce<<<(unsigned)blocks, 256, 0, s>>>(mc, rank, world, n4);
  return check_launch("multimem_allreduce");
}

// ------------------------------------------------------------------------------------------ A.3
__global__ void k_posenc(const float* __restrict__ x, float* __restrict__ out, int64_t M, int L) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int width = 3 + 6 * L;
  if (idx >= M * 3) return;
  const int64_t m = idx / 3;
  const int c = (int)(idx % 3);
  const float v = x[idx];
  float* o = out + m * width;
  o[c] = v;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s, co;
    sincosf(__fmul_rn(v, f), &s, &co);
    o[3 + 6 * k + c] = s;
    o[3 + 6 * k + 3 + c] = co;
    f *= 2.0f;
  }
}

int launch_posenc(const float* x, float* out, int64_t M, int L, cudaStream_t s) {
  if (M == 0) return 0;
  k_posenc<<<(unsigned)((M * 3 + 255) / 256), 256, 0, s>>>(x, out, M, L);
  return check_launch("posenc");
}

// ------------------------------------------------------------------------------------------ A.8
// proj[c][n] = sum_k cond[c][k] * W5[n][63 + k]  (fp32; W5 read from the K-major SIMT section)
__global__ void k_cond_project(const uint8_t* __restrict__ packed, const float* __restrict__ cond,
                               float* __restrict__ proj, int64_t C) {
  __shared__ float s_c[kCond];
  const float* w5t = reinterpret_cast<const float*>(packed + kSecCOffset) + simt_offset_floats(5, 1);
  for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
    __syncthreads();
    s_c[threadIdx.x] = cond[c * kCond + threadIdx.x];
    __syncthreads();
    float acc = 0.0f;
    for (int k = 0; k < kCond; ++k) acc = fmaf(s_c[k], w5t[(int64_t)(kPE + k) * kW + threadIdx.x], acc);
    proj[c * kW + threadIdx.x] = acc;
  }
}

int launch_cond_project(const void* packed, const float* cond, float* proj, int64_t C, cudaStream_t s) {
  if (C == 0) return 0;
  int64_t blocks = C < 65535 ? C : 65535;
  k_cond_project<<<(unsigned)blocks, kW, 0, s>>>(reinterpret_cast<const uint8_t*>(packed), cond, proj, C);
  return check_launch("cond_project");
}

}  // namespace fnerf
