// Micro-benchmark: the epilogue unit of k_mlp_tc (LDTM.x32 -> +bias (LDS) -> ReLU/bf16 pack -> swizzled
// STS.128) in isolation: 16 warps, no MMA / TMA traffic.  Separates intrinsic cost from contention.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../fashion_nerf_b200/csrc/tc_ptx.cuh"
using namespace fnerf::ptx;

template <int VARIANT>
__device__ __forceinline__ void unit(uint32_t taddr, const float* bias_s, uint32_t act_row_addr, uint32_t chunk0, uint32_t row) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[c * 8 + j]);
    float4 b0, b1;
    if (VARIANT == 1) { b0 = make_float4(.1f, .2f, .3f, .4f); b1 = b0; }
    else { b0 = *reinterpret_cast<const float4*>(bias_s + c * 8); b1 = *reinterpret_cast<const float4*>(bias_s + c * 8 + 4); }
    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    uint32_t p0 = pack_bf16_relu(x[0], x[1]), p1 = pack_bf16_relu(x[2], x[3]), p2 = pack_bf16_relu(x[4], x[5]), p3 = pack_bf16_relu(x[6], x[7]);
    const uint32_t c16 = chunk0 + (uint32_t)c;
    st_shared_v4(act_row_addr + ((c16 ^ (row & 7u)) << 4), p0, p1, p2, p3);
  }
}

template <int VARIANT>
__global__ void __launch_bounds__(576, 1) k(int nwarps, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tslot;
  __shared__ __align__(16) float bias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 256) bias[threadIdx.x] = threadIdx.x * 0.001f;
  if (warp == 17) tmem_alloc(smem_u32(&tslot), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t q = warp & 3, grp = warp >> 2, row = q * 32 + lane;
    const uint32_t trow = tmem + ((q * 32u) << 16);
    for (int it = 0; it < iters; ++it) {
      const uint32_t u = (grp + 4u * (it & 1)) & 7u;
      unit<VARIANT>(trow + u * 32, bias + u * 32, base + (u >> 1) * 16384 + row * 128, (u & 1) * 4, row);
      fence_proxy_async_smem();
      tc_fence_before();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int VARIANT>
void run(int nwarps, long long* out, const char* name) {
  const int smem = 65536 + 2048, iters = 4000;
  cudaFuncSetAttribute(k<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<VARIANT><<<148, 576, smem>>>(nwarps, 100, out);
  cudaDeviceSynchronize();
  k<VARIANT><<<148, 576, smem>>>(nwarps, iters, out);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  printf("{\"variant\":\"%s\",\"warps\":%d,\"cycles_per_unit\":%.1f,\"err\":\"%s\"}\n", name, nwarps, cyc / iters, cudaGetErrorString(err));
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  for (int nw : {4, 8, 16}) { run<0>(nw, out, "full"); run<1>(nw, out, "no-bias-lds"); }
  return 0;
}
