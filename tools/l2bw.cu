// Micro-benchmark: L2 -> shared memory bulk-copy (cp.async.bulk, TMA engine) bandwidth when every
// SM streams the same ~1.2 MB weight blob in 32 KB chunks (the access pattern of the fused MLP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bw tools/l2bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(128, 1) k_stream(const uint8_t* __restrict__ blob, int nchunks, int iters, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long acc = 0;
  if (threadIdx.x == 0) {
    const int total = nchunks * iters;
    int issued = 0;
    for (; issued < STAGES && issued < total; ++issued) {
      mbar_expect_tx(smem_u32(&full[issued]), CHUNK);
      bulk_g2s(smem_u32(smem + issued * CHUNK), blob + (size_t)(issued % nchunks) * CHUNK, CHUNK, smem_u32(&full[issued]));
    }
    for (int c = 0; c < total; ++c) {
      const int s = c % STAGES;
      mbar_wait(smem_u32(&full[s]), (c / STAGES) & 1);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + s * CHUNK + 64);
      if (issued < total) {
        mbar_expect_tx(smem_u32(&full[s]), CHUNK);
        bulk_g2s(smem_u32(smem + s * CHUNK), blob + (size_t)(issued % nchunks) * CHUNK, CHUNK, smem_u32(&full[s]));
        ++issued;
      }
    }
    sink[blockIdx.x] = acc;
  }
}

template <int STAGES, int CHUNK>
void run(const uint8_t* blob, int nchunks, int iters, unsigned long long* sink, int grid) {
  size_t smem = (size_t)STAGES * CHUNK + 1024;
  cudaFuncSetAttribute(k_stream<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_stream<STAGES, CHUNK><<<grid, 128, smem>>>(blob, nchunks, 4, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_stream<STAGES, CHUNK><<<grid, 128, smem>>>(blob, nchunks, iters, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = (double)grid * nchunks * iters * CHUNK;
  printf("{\"bench\":\"l2_to_smem_bulk\",\"grid\":%d,\"stages\":%d,\"chunk\":%d,\"ms\":%.3f,\"TBps\":%.3f,\"B_per_clk_per_SM_at_1.9GHz\":%.1f,\"err\":\"%s\"}\n",
         grid, STAGES, CHUNK, ms, bytes / ms / 1e9, bytes / grid / (ms * 1e-3 * 1.9e9), cudaGetErrorString(err));
}

int main() {
  const int nchunks = 36;  // ~1.15 MB of 32 KB chunks
  uint8_t* blob; unsigned long long* sink;
  cudaMalloc(&blob, (size_t)nchunks * 32768);
  cudaMemset(blob, 1, (size_t)nchunks * 32768);
  cudaMalloc(&sink, 4096 * 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<2, 32768>(blob, nchunks, 400, sink, sms);
  run<4, 32768>(blob, nchunks, 400, sink, sms);
  run<6, 32768>(blob, nchunks, 400, sink, sms);
  run<4, 16384>(blob, nchunks * 2, 400, sink, sms);
  run<8, 16384>(blob, nchunks * 2, 400, sink, sms);
  run<4, 32768>(blob, nchunks, 400, sink, sms / 2);
  return 0;
}
