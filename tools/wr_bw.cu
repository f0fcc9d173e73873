// Write-bandwidth probe: (a) cudaMemset, (b) coalesced st.global.v4 grid-stride, (c) 16 KB cp.async.bulk
// shared->global stores from one CTA per SM (k groups in flight), (d) same with 4 CTAs worth of issue threads.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k_st(uint4* p, size_t n) {
  uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
template <int kInflight>
__global__ void k_bulk(uint8_t* p, size_t nblk, uint32_t blk) {   // nblk blocks of blk bytes
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    int k = 0;
    for (size_t b = blockIdx.x; b < nblk; b += gridDim.x, ++k) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + b * blk), "r"(s + (uint32_t)((k & 3) * 16384) % (65536 - blk + 1)), "r"(blk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kInflight) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  size_t bytes = 8ull << 30;
  uint8_t* p; cudaMalloc(&p, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  auto report = [&](const char* name) { cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %8.3f ms  %7.1f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError())); };
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); cudaMemsetAsync(p, 1, bytes); report("cudaMemset");
    cudaEventRecord(e0); k_st<<<148 * 8, 256>>>((uint4*)p, bytes / 16); report("st.global.v4 148x8 CTAs");
    cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_bulk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEventRecord(e0); k_bulk<1><<<148, 128, 65536>>>(p, bytes / 16384, 16384); report("bulk s2g 16KB, 148 CTAs, 2 in flight");
    cudaEventRecord(e0); k_bulk<4><<<148, 128, 65536>>>(p, bytes / 16384, 16384); report("bulk s2g 16KB, 148 CTAs, 5 in flight");
    cudaEventRecord(e0); k_bulk<16><<<148, 128, 65536>>>(p, bytes / 16384, 16384); report("bulk s2g 16KB, 148 CTAs, 17 in flight");
    cudaEventRecord(e0); k_bulk<16><<<148, 128, 65536>>>(p, bytes / 65536, 65536); report("bulk s2g 64KB, 148 CTAs, 17 in flight");
    cudaEventRecord(e0); k_bulk<4><<<296, 128, 65536>>>(p, bytes / 16384, 16384); report("bulk s2g 16KB, 296 CTAs, 5 in flight");
  }
  return 0;
}
