// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM for several shapes and warp counts.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../fashion_nerf_b200/csrc/tc_ptx.cuh"
using namespace fnerf::ptx;

__device__ __forceinline__ void ld_32x32b_x16(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_32x32b_x64(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]),
      "=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]),
      "=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63]) : "r"(a) : "memory");
}
// 16 lanes x 256 bits: each thread gets 4 regs per repetition (.x8 -> 32 regs = 16 lanes x 64 columns)
__device__ __forceinline__ void ld_16x256b_x8(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(a) : "memory");
}

// MODE 0: 32x32b.x32 ; 1: 32x32b.x16 ; 2: 32x32b.x64 ; 3: 16x256b.x8 (two per 32 lanes)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, long long* out, unsigned* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t q = warp & 3, grp = warp >> 2;
  unsigned acc = 0;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t lanebase = tmem + ((q * 32u) << 16);
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = ((grp * 64u) + (it & 1) * 32u) & 511u;
      if (MODE == 0) { uint32_t v[32]; tmem_ld32(lanebase + col, v); tmem_ld_wait(); acc += v[0] + v[31]; }
      if (MODE == 1) { uint32_t v[16]; ld_32x32b_x16(lanebase + col, v); ld_32x32b_x16(lanebase + col + 16, v); tmem_ld_wait(); acc += v[0] + v[15]; }
      if (MODE == 2) { uint32_t v[64]; ld_32x32b_x64(lanebase + (col & 448u), v); tmem_ld_wait(); acc += v[0] + v[63]; }
      if (MODE == 3) { uint32_t v[32]; ld_16x256b_x8(lanebase + col, v); tmem_ld_wait(); acc += v[0] + v[31];
                       ld_16x256b_x8(lanebase + (16u << 16) + col, v); tmem_ld_wait(); acc += v[0] + v[31]; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 512 + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
void run(int nwarps, long long* out, unsigned* sink, const char* name, int bytes_per_iter_per_warp) {
  const int iters = 20000;
  k<MODE><<<148, 512>>>(nwarps, 100, out, sink);
  cudaDeviceSynchronize();
  k<MODE><<<148, 512>>>(nwarps, iters, out, sink);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  printf("{\"shape\":\"%s\",\"warps\":%d,\"cycles_per_iter\":%.1f,\"B_per_clk_per_SM\":%.1f,\"err\":\"%s\"}\n", name, nwarps, cyc / iters,
         (double)bytes_per_iter_per_warp * nwarps * iters / cyc, cudaGetErrorString(err));
}

int main() {
  long long* out; unsigned* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  for (int nw : {1, 4, 8, 16}) {
    run<0>(nw, out, sink, "32x32b.x32", 4096);
    run<1>(nw, out, sink, "32x32b.x16 (x2)", 4096);
    run<2>(nw, out, sink, "32x32b.x64", 8192);
    run<3>(nw, out, sink, "16x256b.x8 (x2)", 8192);
  }
  return 0;
}
