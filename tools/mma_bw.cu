// Micro-benchmark: tcgen05.mma (SS, bf16, M=128) issue rate per SM for N=256 / N=128, with and
// without a concurrent bulk-TMA stream into shared memory and a concurrent st.shared epilogue-like
// stream.  Answers: is the fused MLP limited by shared-memory bandwidth?
#include <cuda_runtime.h>
#include <stdio.h>
#include "../fashion_nerf_b200/csrc/tc_ptx.cuh"
using namespace fnerf::ptx;

template <int N, bool TMA, bool STS>
__global__ void __launch_bounds__(192, 1) k(const uint8_t* blob, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tslot;
  __shared__ int done;
  if (threadIdx.x == 0) done = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tslot), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t a_addr = base, b_addr = base + 65536, w_addr = base + 65536 + 32768;  // A 64K, B 32K, TMA dst 4x16K
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + (it & 1) * 256, umma_desc_sw128(a_addr + kb * 16384 + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc, (kb | ks) ? 1u : 0u);
      const int b = (it & 1) ? 5 : 0;                 // two barriers, alternating: no parity aliasing
      umma_commit(smem_u32(&bars[b]));
      if (it >= 1) mbar_wait(smem_u32(&bars[((it - 1) & 1) ? 5 : 0]), ((it - 1) >> 1) & 1);
    }
    mbar_wait(smem_u32(&bars[((iters - 1) & 1) ? 5 : 0]), ((iters - 1) >> 1) & 1);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    *(volatile int*)&done = 1;
  } else if (warp == 0 && lane == 0 && TMA) {
    // stream 16 KB chunks as fast as allowed: 4 stages, wait for own completion only
    const int total = iters * 8;   // 8 x 16 KB per 16 MMAs(N=256) == the MLP's 64 B/clk at full rate
    for (int c = 0; c < total; ++c) {
      const int s = c & 3;
      if (c >= 4) mbar_wait(smem_u32(&bars[1 + s]), ((c >> 2) - 1) & 1);
      mbar_expect_tx(smem_u32(&bars[1 + s]), 16384);
      bulk_g2s(w_addr + s * 16384, blob + (size_t)(c % 64) * 16384, 16384, smem_u32(&bars[1 + s]));
    }
    for (int s = 0; s < 4; ++s) mbar_wait(smem_u32(&bars[1 + s]), (((total - 4 + s) >> 2)) & 1);
  } else if (warp >= 2 && STS) {
    // epilogue-like traffic: each thread writes 64 x 16 B per "layer" (64 KB per CTA per iteration)
    // run until the MMA thread is done; count 16-byte stores per thread
    const uint32_t row = (warp - 2) * 32 + lane;
    long long n = 0;
    while (*(volatile int*)&done == 0) {
      for (int c = 0; c < 32; ++c)
        st_shared_v4(w_addr + 65536 + row * 128 + (((c & 7) ^ (row & 7)) << 4), (uint32_t)n, c, row, 0);
      n += 32;
    }
    if (lane == 0 && warp == 2) out[148 + blockIdx.x] = n;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TMA, bool STS>
void run(const uint8_t* blob, long long* out, const char* tag) {
  const int smem = 65536 + 32768 + 65536 + 16384 + 2048;
  cudaFuncSetAttribute(k<N, TMA, STS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  k<N, TMA, STS><<<148, 192, smem>>>(blob, 100, out);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<N, TMA, STS><<<148, 192, smem>>>(blob, iters, out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0, sts = 0; for (int i = 0; i < 148; ++i) { cyc += h[i]; sts += h[148 + i]; } cyc /= 148; sts /= 148;
  if (STS) printf("   sts: %.1f B/clk/SM from 4 warps (128 threads x 16 B x %.0f stores / %.0f cycles)\n", sts * 128 * 16 / cyc, sts, cyc);
  double flops = 148.0 * iters * 16 * 2.0 * 128 * N * 16;
  printf("{\"bench\":\"%s\",\"N\":%d,\"tma\":%d,\"sts\":%d,\"ms\":%.3f,\"TFLOPs\":%.1f,\"cycles_per_mma\":%.1f,\"err\":\"%s\"}\n", tag, N, (int)TMA, (int)STS, ms,
         flops / ms / 1e9, cyc / (iters * 16.0), cudaGetErrorString(err));
}

int main() {
  uint8_t* blob; long long* out;
  cudaMalloc(&blob, 64 * 16384); cudaMemset(blob, 0, 64 * 16384);
  cudaMalloc(&out, 296 * 8); cudaMemset(out, 0, 296 * 8);
  run<256, false, false>(blob, out, "mma_only");
  run<128, false, false>(blob, out, "mma_only");
  run<256, true, false>(blob, out, "mma+tma");
  run<128, true, false>(blob, out, "mma+tma");
  run<256, false, true>(blob, out, "mma+sts");
  run<256, true, true>(blob, out, "mma+tma+sts");
  run<128, true, true>(blob, out, "mma+tma+sts");
  return 0;
}
