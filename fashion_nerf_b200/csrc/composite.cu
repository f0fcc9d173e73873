// Alpha compositing forward (A.5, raw2outputs) and backward (A.6).
//
// One warp per ray.  Samples are walked 32 at a time (lane = sample), so raw[R,S,4] is read as one
// float4 per lane (512 contiguous bytes per warp request) and z / weights as 128-byte rows.
// Transmittance is an exclusive product scan: a 5-step shuffle scan inside the 32-sample block
// and a scalar carry between blocks.  The kernels are HBM-bound: 24S+36 B/ray forward,
// 36S+24 B/ray backward.
#include "common.cuh"
#include "composite_math.cuh"

namespace fnerf {

constexpr int kCompWarps = 8;

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------ A.5
template <bool kHasNoise>
__global__ void __launch_bounds__(kCompWarps * 32, 8)     // latency-bound (one block in flight per warp): 32 registers, 8 CTAs per SM
k_composite_fwd(const float4* __restrict__ raw, const float* __restrict__ z,
                const float* __restrict__ dnorm, const float* __restrict__ noise,
                float* __restrict__ rgb_out, float* __restrict__ depth_out,
                float* __restrict__ acc_out, float* __restrict__ disp_out,
                float* __restrict__ weights, int64_t R, int S, int white) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kCompWarps;
  const int nblk = (S + 31) >> 5;

  for (int64_t r = warp_global; r < R; r += warp_stride) {
    const float4* rawr = raw + r * S;
    const float* zr = z + r * S;
    const float dn = dnorm[r];
    float carry = 1.0f;                       // transmittance entering this 32-sample block
    CompSums a = {0.f, 0.f, 0.f, 0.f, 0.f};

    // software pipeline: the next block's loads are issued before this block's scan
    int i = lane;
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    float zv = 0.f, nz = 0.f;
    if (i < S) { rv = ldg_stream4(rawr + i); zv = ldg_stream(zr + i); if (kHasNoise) nz = ldg_stream(noise + r * S + i); }
    for (int b = 0; b < nblk; ++b) {
      const int inext = i + 32;
      float4 rv_n = make_float4(0.f, 0.f, 0.f, 0.f);
      float zv_n = 0.f, nz_n = 0.f;
      if (inext < S) { rv_n = ldg_stream4(rawr + inext); zv_n = ldg_stream(zr + inext); if (kHasNoise) nz_n = ldg_stream(noise + r * S + inext); }

      // z of the following sample: next lane, or lane 0 of the next block
      float z_up = __shfl_down_sync(0xffffffffu, zv, 1);
      const float z_first_next = __shfl_sync(0xffffffffu, zv_n, 0);
      if (lane == 31) z_up = z_first_next;
      const bool valid = i < S;
      float dist = (i == S - 1) ? 1e10f : (z_up - zv);
      dist *= dn;
      float sigma = rv.w;
      if (kHasNoise) sigma += nz;
      const float alpha = comp_alpha(sigma, dist, valid);
      const float p = comp_scan(alpha, valid, lane);
      const float w = comp_weight(alpha, p, carry, lane);
      carry *= __shfl_sync(0xffffffffu, p, 31);

      if (valid) {
        comp_accum(a, w, sigmoidf_(rv.x), sigmoidf_(rv.y), sigmoidf_(rv.z), zv);
        if (weights != nullptr) weights[r * S + i] = w;
      }
      rv = rv_n; zv = zv_n; nz = nz_n; i = inext;
    }
    comp_finish(a, lane, r, white, rgb_out, depth_out, acc_out, disp_out);
  }
}

// Register-resident variant for S <= 32*NB (NB <= 8): every load of the ray is issued before the first
// scan step (NB float4 + NB floats in flight per lane), which is what an HBM-bound kernel needs; the
// software-pipelined kernel above keeps only one 32-sample block in flight per warp.
// A warp walks rays r0, r0 + W, r0 + 2W, ... (W = warps in the grid, so the device sweeps raw linearly) in passes of up to
// 32: after the j-th ray's warp reduction lane j keeps the totals, and the per-ray scalar tail
// (background, two IEEE divisions, six stores -- ~40 instructions that ran with one active lane per ray) is executed once
// per pass with one lane per ray.  Short rays (S <= 64) instead take passes over up to 32 CONSECUTIVE rays (coalesced map
// stores: 0.77 against 0.74 of HBM at S = 64); with longer rays every warp would stream its own 100 KB region and S = 192
// falls from 0.91 to 0.81.  `batch` is 32 for the strided form; for the consecutive form it shrinks with R so that a
// small launch (a 4096-ray training batch) still spreads over the device.  Same arithmetic per ray, same bits.
template <int NB, bool kHasNoise>
__global__ void __launch_bounds__(kCompWarps * 32)
k_composite_fwd_reg(const float4* __restrict__ raw, const float* __restrict__ z,
                    const float* __restrict__ dnorm, const float* __restrict__ noise, float* __restrict__ rgb_out,
                    float* __restrict__ depth_out, float* __restrict__ acc_out,
                    float* __restrict__ disp_out, float* __restrict__ weights, int64_t R, int S, int white, int batch) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kCompWarps;
  // short rays (NB <= 2): passes over `batch` CONSECUTIVE rays, long ones: rays W apart (see above)
  constexpr bool kConsec = NB <= 2;
  const int64_t rstep = kConsec ? 1 : warp_stride;
  for (int64_t r0 = kConsec ? warp_global * batch : warp_global; r0 < R; r0 += warp_stride * batch) {
    const int64_t left = (R - r0 + rstep - 1) / rstep;                  // rays r0, r0 + rstep, ... below R
    const int nr = (int)(left < batch ? left : batch);
    const float dn_l = dnorm[r0 + (lane < nr ? lane : 0) * rstep];      // |d| of this pass's lane-th ray
    CompSums keep = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int j = 0; j < nr; ++j) {
      const int64_t r = r0 + j * rstep;
      const float4* rawr = raw + r * S;
      const float* zr = z + r * S;
      float4 rv[NB];
      float zv[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i = b * 32 + lane;
        rv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        zv[b] = 0.f;
        if (i < S) {
          rv[b] = ldg_stream4(rawr + i); zv[b] = ldg_stream(zr + i);
          if (kHasNoise) rv[b].w += ldg_stream(noise + r * S + i);     // sigma = raw[...,3] + raw_noise (A.5)
        }
      }
      const float dn = __shfl_sync(0xffffffffu, dn_l, j);
      float carry = 1.0f;
      CompSums a = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i = b * 32 + lane;
        if (b * 32 < S) {                         // warp-uniform
          float z_up = __shfl_down_sync(0xffffffffu, zv[b], 1);
          const float z_first_next = (b + 1 < NB) ? __shfl_sync(0xffffffffu, zv[(b + 1 < NB) ? b + 1 : b], 0) : 0.f;
          if (lane == 31) z_up = z_first_next;
          const bool valid = i < S;
          float dist = (i == S - 1) ? 1e10f : (z_up - zv[b]);
          dist *= dn;
          const float alpha = comp_alpha(rv[b].w, dist, valid);
          const float p = comp_scan(alpha, valid, lane);
          const float w = comp_weight(alpha, p, carry, lane);
          carry *= __shfl_sync(0xffffffffu, p, 31);
          if (valid) {
            comp_accum(a, w, sigmoidf_(rv[b].x), sigmoidf_(rv[b].y), sigmoidf_(rv[b].z), zv[b]);
            if (weights != nullptr) weights[r * S + i] = w;
          }
        }
      }
      comp_reduce(a);
      if (lane == j) keep = a;
    }
    if (lane < nr) comp_store(keep, r0 + lane * rstep, white, rgb_out, depth_out, acc_out, disp_out);
  }
}

int launch_composite_fwd(const float* raw, const float* z, const float* dnorm, const float* noise,
                         float* rgb, float* depth, float* acc, float* disp, float* weights,
                         int64_t R, int64_t S, int white, cudaStream_t s) {
  if (R == 0) return 0;
  int64_t blocks = (R + kCompWarps - 1) / kCompWarps;
  const int64_t cap = (int64_t)num_sms() * 8 * 4;   // 8 resident CTAs/SM x 4 waves, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (S <= 256) {
    const float4* raw4 = (const float4*)raw;
    int64_t batch = 32;
    if (S <= 64) {                                                    // consecutive passes (NB <= 2)
      batch = R / ((int64_t)num_sms() * 64);                          // 32 once 64 warps per SM have a full pass
      batch = batch < 1 ? 1 : (batch > 32 ? 32 : batch);
      blocks = (R + batch * kCompWarps - 1) / (batch * kCompWarps);
      if (blocks > cap) blocks = cap;
    }
    const unsigned g = (unsigned)blocks, t = kCompWarps * 32;
#define FN_FWD(NB) do { if (noise) k_composite_fwd_reg<NB, true><<<g, t, 0, s>>>(raw4, z, dnorm, noise, rgb, depth, acc, disp, weights, R, (int)S, white, (int)batch); \
                        else k_composite_fwd_reg<NB, false><<<g, t, 0, s>>>(raw4, z, dnorm, noise, rgb, depth, acc, disp, weights, R, (int)S, white, (int)batch); } while (0)
    switch ((S + 31) / 32) {
      case 1: FN_FWD(1); break;
      case 2: FN_FWD(2); break;
      case 3: case 4: FN_FWD(4); break;
      case 5: case 6: FN_FWD(6); break;
      default: FN_FWD(8); break;
    }
    return check_launch("composite_fwd");
  }
  if (noise != nullptr)
    k_composite_fwd<true><<<(unsigned)blocks, kCompWarps * 32, 0, s>>>(
        (const float4*)raw, z, dnorm, noise, rgb, depth, acc, disp, weights, R, (int)S, white);
  else
    k_composite_fwd<false><<<(unsigned)blocks, kCompWarps * 32, 0, s>>>(
        (const float4*)raw, z, dnorm, nullptr, rgb, depth, acc, disp, weights, R, (int)S, white);
  return check_launch("composite_fwd");
}

// ------------------------------------------------------------------------------------------ A.6
// Pass 1 (forward over the ray) computes T_i and w_i*v_i and parks them in shared memory;
// pass 2 (backward over the ray) runs the exclusive suffix sum S_i = sum_{k>i} w_k v_k as a true
// reverse scan and writes g_raw.  raw/z are re-read in pass 2 (L1/L2 hits; DRAM traffic is one
// read of raw,z and one write of g_raw).
constexpr int kBwdWarps = 4;

__global__ void __launch_bounds__(kBwdWarps * 32)
k_composite_bwd(const float4* __restrict__ raw, const float* __restrict__ z,
                const float* __restrict__ dnorm, const float* __restrict__ noise, const float* __restrict__ g_rgb,
                const float* __restrict__ g_depth, const float* __restrict__ g_acc,
                float4* __restrict__ g_raw, int64_t R, int S, int white) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_T = smem + (size_t)warp * 2 * S;
  float* s_wv = s_T + S;
  const int nblk = (S + 31) >> 5;

  for (int64_t r = (int64_t)blockIdx.x * kBwdWarps + warp; r < R; r += (int64_t)gridDim.x * kBwdWarps) {
    const float4* rawr = raw + r * S;
    const float* zr = z + r * S;
    const float dn = dnorm[r];
    const float gr = g_rgb[3 * r], gg = g_rgb[3 * r + 1], gb = g_rgb[3 * r + 2];
    const float gd = g_depth ? g_depth[r] : 0.0f;
    float ga = g_acc ? g_acc[r] : 0.0f;
    if (white) ga -= (gr + gg + gb);

    float carry = 1.0f;
    for (int b = 0; b < nblk; ++b) {
      const int i = b * 32 + lane;
      const bool valid = i < S;
      float4 rv = valid ? rawr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && noise != nullptr) rv.w += noise[r * S + i];
      const float zv = valid ? zr[i] : 0.f;
      const float z_up = (i + 1 < S) ? zr[i + 1] : 0.f;
      float dist = (i == S - 1) ? 1e10f : (z_up - zv);
      dist *= dn;
      const float alpha = valid ? (1.0f - expf(-fmaxf(rv.w, 0.0f) * dist)) : 0.0f;
      const float om = valid ? (1.0f - alpha + 1e-10f) : 1.0f;
      float p = om;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, p, o);
        if (lane >= o) p *= n;
      }
      float excl = __shfl_up_sync(0xffffffffu, p, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry *= __shfl_sync(0xffffffffu, p, 31);
      if (valid) {
        const float v = gr * sigmoidf_(rv.x) + gg * sigmoidf_(rv.y) + gb * sigmoidf_(rv.z) + gd * zv + ga;
        s_T[i] = T;
        s_wv[i] = alpha * T * v;
      }
    }
    __syncwarp();

    float tail = 0.0f;                        // sum of w*v over all later blocks
    for (int b = nblk - 1; b >= 0; --b) {
      const int i = b * 32 + lane;
      const bool valid = i < S;
      const float wv = valid ? s_wv[i] : 0.0f;
      float q = wv;                           // inclusive suffix scan inside the block
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_down_sync(0xffffffffu, q, o);
        if (lane + o < 32) q += n;
      }
      const float suffix = (q - wv) + tail;   // exclusive
      tail += __shfl_sync(0xffffffffu, q, 0);
      if (valid) {
        float4 rv = rawr[i];
        if (noise != nullptr) rv.w += noise[r * S + i];
        const float zv = zr[i];
        const float z_up = (i + 1 < S) ? zr[i + 1] : 0.f;
        float dist = (i == S - 1) ? 1e10f : (z_up - zv);
        dist *= dn;
        const float alpha = 1.0f - expf(-fmaxf(rv.w, 0.0f) * dist);
        const float om = 1.0f - alpha + 1e-10f;
        const float T = s_T[i];
        const float w = alpha * T;
        const float cr = sigmoidf_(rv.x), cg = sigmoidf_(rv.y), cb = sigmoidf_(rv.z);
        const float v = gr * cr + gg * cg + gb * cb + gd * zv + ga;
        const float g_alpha = T * v - suffix / om;
        const float g_sigma = (rv.w > 0.0f) ? dist * (1.0f - alpha) * g_alpha : 0.0f;
        g_raw[r * S + i] = make_float4(w * gr * cr * (1.0f - cr), w * gg * cg * (1.0f - cg),
                                       w * gb * cb * (1.0f - cb), g_sigma);
      }
    }
    __syncwarp();
  }
}

// Register-resident backward for S <= 32*NB: raw and z are read ONCE (all loads up front), the forward
// scan keeps alpha / T / w*v per block in registers, the reverse scan writes g_raw.  Same arithmetic
// as k_composite_bwd above.
template <int NB, bool kHasNoise>
__global__ void __launch_bounds__(kCompWarps * 32)
k_composite_bwd_reg(const float4* __restrict__ raw, const float* __restrict__ z,
                    const float* __restrict__ dnorm, const float* __restrict__ noise, const float* __restrict__ g_rgb,
                    const float* __restrict__ g_depth, const float* __restrict__ g_acc,
                    float4* __restrict__ g_raw, int64_t R, int S, int white) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kCompWarps;
  for (int64_t r = warp_global; r < R; r += warp_stride) {
    const float4* rawr = raw + r * S;
    const float* zr = z + r * S;
    float4 rv[NB];
    float zv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = b * 32 + lane;
      rv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      zv[b] = 0.f;
      if (i < S) {
        rv[b] = ldg_stream4(rawr + i); zv[b] = ldg_stream(zr + i);
        if (kHasNoise) rv[b].w += ldg_stream(noise + r * S + i);
      }
    }
    const float dn = dnorm[r];
    const float gr = g_rgb[3 * r], gg = g_rgb[3 * r + 1], gb = g_rgb[3 * r + 2];
    const float gd = g_depth ? g_depth[r] : 0.0f;
    float ga = g_acc ? g_acc[r] : 0.0f;
    if (white) ga -= (gr + gg + gb);

    // the sigmoids of the forward pass are kept for the reverse pass when the register budget allows (short rays are
    // instruction-issue bound: 18 fewer instructions per 32 samples)
    constexpr bool kKeepRgb = NB <= 4;
    float Tb[NB], alb[NB], distb[NB], vb[NB];
    [[maybe_unused]] float crb[kKeepRgb ? NB : 1], cgb[kKeepRgb ? NB : 1], cbb[kKeepRgb ? NB : 1];
    float carry = 1.0f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = b * 32 + lane;
      Tb[b] = 0.f; alb[b] = 0.f; distb[b] = 0.f; vb[b] = 0.f;
      if (b * 32 < S) {
        float z_up = __shfl_down_sync(0xffffffffu, zv[b], 1);
        const float z_first_next = (b + 1 < NB) ? __shfl_sync(0xffffffffu, zv[(b + 1 < NB) ? b + 1 : b], 0) : 0.f;
        if (lane == 31) z_up = z_first_next;
        const bool valid = i < S;
        float dist = (i == S - 1) ? 1e10f : (z_up - zv[b]);
        dist *= dn;
        const float alpha = valid ? (1.0f - expf(-fmaxf(rv[b].w, 0.0f) * dist)) : 0.0f;
        float p = valid ? (1.0f - alpha + 1e-10f) : 1.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, p, o);
          if (lane >= o) p *= n;
        }
        float excl = __shfl_up_sync(0xffffffffu, p, 1);
        if (lane == 0) excl = 1.0f;
        Tb[b] = carry * excl;
        carry *= __shfl_sync(0xffffffffu, p, 31);
        alb[b] = alpha;
        distb[b] = dist;
        const float cr = sigmoidf_(rv[b].x), cg = sigmoidf_(rv[b].y), cb = sigmoidf_(rv[b].z);
        if (kKeepRgb) { crb[b] = cr; cgb[b] = cg; cbb[b] = cb; }
        vb[b] = valid ? (gr * cr + gg * cg + gb * cb + gd * zv[b] + ga) : 0.0f;
      }
    }
    float tail = 0.0f;
#pragma unroll
    for (int b = NB - 1; b >= 0; --b) {
      const int i = b * 32 + lane;
      if (b * 32 < S) {
        const bool valid = i < S;
        const float wv = valid ? alb[b] * Tb[b] * vb[b] : 0.0f;
        float q = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_down_sync(0xffffffffu, q, o);
          if (lane + o < 32) q += n;
        }
        const float suffix = (q - wv) + tail;
        tail += __shfl_sync(0xffffffffu, q, 0);
        if (valid) {
          const float om = 1.0f - alb[b] + 1e-10f;
          const float w = alb[b] * Tb[b];
          const float cr = kKeepRgb ? crb[b] : sigmoidf_(rv[b].x), cg = kKeepRgb ? cgb[b] : sigmoidf_(rv[b].y),
                      cb = kKeepRgb ? cbb[b] : sigmoidf_(rv[b].z);
          const float g_alpha = Tb[b] * vb[b] - __fdividef(suffix, om);
          const float g_sigma = (rv[b].w > 0.0f) ? distb[b] * (1.0f - alb[b]) * g_alpha : 0.0f;
          g_raw[r * S + i] = make_float4(w * gr * cr * (1.0f - cr), w * gg * cg * (1.0f - cg),
                                         w * gb * cb * (1.0f - cb), g_sigma);
        }
      }
    }
  }
}

// Backward for 256 < S <= 1024: ceil(S / 128) warps share a ray, each holds one 128-sample chunk in registers exactly like
// k_composite_bwd_reg<4> (raw and z read ONCE, nothing recomputed), and the chunks exchange two scalars through shared
// memory: the product of (1 - alpha + 1e-10) over a chunk (transmittance entering the later chunks; T enters every term
// linearly, so a chunk scans with carry 1 and scales afterwards) and the chunk's sum of w*v (the suffix sum entering the
// earlier chunks).  The shared-memory kernel above re-reads raw and z in its second pass (ncu at S = 1024: 43 % more DRAM
// reads than algorithmic, 8.5 k warp instructions per ray, 0.37 of HBM) and stays for S > 1024.
constexpr int kLongNB = 4;                                   // 32-sample blocks per warp (8 or 2: slower, 3.1 / 3.6 vs 4.1 TB/s at S = 1024)
constexpr int kLongMaxWarps = 8;                             // warps per ray = ceil(S / 128) <= 8
template <int kLongWarps, bool kHasNoise>
__global__ void __launch_bounds__(kLongWarps * 32)
k_composite_bwd_long(const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ dnorm, const float* __restrict__ noise, const float* __restrict__ g_rgb,
                     const float* __restrict__ g_depth, const float* __restrict__ g_acc,
                     float4* __restrict__ g_raw, int64_t R, int S, int white) {
  constexpr int NB = kLongNB;
  __shared__ float s_prod[2][kLongWarps], s_sum[2][kLongWarps];       // double-buffered by ray parity: one barrier per ray
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = warp * (NB * 32);
  int par = 0;
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x, par ^= 1) {
    const float4* rawr = raw + r * S;
    const float* zr = z + r * S;
    float4 rv[NB];
    float zv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = base + b * 32 + lane;
      rv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      zv[b] = 0.f;
      if (i < S) {
        rv[b] = ldg_stream4(rawr + i); zv[b] = ldg_stream(zr + i);
        if (kHasNoise) rv[b].w += ldg_stream(noise + r * S + i);
      }
    }
    const float z_chunk_next = (base + NB * 32 < S) ? zr[base + NB * 32] : 0.f;   // first depth of the next chunk
    const float dn = dnorm[r];
    const float gr = g_rgb[3 * r], gg = g_rgb[3 * r + 1], gb = g_rgb[3 * r + 2];
    const float gd = g_depth ? g_depth[r] : 0.0f;
    float ga = g_acc ? g_acc[r] : 0.0f;
    if (white) ga -= (gr + gg + gb);

    float Tb[NB], alb[NB], distb[NB], vb[NB];
    float carry = 1.0f;                            // relative to the chunk's first sample
    float wsum = 0.0f;                             // this lane's share of sum alpha * T_rel * v
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = base + b * 32 + lane;
      Tb[b] = 0.f; alb[b] = 0.f; distb[b] = 0.f; vb[b] = 0.f;
      if (base + b * 32 < S) {                     // warp-uniform
        float z_up = __shfl_down_sync(0xffffffffu, zv[b], 1);
        const float z_first_next = (b + 1 < NB) ? __shfl_sync(0xffffffffu, zv[(b + 1 < NB) ? b + 1 : b], 0) : z_chunk_next;
        if (lane == 31) z_up = z_first_next;
        const bool valid = i < S;
        float dist = (i == S - 1) ? 1e10f : (z_up - zv[b]);
        dist *= dn;
        const float alpha = valid ? (1.0f - expf(-fmaxf(rv[b].w, 0.0f) * dist)) : 0.0f;
        float p = valid ? (1.0f - alpha + 1e-10f) : 1.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, p, o);
          if (lane >= o) p *= n;
        }
        float excl = __shfl_up_sync(0xffffffffu, p, 1);
        if (lane == 0) excl = 1.0f;
        Tb[b] = carry * excl;
        carry *= __shfl_sync(0xffffffffu, p, 31);
        alb[b] = alpha;
        distb[b] = dist;
        vb[b] = valid ? (gr * sigmoidf_(rv[b].x) + gg * sigmoidf_(rv[b].y) + gb * sigmoidf_(rv[b].z) + gd * zv[b] + ga) : 0.0f;
        wsum += alpha * Tb[b] * vb[b];
      }
    }
    wsum = warp_sum(wsum);
    if (lane == 0) { s_prod[par][warp] = carry; s_sum[par][warp] = wsum; }
    asm volatile("bar.sync 2, %0;" ::"n"(kLongWarps * 32) : "memory");
    float cin = 1.0f;                              // transmittance entering this chunk
    float tail = 0.0f;                             // sum of w*v over the later chunks
    {
      float run = 1.0f;                            // transmittance entering chunk w2
#pragma unroll
      for (int w2 = 0; w2 < kLongWarps; ++w2) {
        if (w2 == warp) cin = run;
        if (w2 > warp) tail += run * s_sum[par][w2];
        run *= s_prod[par][w2];
      }
    }
#pragma unroll
    for (int b = NB - 1; b >= 0; --b) {
      const int i = base + b * 32 + lane;
      if (base + b * 32 < S) {
        const bool valid = i < S;
        const float T = cin * Tb[b];
        const float wv = valid ? alb[b] * T * vb[b] : 0.0f;
        float q = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_down_sync(0xffffffffu, q, o);
          if (lane + o < 32) q += n;
        }
        const float suffix = (q - wv) + tail;
        tail += __shfl_sync(0xffffffffu, q, 0);
        if (valid) {
          const float om = 1.0f - alb[b] + 1e-10f;
          const float w = alb[b] * T;
          const float cr = sigmoidf_(rv[b].x), cg = sigmoidf_(rv[b].y), cb = sigmoidf_(rv[b].z);
          const float g_alpha = T * vb[b] - __fdividef(suffix, om);
          const float g_sigma = (rv[b].w > 0.0f) ? distb[b] * (1.0f - alb[b]) * g_alpha : 0.0f;
          g_raw[r * S + i] = make_float4(w * gr * cr * (1.0f - cr), w * gg * cg * (1.0f - cg),
                                         w * gb * cb * (1.0f - cb), g_sigma);
        }
      }
    }
  }
}

int launch_composite_bwd(const float* raw, const float* z, const float* dnorm, const float* noise, const float* g_rgb,
                         const float* g_depth, const float* g_acc, float* g_raw, int64_t R,
                         int64_t S, int white, cudaStream_t s) {
  if (R == 0) return 0;
  if (S <= 256) {
    int64_t nb = (R + kCompWarps - 1) / kCompWarps;
    const int64_t capr = (int64_t)num_sms() * 8 * 4;
    if (nb > capr) nb = capr;
    const unsigned g = (unsigned)nb, t = kCompWarps * 32;
    const float4* raw4 = (const float4*)raw;
    float4* g4 = (float4*)g_raw;
#define FN_BWD(NB) do { if (noise) k_composite_bwd_reg<NB, true><<<g, t, 0, s>>>(raw4, z, dnorm, noise, g_rgb, g_depth, g_acc, g4, R, (int)S, white); \
                        else k_composite_bwd_reg<NB, false><<<g, t, 0, s>>>(raw4, z, dnorm, noise, g_rgb, g_depth, g_acc, g4, R, (int)S, white); } while (0)
    switch ((S + 31) / 32) {
      case 1: FN_BWD(1); break;
      case 2: FN_BWD(2); break;
      case 3: case 4: FN_BWD(4); break;
      case 5: case 6: FN_BWD(6); break;
      default: FN_BWD(8); break;
    }
    return check_launch("composite_bwd");
  }
  if (S <= kLongMaxWarps * kLongNB * 32) {
    int64_t nb = R;
    const int64_t capl = (int64_t)num_sms() * 32;
    if (nb > capl) nb = capl;
    const int W = (int)((S + kLongNB * 32 - 1) / (kLongNB * 32));     // 3..8
#define FN_LONG(WW) case WW: \
      if (noise) k_composite_bwd_long<WW, true><<<(unsigned)nb, WW * 32, 0, s>>>((const float4*)raw, z, dnorm, noise, g_rgb, g_depth, g_acc, (float4*)g_raw, R, (int)S, white); \
      else k_composite_bwd_long<WW, false><<<(unsigned)nb, WW * 32, 0, s>>>((const float4*)raw, z, dnorm, noise, g_rgb, g_depth, g_acc, (float4*)g_raw, R, (int)S, white); \
      break
    switch (W) { FN_LONG(3); FN_LONG(4); FN_LONG(5); FN_LONG(6); FN_LONG(7); default: FN_LONG(8); }
#undef FN_LONG
    return check_launch("composite_bwd");
  }
  const size_t smem = (size_t)kBwdWarps * 2 * S * sizeof(float);
  if (smem > 200 * 1024) return set_error(FNERF_ERR_SIZE, "composite_bwd: S too large");
  static DeviceOnce once;                            // opt in once to the 200 KB cap checked above
  if (smem > 48 * 1024)
    if (cudaError_t e = opt_in_smem_once(once, k_composite_bwd, 200 * 1024)) return set_error((int)e, "composite_bwd: %s", cudaGetErrorString(e));
  int64_t blocks = (R + kBwdWarps - 1) / kBwdWarps;
  const int64_t cap = (int64_t)num_sms() * 64;
  if (blocks > cap) blocks = cap;
  k_composite_bwd<<<(unsigned)blocks, kBwdWarps * 32, smem, s>>>(
      (const float4*)raw, z, dnorm, noise, g_rgb, g_depth, g_acc, (float4*)g_raw, R, (int)S, white);
  return check_launch("composite_bwd");
}

}  // namespace fnerf
