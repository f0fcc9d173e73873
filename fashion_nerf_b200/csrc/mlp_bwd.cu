// MLP backward (A.4/A.8 gradients w.r.t. the parameters), fp32 SIMT path of ABI v1.
//
// Layer-major schedule over chunks of samples (the shape wgrad wants: dW accumulates over many
// samples before it is flushed): per chunk the activations are recomputed and parked in the
// caller's workspace, then every layer runs dgrad and wgrad as tiled SGEMMs.
//   forward : Y = act(X . W^T + b)          NN gemm against the K-major weights (packed section C)
//   dgrad   : dX = (dZ . W) [+ g_sigma (x) w_alpha] (.) [X > 0]      NT gemm, mask fused
//   wgrad   : dW += dZ^T . X                TN gemm, split over samples, atomicAdd into flat_grad
// Inputs (encodings, cond) carry no gradient (A.4); sample positions are detached (A.7).
// flat_grad uses the flat nn.Linear layout of include/fnerf.h, i.e. it is directly the buffer the
// data-parallel step all-reduces.
#include "common.cuh"

namespace fnerf {

constexpr int64_t kBwdChunk = 32768;   // samples per chunk
constexpr int kBM = 128, kBN = 64, kBK = 16, kGemmThreads = 256;

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int I, J; int64_t K;
  // forward epilogue
  const float* bias; int relu;
  // dgrad epilogue
  const float* mask; int64_t ldmask;           // multiply by (mask[i][j] > 0) when non-null
  const float* rank1_col; int64_t rank1_ld;    // + rank1_col[i*rank1_ld] * rank1_row[j] when non-null
  const float* rank1_row;
  int64_t k_split;                             // wgrad: K rows per blockIdx.z
};

// C[I,J] (+)= sum_k A(i,k) B(k,j);  A(i,k) = TA ? A[k*lda+i] : A[i*lda+k];  B(k,j) = TB ? B[j*ldb+k] : B[k*ldb+j]
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(kGemmThreads) k_sgemm(const GemmArgs g) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * kBM, j0 = blockIdx.x * kBN;
  int64_t kbeg = 0, kend = g.K;
  if (EPI == EPI_WGRAD) {
    kbeg = (int64_t)blockIdx.z * g.k_split;
    kend = kbeg + g.k_split < g.K ? kbeg + g.k_split : g.K;
  }
  const int ty = tid / 16, tx = tid % 16;      // 16 x 16 threads, 8 x 4 outputs each
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += kBK) {
    // ---- stage A tile [kBK][kBM]
    if (TA) {       // A[k*lda + i]: contiguous along i
      for (int e = tid; e < kBK * kBM; e += kGemmThreads) {
        const int kk = e / kBM, ii = e % kBM;
        const int64_t k = k0 + kk; const int i = i0 + ii;
        As[kk][ii] = (k < kend && i < g.I) ? g.A[k * g.lda + i] : 0.0f;
      }
    } else {        // A[i*lda + k]: contiguous along k
      for (int e = tid; e < kBK * kBM; e += kGemmThreads) {
        const int ii = e / kBK, kk = e % kBK;
        const int64_t k = k0 + kk; const int i = i0 + ii;
        As[kk][ii] = (k < kend && i < g.I) ? g.A[(int64_t)i * g.lda + k] : 0.0f;
      }
    }
    // ---- stage B tile [kBK][kBN]
    if (TB) {       // B[j*ldb + k]
      for (int e = tid; e < kBK * kBN; e += kGemmThreads) {
        const int jj = e / kBK, kk = e % kBK;
        const int64_t k = k0 + kk; const int j = j0 + jj;
        Bs[kk][jj] = (k < kend && j < g.J) ? g.B[(int64_t)j * g.ldb + k] : 0.0f;
      }
    } else {        // B[k*ldb + j]
      for (int e = tid; e < kBK * kBN; e += kGemmThreads) {
        const int kk = e / kBN, jj = e % kBN;
        const int64_t k = k0 + kk; const int j = j0 + jj;
        Bs[kk][jj] = (k < kend && j < g.J) ? g.B[k * g.ldb + j] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + ty * 8 + a;
    if (i >= g.I) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + tx * 4 + b;
      if (j >= g.J) continue;
      float v = acc[a][b];
      float* c = g.C + (int64_t)i * g.ldc + j;
      if (EPI == EPI_FWD) {
        if (g.bias) v += g.bias[j];
        if (g.relu) v = fmaxf(v, 0.0f);
        *c = v;
      } else if (EPI == EPI_DGRAD) {
        if (g.rank1_col) v = fmaf(g.rank1_col[(int64_t)i * g.rank1_ld], g.rank1_row[j], v);
        if (g.mask) v = (g.mask[(int64_t)i * g.ldmask + j] > 0.0f) ? v : 0.0f;
        *c = v;
      } else {
        atomicAdd(c, v);
      }
    }
  }
}

template <bool TA, bool TB, int EPI>
static void gemm(const GemmArgs& g, cudaStream_t s) {
  dim3 grid((g.J + kBN - 1) / kBN, (g.I + kBM - 1) / kBM, 1);
  if (EPI == EPI_WGRAD) grid.z = (unsigned)((g.K + g.k_split - 1) / g.k_split);
  k_sgemm<TA, TB, EPI><<<grid, kGemmThreads, 0, s>>>(g);
  note_launches(1);
}

// column sums of dZ[M,N] (ld) added into out[N] (bias gradients)
__global__ void k_colsum(const float* __restrict__ dz, int64_t ld, int64_t M, int N, float* __restrict__ out, int64_t rows_per_block) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
  float acc = 0.0f;
  for (int64_t m = m0; m < m1; ++m) acc += dz[m * ld + j];
  atomicAdd(out + j, acc);
}
static void colsum(const float* dz, int64_t ld, int64_t M, int N, float* out, cudaStream_t s) {
  const int64_t rpb = 512;
  dim3 grid((N + 127) / 128, (unsigned)((M + rpb - 1) / rpb));
  k_colsum<<<grid, 128, 0, s>>>(dz, ld, M, N, out, rpb);
  note_launches(1);
}

// encodings (A.3) of a chunk of samples into the concatenated layer inputs:
//   X5[m] = [pe(63) | cond(256, optional) | (h4 written later by the L4 gemm)],  Xv[m] = [(feat) | pe_dir(27)]
__global__ void k_encode_chunk(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                               const float* __restrict__ viewdirs, const float* __restrict__ z,
                               const float* __restrict__ cond_rows, const int32_t* __restrict__ cond_index,
                               int64_t C, int S, int64_t m_begin, int64_t m_count, float* __restrict__ X5,
                               int64_t ld5, float* __restrict__ Xv, int64_t ldv) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / 3;
  const int c = (int)(t % 3);
  if (m >= m_count) return;
  const int64_t gidx = m_begin + m;
  const int64_t ray = gidx / S;
  const float p = __fadd_rn(rays_o[3 * ray + c], __fmul_rn(rays_d[3 * ray + c], z[gidx]));
  float* x5 = X5 + m * ld5;
  x5[c] = p;
  for (int k = 0; k < kLX; ++k) {
    float sn, cs;
    sincosf(__fmul_rn(p, (float)(1 << k)), &sn, &cs);
    x5[3 + 6 * k + c] = sn;
    x5[3 + 6 * k + 3 + c] = cs;
  }
  const float d = viewdirs[3 * ray + c];
  float* xv = Xv + m * ldv + kW;
  xv[c] = d;
  for (int k = 0; k < kLD; ++k) {
    float sn, cs;
    sincosf(__fmul_rn(d, (float)(1 << k)), &sn, &cs);
    xv[3 + 6 * k + c] = sn;
    xv[3 + 6 * k + 3 + c] = cs;
  }
  if (cond_rows != nullptr) {   // gather this sample's garment code into columns 63..318
    const int64_t row = cond_row(cond_index, C, ray);
    const float* src = cond_rows + row * kCond;
    for (int k = c; k < kCond; k += 3) x5[kPE + k] = src[k];
  }
}

// per-sample workspace floats (see launch_mlp_bwd_fp32 for the carve-up)
static int64_t ws_floats_per_sample(int cond) {
  const int in5 = kPE + (cond ? kCond : 0) + kW;
  return in5 + 7 * kW + (kW + kPED) + kWV + 2 * kW;
}
int64_t mlp_bwd_workspace_bytes(int64_t R, int64_t S) {
  int64_t M = R * S;
  if (M > kBwdChunk) M = kBwdChunk;
  if (M < 1) M = 1;
  return M * ws_floats_per_sample(1) * 4 + 256;   // sized for the conditioned variant
}

// NOTE: for the conditioned variant this entry needs the RAW codes (cond_rows [C,256]) in a.cond_proj.
int launch_mlp_bwd_fp32(const MlpArgs& a, const float* g_raw, float* flat_grad, void* ws, int64_t ws_bytes,
                        cudaStream_t s) {
  const int cond = a.cond;
  const int64_t M = a.R * a.S;
  const int in5 = kPE + (cond ? kCond : 0) + kW, hoff = in5 - kW, inv = kW + kPED;
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(a.packed);
  const float* aux = reinterpret_cast<const float*>(packed + kSecBOffset);
  const float* secC = reinterpret_cast<const float*>(packed + kSecCOffset);
  auto wt = [&](int j) { return secC + simt_offset_floats(j, cond); };       // K-major weights [in][out]
  auto gw = [&](int l) { return flat_grad + flat_weight_offset(l, cond); };   // grads, flat layout
  auto gb = [&](int l) { return flat_grad + flat_bias_offset(l, cond); };
  (void)ws_bytes;

  for (int64_t m0 = 0; m0 < M; m0 += kBwdChunk) {
    const int64_t mc = (M - m0 < kBwdChunk) ? (M - m0) : kBwdChunk;
    float* p = reinterpret_cast<float*>(ws);
    float* X5 = p;            p += mc * in5;       // [pe | cond | h4]
    float* H[8];
    for (int l = 0; l < 8; ++l) { if (l == 4) { H[4] = X5 + hoff; continue; } H[l] = p; p += mc * kW; }
    float* Xv = p;            p += mc * inv;       // [feat | pe_dir]
    float* HV = p;            p += mc * kWV;
    float* dA = p;            p += mc * kW;
    float* dB = p;            p += mc * kW;
    auto ldH = [&](int l) -> int64_t { return l == 4 ? in5 : kW; };
    const float* gr = g_raw + m0 * 4;

    {
      const int64_t threads = mc * 3;
      k_encode_chunk<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(a.rays_o, a.rays_d, a.viewdirs, a.z,
          cond ? a.cond_proj : nullptr, a.cond_index, a.C, (int)a.S, m0, mc, X5, in5, Xv, inv);
      note_launches(1);
    }
    // ---- forward recompute -------------------------------------------------------------------------
    auto fwd = [&](const float* X, int64_t ldx, int K, const float* W, int N, const float* bias, int relu, float* Y, int64_t ldy) {
      GemmArgs g{}; g.A = X; g.lda = ldx; g.B = W; g.ldb = N; g.C = Y; g.ldc = ldy; g.I = (int)mc; g.J = N; g.K = K;
      g.bias = bias; g.relu = relu;
      gemm<false, false, EPI_FWD>(g, s);
    };
    fwd(X5, in5, kPE, wt(0), kW, aux + kAuxBiasPts, 1, H[0], kW);
    for (int l = 1; l <= 4; ++l) fwd(H[l - 1], kW, kW, wt(l), kW, aux + kAuxBiasPts + 256 * l, 1, H[l], ldH(l));
    fwd(X5, in5, in5, wt(5), kW, aux + kAuxBiasPts + 256 * 5, 1, H[5], kW);
    fwd(H[5], kW, kW, wt(6), kW, aux + kAuxBiasPts + 256 * 6, 1, H[6], kW);
    fwd(H[6], kW, kW, wt(7), kW, aux + kAuxBiasPts + 256 * 7, 1, H[7], kW);
    fwd(H[7], kW, kW, wt(8), kW, aux + kAuxBiasFeat, 0, Xv, inv);                      // feature -> Xv[:, :256]
    fwd(Xv, inv, inv, wt(9), kWV, aux + kAuxBiasViews, 1, HV, kWV);                    // views

    // ---- backward -------------------------------------------------------------------------------------
    auto wgrad = [&](const float* dZ, int64_t ldz, int N, const float* X, int64_t ldx, int K, float* dW, int64_t lddw) {
      GemmArgs g{}; g.A = dZ; g.lda = ldz; g.B = X; g.ldb = ldx; g.C = dW; g.ldc = lddw; g.I = N; g.J = K; g.K = mc;
      g.k_split = 1024;
      gemm<true, false, EPI_WGRAD>(g, s);
    };
    auto dgrad = [&](const float* dZ, int64_t ldz, int N, const float* W, int64_t ldw, int K, float* dX, int64_t lddx,
                     const float* mask, int64_t ldmask, const float* r1c, int64_t r1ld, const float* r1r) {
      // dX[m][k] = sum_n dZ[m][n] * W[n][k];  K-major storage: W[n][k] = Wt[k*ldw + n]  -> B(kk=n, j=k) = Wt[j*ldw + kk]
      GemmArgs g{}; g.A = dZ; g.lda = ldz; g.B = W; g.ldb = ldw; g.C = dX; g.ldc = lddx; g.I = (int)mc; g.J = K; g.K = N;
      g.mask = mask; g.ldmask = ldmask; g.rank1_col = r1c; g.rank1_ld = r1ld; g.rank1_row = r1r;
      gemm<false, true, EPI_DGRAD>(g, s);
    };
    // rgb head: raw[:, :3] = HV . Wrgb^T + b
    wgrad(gr, 4, 3, HV, kWV, kWV, gw(11), kWV);
    colsum(gr, 4, mc, 3, gb(11), s);
    {   // dZv = (g_rgb . Wrgb) (.) [HV > 0]   (Wrgb is [3][128] row-major in aux: B(kk=c, j) = Wrgb[c*128 + j])
      GemmArgs g{}; g.A = gr; g.lda = 4; g.B = aux + kAuxWRgb; g.ldb = kWV; g.C = dA; g.ldc = kWV; g.I = (int)mc; g.J = kWV; g.K = 3;
      g.mask = HV; g.ldmask = kWV;
      gemm<false, false, EPI_DGRAD>(g, s);
    }
    // views layer
    wgrad(dA, kWV, kWV, Xv, inv, inv, gw(10), inv);
    colsum(dA, kWV, mc, kWV, gb(10), s);
    dgrad(dA, kWV, kWV, wt(9), kWV, kW, dB, kW, nullptr, 0, nullptr, 0, nullptr);      // dFEAT (no activation)
    // alpha head (sigma = H7 . w_alpha + b) and feature layer both feed dH7
    wgrad(gr + 3, 4, 1, H[7], kW, kW, gw(8), kW);
    colsum(gr + 3, 4, mc, 1, gb(8), s);
    wgrad(dB, kW, kW, H[7], kW, kW, gw(9), kW);
    colsum(dB, kW, mc, kW, gb(9), s);
    dgrad(dB, kW, kW, wt(8), kW, kW, dA, kW, H[7], kW, gr + 3, 4, aux + kAuxWAlpha);  // dZ7
    // trunk layers 7 .. 1
    float* dz = dA; float* dnext = dB;
    for (int l = 7; l >= 1; --l) {
      if (l == 5) {
        wgrad(dz, kW, kW, X5, in5, in5, gw(5), in5);
        colsum(dz, kW, mc, kW, gb(5), s);
        dgrad(dz, kW, kW, wt(5) + (size_t)hoff * kW, kW, kW, dnext, kW, H[4], in5, nullptr, 0, nullptr);
      } else {
        wgrad(dz, kW, kW, H[l - 1], ldH(l - 1), kW, gw(l), kW);
        colsum(dz, kW, mc, kW, gb(l), s);
        dgrad(dz, kW, kW, wt(l), kW, kW, dnext, kW, H[l - 1], ldH(l - 1), nullptr, 0, nullptr);
      }
      float* t = dz; dz = dnext; dnext = t;
    }
    // layer 0
    wgrad(dz, kW, kW, X5, in5, kPE, gw(0), kPE);
    colsum(dz, kW, mc, kW, gb(0), s);
  }
  return check_launch("mlp_bwd", 0);
}

}  // namespace fnerf
