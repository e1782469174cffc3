// Generic batched fp32 SIMT GEMM used by the encoder ops and by the exact-fp32 reference variants of the
// corpus contractions.  C[b][m][n] = epilogue( sum_k A[b][m][k] * B[b][n][k] )   (B k-contiguous, "NT")
//                                 or epilogue( sum_k A[b][m][k] * B[b][k][n] )   (B n-contiguous, "NN")
// The tcgen05 kernels (gemm_tc.cu) replace this on the hot contractions; this one stays as the
// exact-fp32 path for shapes the tensor-core tiles do not cover.
#pragma once
#include "common.cuh"

enum GemmEpilogue {
  EPI_STORE = 0,   // (acc / div) + att_mask_add + bias + residual, optional relu -> C
  EPI_VRMAX = 1,   // per (row, video) masked max over the video's clips -> atomic max into C[m][n / L]
};

struct GemmParams {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  long long lda, ldb, ldc;
  int b_is_kn;
  int a_is_km;  // A stored [k][m] (lda = pitch between k rows): C = A^T-stored . B, the dW / dV / dK products of backward
  int batch1;  // blockIdx.z = b0 * batch1 + b1
  long long sA0, sA1, sB0, sB1, sC0, sC1;
  int epilogue;
  float div;              // 0 -> no division; else acc / div (reference divides scores by sqrt(dh))
  const float* bias;      // [N] or null
  const float* residual;  // laid out like C or null
  int relu;
  const float* att_mask;  // null, or mask[b0][m * mask_sm][n]; adds (1 - mask) * -10000
  long long mask_s0, mask_sm;
  // EPI_VRMAX
  const float* clip_mask;  // [N] float {0,1}: column n = video n / L, clip n % L
  int L;
  int n_videos;
};

template <int BM, int BN, int RM, int RN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  constexpr int BK = 16;
  constexpr int TM = 4 * RM, TN = 4 * RN;
  constexpr int TX = BN / TN;  // threads along n
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads per CTA");
  constexpr int SA = BM + 4, SB = BN + 4;
  __shared__ __align__(16) float As[2][BK][SA];
  __shared__ __align__(16) float Bs[2][BK][SB];

  const int t = threadIdx.x;
  const int tx = t % TX, ty = t / TX;
  const int b0 = blockIdx.z / p.batch1, b1 = blockIdx.z % p.batch1;
  const float* __restrict__ A = p.A + b0 * p.sA0 + b1 * p.sA1;
  const float* __restrict__ B = p.B + b0 * p.sB0 + b1 * p.sB1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  constexpr int A_PER = BM * BK / 256, B_PER = BN * BK / 256;
  float ra[A_PER], rb[B_PER];

  auto load_tiles = [&](int k0) {
    if (!p.a_is_km) {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int idx = t + i * 256;
        int k = idx % BK, m = idx / BK;
        int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < p.M && gk < p.K) ? __ldg(A + (long long)gm * p.lda + gk) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int idx = t + i * 256;
        int m = idx % BM, k = idx / BM;
        int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < p.M && gk < p.K) ? __ldg(A + (long long)gk * p.lda + gm) : 0.f;
      }
    }
    if (!p.b_is_kn) {
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        int idx = t + i * 256;
        int k = idx % BK, n = idx / BK;
        int gn = n0 + n, gk = k0 + k;
        rb[i] = (gn < p.N && gk < p.K) ? __ldg(B + (long long)gn * p.ldb + gk) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        int idx = t + i * 256;
        int n = idx % BN, k = idx / BN;
        int gn = n0 + n, gk = k0 + k;
        rb[i] = (gn < p.N && gk < p.K) ? __ldg(B + (long long)gk * p.ldb + gn) : 0.f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (!p.a_is_km) {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int idx = t + i * 256;
        As[buf][idx % BK][idx / BK] = ra[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int idx = t + i * 256;
        As[buf][idx / BM][idx % BM] = ra[i];
      }
    }
    if (!p.b_is_kn) {
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        int idx = t + i * 256;
        Bs[buf][idx % BK][idx / BK] = rb[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        int idx = t + i * 256;
        Bs[buf][idx / BN][idx % BN] = rb[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int n_k = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < n_k; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_k) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * (BM / RM) + ty * 4]);
        a[r * 4 + 0] = v.x, a[r * 4 + 1] = v.y, a[r * 4 + 2] = v.z, a[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][r * (BN / RN) + tx * 4]);
        b[r * 4 + 0] = v.x, b[r * 4 + 1] = v.y, b[r * 4 + 2] = v.z, b[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_k) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---------------- epilogue ----------------
  if (p.epilogue == EPI_STORE) {
    float* __restrict__ C = p.C + b0 * p.sC0 + b1 * p.sC1;
    const float* __restrict__ R = p.residual ? p.residual + b0 * p.sC0 + b1 * p.sC1 : nullptr;
    const float* __restrict__ MK = p.att_mask ? p.att_mask + b0 * p.mask_s0 : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int gm = m0 + (i / 4) * (BM / RM) + ty * 4 + (i % 4);
      if (gm >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int gn = n0 + (j / 4) * (BN / RN) + tx * 4 + (j % 4);
        if (gn >= p.N) continue;
        float v = acc[i][j];
        if (p.div != 0.f) v = __fdiv_rn(v, p.div);
        if (MK) v = __fadd_rn(v, __fmul_rn(__fsub_rn(1.f, __ldg(MK + gm * p.mask_sm + gn)), ATT_MASK_FILL));
        if (p.bias) v += __ldg(p.bias + gn);
        if (R) v += __ldg(R + (long long)gm * p.ldc + gn);
        if (p.relu) v = fmaxf(v, 0.f);
        C[(long long)gm * p.ldc + gn] = v;
      }
    }
  } else {  // EPI_VRMAX: columns are corpus clips; reduce to per-video masked max
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int gm = m0 + (i / 4) * (BM / RM) + ty * 4 + (i % 4);
      if (gm >= p.M) continue;
      int cur_v = -1;
      float cur = 0.f;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int gn = n0 + (j / 4) * (BN / RN) + tx * 4 + (j % 4);
        if (gn >= p.N) continue;
        if (__ldg(p.clip_mask + gn) == 0.f) continue;  // masked clip contributes -1e10 (the initial value)
        const int v = gn / p.L;
        if (v != cur_v) {
          if (cur_v >= 0) atomic_max_float(p.C + (long long)gm * p.n_videos + cur_v, cur);
          cur_v = v;
          cur = acc[i][j];
        } else {
          cur = fmaxf(cur, acc[i][j]);
        }
      }
      if (cur_v >= 0) atomic_max_float(p.C + (long long)gm * p.n_videos + cur_v, cur);
    }
  }
}

// host-side launcher (defined in gemm_simt.cu)
int xmlb_gemm_launch(const GemmParams& p, int batch0, cudaStream_t stream);
