// Backward kernels of the training step (BASELINE config #4; reference model_xml.py:212-251 differentiated by
// torch.autograd, train.py:77-85): everything that is not a Linear layer -- LayerNorm, the attention core, the
// modular query pooling, L2 normalisation, the in-batch masked-max video scores and the similarity + ConvSE span
// logits of each query on its own video -- plus the small reductions (bias / LayerNorm / position-table / ConvSE
// gradients).  All fp32, all deterministic (fixed summation order: no floating-point atomics).
// The Linear layers' dX / dW products run on the GEMM kernels (linear_tc.cu / gemm_simt.cu), see autograd.py.
#include <math.h>
#include "gemm_simt.cuh"
#include "xmlb200.h"

extern "C" int xmlb_softmax_rows(const float* x, float* out, long long rows, int dim, void* stream);
extern "C" int xmlb_dropout(const float* x, float* out, long long n, float p, unsigned long long seed,
                            unsigned long long index0, void* stream);

namespace {

// ---------------------------------------------------------------------------------------------- strided row sums
// out[chunk][p][c] = sum over the chunk's groups g of in[(g * group_rows + p) * dim + c], groups in order.
constexpr int SUM_CHUNK = 64;
__global__ void __launch_bounds__(256) sum_rows_kernel(const float* __restrict__ in, long long n_groups, int group_rows,
                                                       int dim, float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int p = blockIdx.y;
  const long long g0 = (long long)blockIdx.z * SUM_CHUNK, g1 = min(n_groups, g0 + SUM_CHUNK);
  if (c >= dim) return;
  float s = 0.f;
  for (long long g = g0; g < g1; ++g) s += __ldg(in + (g * group_rows + p) * dim + c);
  out[((long long)blockIdx.z * group_rows + p) * dim + c] = s;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// y = LN(x + add[r % add_rows]) * gamma + beta.  dx = gradient w.r.t. (x + add); dy_xhat = dy * xhat (its column sums
// are dgamma; the column sums of dy are dbeta).  One warp per row.
__global__ void __launch_bounds__(256) layernorm_backward_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                                 long long add_rows, const float* __restrict__ gamma,
                                                                 const float* __restrict__ dy, long long rows, int dim,
                                                                 float eps, float* __restrict__ dx,
                                                                 float* __restrict__ dy_xhat) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  const float* ar = add ? add + (row % add_rows) * dim : nullptr;
  const float* g = dy + row * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += ar ? xr[i] + ar[i] : xr[i];
  const float mean = warp_sum(s) / (float)dim;
  float v = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float d = (ar ? xr[i] + ar[i] : xr[i]) - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = 1.f / sqrtf(warp_sum(v) / (float)dim + eps);
  float s1 = 0.f, s2 = 0.f;  // sum(g * gamma), sum(g * gamma * xhat)
  for (int i = lane; i < dim; i += 32) {
    const float xh = ((ar ? xr[i] + ar[i] : xr[i]) - mean) * rstd;
    const float gg = g[i] * __ldg(gamma + i);
    s1 += gg, s2 = fmaf(gg, xh, s2);
  }
  s1 = warp_sum(s1) / (float)dim, s2 = warp_sum(s2) / (float)dim;
  for (int i = lane; i < dim; i += 32) {
    const float xh = ((ar ? xr[i] + ar[i] : xr[i]) - mean) * rstd;
    const float gg = g[i] * __ldg(gamma + i);
    dx[row * dim + i] = rstd * (gg - s1 - xh * s2);
    dy_xhat[row * dim + i] = g[i] * xh;
  }
}

// y = x / max(||x||, eps):  dx = (g - y (y . g)) / max(||x||, eps)   (rows with ||x|| < eps: dx = g / eps)
__global__ void __launch_bounds__(256) l2norm_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              long long rows, int dim, float eps, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  const float* g = dy + row * dim;
  float s = 0.f, d = 0.f;
  for (int i = lane; i < dim; i += 32) s = fmaf(xr[i], xr[i], s), d = fmaf(xr[i], g[i], d);
  const float nrm = sqrtf(warp_sum(s));
  d = warp_sum(d);
  if (nrm >= eps) {
    const float inv = 1.f / nrm, c = d * inv * inv * inv;  // (x . g) / ||x||^3
    for (int i = lane; i < dim; i += 32) dx[row * dim + i] = g[i] * inv - xr[i] * c;
  } else {
    for (int i = lane; i < dim; i += 32) dx[row * dim + i] = g[i] / eps;
  }
}

__global__ void relu_backward_kernel(const float* __restrict__ g, const float* __restrict__ out, long long n,
                                     float* __restrict__ dx) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dx[i] = out[i] > 0.f ? g[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------- modular pooling
// forward (model_xml.py:410-423): logit[t][m] = e_t . w_m; a = softmax_t(mask_logits(logit, mask)); out_m = sum_t a e_t.
// backward per query (one CTA): da[t][m] = dout_m . e_t; dlogit = mask * a * (da - sum_t a da);
//   de_t = sum_m (a[t][m] dout_m + dlogit[t][m] w_m).  dlogit (n, len, 2) is returned for dW = dlogit^T . E (a GEMM).
__global__ void __launch_bounds__(128) modular_pool_backward_kernel(const float* __restrict__ enc,
                                                                    const float* __restrict__ mask,
                                                                    const float* __restrict__ w_mod,
                                                                    const float* __restrict__ dout0,
                                                                    const float* __restrict__ dout1, int len, int hidden,
                                                                    int n_mod, float* __restrict__ d_enc,
                                                                    float* __restrict__ dlogit) {
  extern __shared__ float sm[];  // att[len][2], da[len][2], dl[len][2]
  float* att = sm;
  float* da = sm + 2 * len;
  float* dl = da + 2 * len;
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* e = enc + (long long)n * len * hidden;
  const float* g0 = dout0 + (long long)n * hidden;
  const float* g1 = n_mod == 2 ? dout1 + (long long)n * hidden : nullptr;
  for (int t = warp; t < len; t += 4) {
    float s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;
    for (int d = lane; d < hidden; d += 32) {
      const float x = e[(long long)t * hidden + d];
      s0 = fmaf(x, __ldg(w_mod + d), s0), a0 = fmaf(x, __ldg(g0 + d), a0);
      if (n_mod == 2) s1 = fmaf(x, __ldg(w_mod + hidden + d), s1), a1 = fmaf(x, __ldg(g1 + d), a1);
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1), a0 = warp_sum(a0), a1 = warp_sum(a1);
    if (lane == 0) {
      const float m = mask[(long long)n * len + t];
      att[t * 2 + 0] = mask_logit(s0, m), att[t * 2 + 1] = mask_logit(s1, m);
      da[t * 2 + 0] = a0, da[t * 2 + 1] = a1;
    }
  }
  __syncthreads();
  if (warp < n_mod) {  // softmax over tokens and its backward, one warp per modular vector
    float mx = -INFINITY;
    for (int t = lane; t < len; t += 32) mx = fmaxf(mx, att[t * 2 + warp]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < len; t += 32) s += expf(att[t * 2 + warp] - mx);
    s = warp_sum(s);
    float dot = 0.f;
    for (int t = lane; t < len; t += 32) {
      const float a = __fdiv_rn(expf(att[t * 2 + warp] - mx), s);
      att[t * 2 + warp] = a;
      dot = fmaf(a, da[t * 2 + warp], dot);
    }
    dot = warp_sum(dot);
    for (int t = lane; t < len; t += 32) {
      const float m = mask[(long long)n * len + t];
      const float v = m * att[t * 2 + warp] * (da[t * 2 + warp] - dot);  // d mask_logits / d logit = mask
      dl[t * 2 + warp] = v;
      dlogit[((long long)n * len + t) * 2 + warp] = v;
    }
  } else if (warp == 1 && n_mod == 1) {
    for (int t = lane; t < len; t += 32) dl[t * 2 + 1] = 0.f, dlogit[((long long)n * len + t) * 2 + 1] = 0.f;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < hidden; d += 128) {
    const float w0 = __ldg(w_mod + d), w1 = n_mod == 2 ? __ldg(w_mod + hidden + d) : 0.f;
    const float q0 = __ldg(g0 + d), q1 = n_mod == 2 ? __ldg(g1 + d) : 0.f;
    for (int t = 0; t < len; ++t)
      d_enc[((long long)n * len + t) * hidden + d] =
          att[t * 2] * q0 + att[t * 2 + 1] * q1 + dl[t * 2] * w0 + dl[t * 2 + 1] * w1;
  }
}

// ---------------------------------------------------------------------------------------------- in-batch video scores
// forward (model_xml.py:446-452): s[m][n] = max over valid clips l of q_m . c[n][l].  argmax[m][n] = that clip (-1: no
// valid clip).  One warp per (m, n) pair.
__global__ void __launch_bounds__(256) vr_argmax_kernel(const float* __restrict__ q, const float* __restrict__ c,
                                                        const float* __restrict__ mask, int nq, int nv, int len,
                                                        int hidden, int* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const long long pair = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (pair >= (long long)nq * nv) return;
  const int m = (int)(pair / nv), n = (int)(pair % nv);
  const float* qm = q + (long long)m * hidden;
  float best = -INFINITY;
  int arg = -1;
  for (int l = 0; l < len; ++l) {
    if (mask[(long long)n * len + l] == 0.f) continue;
    const float* cl = c + ((long long)n * len + l) * hidden;
    float s = 0.f;
    for (int d = lane; d < hidden; d += 32) s = fmaf(qm[d], cl[d], s);
    s = warp_sum(s);
    if (s > best) best = s, arg = l;
  }
  if (lane == 0) argmax[pair] = arg;
}
// dq[m] (+)= scale * sum_n g[m][n] c[n][argmax[m][n]]     (one CTA per query, n in order)
__global__ void __launch_bounds__(256) vr_dq_kernel(const float* __restrict__ g, const float* __restrict__ c,
                                                    const int* __restrict__ argmax, int nv, int len, int hidden,
                                                    float scale, float* __restrict__ dq) {
  const int m = blockIdx.x;
  for (int d = threadIdx.x; d < hidden; d += 256) {
    float s = 0.f;
    for (int n = 0; n < nv; ++n) {
      const int l = argmax[(long long)m * nv + n];
      if (l >= 0) s = fmaf(g[(long long)m * nv + n], c[((long long)n * len + l) * hidden + d], s);
    }
    dq[(long long)m * hidden + d] = s * scale;
  }
}
// dc[n][l] = scale * sum over m with argmax[m][n] == l of g[m][n] q[m]     (one CTA per video, m in order)
__global__ void __launch_bounds__(256) vr_dc_kernel(const float* __restrict__ g, const float* __restrict__ q,
                                                    const int* __restrict__ argmax, int nq, int nv, int len, int hidden,
                                                    float scale, float* __restrict__ dc) {
  const int n = blockIdx.x;
  float* out = dc + (long long)n * len * hidden;
  for (long long i = threadIdx.x; i < (long long)len * hidden; i += 256) out[i] = 0.f;
  __syncthreads();
  for (int m = 0; m < nq; ++m) {
    const int l = argmax[(long long)m * nv + n];
    if (l < 0) continue;  // (uniform over the CTA)
    const float w = g[(long long)m * nv + n] * scale;
    for (int d = threadIdx.x; d < hidden; d += 256) out[(long long)l * hidden + d] = fmaf(w, q[(long long)m * hidden + d], out[(long long)l * hidden + d]);
  }
}

// ---------------------------------------------------------------------------------------------- span logits, diagonal
// forward (model_xml.py:455-502 / 512-551, cross=False): item b scores ITS OWN video.  Per stream x: sim_x[l] =
// q_x[b] . f2_x[b][l]; merged: sim = (sim_a + sim_b) / 2, st = mask_logits(conv(w_st_a, sim), mask_a); separate:
// st = mean_x mask_logits(conv(w_st_x, sim_x), mask_x).  conv = cross-correlation, zero padding ksize / 2.
// backward, one CTA per item: dq_x, df2_x, and per-item partial ConvSE weight gradients dw[b][stream][st|ed][ksize]
// (summed over the batch by xmlb_sum_rows).
constexpr int MAX_KSIZE = 31;
__global__ void __launch_bounds__(128) span_diag_backward_kernel(
    const float* __restrict__ q_a, const float* __restrict__ q_b, const float* __restrict__ f2_a,
    const float* __restrict__ f2_b, const float* __restrict__ mask_a, const float* __restrict__ mask_b,
    const float* __restrict__ w_st_a, const float* __restrict__ w_ed_a, const float* __restrict__ w_st_b,
    const float* __restrict__ w_ed_b, int ksize, int merged, int len, int hidden, const float* __restrict__ dst,
    const float* __restrict__ ded, float* __restrict__ dq_a, float* __restrict__ dq_b, float* __restrict__ df2_a,
    float* __restrict__ df2_b, float* __restrict__ dw) {
  extern __shared__ float sm[];  // sim[2][len], gs[2][len] (masked dst per stream), ge[2][len], dsim[2][len]
  float* sim = sm;
  float* gs = sm + 2 * len;
  float* ge = gs + 2 * len;
  float* dsim = ge + 2 * len;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pad = ksize / 2;
  const int n_str = q_b ? 2 : 1;
  const float* qs[2] = {q_a + (long long)b * hidden, q_b ? q_b + (long long)b * hidden : nullptr};
  const float* fs[2] = {f2_a + (long long)b * len * hidden, f2_b ? f2_b + (long long)b * len * hidden : nullptr};
  for (int x = 0; x < n_str; ++x)
    for (int l = warp; l < len; l += 4) {
      float s = 0.f;
      for (int d = lane; d < hidden; d += 32) s = fmaf(qs[x][d], fs[x][(long long)l * hidden + d], s);
      s = warp_sum(s);
      if (lane == 0) sim[x * len + l] = s;
    }
  __syncthreads();
  if (merged) {  // one curve: (sim_a + sim_b) / 2, one pair of predictors, mask_a
    for (int l = threadIdx.x; l < len; l += 128) {
      sim[l] = (sim[l] + sim[len + l]) * 0.5f;
      const float m = mask_a[(long long)b * len + l];
      gs[l] = dst[(long long)b * len + l] * m, ge[l] = ded[(long long)b * len + l] * m;
    }
  } else {
    const float inv = 1.f / (float)n_str;
    for (int x = 0; x < n_str; ++x)
      for (int l = threadIdx.x; l < len; l += 128) {
        const float m = (x == 0 ? mask_a : mask_b)[(long long)b * len + l] * inv;
        gs[x * len + l] = dst[(long long)b * len + l] * m, ge[x * len + l] = ded[(long long)b * len + l] * m;
      }
  }
  __syncthreads();
  const int n_curves = merged ? 1 : n_str;
  for (int x = 0; x < n_curves; ++x) {
    const float* ws = x == 0 ? w_st_a : w_st_b;
    const float* we = x == 0 ? w_ed_a : w_ed_b;
    // st[l'] = sum_t w[t] sim[l' + t - pad]  =>  dsim[l] = sum_t w[t] g[l - t + pad]
    for (int l = threadIdx.x; l < len; l += 128) {
      float s = 0.f;
      for (int t = 0; t < ksize; ++t) {
        const int lp = l - t + pad;
        if (lp >= 0 && lp < len) s = fmaf(__ldg(ws + t), gs[x * len + lp], fmaf(__ldg(we + t), ge[x * len + lp], s));
      }
      dsim[x * len + l] = s;
    }
    // dw[t] = sum_l g[l] sim[l + t - pad]
    if (threadIdx.x < 2 * ksize) {
      const int t = threadIdx.x % ksize, which = threadIdx.x / ksize;  // 0 = start, 1 = end predictor
      const float* g = which == 0 ? gs + x * len : ge + x * len;
      float s = 0.f;
      for (int l = 0; l < len; ++l) {
        const int src = l + t - pad;
        if (src >= 0 && src < len) s = fmaf(g[l], sim[x * len + src], s);
      }
      dw[(((long long)b * 2 + x) * 2 + which) * MAX_KSIZE + t] = s;
    }
  }
  __syncthreads();
  if (merged)
    for (int l = threadIdx.x; l < len; l += 128) dsim[l] *= 0.5f, dsim[len + l] = dsim[l];
  __syncthreads();
  float* dqs[2] = {dq_a + (long long)b * hidden, dq_b ? dq_b + (long long)b * hidden : nullptr};
  float* dfs[2] = {df2_a + (long long)b * len * hidden, df2_b ? df2_b + (long long)b * len * hidden : nullptr};
  for (int x = 0; x < n_str; ++x)
    for (int d = threadIdx.x; d < hidden; d += 128) {
      const float qd = qs[x][d];
      float s = 0.f;
      for (int l = 0; l < len; ++l) {
        const float ds = dsim[x * len + l];
        s = fmaf(ds, fs[x][(long long)l * hidden + d], s);
        dfs[x][(long long)l * hidden + d] = ds * qd;
      }
      dqs[x][d] = s;
    }
}

// ---------------------------------------------------------------------------------------------- attention core
// dS = P * (dP - rowsum(dP * P)) with dP = keep ? dPd / (1 - p) : 0 (the forward's dropout mask re-derived from the
// counter-based hash).  In place on dpd; one warp per (batch, head, query) row.
__global__ void __launch_bounds__(256) softmax_dropout_backward_kernel(const float* __restrict__ p_, float* __restrict__ dpd,
                                                                       long long rows, int dim, uint32_t threshold,
                                                                       float scale, unsigned long long seed,
                                                                       unsigned long long index0) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pr = p_ + row * dim;
  float* gr = dpd + row * dim;
  float dot = 0.f;
  for (int i = lane; i < dim; i += 32) {
    float g = gr[i];
    if (threshold) g = dropout_keep(seed, index0 + row * dim + i, threshold) ? g * scale : 0.f;
    gr[i] = g;
    dot = fmaf(g, pr[i], dot);
  }
  dot = warp_sum(dot);
  for (int i = lane; i < dim; i += 32) gr[i] = pr[i] * (gr[i] - dot);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int xmlb_sum_rows(const float* in, long long n_groups, int group_rows, int dim, float* out, float* ws,
                             void* stream) {
  XMLB_REQUIRE(in && out && n_groups >= 1 && group_rows >= 1 && dim >= 1, "xmlb_sum_rows: bad argument");
  XMLB_REQUIRE(group_rows <= 65535, "xmlb_sum_rows: group_rows too large");
  XMLB_REQUIRE(n_groups <= SUM_CHUNK || ws, "xmlb_sum_rows: workspace required for more than %d groups", SUM_CHUNK);
  const float* src = in;
  long long n = n_groups;
  float* bufs[2] = {ws, ws ? ws + ((n_groups + SUM_CHUNK - 1) / SUM_CHUNK) * group_rows * (long long)dim : nullptr};
  int which = 0;
  for (;;) {
    const long long chunks = (n + SUM_CHUNK - 1) / SUM_CHUNK;
    XMLB_REQUIRE(chunks <= 65535, "xmlb_sum_rows: too many groups");
    float* dst = chunks == 1 ? out : bufs[which];
    sum_rows_kernel<<<dim3(ceil_div(dim, 256), group_rows, (unsigned)chunks), 256, 0, (cudaStream_t)stream>>>(
        src, n, group_rows, dim, dst);
    xmlb_count_launch(1);
    XMLB_LAUNCH_CHECK();
    if (chunks == 1) break;
    src = dst, n = chunks, which ^= 1;
  }
  return XMLB_OK;
}

extern "C" int xmlb_layernorm_backward(const float* x, const float* add, long long add_rows, const float* gamma,
                                       const float* dy, long long rows, int dim, float eps, float* dx, float* dy_xhat,
                                       void* stream) {
  XMLB_REQUIRE(x && gamma && dy && dx && dy_xhat && dim > 0 && rows >= 0, "xmlb_layernorm_backward: bad argument");
  XMLB_REQUIRE(!add || add_rows > 0, "xmlb_layernorm_backward: add_rows must be > 0 when add is given");
  if (rows == 0) return XMLB_OK;
  layernorm_backward_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, add, add_rows, gamma, dy, rows, dim,
                                                                                eps, dx, dy_xhat);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_l2norm_backward(const float* x, const float* dy, long long rows, int dim, float eps, float* dx,
                                    void* stream) {
  XMLB_REQUIRE(x && dy && dx && dim > 0 && rows >= 0, "xmlb_l2norm_backward: bad argument");
  if (rows == 0) return XMLB_OK;
  l2norm_backward_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, dy, rows, dim, eps, dx);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_relu_backward(const float* g, const float* out, long long n, float* dx, void* stream) {
  XMLB_REQUIRE(g && out && dx && n >= 0, "xmlb_relu_backward: bad argument");
  if (n == 0) return XMLB_OK;
  const long long blocks = (n + 255) / 256;
  relu_backward_kernel<<<(int)(blocks < 1184 * 4 ? blocks : 1184 * 4), 256, 0, (cudaStream_t)stream>>>(g, out, n, dx);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_modular_pool_backward(const float* encoded, const float* mask, const float* w_mod, const float* dout0,
                                          const float* dout1, int n_queries, int len, int hidden, int n_mod,
                                          float* d_encoded, float* dlogit, void* stream) {
  XMLB_REQUIRE(encoded && mask && w_mod && dout0 && d_encoded && dlogit, "xmlb_modular_pool_backward: null pointer");
  XMLB_REQUIRE(n_mod == 1 || (n_mod == 2 && dout1), "xmlb_modular_pool_backward: n_mod must be 1 or 2 (with dout1)");
  XMLB_REQUIRE(len > 0 && len <= 2048, "xmlb_modular_pool_backward: len out of range");
  if (n_queries == 0) return XMLB_OK;
  modular_pool_backward_kernel<<<n_queries, 128, 6 * len * sizeof(float), (cudaStream_t)stream>>>(
      encoded, mask, w_mod, dout0, dout1, len, hidden, n_mod, d_encoded, dlogit);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_vr_scores_backward(const float* q_n, const float* c_n, const float* mask, const float* g, float scale,
                                       int n_queries, int n_videos, int ctx_len, int hidden, int* argmax_ws, float* dq,
                                       float* dc, void* stream) {
  XMLB_REQUIRE(q_n && c_n && mask && g && argmax_ws && dq && dc, "xmlb_vr_scores_backward: null pointer");
  if (n_queries == 0 || n_videos == 0) return XMLB_OK;
  const long long pairs = (long long)n_queries * n_videos;
  vr_argmax_kernel<<<ceil_div(pairs, 8), 256, 0, (cudaStream_t)stream>>>(q_n, c_n, mask, n_queries, n_videos, ctx_len,
                                                                        hidden, argmax_ws);
  vr_dq_kernel<<<n_queries, 256, 0, (cudaStream_t)stream>>>(g, c_n, argmax_ws, n_videos, ctx_len, hidden, scale, dq);
  vr_dc_kernel<<<n_videos, 256, 0, (cudaStream_t)stream>>>(g, q_n, argmax_ws, n_queries, n_videos, ctx_len, hidden,
                                                          scale, dc);
  xmlb_count_launch(3);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_span_logits_diag_backward(const float* q_a, const float* q_b, const float* feat2_a,
                                              const float* feat2_b, const float* mask_a, const float* mask_b,
                                              const float* w_st_a, const float* w_ed_a, const float* w_st_b,
                                              const float* w_ed_b, int ksize, int merged, int n, int ctx_len, int hidden,
                                              const float* dst, const float* ded, float* dq_a, float* dq_b,
                                              float* dfeat2_a, float* dfeat2_b, float* dw_partial, void* stream) {
  XMLB_REQUIRE(q_a && feat2_a && mask_a && w_st_a && w_ed_a && dst && ded && dq_a && dfeat2_a && dw_partial,
               "xmlb_span_logits_diag_backward: null pointer");
  XMLB_REQUIRE(!q_b || (feat2_b && dq_b && dfeat2_b && (merged || (mask_b && w_st_b && w_ed_b))),
               "xmlb_span_logits_diag_backward: incomplete second stream");
  XMLB_REQUIRE(!merged || q_b, "xmlb_span_logits_diag_backward: the merged predictor needs two streams");
  XMLB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= MAX_KSIZE && 2 * ksize <= 128,
               "xmlb_span_logits_diag_backward: ksize must be odd and <= %d", MAX_KSIZE);
  if (n == 0) return XMLB_OK;
  XMLB_CUDA(cudaMemsetAsync(dw_partial, 0, sizeof(float) * (size_t)n * 4 * MAX_KSIZE, (cudaStream_t)stream));
  span_diag_backward_kernel<<<n, 128, 8 * ctx_len * sizeof(float), (cudaStream_t)stream>>>(
      q_a, q_b, feat2_a, feat2_b, mask_a, mask_b, w_st_a, w_ed_a, w_st_b, w_ed_b, ksize, merged, ctx_len, hidden, dst, ded,
      dq_a, dq_b, dfeat2_a, dfeat2_b, dw_partial);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

// Backward of xmlb_attention / xmlb_attention_train (model_components.py:277-303 after the projections), exact fp32:
// the probabilities are recomputed (S = Q K^T / sqrt(dh) + mask, softmax; the forward kept nothing), then
//   dV = Pd^T dO,  dPd = dO V^T,  dS = P * (dP - rowsum(dP * P)),  dQ = dS K / sqrt(dh),  dK = dS^T Q / sqrt(dh)
// with Pd = dropout(P) from the same counter-based mask.  ws_p, ws_g: batch * n_heads * len_q * len_k floats each.
extern "C" int xmlb_attention_backward(const float* q, const float* k, const float* v, const float* mask,
                                       long long mask_batch_stride, long long mask_q_stride, const float* dout, float* dq,
                                       float* dk, float* dv, float* ws_p, float* ws_g, int batch, int len_q, int len_k,
                                       int hidden, int n_heads, float dropout_p, unsigned long long seed,
                                       unsigned long long index0, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XMLB_REQUIRE(q && k && v && mask && dout && dq && dk && dv && ws_p && ws_g, "xmlb_attention_backward: null pointer");
  XMLB_REQUIRE(n_heads > 0 && hidden % n_heads == 0, "xmlb_attention_backward: hidden %% n_heads != 0");
  XMLB_REQUIRE((long long)batch * n_heads <= 65535, "xmlb_attention_backward: batch*n_heads > 65535, split the batch");
  XMLB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "xmlb_attention_backward: dropout_p must be in [0, 1)");
  if (batch == 0 || len_q == 0) return XMLB_OK;
  const int dh = hidden / n_heads;
  const long long n_sc = (long long)batch * n_heads * len_q * len_k;
  const float sd = sqrtf((float)dh);
  int rc;
  GemmParams s = {};  // S = Q K^T / sqrt(dh) + (1 - mask) * -10000   -> ws_p
  s.A = q, s.B = k, s.C = ws_p, s.M = len_q, s.N = len_k, s.K = dh;
  s.lda = hidden, s.ldb = hidden, s.ldc = len_k, s.batch1 = n_heads;
  s.sA0 = (long long)len_q * hidden, s.sA1 = dh, s.sB0 = (long long)len_k * hidden, s.sB1 = dh;
  s.sC0 = (long long)n_heads * len_q * len_k, s.sC1 = (long long)len_q * len_k;
  s.epilogue = EPI_STORE, s.div = sd, s.att_mask = mask, s.mask_s0 = mask_batch_stride, s.mask_sm = mask_q_stride;
  if ((rc = xmlb_gemm_launch(s, batch, stream))) return rc;
  if ((rc = xmlb_softmax_rows(ws_p, ws_p, (long long)batch * n_heads * len_q, len_k, stream_))) return rc;
  const float* pd = ws_p;
  if (dropout_p > 0.f) {  // Pd = dropout(P) -> ws_g (consumed by the dV product before dPd overwrites it)
    if ((rc = xmlb_dropout(ws_p, ws_g, n_sc, dropout_p, seed, index0, stream_))) return rc;
    pd = ws_g;
  }
  GemmParams a = {};  // dV[b][h] (len_k x dh) = Pd^T (stored [len_q][len_k]) . dO_h (len_q x dh)
  a.A = pd, a.B = dout, a.C = dv, a.M = len_k, a.N = dh, a.K = len_q;
  a.lda = len_k, a.a_is_km = 1, a.ldb = hidden, a.b_is_kn = 1, a.ldc = hidden, a.batch1 = n_heads;
  a.sA0 = s.sC0, a.sA1 = s.sC1, a.sB0 = (long long)len_q * hidden, a.sB1 = dh, a.sC0 = (long long)len_k * hidden, a.sC1 = dh;
  a.epilogue = EPI_STORE;
  if ((rc = xmlb_gemm_launch(a, batch, stream))) return rc;
  GemmParams g = {};  // dPd = dO_h (len_q x dh) . V_h^T   -> ws_g
  g.A = dout, g.B = v, g.C = ws_g, g.M = len_q, g.N = len_k, g.K = dh;
  g.lda = hidden, g.ldb = hidden, g.ldc = len_k, g.batch1 = n_heads;
  g.sA0 = (long long)len_q * hidden, g.sA1 = dh, g.sB0 = (long long)len_k * hidden, g.sB1 = dh, g.sC0 = s.sC0, g.sC1 = s.sC1;
  g.epilogue = EPI_STORE;
  if ((rc = xmlb_gemm_launch(g, batch, stream))) return rc;
  softmax_dropout_backward_kernel<<<ceil_div((long long)batch * n_heads * len_q, 8), 256, 0, stream>>>(
      ws_p, ws_g, (long long)batch * n_heads * len_q, len_k, dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u,
      1.f / (1.f - dropout_p), seed, index0);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  GemmParams dqp = {};  // dQ_h = dS (len_q x len_k) . K_h (len_k x dh) / sqrt(dh)
  dqp.A = ws_g, dqp.B = k, dqp.C = dq, dqp.M = len_q, dqp.N = dh, dqp.K = len_k;
  dqp.lda = len_k, dqp.ldb = hidden, dqp.b_is_kn = 1, dqp.ldc = hidden, dqp.batch1 = n_heads;
  dqp.sA0 = s.sC0, dqp.sA1 = s.sC1, dqp.sB0 = (long long)len_k * hidden, dqp.sB1 = dh;
  dqp.sC0 = (long long)len_q * hidden, dqp.sC1 = dh, dqp.epilogue = EPI_STORE, dqp.div = sd;
  if ((rc = xmlb_gemm_launch(dqp, batch, stream))) return rc;
  GemmParams dkp = {};  // dK_h = dS^T (stored [len_q][len_k]) . Q_h (len_q x dh) / sqrt(dh)
  dkp.A = ws_g, dkp.B = q, dkp.C = dk, dkp.M = len_k, dkp.N = dh, dkp.K = len_q;
  dkp.lda = len_k, dkp.a_is_km = 1, dkp.ldb = hidden, dkp.b_is_kn = 1, dkp.ldc = hidden, dkp.batch1 = n_heads;
  dkp.sA0 = s.sC0, dkp.sA1 = s.sC1, dkp.sB0 = (long long)len_q * hidden, dkp.sB1 = dh;
  dkp.sC0 = (long long)len_k * hidden, dkp.sC1 = dh, dkp.epilogue = EPI_STORE, dkp.div = sd;
  return xmlb_gemm_launch(dkp, batch, stream);
}

// dW (n_mod x hidden) = dlogit^T (stored [n * len][2]) . encoded ([n * len][hidden]) -- the modular mapping's gradient
extern "C" int xmlb_modular_mapping_grad(const float* dlogit, const float* encoded, long long rows, int hidden, int n_mod,
                                         float* dw, void* stream) {
  XMLB_REQUIRE(dlogit && encoded && dw && rows >= 1 && rows < (1ll << 31), "xmlb_modular_mapping_grad: bad argument");
  GemmParams p = {};
  p.A = dlogit, p.B = encoded, p.C = dw, p.M = n_mod, p.N = hidden, p.K = (int)rows;
  p.lda = 2, p.a_is_km = 1, p.ldb = hidden, p.b_is_kn = 1, p.ldc = hidden, p.batch1 = 1, p.epilogue = EPI_STORE;
  return xmlb_gemm_launch(p, 1, (cudaStream_t)stream);
}
