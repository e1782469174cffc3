// Multi-head attention core and the modular query pooling of the XML encoders (exact-fp32 path).
//   xmlb_attention    : reference BertSelfAttention.forward after the Q/K/V projections
//                       (model_components.py:277-303):  softmax(QK^T / sqrt(dh) + (1-mask)*-10000) V
//   xmlb_modular_pool : reference XML.get_modularized_queries (model_xml.py:410-423)
#include <math.h>
#include "gemm_simt.cuh"
#include "xmlb200.h"

extern "C" int xmlb_softmax_rows(const float* x, float* out, long long rows, int dim, void* stream);

extern "C" int xmlb_dropout(const float* x, float* out, long long n, float p, unsigned long long seed,
                            unsigned long long index0, void* stream);

static int attention_impl(const float* q, const float* k, const float* v, const float* mask,
                          long long mask_batch_stride, long long mask_q_stride, float* out, float* scores_ws,
                          int batch, int len_q, int len_k, int hidden, int n_heads, float dropout_p,
                          unsigned long long seed, unsigned long long index0, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XMLB_REQUIRE(q && k && v && mask && out && scores_ws, "xmlb_attention: null pointer");
  XMLB_REQUIRE(n_heads > 0 && hidden % n_heads == 0, "xmlb_attention: hidden %% n_heads != 0");
  XMLB_REQUIRE((long long)batch * n_heads <= 65535, "xmlb_attention: batch*n_heads > 65535, split the batch");
  if (batch == 0 || len_q == 0) return XMLB_OK;
  const int dh = hidden / n_heads;
  // 1) scores[b][h] = Q_h K_h^T / sqrt(dh) + (1 - mask) * -10000
  GemmParams p = {};
  p.A = q, p.B = k, p.C = scores_ws;
  p.M = len_q, p.N = len_k, p.K = dh;
  p.lda = hidden, p.ldb = hidden, p.ldc = len_k;
  p.b_is_kn = 0, p.batch1 = n_heads;
  p.sA0 = (long long)len_q * hidden, p.sA1 = dh;
  p.sB0 = (long long)len_k * hidden, p.sB1 = dh;
  p.sC0 = (long long)n_heads * len_q * len_k, p.sC1 = (long long)len_q * len_k;
  p.epilogue = EPI_STORE;
  p.div = sqrtf((float)dh);
  p.att_mask = mask, p.mask_s0 = mask_batch_stride, p.mask_sm = mask_q_stride;
  int rc = xmlb_gemm_launch(p, batch, stream);
  if (rc) return rc;
  // 2) softmax over keys, in place
  rc = xmlb_softmax_rows(scores_ws, scores_ws, (long long)batch * n_heads * len_q, len_k, stream_);
  if (rc) return rc;
  if (dropout_p > 0.f) {  // train mode: dropout on the probabilities (model_components.py:296)
    rc = xmlb_dropout(scores_ws, scores_ws, (long long)batch * n_heads * len_q * len_k, dropout_p, seed, index0,
                      stream_);
    if (rc) return rc;
  }
  // 3) out[b][:, h*dh:(h+1)*dh] = P_h V_h
  GemmParams o = {};
  o.A = scores_ws, o.B = v, o.C = out;
  o.M = len_q, o.N = dh, o.K = len_k;
  o.lda = len_k, o.ldb = hidden, o.ldc = hidden;
  o.b_is_kn = 1, o.batch1 = n_heads;
  o.sA0 = p.sC0, o.sA1 = p.sC1;
  o.sB0 = (long long)len_k * hidden, o.sB1 = dh;
  o.sC0 = (long long)len_q * hidden, o.sC1 = dh;
  o.epilogue = EPI_STORE;
  return xmlb_gemm_launch(o, batch, stream);
}

extern "C" int xmlb_attention(const float* q, const float* k, const float* v, const float* mask,
                              long long mask_batch_stride, long long mask_q_stride, float* out,
                              float* scores_ws, int batch, int len_q, int len_k, int hidden, int n_heads,
                              void* stream) {
  return attention_impl(q, k, v, mask, mask_batch_stride, mask_q_stride, out, scores_ws, batch, len_q, len_k,
                        hidden, n_heads, 0.f, 0ull, 0ull, stream);
}

extern "C" int xmlb_attention_train(const float* q, const float* k, const float* v, const float* mask,
                                    long long mask_batch_stride, long long mask_q_stride, float* out,
                                    float* scores_ws, int batch, int len_q, int len_k, int hidden, int n_heads,
                                    float dropout_p, unsigned long long seed, unsigned long long index0,
                                    void* stream) {
  XMLB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "xmlb_attention_train: dropout_p must be in [0, 1)");
  return attention_impl(q, k, v, mask, mask_batch_stride, mask_q_stride, out, scores_ws, batch, len_q, len_k,
                        hidden, n_heads, dropout_p, seed, index0, stream);
}

// One CTA (128 threads) per query.  smem: att[len][n_mod]
__global__ void __launch_bounds__(128) modular_pool_kernel(const float* __restrict__ enc, const float* __restrict__ mask,
                                                           const float* __restrict__ w_mod, float* __restrict__ out0,
                                                           float* __restrict__ out1, int len, int hidden, int n_mod,
                                                           const int* __restrict__ cu_seqlens) {
  extern __shared__ float att[];  // [len][2]
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* e = enc + (long long)n * len * hidden;
  if (cu_seqlens) {  // packed tokens: sequence n owns rows [cu[n], cu[n+1]), all valid
    const int r0 = __ldg(cu_seqlens + n);
    e = enc + (long long)r0 * hidden;
    len = min(len, __ldg(cu_seqlens + n + 1) - r0);
  }
  for (int t = warp; t < len; t += 4) {
    float s0 = 0.f, s1 = 0.f;
    for (int d = lane; d < hidden; d += 32) {
      const float x = e[(long long)t * hidden + d];
      s0 = fmaf(x, __ldg(w_mod + d), s0);
      if (n_mod == 2) s1 = fmaf(x, __ldg(w_mod + hidden + d), s1);
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1);
    if (lane == 0) {
      const float m = cu_seqlens ? 1.f : mask[(long long)n * len + t];
      att[t * 2 + 0] = mask_logit(s0, m);
      att[t * 2 + 1] = mask_logit(s1, m);
    }
  }
  __syncthreads();
  if (warp < n_mod) {  // softmax over tokens, one warp per modular vector
    float mx = -INFINITY;
    for (int t = lane; t < len; t += 32) mx = fmaxf(mx, att[t * 2 + warp]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < len; t += 32) s += expf(att[t * 2 + warp] - mx);
    s = warp_sum(s);
    for (int t = lane; t < len; t += 32) att[t * 2 + warp] = __fdiv_rn(expf(att[t * 2 + warp] - mx), s);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < hidden; d += 128) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < len; ++t) {
      const float x = e[(long long)t * hidden + d];
      a0 = fmaf(att[t * 2 + 0], x, a0);
      a1 = fmaf(att[t * 2 + 1], x, a1);
    }
    out0[(long long)n * hidden + d] = a0;
    if (n_mod == 2) out1[(long long)n * hidden + d] = a1;
  }
}

static int modular_pool_launch(const float* encoded, const float* mask, const int* cu_seqlens, const float* w_mod,
                               float* out0, float* out1, int n_queries, int len, int hidden, int n_mod, void* stream) {
  XMLB_REQUIRE(encoded && (mask || cu_seqlens) && w_mod && out0, "xmlb_modular_pool: null pointer");
  XMLB_REQUIRE(n_mod == 1 || (n_mod == 2 && out1), "xmlb_modular_pool: n_mod must be 1 or 2 (with out1)");
  XMLB_REQUIRE(len > 0 && len <= 4096, "xmlb_modular_pool: len out of range");
  if (n_queries == 0) return XMLB_OK;
  modular_pool_kernel<<<n_queries, 128, len * 2 * sizeof(float), (cudaStream_t)stream>>>(
      encoded, mask, w_mod, out0, out1, len, hidden, n_mod, cu_seqlens);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_modular_pool(const float* encoded, const float* mask, const float* w_mod, float* out0,
                                 float* out1, int n_queries, int len, int hidden, int n_mod, void* stream) {
  XMLB_REQUIRE(mask, "xmlb_modular_pool: null pointer");
  return modular_pool_launch(encoded, mask, nullptr, w_mod, out0, out1, n_queries, len, hidden, n_mod, stream);
}

extern "C" int xmlb_modular_pool_ragged(const float* encoded, const int* cu_seqlens, const float* w_mod, float* out0,
                                        float* out1, int n_queries, int max_len, int hidden, int n_mod,
                                        void* stream) {
  XMLB_REQUIRE(cu_seqlens, "xmlb_modular_pool_ragged: null pointer");
  return modular_pool_launch(encoded, nullptr, cu_seqlens, w_mod, out0, out1, n_queries, max_len, hidden, n_mod,
                             stream);
}
