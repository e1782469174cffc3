// Training-step kernels (BASELINE config #4), all HBM-bound:
//   xmlb_dropout         : nn.Dropout in train mode (model_components.py:77,152,263,311), counter-based mask so the
//                          backward pass re-applies the SAME mask from (seed, element index) -- nothing is stored
//   xmlb_bert_adam_step  : BertAdam.step (optimization.py:273-338) for ALL parameter tensors in two launches:
//                          per-tensor gradient-norm clip, moment update, decoupled weight decay, no bias correction
#include "common.cuh"
#include "xmlb200.h"

// ---------------------------------------------------------------------------------------------- dropout
// keep(i) depends only on (seed, i): dropout_keep in common.cuh
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                                      uint32_t threshold, float scale, unsigned long long seed,
                                                      unsigned long long index0) {
  // 4 consecutive elements per thread; x may be NULL (mask of ones * scale)
  const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (i0 + 4 <= n && ((((uintptr_t)x | (uintptr_t)out) & 15) == 0)) {
    float4 v = x ? *reinterpret_cast<const float4*>(x + i0) : make_float4(1.f, 1.f, 1.f, 1.f);
    v.x = dropout_keep(seed, index0 + i0 + 0, threshold) ? v.x * scale : 0.f;
    v.y = dropout_keep(seed, index0 + i0 + 1, threshold) ? v.y * scale : 0.f;
    v.z = dropout_keep(seed, index0 + i0 + 2, threshold) ? v.z * scale : 0.f;
    v.w = dropout_keep(seed, index0 + i0 + 3, threshold) ? v.w * scale : 0.f;
    *reinterpret_cast<float4*>(out + i0) = v;
  } else {
    for (long long i = i0; i < n && i < i0 + 4; ++i)
      out[i] = dropout_keep(seed, index0 + i, threshold) ? (x ? x[i] : 1.f) * scale : 0.f;
  }
}

extern "C" int xmlb_dropout(const float* x, float* out, long long n, float p, unsigned long long seed,
                            unsigned long long index0, void* stream) {
  XMLB_REQUIRE(out && n >= 0, "xmlb_dropout: bad argument");
  XMLB_REQUIRE(p >= 0.f && p < 1.f, "xmlb_dropout: p must be in [0, 1)");
  if (n == 0) return XMLB_OK;
  const double t = (double)p * 4294967296.0;
  const uint32_t threshold = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  dropout_kernel<<<ceil_div(n, 1024), 256, 0, (cudaStream_t)stream>>>(x, out, n, threshold, 1.f / (1.f - p), seed,
                                                                     index0);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

// ---------------------------------------------------------------------------------------------- BertAdam
// Work is cut into chunks of <= BA_CHUNK elements of ONE tensor; chunk table row = {param, grad, m, v} pointers
// (already offset to the chunk), element count, tensor id.  Launch 1 writes the sum of squared gradients of every
// chunk; launch 2 re-derives each tensor's norm from its chunks' partial sums IN FIXED ORDER (deterministic, no
// float atomics) and applies the update.
struct BaChunk {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
  long long tensor;
};
static_assert(sizeof(BaChunk) == 48, "BaChunk is 6 x 8 bytes (the host packs it as int64[6])");

__global__ void __launch_bounds__(256) bert_adam_sqnorm_kernel(const BaChunk* __restrict__ chunks,
                                                               float* __restrict__ partial) {
  __shared__ float red[8];
  const BaChunk c = chunks[blockIdx.x];
  float s = 0.f;
  const long long n4 = ((uintptr_t)c.g & 15) == 0 ? c.n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += 256) {
    const float4 g = reinterpret_cast<const float4*>(c.g)[i];
    s = fmaf(g.x, g.x, s), s = fmaf(g.y, g.y, s), s = fmaf(g.z, g.z, s), s = fmaf(g.w, g.w, s);
  }
  for (long long i = n4 * 4 + threadIdx.x; i < c.n; i += 256) s = fmaf(c.g[i], c.g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

struct BaTensor {
  int first_chunk, n_chunks;
  float lr, weight_decay;  // lr already multiplied by the schedule (optimization.py:321-322)
};

__global__ void __launch_bounds__(256) bert_adam_update_kernel(const BaChunk* __restrict__ chunks,
                                                               const BaTensor* __restrict__ tensors,
                                                               const float* __restrict__ partial, float b1,
                                                               float one_minus_b1, float b2, float one_minus_b2,
                                                               float eps, float max_grad_norm) {
  __shared__ float clip_s;
  const BaChunk c = chunks[blockIdx.x];
  const BaTensor t = tensors[c.tensor];
  if (threadIdx.x < 32) {
    float clip = 1.f;
    if (max_grad_norm > 0.f) {  // torch.nn.utils.clip_grad_norm_(p, max_norm): coef = max_norm / (norm + 1e-6)
      float s = 0.f;
      for (int i = threadIdx.x; i < t.n_chunks; i += 32) s += partial[t.first_chunk + i];
      s = warp_sum(s);
      const float coef = max_grad_norm / (sqrtf(s) + 1e-6f);
      if (coef < 1.f) clip = coef;
    }
    if (threadIdx.x == 0) clip_s = clip;
  }
  __syncthreads();
  const float clip = clip_s;
  const bool scale_grad = clip != 1.f;
  for (long long i = threadIdx.x; i < c.n; i += 256) {
    float g = c.g[i];
    if (scale_grad) {
      g *= clip;
      c.g[i] = g;  // clip_grad_norm_ rescales the stored gradient in place
    }
    const float m = __fadd_rn(__fmul_rn(c.m[i], b1), __fmul_rn(one_minus_b1, g));
    const float v = __fadd_rn(__fmul_rn(c.v[i], b2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
    float update = __fdiv_rn(m, __fadd_rn(sqrtf(v), eps));
    const float p = c.p[i];
    if (t.weight_decay > 0.f) update = __fadd_rn(update, __fmul_rn(t.weight_decay, p));
    c.m[i] = m, c.v[i] = v;
    c.p[i] = __fsub_rn(p, __fmul_rn(t.lr, update));
  }
}

extern "C" int xmlb_bert_adam_step(const long long* chunk_table, int n_chunks, const int* tensor_table,
                                   int n_tensors, float* partial_ws, double b1, double b2, double eps,
                                   double max_grad_norm, void* stream) {
  XMLB_REQUIRE(chunk_table && tensor_table && partial_ws, "xmlb_bert_adam_step: null pointer");
  XMLB_REQUIRE(n_chunks >= 0 && n_tensors >= 0, "xmlb_bert_adam_step: negative count");
  XMLB_REQUIRE(b1 >= 0. && b1 < 1. && b2 >= 0. && b2 < 1. && eps >= 0., "xmlb_bert_adam_step: bad hyper-parameter");
  if (n_chunks == 0) return XMLB_OK;
  const BaChunk* chunks = reinterpret_cast<const BaChunk*>(chunk_table);
  const BaTensor* tensors = reinterpret_cast<const BaTensor*>(tensor_table);
  int launches = 1;
  if (max_grad_norm > 0.) {
    bert_adam_sqnorm_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(chunks, partial_ws);
    ++launches;
  }
  // hyper-parameters arrive as doubles and are rounded to fp32 ONCE each, like the Python scalars of the reference
  // (1 - b2 evaluated in fp32 from 0.999f would be off by 4.7e-5 relative)
  bert_adam_update_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(chunks, tensors, partial_ws, (float)b1,
                                                                      (float)(1.0 - b1), (float)b2, (float)(1.0 - b2),
                                                                      (float)eps, (float)max_grad_norm);
  xmlb_count_launch(launches);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
