// Row-wise HBM-bound kernels: (add +) LayerNorm, L2 normalisation, softmax.
//   xmlb_add_layernorm : reference nn.LayerNorm(eps=1e-5) with the fused adds of
//                        TrainablePositionalEncoding.forward (model_components.py:81-88),
//                        BertSelfOutput.forward (:313-317) and the cross-attention residual (model_xml.py:371)
//   xmlb_l2norm_rows   : F.normalize(x, dim=-1) (model_xml.py:446-447): x / max(||x||_2, 1e-12)
//   xmlb_softmax_rows  : nn.Softmax(dim=-1) (model_components.py:293), F.softmax (inference.py:321-322)
// One warp per row, 8 rows per CTA; rows are read through L1 so the 2nd/3rd pass never touch HBM.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "xmlb200.h"

__device__ __forceinline__ void ln_split16(float x, int is_bf16, unsigned short& hi, unsigned short& lo) {
  if (is_bf16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h), lo = __bfloat16_as_ushort(l);
  } else {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h), lo = __half_as_ushort(l);
  }
}

// out (fp32) and / or (out_hi, out_lo) = the 16-bit split of the same values, rows of kpad elements (zero padded):
// the operand format of the tensor-core Linear that follows every LayerNorm of the encoders.
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                            long long add_rows, const int* __restrict__ add_index,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out,
                                                            unsigned short* __restrict__ out_hi,
                                                            unsigned short* __restrict__ out_lo, int kpad, int is_bf16,
                                                            long long rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  const float* ar = add ? add + (add_index ? (long long)__ldg(add_index + row) : row % add_rows) * dim : nullptr;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += ar ? xr[i] + ar[i] : xr[i];
  const float mean = warp_sum(s) / (float)dim;
  float v = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float d = (ar ? xr[i] + ar[i] : xr[i]) - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = 1.f / sqrtf(warp_sum(v) / (float)dim + eps);
  float* o = out ? out + row * dim : nullptr;
  unsigned short* oh = out_hi ? out_hi + row * kpad : nullptr;
  unsigned short* ol = out_hi ? out_lo + row * kpad : nullptr;
  for (int i = lane; i < (oh ? kpad : dim); i += 32) {
    float y = 0.f;
    if (i < dim) {
      const float d = (ar ? xr[i] + ar[i] : xr[i]) - mean;
      y = fmaf(d * rstd, __ldg(gamma + i), __ldg(beta + i));
      if (o) o[i] = y;
    }
    if (oh) ln_split16(y, is_bf16, oh[i], ol[i]);
  }
}

static int add_layernorm_launch(const float* x, const float* add, long long add_rows, const int* add_index,
                                const float* gamma, const float* beta, float* out, unsigned short* out_hi,
                                unsigned short* out_lo, int kpad, int is_bf16, long long rows, int dim, float eps,
                                void* stream) {
  XMLB_REQUIRE(x && gamma && beta && (out || out_hi) && dim > 0 && rows >= 0, "xmlb_add_layernorm: bad argument");
  XMLB_REQUIRE(!add || add_rows > 0, "xmlb_add_layernorm: add_rows must be > 0 when add is given");
  XMLB_REQUIRE(!out_hi || (out_lo && kpad >= dim), "xmlb_add_layernorm: split output needs out_lo and kpad >= dim");
  if (rows == 0) return XMLB_OK;
  add_layernorm_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, add, add_rows, add_index, gamma, beta,
                                                                           out, out_hi, out_lo, kpad, is_bf16, rows,
                                                                           dim, eps);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_add_layernorm(const float* x, const float* add, long long add_rows, const float* gamma,
                                  const float* beta, float* out, long long rows, int dim, float eps, void* stream) {
  return add_layernorm_launch(x, add, add_rows, nullptr, gamma, beta, out, nullptr, nullptr, 0, 0, rows, dim, eps,
                              stream);
}

extern "C" int xmlb_add_layernorm_split(const float* x, const float* add, long long add_rows, const int* add_index,
                                        const float* gamma, const float* beta, float* out, unsigned short* out_hi,
                                        unsigned short* out_lo, int kpad, int is_bf16, long long rows, int dim,
                                        float eps, void* stream) {
  XMLB_REQUIRE(out_hi && out_lo, "xmlb_add_layernorm_split: null pointer");
  return add_layernorm_launch(x, add, add_rows, add_index, gamma, beta, out, out_hi, out_lo, kpad, is_bf16, rows, dim,
                              eps, stream);
}

extern "C" int xmlb_add_layernorm_indexed(const float* x, const float* add, const int* add_index, long long add_rows,
                                          const float* gamma, const float* beta, float* out, long long rows, int dim,
                                          float eps, void* stream) {
  XMLB_REQUIRE(add && add_index, "xmlb_add_layernorm_indexed: add and add_index are required");
  return add_layernorm_launch(x, add, add_rows, add_index, gamma, beta, out, nullptr, nullptr, 0, 0, rows, dim, eps,
                              stream);
}

__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                          long long rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s = fmaf(xr[i], xr[i], s);
  const float denom = fmaxf(sqrtf(warp_sum(s)), eps);
  float* o = out + row * dim;
  for (int i = lane; i < dim; i += 32) o[i] = __fdiv_rn(xr[i], denom);
}

extern "C" int xmlb_l2norm_rows(const float* x, float* out, long long rows, int dim, float eps, void* stream) {
  XMLB_REQUIRE(x && out && dim > 0 && rows >= 0, "xmlb_l2norm_rows: bad argument");
  if (rows == 0) return XMLB_OK;
  l2norm_rows_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, out, rows, dim, eps);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* x, float* out,  // in-place allowed
                                                          
                                                           long long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float m = -INFINITY;
  for (int i = lane; i < dim; i += 32) m = fmaxf(m, xr[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += expf(xr[i] - m);
  s = warp_sum(s);
  float* o = out + row * dim;
  for (int i = lane; i < dim; i += 32) o[i] = __fdiv_rn(expf(xr[i] - m), s);
}

extern "C" int xmlb_softmax_rows(const float* x, float* out, long long rows, int dim, void* stream) {
  XMLB_REQUIRE(x && out && dim > 0 && rows >= 0, "xmlb_softmax_rows: bad argument");
  if (rows == 0) return XMLB_OK;
  softmax_rows_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, out, rows, dim);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
