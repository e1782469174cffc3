// Exact-fp32 SIMT GEMM launcher + the C-ABI entry points built on it:
//   xmlb_linear         : reference nn.Linear (+ReLU, +residual)   -- model_components.py:156-163, 313-317
//   xmlb_vr_scores_f32  : reference XML.get_video_level_scores       -- model_xml.py:446-452 (exact-fp32 variant)
#include "gemm_simt.cuh"
#include "xmlb200.h"

int xmlb_gemm_launch(const GemmParams& p, int batch0, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || batch0 <= 0) return XMLB_OK;
  const int nb = batch0 * p.batch1;
  if (p.M <= 64 || p.N <= 64) {
    dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), nb);
    gemm_simt_kernel<64, 64, 1, 1><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128), nb);
    gemm_simt_kernel<128, 128, 2, 2><<<grid, 256, 0, stream>>>(p);
  }
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_linear(const float* x, const float* weight, const float* bias, const float* residual,
                           float* out, long long rows, int out_dim, int in_dim, int relu, void* stream) {
  XMLB_REQUIRE(x && weight && out, "xmlb_linear: null pointer");
  XMLB_REQUIRE(rows >= 0 && rows < (1ll << 31) && out_dim > 0 && in_dim > 0, "xmlb_linear: bad shape");
  GemmParams p = {};
  p.A = x, p.B = weight, p.C = out;
  p.M = (int)rows, p.N = out_dim, p.K = in_dim;
  p.lda = in_dim, p.ldb = in_dim, p.ldc = out_dim;
  p.b_is_kn = 0, p.batch1 = 1;
  p.epilogue = EPI_STORE;
  p.bias = bias, p.residual = residual, p.relu = relu;
  return xmlb_gemm_launch(p, 1, (cudaStream_t)stream);
}

__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}

// out[q][v] (+)= (max over valid clips l of  qn[q] . c1n[v][l]) / divisor, accumulated over modalities.
__global__ void vr_combine_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                  float* __restrict__ out, long long n, float divisor) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s = a[i];
    if (b) s = __fadd_rn(s, b[i]);
    out[i] = __fdiv_rn(s, divisor);
  }
}

extern "C" int xmlb_vr_scores_f32(const float* q_video_n, const float* q_sub_n, const float* feat1_video_n,
                                  const float* feat1_sub_n, const float* video_mask, const float* sub_mask,
                                  float* q2c, float* workspace, int n_queries, int n_videos, int ctx_len,
                                  int hidden, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool use_v = q_video_n && feat1_video_n, use_s = q_sub_n && feat1_sub_n;
  XMLB_REQUIRE(use_v || use_s, "xmlb_vr_scores_f32: no modality given");
  XMLB_REQUIRE(q2c && workspace, "xmlb_vr_scores_f32: null output/workspace");
  XMLB_REQUIRE((!use_v || video_mask) && (!use_s || sub_mask), "xmlb_vr_scores_f32: mask missing");
  XMLB_REQUIRE((long long)n_videos * ctx_len < (1ll << 31), "xmlb_vr_scores_f32: corpus too large for one call");
  if (n_queries == 0 || n_videos == 0) return XMLB_OK;
  const long long n_out = (long long)n_queries * n_videos;
  float* part[2] = {workspace, workspace + n_out};
  const int n_mod = (use_v ? 1 : 0) + (use_s ? 1 : 0);
  const int blocks = (int)((n_out * n_mod + 255) / 256 < 148 * 8 ? (n_out * n_mod + 255) / 256 : 148 * 8);
  fill_kernel<<<blocks, 256, 0, stream>>>(workspace, n_out * n_mod, MASK_FILL);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  int slot = 0;
  for (int mod = 0; mod < 2; ++mod) {
    if ((mod == 0 && !use_v) || (mod == 1 && !use_s)) continue;
    GemmParams p = {};
    p.A = mod == 0 ? q_video_n : q_sub_n;
    p.B = mod == 0 ? feat1_video_n : feat1_sub_n;
    p.C = part[slot++];
    p.M = n_queries, p.N = n_videos * ctx_len, p.K = hidden;
    p.lda = hidden, p.ldb = hidden, p.ldc = 0;
    p.b_is_kn = 0, p.batch1 = 1;
    p.epilogue = EPI_VRMAX;
    p.clip_mask = mod == 0 ? video_mask : sub_mask;
    p.L = ctx_len, p.n_videos = n_videos;
    int rc = xmlb_gemm_launch(p, 1, stream);
    if (rc) return rc;
  }
  vr_combine_kernel<<<blocks, 256, 0, stream>>>(part[0], n_mod == 2 ? part[1] : nullptr, q2c, n_out,
                                                (float)n_mod);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
