// Fused multi-head self-attention for SHORT ragged sequences (the query encoder: <= 32 tokens per query):
//   out = merge_heads(softmax(Q_h K_h^T / sqrt(dh)) V_h)        reference model_components.py:277-303
// on a PACKED token layout: sequence s owns rows [cu_seqlens[s], cu_seqlens[s+1]) of q / k / v / out (T, hidden).
// Only valid tokens exist, so the reference's additive key mask (1 - mask) * -10000 never applies: a padded key has
// probability exp(-10000 - max) == 0 exactly in fp32 and a padded query row is never read by the pooling, hence the
// packed computation returns the reference's values for every valid token.
// One CTA (128 threads) per (sequence, head): Q_h, K_h, V_h (<= 32 x dh fp32 each) staged in shared memory, the
// 32 x 32 score tile and the 32 x dh output register-blocked 2 x 4 per thread -- no (T, T) workspace, no separate
// softmax pass (the unfused path spends 27 % of the query-encoder time in two batched 30 x 30 GEMMs + a softmax).
#include <math.h>
#include "common.cuh"
#include "xmlb200.h"

namespace {

constexpr int RL = 32;   // max tokens per sequence
constexpr int PAD = 4;   // row stride dh + 4 floats: for dh % 32 == 0 consecutive rows start 4 banks apart
                         // (conflict-free float4 reads); other head sizes are merely slower

__global__ void __launch_bounds__(128) attention_ragged_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v,
                                                               const int* __restrict__ cu_seqlens,
                                                               float* __restrict__ out, int hidden, int dh, int in_ld) {
  extern __shared__ __align__(16) float smem[];
  const int ld = dh + PAD;
  float* Qs = smem;                 // [RL][ld]
  float* Ks = Qs + RL * ld;         // [RL][ld]
  float* Vs = Ks + RL * ld;         // [RL][dh]
  float* Ps = Vs + RL * dh;         // [RL][RL + 1]
  const int s = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int row0 = __ldg(cu_seqlens + s);
  const int len = min(__ldg(cu_seqlens + s + 1) - row0, RL);
  if (len <= 0) return;
  const int vec = dh / 4;
  // ---- stage the head's Q, K, V rows (zero rows beyond the sequence) ----
  for (int i = tid; i < RL * vec; i += 128) {
    const int r = i / vec, c = (i - r * vec) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, d = a;
    if (r < len) {
      const long long g = (long long)(row0 + r) * in_ld + h * dh + c;  // q / k / v rows are in_ld floats apart
      a = __ldg(reinterpret_cast<const float4*>(q + g));
      b = __ldg(reinterpret_cast<const float4*>(k + g));
      d = __ldg(reinterpret_cast<const float4*>(v + g));
    }
    *reinterpret_cast<float4*>(Qs + r * ld + c) = a;
    *reinterpret_cast<float4*>(Ks + r * ld + c) = b;
    *reinterpret_cast<float4*>(Vs + r * dh + c) = d;
  }
  __syncthreads();
  // ---- S = Q K^T: thread (ti, tj) owns rows {2 ti, 2 ti + 1} x columns {tj, tj + 8, tj + 16, tj + 24}; rows and
  // columns beyond the sequence are skipped (a warp covers rows 8 w .. 8 w + 7: whole warps drop out for short queries)
  const int ti = tid >> 3, tj = tid & 7;
  if (2 * ti < len) {
    float acc[2][4] = {};
    const float* q0 = Qs + (2 * ti) * ld;
    const float* q1 = q0 + ld;
    const int nc = (len - tj + 7) >> 3;  // columns tj + 8 c < len  <=>  c < nc  (0 when tj >= len)
    for (int d = 0; d < dh; d += 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(q0 + d);
      const float4 a1 = *reinterpret_cast<const float4*>(q1 + d);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nc) {
          const float4 b = *reinterpret_cast<const float4*>(Ks + (tj + 8 * c) * ld + d);
          acc[0][c] = fmaf(a0.x, b.x, acc[0][c]), acc[0][c] = fmaf(a0.y, b.y, acc[0][c]);
          acc[0][c] = fmaf(a0.z, b.z, acc[0][c]), acc[0][c] = fmaf(a0.w, b.w, acc[0][c]);
          acc[1][c] = fmaf(a1.x, b.x, acc[1][c]), acc[1][c] = fmaf(a1.y, b.y, acc[1][c]);
          acc[1][c] = fmaf(a1.z, b.z, acc[1][c]), acc[1][c] = fmaf(a1.w, b.w, acc[1][c]);
        }
      }
    }
    const float div = sqrtf((float)dh);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nc) Ps[(2 * ti + r) * (RL + 1) + tj + 8 * c] = __fdiv_rn(acc[r][c], div);
  }
  __syncthreads();
  // ---- softmax over the valid keys: warp w handles rows 8 w .. 8 w + 7, lane = key ----
  {
    const int w = tid >> 5, lane = tid & 31;
    for (int r = 8 * w; r < 8 * w + 8; ++r) {
      if (r >= len) break;  // warp-uniform
      const float x = lane < len ? Ps[r * (RL + 1) + lane] : -INFINITY;
      const float m = warp_max(x);
      const float e = lane < len ? expf(x - m) : 0.f;
      const float sum = warp_sum(e);
      Ps[r * (RL + 1) + lane] = __fdiv_rn(e, sum);  // exactly 0 for lane >= len
    }
  }
  __syncthreads();
  // ---- O = P V: thread (ti, tj) owns rows {2 ti, 2 ti + 1} x columns {4 tj + 32 m .. + 3} ----
  {
    const int r0 = 2 * ti;
    if (r0 >= len) return;
    const float* p0 = Ps + r0 * (RL + 1);
    const float* p1 = p0 + (RL + 1);
    for (int c0 = 4 * tj; c0 < dh; c0 += 32) {
      float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
      for (int j = 0; j < len; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(Vs + j * dh + c0);
        const float a0 = p0[j], a1 = p1[j];
        o0.x = fmaf(a0, b.x, o0.x), o0.y = fmaf(a0, b.y, o0.y), o0.z = fmaf(a0, b.z, o0.z), o0.w = fmaf(a0, b.w, o0.w);
        o1.x = fmaf(a1, b.x, o1.x), o1.y = fmaf(a1, b.y, o1.y), o1.z = fmaf(a1, b.z, o1.z), o1.w = fmaf(a1, b.w, o1.w);
      }
      float* g = out + (long long)(row0 + r0) * hidden + h * dh + c0;
      *reinterpret_cast<float4*>(g) = o0;
      if (r0 + 1 < len) *reinterpret_cast<float4*>(g + hidden) = o1;
    }
  }
}

}  // namespace

static int attention_ragged_launch(const float* q, const float* k, const float* v, int in_ld, const int* cu_seqlens,
                                   float* out, int n_seqs, int max_len, int hidden, int n_heads, void* stream) {
  XMLB_REQUIRE(q && k && v && cu_seqlens && out, "xmlb_attention_ragged: null pointer");
  XMLB_REQUIRE(in_ld >= hidden && in_ld % 4 == 0, "xmlb_attention_ragged: in_ld must be >= hidden and a multiple of 4");
  XMLB_REQUIRE(n_heads > 0 && hidden % n_heads == 0, "xmlb_attention_ragged: hidden %% n_heads != 0");
  const int dh = hidden / n_heads;
  XMLB_REQUIRE(dh % 4 == 0, "xmlb_attention_ragged: head size must be a multiple of 4 (16-byte row pieces)");
  XMLB_REQUIRE(max_len >= 1 && max_len <= RL, "xmlb_attention_ragged: sequences longer than 32 tokens are not supported");
  XMLB_REQUIRE(n_heads <= 65535, "xmlb_attention_ragged: too many heads");
  XMLB_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0 && hidden % 4 == 0,
               "xmlb_attention_ragged: rows must be 16-byte aligned");
  if (n_seqs == 0) return XMLB_OK;
  const size_t smem = sizeof(float) * ((size_t)2 * RL * (dh + PAD) + (size_t)RL * dh + (size_t)RL * (RL + 1));
  XMLB_REQUIRE(smem <= 227 * 1024, "xmlb_attention_ragged: head size too large for shared memory");
  XMLB_CUDA(cudaFuncSetAttribute(attention_ragged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_ragged_kernel<<<dim3(n_seqs, n_heads), 128, smem, (cudaStream_t)stream>>>(q, k, v, cu_seqlens, out, hidden,
                                                                                    dh, in_ld);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_attention_ragged(const float* q, const float* k, const float* v, const int* cu_seqlens,
                                     float* out, int n_seqs, int max_len, int hidden, int n_heads, void* stream) {
  return attention_ragged_launch(q, k, v, hidden, cu_seqlens, out, n_seqs, max_len, hidden, n_heads, stream);
}

extern "C" int xmlb_attention_ragged_qkv(const float* qkv, const int* cu_seqlens, float* out, int n_seqs, int max_len,
                                         int hidden, int n_heads, void* stream) {
  XMLB_REQUIRE(qkv, "xmlb_attention_ragged_qkv: null pointer");
  return attention_ragged_launch(qkv, qkv + hidden, qkv + 2 * hidden, 3 * hidden, cu_seqlens, out, n_seqs, max_len,
                                 hidden, n_heads, stream);
}
