// Fused multi-head self-attention for SHORT ragged sequences (the query encoder: <= 32 tokens per query):
//   out = merge_heads(softmax(Q_h K_h^T / sqrt(dh)) V_h)        reference model_components.py:277-303
// on a PACKED token layout: sequence s owns rows [cu_seqlens[s], cu_seqlens[s+1]) of q / k / v / out (T, hidden).
// Only valid tokens exist, so the reference's additive key mask (1 - mask) * -10000 never applies: a padded key has
// probability exp(-10000 - max) == 0 exactly in fp32 and a padded query row is never read by the pooling, hence the
// packed computation returns the reference's values for every valid token.
// One CTA (128 threads) per (sequence, head).  The head's Q / K / V rows pass through shared memory in CHUNKS of 64
// head dimensions (16-byte cp.async copies of the valid rows only, all of a chunk's copies in flight at once): S
// accumulates over the Q / K chunks in registers (2 x 4 per thread), the 32 x 32 probabilities stay in shared memory,
// O = P V is produced 64 columns at a time from double-buffered V chunks.  21.6 KB of shared memory per CTA keeps
// 10 CTAs per SM in flight -- the kernel is a chain of global-load latencies, and the first version (whole head
// staged at once: 79 KB, 2 CTAs per SM, a serial staging loop) spent 2.6 ms per 10 K queries waiting on them.
// Same accumulation order over d and over the keys as before: bit-identical outputs.  No (T, T) workspace.
#include <math.h>
#include "common.cuh"
#include "xmlb200.h"

namespace {

constexpr int RL = 32;        // max tokens per sequence
constexpr int DC = 64;        // head dimensions per staged chunk
constexpr int LDC = DC + 4;   // Q / K chunk row stride: consecutive rows start 4 banks apart (conflict-free float4 reads)

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(128) attention_ragged_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v,
                                                               const int* __restrict__ cu_seqlens,
                                                               float* __restrict__ out, int hidden, int dh, int in_ld) {
  __shared__ __align__(16) float Qs[RL * LDC];       // Q chunk [RL][LDC]; later V chunks (even) as [RL][DC]
  __shared__ __align__(16) float Ks[RL * LDC];       // K chunk [RL][LDC]; later V chunks (odd)
  __shared__ float Ps[RL * (RL + 1)];
  const int s = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int row0 = __ldg(cu_seqlens + s);
  const int len = min(__ldg(cu_seqlens + s + 1) - row0, RL);
  if (len <= 0) return;
  const long long g0 = (long long)row0 * in_ld + h * dh;  // q / k / v rows are in_ld floats apart
  // ---- S = Q K^T: thread (ti, tj) owns rows {2 ti, 2 ti + 1} x columns {tj, tj + 8, tj + 16, tj + 24}; rows and
  // columns beyond the sequence are skipped (a warp covers rows 8 w .. 8 w + 7: whole warps drop out for short
  // queries).  Shared-memory rows >= len are never written: whatever they hold only reaches accumulators / outputs
  // of rows >= len, which are discarded.
  const int ti = tid >> 3, tj = tid & 7;
  const int nc = (len - tj + 7) >> 3;  // columns tj + 8 c < len  <=>  c < nc  (0 when tj >= len)
  float acc[2][4] = {};
  for (int d0 = 0; d0 < dh; d0 += DC) {
    const int dc = min(DC, dh - d0), vec = dc >> 2;
    for (int i = tid; i < len * vec; i += 128) {
      const int r = i / vec, c = (i - r * vec) * 4;
      const long long g = g0 + (long long)r * in_ld + d0 + c;
      cp_async16(Qs + r * LDC + c, q + g);
      cp_async16(Ks + r * LDC + c, k + g);
    }
    cp_async_wait_all();
    __syncthreads();
    if (2 * ti < len) {
      const float* q0 = Qs + (2 * ti) * LDC;
      const float* q1 = q0 + LDC;
      for (int d = 0; d < dc; d += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(q0 + d);
        const float4 a1 = *reinterpret_cast<const float4*>(q1 + d);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nc) {
            const float4 b = *reinterpret_cast<const float4*>(Ks + (tj + 8 * c) * LDC + d);
            acc[0][c] = fmaf(a0.x, b.x, acc[0][c]), acc[0][c] = fmaf(a0.y, b.y, acc[0][c]);
            acc[0][c] = fmaf(a0.z, b.z, acc[0][c]), acc[0][c] = fmaf(a0.w, b.w, acc[0][c]);
            acc[1][c] = fmaf(a1.x, b.x, acc[1][c]), acc[1][c] = fmaf(a1.y, b.y, acc[1][c]);
            acc[1][c] = fmaf(a1.z, b.z, acc[1][c]), acc[1][c] = fmaf(a1.w, b.w, acc[1][c]);
          }
        }
      }
    }
    __syncthreads();  // the chunk buffers are overwritten by the next chunk / the first V chunks
  }
  // the first two V chunks travel while the softmax runs
  const int n_vchunks = (dh + DC - 1) / DC;
  auto stage_v = [&](int ch) {
    float* Vs = (ch & 1) ? Ks : Qs;
    const int d0 = ch * DC, vec = min(DC, dh - d0) >> 2;
    for (int i = tid; i < len * vec; i += 128) {
      const int r = i / vec, c = (i - r * vec) * 4;
      cp_async16(Vs + r * DC + c, v + g0 + (long long)r * in_ld + d0 + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_v(0);
  if (n_vchunks > 1) stage_v(1);
  if (2 * ti < len) {
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
  // ---- O = P V, 64 columns per chunk: thread (ti, tj) owns rows {2 ti, 2 ti + 1} x columns {4 tj + 32 m .. + 3} ----
  const int r0 = 2 * ti;
  const float* p0 = Ps + r0 * (RL + 1);
  const float* p1 = p0 + (RL + 1);
  for (int ch = 0; ch < n_vchunks; ++ch) {
    // chunk ch has landed once at most one younger group (chunk ch + 1) is still pending
    if (ch + 1 < n_vchunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // (first pass: also publishes the probabilities)
    const float* Vs = (ch & 1) ? Ks : Qs;
    const int d0 = ch * DC, dc = min(DC, dh - d0);
    if (r0 < len) {
      for (int c0 = 4 * tj; c0 < dc; c0 += 32) {
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        for (int j = 0; j < len; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(Vs + j * DC + c0);
          const float a0 = p0[j], a1 = p1[j];
          o0.x = fmaf(a0, b.x, o0.x), o0.y = fmaf(a0, b.y, o0.y), o0.z = fmaf(a0, b.z, o0.z), o0.w = fmaf(a0, b.w, o0.w);
          o1.x = fmaf(a1, b.x, o1.x), o1.y = fmaf(a1, b.y, o1.y), o1.z = fmaf(a1, b.z, o1.z), o1.w = fmaf(a1, b.w, o1.w);
        }
        float* g = out + (long long)(row0 + r0) * hidden + h * dh + d0 + c0;
        *reinterpret_cast<float4*>(g) = o0;
        if (r0 + 1 < len) *reinterpret_cast<float4*>(g + hidden) = o1;
      }
    }
    if (ch + 2 < n_vchunks) {
      __syncthreads();  // everyone is done with this buffer before chunk ch + 2 overwrites it
      stage_v(ch + 2);
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
  attention_ragged_kernel<<<dim3(n_seqs, n_heads), 128, 0, (cudaStream_t)stream>>>(q, k, v, cu_seqlens, out, hidden, dh,
                                                                                 in_ld);
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
