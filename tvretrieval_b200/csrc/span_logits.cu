// Query x clip similarity curves + ConvSE start/end detectors + mask (+ softmax over clips).
//   reference XML.get_merged_st_ed_prob (model_xml.py:455-502), XML._get_st_ed_prob (:512-551) and the
//   softmax of the driver (inference.py:153-154, 321-322).
// Exact-fp32 SIMT kernel.  One CTA = one video x up to 32 queries, so every video's (L, H) tiles are read
// once per 32 queries that selected it ("inverted lists", xmlb_build_pair_lists) instead of once per
// (query, video) pair; the dense variant (all queries x all videos) serves get_pred_from_raw_query.
#include "common.cuh"
#include "xmlb200.h"

struct SpanParams {
  const float* q[2];       // projected query vectors per stream (Nq, H)         (stream 1 may be null)
  const float* f2[2];      // second-level context features per stream (Nv, L, H)
  const float* mask[2];    // (Nv, L) float {0,1}
  const float* w_st[2];    // ConvSE taps per stream [ksize]; merged mode uses index 0 only
  const float* w_ed[2];
  int n_streams, merged, ksize, softmax;
  int Nq, Nv, L, H;
  const int* chunk_ptr;    // [Nv+1] list mode; null -> dense
  const int* vid_ptr;      // [Nv+1]
  const int* entry_q;      // [E]
  const int* entry_out;    // [E] output row
  float* out_st;
  float* out_ed;           // [rows][L]
};

template <int CL>
__global__ void __launch_bounds__(256) span_logits_kernel(const SpanParams p) {
  constexpr int QT = 32, BK = 32, LP = 32 * CL, QS = 36;
  extern __shared__ __align__(16) float smem[];
  float* Fs = smem;                  // [BK][LP + 1]
  float* Qs = Fs + BK * (LP + 1);    // [QT][QS]
  float* S0 = Qs + QT * QS;          // [QT][LP]  similarity of stream 0 (or merged)
  float* S1 = S0 + QT * LP;          // [QT][LP]  similarity of stream 1
  __shared__ int q_idx[QT];
  __shared__ int out_row[QT];
  __shared__ int s_video;

  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int v;
  if (p.chunk_ptr == nullptr) {  // dense
    v = blockIdx.y;
    const int q0 = blockIdx.x * QT;
    if (t < QT) {
      const int q = q0 + t;
      q_idx[t] = q < p.Nq ? q : -1;
      out_row[t] = q < p.Nq ? q * p.Nv + v : -1;
    }
  } else {
    const int c = blockIdx.x;
    if (c >= p.chunk_ptr[p.Nv]) return;
    if (t == 0) {  // largest v with chunk_ptr[v] <= c
      int lo = 0, hi = p.Nv - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.chunk_ptr[mid] <= c) lo = mid; else hi = mid - 1;
      }
      s_video = lo;
    }
    __syncthreads();
    v = s_video;
    const int e0 = p.vid_ptr[v] + (c - p.chunk_ptr[v]) * QT;
    if (t < QT) {
      const bool ok = e0 + t < p.vid_ptr[v + 1];
      q_idx[t] = ok ? p.entry_q[e0 + t] : -1;
      out_row[t] = ok ? p.entry_out[e0 + t] : -1;
    }
  }
  __syncthreads();

  float acc[2][4][CL];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < CL; ++c) acc[s][i][c] = 0.f;

#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (s >= p.n_streams) break;
    const float* __restrict__ F = p.f2[s] + (long long)v * p.L * p.H;
    const float* __restrict__ Q = p.q[s];
    for (int k0 = 0; k0 < p.H; k0 += BK) {
      __syncthreads();
      const bool kok = k0 + lane < p.H;
      for (int l = w; l < LP; l += 8)
        Fs[lane * (LP + 1) + l] = (l < p.L && kok) ? __ldg(F + (long long)l * p.H + k0 + lane) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = q_idx[w * 4 + i];
        Qs[(w * 4 + i) * QS + lane] = (q >= 0 && kok) ? __ldg(Q + (long long)q * p.H + k0 + lane) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        float4 qa[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4*>(&Qs[(w * 4 + i) * QS + 4 * k4]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float f[CL];
#pragma unroll
          for (int c = 0; c < CL; ++c) f[c] = Fs[(4 * k4 + kk) * (LP + 1) + lane + 32 * c];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = kk == 0 ? qa[i].x : kk == 1 ? qa[i].y : kk == 2 ? qa[i].z : qa[i].w;
#pragma unroll
            for (int c = 0; c < CL; ++c) acc[s][i][c] = fmaf(a, f[c], acc[s][i][c]);
          }
        }
      }
    }
  }

  // similarity rows are private to the warp that owns the 4 queries
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = w * 4 + i;
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      const int l = lane + 32 * c;
      if (p.merged) {
        S0[qi * LP + l] = __fdiv_rn(__fadd_rn(acc[0][i][c], acc[1][i][c]), 2.f);  // model_xml.py:466
      } else {
        S0[qi * LP + l] = acc[0][i][c];
        S1[qi * LP + l] = acc[1][i][c];
      }
    }
  }
  __syncwarp();

  const int pad = p.ksize / 2;
  const float divisor = (float)p.n_streams;
  for (int i = 0; i < 4; ++i) {
    const int qi = w * 4 + i;
    const int row = out_row[qi];
    if (row < 0) continue;  // warp-uniform
    float st[CL], ed[CL];
    float mx_st = -INFINITY, mx_ed = -INFINITY;
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      const int l = lane + 32 * c;
      st[c] = 0.f, ed[c] = 0.f;
      if (l < p.L) {
        const int n_conv = p.merged ? 1 : p.n_streams;
        for (int s = 0; s < n_conv; ++s) {
          const float* S = s == 0 ? S0 : S1;
          float a = 0.f, b = 0.f;
          for (int j = 0; j < p.ksize; ++j) {
            const int src = l + j - pad;
            const float x = (src >= 0 && src < p.L) ? S[qi * LP + src] : 0.f;
            a = fmaf(__ldg(p.w_st[s] + j), x, a);
            b = fmaf(__ldg(p.w_ed[s] + j), x, b);
          }
          const float m = __ldg(p.mask[s] + (long long)v * p.L + l);
          st[c] += mask_logit(a, m);
          ed[c] += mask_logit(b, m);
        }
        if (!p.merged) {  // model_xml.py:584-585: average of the per-stream masked logits
          st[c] = __fdiv_rn(st[c], divisor);
          ed[c] = __fdiv_rn(ed[c], divisor);
        }
        mx_st = fmaxf(mx_st, st[c]);
        mx_ed = fmaxf(mx_ed, ed[c]);
      }
    }
    if (p.softmax) {
      mx_st = warp_max(mx_st), mx_ed = warp_max(mx_ed);
      float s_st = 0.f, s_ed = 0.f;
#pragma unroll
      for (int c = 0; c < CL; ++c) {
        if (lane + 32 * c < p.L) {
          st[c] = expf(st[c] - mx_st), ed[c] = expf(ed[c] - mx_ed);
          s_st += st[c], s_ed += ed[c];
        }
      }
      s_st = warp_sum(s_st), s_ed = warp_sum(s_ed);
#pragma unroll
      for (int c = 0; c < CL; ++c) st[c] = __fdiv_rn(st[c], s_st), ed[c] = __fdiv_rn(ed[c], s_ed);
    }
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      const int l = lane + 32 * c;
      if (l < p.L) {
        p.out_st[(long long)row * p.L + l] = st[c];
        p.out_ed[(long long)row * p.L + l] = ed[c];
      }
    }
  }
}

template <int CL>
static int launch_span(const SpanParams& p, int max_chunks, cudaStream_t stream) {
  constexpr int LP = 32 * CL;
  const size_t smem = (size_t)(32 * (LP + 1) + 32 * 36 + 2 * 32 * LP) * sizeof(float);
  XMLB_CUDA(cudaFuncSetAttribute(span_logits_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid;
  if (p.chunk_ptr == nullptr) {
    XMLB_REQUIRE(p.Nv <= 65535, "xmlb_span_logits: dense mode supports at most 65535 videos per call");
    grid = dim3(ceil_div(p.Nq, 32), p.Nv, 1);
  } else {
    grid = dim3(max_chunks, 1, 1);
  }
  span_logits_kernel<CL><<<grid, 256, smem, stream>>>(p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_span_logits(const float* q_a, const float* q_b, const float* feat2_a, const float* feat2_b,
                                const float* mask_a, const float* mask_b, const float* w_st_a, const float* w_ed_a,
                                const float* w_st_b, const float* w_ed_b, int ksize, int merged, int apply_softmax,
                                int n_queries, int n_videos, int ctx_len, int hidden, const int* chunk_ptr,
                                const int* vid_ptr, const int* entry_q, const int* entry_out, int max_chunks,
                                float* out_st, float* out_ed, void* stream) {
  XMLB_REQUIRE(q_a && feat2_a && mask_a && w_st_a && w_ed_a && out_st && out_ed, "xmlb_span_logits: null pointer");
  XMLB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 31, "xmlb_span_logits: ksize must be odd and <= 31");
  XMLB_REQUIRE(ctx_len >= 1 && ctx_len <= 256, "xmlb_span_logits: ctx_len must be in [1, 256]");
  const bool two = q_b && feat2_b;
  XMLB_REQUIRE(!merged || two, "xmlb_span_logits: merged mode needs both streams");
  XMLB_REQUIRE(!two || mask_b, "xmlb_span_logits: second stream needs its mask");
  XMLB_REQUIRE(merged || !two || (w_st_b && w_ed_b), "xmlb_span_logits: second stream needs its ConvSE taps");
  XMLB_REQUIRE(!chunk_ptr || (vid_ptr && entry_q && entry_out && max_chunks >= 0), "xmlb_span_logits: bad pair lists");
  if (n_videos == 0 || n_queries == 0 || (chunk_ptr && max_chunks == 0)) return XMLB_OK;
  SpanParams p = {};
  p.q[0] = q_a, p.q[1] = q_b, p.f2[0] = feat2_a, p.f2[1] = feat2_b;
  p.mask[0] = mask_a, p.mask[1] = two ? mask_b : mask_a;
  p.w_st[0] = w_st_a, p.w_ed[0] = w_ed_a, p.w_st[1] = w_st_b, p.w_ed[1] = w_ed_b;
  p.n_streams = two ? 2 : 1, p.merged = merged, p.ksize = ksize, p.softmax = apply_softmax;
  p.Nq = n_queries, p.Nv = n_videos, p.L = ctx_len, p.H = hidden;
  p.chunk_ptr = chunk_ptr, p.vid_ptr = vid_ptr, p.entry_q = entry_q, p.entry_out = entry_out;
  p.out_st = out_st, p.out_ed = out_ed;
  cudaStream_t s = (cudaStream_t)stream;
  if (ctx_len <= 32) return launch_span<1>(p, max_chunks, s);
  if (ctx_len <= 64) return launch_span<2>(p, max_chunks, s);
  if (ctx_len <= 128) return launch_span<4>(p, max_chunks, s);
  return launch_span<8>(p, max_chunks, s);
}

// ------------------------------------------------------------------------------------------------
// Inverted pair lists: (query, slot) -> selected video  ==>  per-video list of (query, output row).
// ------------------------------------------------------------------------------------------------
__global__ void pair_count_kernel(const int* __restrict__ top_idx, const unsigned char* __restrict__ slot_valid,
                                  long long n_pairs, int vid_lo, int n_videos, int* __restrict__ counts) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  if (slot_valid && !slot_valid[i]) return;
  const int v = top_idx[i] - vid_lo;
  if (v >= 0 && v < n_videos) atomicAdd(counts + v, 1);
}

// single CTA, 1024 threads: exclusive scans of counts -> vid_ptr and of ceil(counts/chunk) -> chunk_ptr
__global__ void __launch_bounds__(1024) pair_scan_kernel(const int* __restrict__ counts, int n_videos, int chunk,
                                                         int* __restrict__ vid_ptr, int* __restrict__ chunk_ptr,
                                                         int* __restrict__ cursor) {
  __shared__ int warp_tot[2][32];
  __shared__ int carry[2];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  for (int base = 0; base < n_videos; base += 1024) {
    const int i = base + t;
    const int c = i < n_videos ? counts[i] : 0;
    int val[2] = {c, (c + chunk - 1) / chunk};
    int inc[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int x = val[s];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      inc[s] = x;
      if (lane == 31) warp_tot[s][w] = x;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        int x = warp_tot[s][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        warp_tot[s][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
    const int off0 = carry[0] + (w ? warp_tot[0][w - 1] : 0) + inc[0] - val[0];
    const int off1 = carry[1] + (w ? warp_tot[1][w - 1] : 0) + inc[1] - val[1];
    if (i < n_videos) {
      vid_ptr[i] = off0;
      cursor[i] = off0;
      chunk_ptr[i] = off1;
    }
    __syncthreads();
    if (t == 0) {
      carry[0] += warp_tot[0][31];
      carry[1] += warp_tot[1][31];
    }
    __syncthreads();
  }
  if (t == 0) {
    vid_ptr[n_videos] = carry[0];
    chunk_ptr[n_videos] = carry[1];
  }
}

__global__ void pair_fill_kernel(const int* __restrict__ top_idx, const unsigned char* __restrict__ slot_valid,
                                 long long n_pairs, int n_slots, int vid_lo, int n_videos, int* __restrict__ cursor,
                                 int* __restrict__ entry_q, int* __restrict__ entry_out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  if (slot_valid && !slot_valid[i]) return;
  const int v = top_idx[i] - vid_lo;
  if (v < 0 || v >= n_videos) return;
  const int pos = atomicAdd(cursor + v, 1);
  entry_q[pos] = (int)(i / n_slots);
  entry_out[pos] = (int)i;
}

extern "C" int xmlb_build_pair_lists(const int* top_idx, const unsigned char* slot_valid, int n_queries, int n_slots,
                                     int vid_lo, int n_videos, int chunk, int* counts_ws, int* cursor_ws,
                                     int* vid_ptr, int* chunk_ptr, int* entry_q, int* entry_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XMLB_REQUIRE(top_idx && counts_ws && cursor_ws && vid_ptr && chunk_ptr && entry_q && entry_out,
               "xmlb_build_pair_lists: null pointer");
  XMLB_REQUIRE(n_videos > 0 && n_slots > 0 && chunk > 0, "xmlb_build_pair_lists: bad shape");
  const long long n_pairs = (long long)n_queries * n_slots;
  XMLB_REQUIRE(n_pairs < (1ll << 31), "xmlb_build_pair_lists: too many pairs");
  XMLB_CUDA(cudaMemsetAsync(counts_ws, 0, sizeof(int) * (size_t)n_videos, stream));
  if (n_pairs > 0) {
    pair_count_kernel<<<ceil_div(n_pairs, 256), 256, 0, stream>>>(top_idx, slot_valid, n_pairs, vid_lo, n_videos,
                                                                 counts_ws);
    XMLB_LAUNCH_CHECK();
  }
  pair_scan_kernel<<<1, 1024, 0, stream>>>(counts_ws, n_videos, chunk, vid_ptr, chunk_ptr, cursor_ws);
  XMLB_LAUNCH_CHECK();
  if (n_pairs > 0) {
    pair_fill_kernel<<<ceil_div(n_pairs, 256), 256, 0, stream>>>(top_idx, slot_valid, n_pairs, n_slots, vid_lo,
                                                                n_videos, cursor_ws, entry_q, entry_out);
    XMLB_LAUNCH_CHECK();
  }
  xmlb_count_launch(3);
  return XMLB_OK;
}
