// Similarity curves of the selected (query, video) pairs on the tcgen05 tensor cores + ConvSE + mask + softmax.
//   replaces XML.get_merged_st_ed_prob(cross=True) on gathered pairs and the driver's softmax
//   (reference model_xml.py:459-471,496-497; inference.py:321-322,365-367) for the merged two-stream model.
//
// Grouped GEMM, one unit = one video x up to BLOCK_N of the queries that selected it (inverted lists):
//   M = the video's clips (128 TMEM lanes),  N = queries of the list chunk,  K = 2 * Kpad: the video and subtitle
//   streams are concatenated along K so that  q'_v . f2_v + q'_s . f2_s  is produced by the accumulator; sim = acc / 2.
// A = feat2 (video || sub) of the corpus as 16-bit hi/lo halves, B = the projected queries gathered in list order
// (so a list chunk is a contiguous row range for TMA).  Split precision: 3 MMAs per product, fp32 accumulate.
// Epilogue: each thread owns one clip row of the accumulator, writes it transposed into shared memory; then each
// warp takes whole queries: 5-tap ConvSE start/end detectors, mask_logits, softmax over the clips, coalesced store.
#include "tc_pipeline.cuh"
#include "xmlb200.h"

namespace {

using tc::BLOCK_K;
using tc::BLOCK_M;
constexpr int MAX_L = 256;  // clips per video: up to two 128-row accumulator halves per (video, query chunk)

constexpr int N_BOXES = 8;  // A boxes of 16, 32, ..., 128 clip rows
struct SpanTcMaps {
  CUtensorMap a_hi[N_BOXES], a_lo[N_BOXES], b_hi, b_lo;
};

struct SpanTcParams {
  int n_videos, ctx_len, k_blocks, block_n, stages, ksize, softmax;
  int n_halves;          // 128-clip halves per video (2 when ctx_len > 128)
  const int4* units;     // {video, first entry row, entries in this chunk, clip rows to load (clip_boxes) or 0}
  int a_kb_rows;         // > 0: f2 is stored k-blocked, [kcat / 32][a_kb_rows = n_videos * ctx_len][32]
  const unsigned short* f2_hi_bulk;  // != null: ... as the shared-memory image (pieces pre-swizzled): plain bulk copies
  const unsigned short* f2_lo_bulk;
  int clip_boxes;        // 1: units[].w = number of leading clip rows of the video the epilogue can need (every
                         // unmasked clip and its ConvSE neighbours); only those are loaded, in 16-row steps
  const int* n_units;    // device scalar
  const int* entry_out;  // [E] output row of each list entry
  const int* entry_q;    // [E] query of each list entry: the B rows are gathered in the kernel (null: pre-gathered)
  int gather_warps;      // with entry_q: 1 = gather warps (ld.global -> swizzled st.shared), 0 = TMA gather4 producer
  const unsigned short* q_hi;  // gather warps: the un-gathered (n_queries, kcat) halves
  const unsigned short* q_lo;
  int kcat;
  const float* mask;     // [Nv][L]
  const float* w_st;
  const float* w_ed;
  float* out_st;
  float* out_ed;
  int* unit_counter;  // zeroed before the launch
  unsigned int idesc;
};

struct SpanSched {
  const SpanTcMaps* maps;
  const SpanTcParams* p;
  int n_units, u, half;
  int4 m;
  __device__ SpanSched(const SpanTcMaps* mp, const SpanTcParams* pp)
      : maps(mp), p(pp), n_units(__ldg(pp->n_units)), u(0), half(0) {}
  __device__ bool next(tc::UnitDesc& d) {
    if (half == 0) {
      u = atomicAdd(p->unit_counter, 1);
      if (u >= n_units) return false;
      m = __ldg(p->units + u);
    }
    int box = N_BOXES - 1;
    if (p->clip_boxes) box = min(N_BOXES - 1, max(0, (m.w - half * BLOCK_M + 15) / 16 - 1));
    d.a_hi = &maps->a_hi[box], d.a_lo = &maps->a_lo[box], d.b_hi = &maps->b_hi, d.b_lo = &maps->b_lo;
    d.a_bytes = (box + 1) * 16 * tc::SWIZZLE_BYTES;
    d.a_kb_rows = p->a_kb_rows;
    d.a_hi_bulk = p->f2_hi_bulk, d.a_lo_bulk = p->f2_lo_bulk;
    if (p->f2_hi_bulk)  // a plain copy has no out-of-bounds fill: never past the video's own rows
      d.a_bytes = min((box + 1) * 16, p->ctx_len - half * BLOCK_M) * tc::SWIZZLE_BYTES;
    d.a_row = m.x * p->ctx_len + half * BLOCK_M;
    d.b_row = m.y;
    d.g_count = m.z;
    d.k_blocks = p->k_blocks;
    d.idesc = p->idesc;
    d.tag0 = u, d.tag1 = half;
    if (++half == p->n_halves) half = 0;
    return true;
  }
};

constexpr int EPI_WARPS = 8;  // two per TMEM lane quadrant
constexpr int S_OFF = 2;      // zero floats on both sides of a similarity row (the reach of the common 5-tap ConvSE)

// NC = 32-clip groups per video handled by a lane in the ConvSE phase: 4 (ctx_len <= 128, one accumulator half) or 8.
// KS = 5: the usual 5-tap detectors, taps in registers, neighbours read from the zero-margined row without bounds
// checks; KS = 0: any odd ksize <= 31.
// The epilogue used to be the kernel's bottleneck (ncu source view: its four warps, one per scheduler with nothing to
// hide a latency behind, never waited for an accumulator: ~5000 cycles per query, chains of L1 loads of taps, mask and
// output row inside the tap loop).  Now: eight warps, everything that does not depend on the query (mask values,
// output rows of the list chunk, taps) is fetched before the accumulator wait, and two queries are processed per
// iteration as one straight-line block.  Same operations in the same order per output: bit-identical results.
template <int NC, int KS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS + 32 * tc::GATHER_WARPS, 1)
span_probs_tc_kernel(const __grid_constant__ SpanTcMaps maps, const __grid_constant__ SpanTcParams p) {
  constexpr int S_LD = 32 * NC + 2 * S_OFF;  // row pitch (floats) of the transposed similarity tile
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const bool gw = p.entry_q && p.gather_warps;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, p.block_n, 3, EPI_WARPS, gw ? 2 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (p.entry_q && !gw) {
      tc::tc_producer_loop_gather(SpanSched(&maps, &p), pipe, lane, p.entry_q, 0, p.block_n);
    } else if (lane == 0) {
      tc::tc_producer_loop(SpanSched(&maps, &p), pipe);
    }
  } else if (warp == 1) {
    tc::tc_mma_loop_warp(pipe, tmem_base);
  } else if (warp >= 2 + EPI_WARPS) {  // ============ gather warps: the listed queries -> B tile ============
    tc::tc_gather_loop(pipe, threadIdx.x - 32 * (2 + EPI_WARPS), p.entry_q, p.block_n, p.kcat,
                       [&](int u, int, int& e0, int& ne, const unsigned short*& hi, const unsigned short*& lo) {
                         const int4 m = __ldg(p.units + u);
                         e0 = m.y, ne = m.z, hi = p.q_hi, lo = p.q_lo;
                       });
  } else {  // ===================== epilogue warps 2 .. 9 =====================
    float* S = reinterpret_cast<float*>(smem_raw + (pipe.extra() - tc::smem_u32(smem_raw)));  // [block_n][S_LD]
    float* W = S + p.block_n * S_LD;                 // [2][32] taps: start | end
    const int row_in_half = (warp & 3) * 32 + lane;  // accumulator row (TMEM lane) owned in phase A
    const int ew = warp - 2;                         // 0..7: queries handled in the ConvSE phase
    const int col_half = ew >> 2;                    // the two warps of a lane quadrant alternate 32-column groups
    const int L = p.ctx_len, pad = p.ksize / 2;
    {  // once: zero margins of every row, taps
      const int et = threadIdx.x - 64;
      for (int r = et; r < p.block_n; r += 32 * EPI_WARPS) {
#pragma unroll
        for (int i = 0; i < S_OFF; ++i) S[r * S_LD + i] = 0.f, S[r * S_LD + S_LD - 1 - i] = 0.f;
      }
      if (et < 64) W[et] = (et & 31) < p.ksize ? __ldg((et < 32 ? p.w_st : p.w_ed) + (et & 31)) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    }
    float wst[5], wed[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) wst[t] = W[t], wed[t] = W[32 + t];
    int u, half;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, u, half); ++unit) {
      const int4 m = __ldg(p.units + u);
      const int v = m.x, e0 = m.y, ne = m.z;
      // query-independent inputs of phase B, in flight while the accumulator is still being produced
      float mk[NC];
      int orow[4];
#pragma unroll
      for (int c = 0; c < NC; ++c) mk[c] = lane + 32 * c < L ? __ldg(p.mask + (long long)v * L + lane + 32 * c) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) orow[k] = lane + 32 * k < ne ? __ldg(p.entry_out + e0 + lane + 32 * k) : 0;
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base);
      // ---- phase A: accumulator (clip x query) -> shared memory, transposed to (query x clip), halved; clips
      // beyond ctx_len read as zero (the ConvSE zero padding)
      const int clip = (NC > 4 ? half * BLOCK_M : 0) + row_in_half;
      for (int c = col_half; c * 32 < ne; c += 2) {  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + c * 32, r);
        tc::tmem_ld_wait();
        if (clip < 32 * NC) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            S[(c * 32 + i) * S_LD + S_OFF + clip] = clip < L ? __fmul_rn(__uint_as_float(r[i]), 0.5f) : 0.f;
        }
      }
      tc::epi_release(pipe, unit);
      if (NC > 4 && half != p.n_halves - 1) continue;  // the video's second half of clips follows as the next unit
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      // ---- phase B: per query: ConvSE start / end, mask_logits, softmax over clips; two queries per iteration
      for (int j0 = ew; j0 < ne; j0 += 2 * EPI_WARPS) {
        const bool has2 = j0 + EPI_WARPS < ne;  // warp-uniform
        const int jq[2] = {j0, has2 ? j0 + EPI_WARPS : j0};
        float st[2][NC], ed[2][NC];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float* s = S + jq[q] * S_LD + S_OFF;  // s[l] = similarity of clip l
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int l = lane + 32 * c;
            float a = 0.f, b = 0.f;
            if (KS == 5) {
#pragma unroll
              for (int t = 0; t < 5; ++t) {
                const float x = s[l + t - 2];
                a = fmaf(wst[t], x, a), b = fmaf(wed[t], x, b);
              }
            } else if (l < L) {
              for (int t = 0; t < p.ksize; ++t) {
                const int src = l + t - pad;
                const float x = (src >= 0 && src < L) ? s[src] : 0.f;
                a = fmaf(W[t], x, a), b = fmaf(W[32 + t], x, b);
              }
            }
            // a masked clip is -1e10 whatever its (finite) logit; its neighbourhood may not even have been loaded
            // (clip_boxes), so whatever was accumulated from it is dropped
            a = mk[c] != 0.f ? a : 0.f, b = mk[c] != 0.f ? b : 0.f;
            st[q][c] = mask_logit(a, mk[c]), ed[q][c] = mask_logit(b, mk[c]);
          }
        }
        if (p.softmax) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float mx_st = -INFINITY, mx_ed = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (lane + 32 * c < L) mx_st = fmaxf(mx_st, st[q][c]), mx_ed = fmaxf(mx_ed, ed[q][c]);
            mx_st = warp_max(mx_st), mx_ed = warp_max(mx_ed);
            float s_st = 0.f, s_ed = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              if (lane + 32 * c < L) {
                st[q][c] = expf(st[q][c] - mx_st), ed[q][c] = expf(ed[q][c] - mx_ed);
                s_st += st[q][c], s_ed += ed[q][c];
              }
            }
            s_st = warp_sum(s_st), s_ed = warp_sum(s_ed);
#pragma unroll
            for (int c = 0; c < NC; ++c) st[q][c] = __fdiv_rn(st[q][c], s_st), ed[q][c] = __fdiv_rn(ed[q][c], s_ed);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          int row = 0;  // output row of query jq[q]: held by lane jq & 31 in orow[jq >> 5]
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = __shfl_sync(0xffffffffu, orow[k], jq[q] & 31);
            if ((jq[q] >> 5) == k) row = r;
          }
          if (q == 0 || has2) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const int l = lane + 32 * c;
              if (l < L) {
                p.out_st[(long long)row * L + l] = st[q][c];
                p.out_ed[(long long)row * L + l] = ed[q][c];
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // S may be overwritten by the next unit
    }
  }
  tc::pipe_teardown(tmem_base);
}

// units[chunk_ptr[v] + c] = {v, vid_ptr[v] + c * chunk, min(chunk, remaining), video_rows ? video_rows[v] : 0}
__global__ void span_units_kernel(const int* __restrict__ vid_ptr, const int* __restrict__ chunk_ptr, int n_videos,
                                  int chunk, const int* __restrict__ video_rows, int4* __restrict__ units) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_videos) return;
  const int e_lo = vid_ptr[v], e_hi = vid_ptr[v + 1];
  const int rows = video_rows ? video_rows[v] : 0;
  int u = chunk_ptr[v];
  for (int e = e_lo; e < e_hi; e += chunk, ++u) units[u] = make_int4(v, e, min(chunk, e_hi - e), rows);
}

}  // namespace

extern "C" int xmlb_build_span_units_rows(const int* vid_ptr, const int* chunk_ptr, int n_videos, int chunk,
                                          const int* video_rows, int* units, void* stream) {
  XMLB_REQUIRE(vid_ptr && chunk_ptr && units && n_videos > 0 && chunk > 0, "xmlb_build_span_units: bad argument");
  XMLB_REQUIRE(((uintptr_t)units & 15) == 0, "xmlb_build_span_units: units must be 16-byte aligned");
  span_units_kernel<<<ceil_div(n_videos, 256), 256, 0, (cudaStream_t)stream>>>(vid_ptr, chunk_ptr, n_videos, chunk,
                                                                              video_rows, reinterpret_cast<int4*>(units));
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_build_span_units(const int* vid_ptr, const int* chunk_ptr, int n_videos, int chunk, int* units,
                                     void* stream) {
  return xmlb_build_span_units_rows(vid_ptr, chunk_ptr, n_videos, chunk, nullptr, units, stream);
}

extern "C" int xmlb_span_probs_tc_clipped(const unsigned short* f2_hi, const unsigned short* f2_lo,
                                          const unsigned short* qg_hi, const unsigned short* qg_lo, const float* mask,
                                          const float* w_st, const float* w_ed, int ksize, int apply_softmax,
                                          int n_videos, int ctx_len, int kcat, long long n_entries, int block_n,
                                          const int* units, const int* n_units, int max_units, const int* entry_out,
                                          const int* entry_q, int gather_warps, long long n_query_rows, int clip_boxes,
                                          int f2_kblocked, float* out_st, float* out_ed, int* sched_ws, int is_bf16,
                                          void* stream) {
  XMLB_REQUIRE(f2_hi && f2_lo && qg_hi && qg_lo && mask && w_st && w_ed && units && n_units && entry_out && out_st &&
                   out_ed && sched_ws, "xmlb_span_probs_tc: null pointer");
  XMLB_REQUIRE(ctx_len >= 1 && ctx_len <= MAX_L, "xmlb_span_probs_tc: ctx_len must be <= 256");
  XMLB_REQUIRE(kcat >= 64 && kcat % 64 == 0, "xmlb_span_probs_tc: kcat must be a multiple of 64");
  XMLB_REQUIRE(block_n == 32 || block_n == 64 || block_n == 128, "xmlb_span_probs_tc: block_n must be 32, 64 or 128");
  XMLB_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 31, "xmlb_span_probs_tc: ksize must be odd and <= 31");
  XMLB_REQUIRE(((uintptr_t)units & 15) == 0, "xmlb_span_probs_tc: units must be 16-byte aligned");
  XMLB_REQUIRE(!entry_q || n_query_rows > 0, "xmlb_span_probs_tc: n_query_rows is required with entry_q");
  if (n_entries == 0 || max_units == 0 || n_videos == 0) return XMLB_OK;
  SpanTcParams p = {};
  p.n_videos = n_videos, p.ctx_len = ctx_len, p.k_blocks = kcat / BLOCK_K, p.block_n = block_n;
  p.ksize = ksize, p.softmax = apply_softmax;
  p.n_halves = ceil_div(ctx_len, BLOCK_M);
  const int s_ld = (ctx_len <= 128 ? 128 : 256) + 2 * S_OFF;
  p.units = reinterpret_cast<const int4*>(units), p.n_units = n_units, p.entry_out = entry_out, p.entry_q = entry_q;
  p.gather_warps = entry_q && gather_warps ? 1 : 0, p.q_hi = qg_hi, p.q_lo = qg_lo, p.kcat = kcat;
  XMLB_REQUIRE(!clip_boxes || p.gather_warps, "xmlb_span_probs_tc: clip_boxes needs the gather-warps mode");
  XMLB_REQUIRE(!f2_kblocked || (long long)n_videos * ctx_len * (kcat / BLOCK_K) < (1ll << 31),
               "xmlb_span_probs_tc: k-blocked f2 too large for 32-bit TMA coordinates");
  p.clip_boxes = clip_boxes ? 1 : 0;
  p.a_kb_rows = f2_kblocked ? n_videos * ctx_len : 0;
  if (f2_kblocked == 2) {
    XMLB_REQUIRE(p.gather_warps && ctx_len % 8 == 0,
                 "xmlb_span_probs_tc: the pre-swizzled f2 layout needs the gather-warps mode and ctx_len %% 8 == 0");
    p.f2_hi_bulk = f2_hi, p.f2_lo_bulk = f2_lo;
  }
  p.mask = mask, p.w_st = w_st, p.w_ed = w_ed, p.out_st = out_st, p.out_ed = out_ed, p.unit_counter = sched_ws;
  p.idesc = tc::idesc_f16(BLOCK_M, block_n, is_bf16 ? 1 : 0);
  const int extra = block_n * s_ld * (int)sizeof(float) + 64 * (int)sizeof(float);  // similarity tile + taps
  p.stages = tc::pipe_stages(block_n, extra);
  XMLB_REQUIRE(p.stages >= 2, "xmlb_span_probs_tc: tile does not fit in shared memory");
  const size_t smem = tc::pipe_smem_bytes(block_n, p.stages, extra);

  SpanTcMaps maps;
  int rc;
  const unsigned long long corpus_rows = (unsigned long long)n_videos * ctx_len;
  // gather mode: the B maps address single rows of the (n_query_rows, kcat) query arrays (TMA gather4)
  const unsigned long long b_rows = entry_q ? (unsigned long long)n_query_rows : (unsigned long long)n_entries;
  const unsigned int b_box = entry_q ? 1u : (unsigned int)block_n;
  for (int b = clip_boxes ? 0 : N_BOXES - 1; b < N_BOXES; ++b) {  // box b: the first 16 (b + 1) clip rows of a tile
    const unsigned long long a_rows = f2_kblocked ? corpus_rows * (kcat / BLOCK_K) : corpus_rows;
    const unsigned long long a_cols = f2_kblocked ? BLOCK_K : kcat;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_hi[b], f2_hi, a_rows, a_cols, 16 * (b + 1), BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_lo[b], f2_lo, a_rows, a_cols, 16 * (b + 1), BLOCK_K))) return rc;
  }
  if ((rc = xmlb_make_tmap_2d_u16(&maps.b_hi, qg_hi, b_rows, kcat, b_box, BLOCK_K))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.b_lo, qg_lo, b_rows, kcat, b_box, BLOCK_K))) return rc;

  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = max_units < sms ? max_units : sms;
  XMLB_CUDA(cudaMemsetAsync(sched_ws, 0, sizeof(int), (cudaStream_t)stream));
  const int threads = 64 + 32 * EPI_WARPS + (p.gather_warps ? 32 * tc::GATHER_WARPS : 0);
  auto launch = [&](auto kernel) -> int {
    XMLB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(maps, p);
    return XMLB_OK;
  };
  if (p.n_halves == 1) {
    if ((rc = ksize == 5 ? launch(span_probs_tc_kernel<4, 5>) : launch(span_probs_tc_kernel<4, 0>))) return rc;
  } else {
    if ((rc = ksize == 5 ? launch(span_probs_tc_kernel<8, 5>) : launch(span_probs_tc_kernel<8, 0>))) return rc;
  }
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_span_probs_tc_ex(const unsigned short* f2_hi, const unsigned short* f2_lo,
                                     const unsigned short* qg_hi, const unsigned short* qg_lo, const float* mask,
                                     const float* w_st, const float* w_ed, int ksize, int apply_softmax, int n_videos,
                                     int ctx_len, int kcat, long long n_entries, int block_n, const int* units,
                                     const int* n_units, int max_units, const int* entry_out, const int* entry_q,
                                     int gather_warps, long long n_query_rows, float* out_st, float* out_ed,
                                     int* sched_ws, int is_bf16, void* stream) {
  return xmlb_span_probs_tc_clipped(f2_hi, f2_lo, qg_hi, qg_lo, mask, w_st, w_ed, ksize, apply_softmax, n_videos,
                                    ctx_len, kcat, n_entries, block_n, units, n_units, max_units, entry_out, entry_q,
                                    gather_warps, n_query_rows, 0, 0, out_st, out_ed, sched_ws, is_bf16, stream);
}

extern "C" int xmlb_span_probs_tc(const unsigned short* f2_hi, const unsigned short* f2_lo,
                                  const unsigned short* qg_hi, const unsigned short* qg_lo, const float* mask,
                                  const float* w_st, const float* w_ed, int ksize, int apply_softmax, int n_videos,
                                  int ctx_len, int kcat, long long n_entries, int block_n, const int* units,
                                  const int* n_units, int max_units, const int* entry_out, float* out_st,
                                  float* out_ed, int* sched_ws, int is_bf16, void* stream) {
  return xmlb_span_probs_tc_ex(f2_hi, f2_lo, qg_hi, qg_lo, mask, w_st, w_ed, ksize, apply_softmax, n_videos, ctx_len,
                               kcat, n_entries, block_n, units, n_units, max_units, entry_out, nullptr, 0, 0, out_st,
                               out_ed, sched_ws, is_bf16, stream);
}
