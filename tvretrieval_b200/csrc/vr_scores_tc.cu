// Video-level retrieval scores on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), split-precision.
//   replaces XML.get_video_level_scores x modalities + their average (reference model_xml.py:446-452, 572-574):
//   q2c[q][v] = mean_mod  max_{l valid}  qn_mod[q] . c1n_mod[v][l]
//
// GEMM view: M = queries, N = corpus clips (a tile holds whole videos), K = hidden.  fp32-accurate products are
// obtained from 16-bit tensor-core MMAs by splitting both operands x = hi + lo (hi = rn16(x), lo = rn16(x - hi))
// and accumulating  A_hi B_hi + A_hi B_lo + A_lo B_hi  in the fp32 TMEM accumulator (3 MMAs per k-step; the
// dropped lo*lo term is ~2^-16 (bf16) / 2^-22 (fp16) relative).  SURVEY.md section 7 "hard part 1" shows that
// single bf16/tf32 products break the rank-exactness contract while this scheme keeps it.
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: 4 tile loads per k-block (A_hi, A_lo, B_hi, B_lo; 128B swizzle) into a smem ring
//   warp 1      MMA issuer (one elected thread): 12 tcgen05.mma per k-block into one of two TMEM accumulators
//   warps 2..5  epilogue: tcgen05.ld the 128 x BLOCK_N fp32 accumulator, masked max over each video's clips
//               (one thread = one query row, no shuffles), combine the two modalities, store q2c
// The two modalities of a tile alternate between the two TMEM accumulators, so the epilogue of one overlaps the
// MMAs of the other.
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_pipeline.cuh"
#include "vr_common.cuh"
#include "xmlb200.h"

namespace {

using tc::BLOCK_K;
using tc::BLOCK_M;
constexpr int MAX_VPT = 8;  // videos per tile (BLOCK_N = vpt * lp <= 256, lp multiple of 32)
using vr::MAX_TILE_VIDEOS;
using vr::PACKED_EXTRA_SMEM;

struct VrMaps {
  CUtensorMap a_hi[2], a_lo[2], b_hi[2], b_lo[2];
};

struct VrTcParams {
  int n_queries, n_videos, lp, vpt, block_n, k_blocks, n_mod, m_tiles, n_tiles, stages;
  const unsigned int* mask_bits[2];
  float* out;
  int* tile_counter;  // zeroed before the launch
  float divisor;
  unsigned int idesc;
};

// tiles are claimed from a global counter (query tile fastest), each tile = n_mod consecutive units
struct VrSched {
  const VrMaps* maps;
  const VrTcParams* p;
  int tile, mod;
  __device__ VrSched(const VrMaps* m, const VrTcParams* pp) : maps(m), p(pp), tile(0), mod(0) {}
  __device__ bool next(tc::UnitDesc& u) {
    if (mod == 0) {
      tile = atomicAdd(p->tile_counter, 1);
      if (tile >= p->m_tiles * p->n_tiles) return false;
    }
    u.a_hi = &maps->a_hi[mod], u.a_lo = &maps->a_lo[mod], u.b_hi = &maps->b_hi[mod], u.b_lo = &maps->b_lo[mod];
    u.a_row = (tile % p->m_tiles) * BLOCK_M;
    u.b_row = (tile / p->m_tiles) * p->block_n;
    u.k_blocks = p->k_blocks;
    u.idesc = p->idesc;
    u.tag0 = tile, u.tag1 = mod;
    if (++mod == p->n_mod) mod = 0;
    return true;
  }
};

__global__ void __launch_bounds__(192, 1)
vr_scores_tc_kernel(const __grid_constant__ VrMaps maps, const __grid_constant__ VrTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, p.block_n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) tc::tc_producer_loop(VrSched(&maps, &p), pipe);
  } else if (warp == 1) {
    tc::tc_mma_loop_warp(pipe, tmem_base);
  } else {  // ===================== epilogue warps 2..5 =====================
    const int row = (warp & 3) * 32 + lane;
    const int chunks = p.lp >> 5;
    float best[MAX_VPT];
    int tile, mod;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, tile, mod); ++unit) {
      const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base);
      const unsigned int* __restrict__ bits = p.mask_bits[mod];
#pragma unroll
      for (int j = 0; j < MAX_VPT; ++j) {
        if (j < p.vpt) {
          const int v = n_tile * p.vpt + j;
          float m = MASK_FILL;
          for (int c = 0; c < chunks; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(taddr + j * p.lp + c * 32, r);
            tc::tmem_ld_wait();
            const unsigned int b = v < p.n_videos ? __ldg(bits + (long long)v * chunks + c) : 0u;
            if (b == 0xffffffffu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) m = (b >> i) & 1u ? fmaxf(m, __uint_as_float(r[i])) : m;
            }
          }
          best[j] = mod == 0 ? m : __fadd_rn(best[j], m);
        }
      }
      tc::epi_release(pipe, unit);  // accumulator fully read: hand it back to the MMA issuer
      const int q = m_tile * BLOCK_M + row;
      if (mod == p.n_mod - 1 && q < p.n_queries) {
#pragma unroll
        for (int j = 0; j < MAX_VPT; ++j) {
          const int v = n_tile * p.vpt + j;
          if (j < p.vpt && v < p.n_videos) p.out[(long long)q * p.n_videos + v] = __fdiv_rn(best[j], p.divisor);
        }
      }
    }
  }
  tc::pipe_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// Packed ("ragged") variant: only VALID clips are stored, whole videos are packed greedily into tiles of <= 256
// consecutive rows (longest videos first), so no MMA work is spent on padded clips (44 % of the padded corpus at
// the TVR length distribution).  Per tile: first packed row, first packed-video ordinal, number of used columns,
// and a 256-bit map of the columns where a new video starts.  Both modalities share the packing.
struct VrPackedParams {
  int n_queries, n_videos, n_tiles, m_tiles, k_blocks, n_mod, stages, terms;
  const int* tile_meta;          // [n_tiles][4]: row_start, first ordinal, used columns, number of videos
  const unsigned int* tile_starts;  // [n_tiles][8]
  float* out;
  int* tile_counter;  // zeroed before the launch
  // restricted mode (fallback of the two-pass search): only the query tiles m_tile_list[0 .. *n_m_tiles) are computed
  const int* m_tile_list;
  const int* n_m_tiles;
  float divisor;
  unsigned int idesc;
  int probe;  // XMLB_VR_PROBE (limiter experiments): bit 0 = epilogue only waits and releases, bit 1 = no B loads
};

struct VrPackedSched {
  const VrMaps* maps;
  const VrPackedParams* p;
  int m_tiles, m_tile, n_tile, mod;
  __device__ VrPackedSched(const VrMaps* m, const VrPackedParams* pp)
      : maps(m), p(pp), m_tiles(pp->m_tile_list ? __ldg(pp->n_m_tiles) : pp->m_tiles), m_tile(0), n_tile(0), mod(0) {}
  __device__ bool next(tc::UnitDesc& u) {
    if (mod == 0) {
      if (m_tiles <= 0) return false;
      const int tile = atomicAdd(p->tile_counter, 1);
      if (tile >= m_tiles * p->n_tiles) return false;
      m_tile = tile % m_tiles, n_tile = tile / m_tiles;  // query tile fastest: the corpus tile is shared through L2
      if (p->m_tile_list) m_tile = __ldg(p->m_tile_list + m_tile);
    }
    u.a_hi = &maps->a_hi[mod], u.a_lo = &maps->a_lo[mod], u.b_hi = &maps->b_hi[mod], u.b_lo = &maps->b_lo[mod];
    u.a_row = m_tile * BLOCK_M;
    u.b_row = __ldg(p->tile_meta + 4 * n_tile);
    u.k_blocks = p->k_blocks;
    u.idesc = p->idesc;
    u.tag0 = n_tile, u.tag1 = m_tile * 2 + mod;
    if (++mod == p->n_mod) mod = 0;
    return true;
  }
};

__global__ void __launch_bounds__(192, 1)
vr_scores_tc_packed_kernel(const __grid_constant__ VrMaps maps, const __grid_constant__ VrPackedParams p) {
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, 256, p.terms);
  pipe.probe = p.probe;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) tc::tc_producer_loop(VrPackedSched(&maps, &p), pipe);
  } else if (warp == 1) {
    tc::tc_mma_loop_warp(pipe, tmem_base);
  } else {  // ===================== epilogue warps 2..5 =====================
    // per query row, the first modality's maxima are parked in shared memory (nothing on the accumulator-release
    // path waits on global memory).  Scores are written in PACKED ORDINAL order: the videos of a tile are adjacent
    // columns, so the 4-byte stores of neighbouring tiles complete whole sectors (writing by original video id
    // scattered them: 7x DRAM write amplification and L2 pollution).
    const uint32_t best_s = pipe.extra();  // float [128][MAX_TILE_VIDEOS + 1]
    const int row = (warp & 3) * 32 + lane;
    const uint32_t my_best = best_s + (uint32_t)row * (MAX_TILE_VIDEOS + 1) * 4;
    int n_tile, tag1;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, n_tile, tag1); ++unit) {
      const int m_tile = tag1 >> 1, mod = tag1 & 1;
      const int q = m_tile * BLOCK_M + row;
      const bool q_ok = q < p.n_queries;
      const int4 meta = __ldg(reinterpret_cast<const int4*>(p.tile_meta) + n_tile);  // row, first ordinal, used, n videos
      float* __restrict__ out_row = p.out + (long long)q * p.n_videos + meta.y;
      const unsigned int* __restrict__ starts = p.tile_starts + 8 * n_tile;
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base);
      if (p.probe & 1) {
        tc::epi_release(pipe, unit);
        continue;
      }
      vr::packed_epilogue(taddr, meta.z, starts, out_row, q_ok, mod, p.n_mod, p.divisor, my_best,
                          [&]() { tc::epi_release(pipe, unit); });
    }
  }
  tc::pipe_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// Exact (split-precision) re-scoring of the candidate (query, video) pairs of the two-pass search.  Grouped GEMM
// over inverted lists: one unit = one packed video x up to 128 of the queries that kept it as a candidate, per
// modality.  A = the queries gathered in list order (xmlb_split_rows with row_index = entry_q), B = the video's
// clips in the packed corpus (the tile is loaded with a fixed box of block_n rows; columns beyond the video's
// length belong to the next videos and are ignored).  Same arithmetic per product as the one-pass kernel.
constexpr int RESCORE_BOXES = 4;  // B boxes of 1/4, 2/4, 3/4, 4/4 of block_n rows
struct VrRescoreMaps {
  CUtensorMap a_hi[2], a_lo[2], b_hi[2][RESCORE_BOXES], b_lo[2][RESCORE_BOXES];
};

struct VrRescoreParams {
  int n_mod, k_blocks, stages, block_n;
  int c_kb_rows;         // > 0: the corpus halves are stored k-blocked, [kpad / 32][c_kb_rows = n_packed_rows][32]
  int clip_boxes;        // 1: only the quarter-tiles of the B box that hold the video's clips are loaded
  const int4* units;     // {packed ordinal, first list entry, entries in this chunk, 0}
  const int* n_units;    // device scalar
  const int* entry_out;  // [E] slot (row * max_cand + position) of each list entry
  const int* entry_q;    // [E] query of each list entry: the A rows are gathered in the kernel (null: pre-gathered)
  int gather_warps;      // with entry_q: 1 = gather warps (ld.global -> swizzled st.shared), 0 = TMA gather4 producer
  int is_bf16;
  const unsigned short* q_hi[2];  // gather warps: the un-gathered (n_queries, kpad) halves per modality
  const unsigned short* q_lo[2];
  int kpad;
  const int* row_start;  // [n_packed + 1] first packed row of each ordinal
  float* out;            // candidate scores, indexed by entry_out
  int* unit_counter;     // zeroed before the launch
  float divisor;
  unsigned int idesc;
};

struct VrRescoreSched {
  const VrRescoreMaps* maps;
  const VrRescoreParams* p;
  int n_units, u, mod;
  int4 m;
  __device__ VrRescoreSched(const VrRescoreMaps* mp, const VrRescoreParams* pp)
      : maps(mp), p(pp), n_units(__ldg(pp->n_units)), u(0), mod(0) {}
  __device__ bool next(tc::UnitDesc& d) {
    if (mod == 0) {
      u = atomicAdd(p->unit_counter, 1);
      if (u >= n_units) return false;
      m = __ldg(p->units + u);
    }
    d.a_row = m.y;
    d.g_count = m.z;
    d.b_row = __ldg(p->row_start + m.x);
    d.k_blocks = p->k_blocks;
    // N of this unit's MMAs = the video's clips rounded up to 16; the B box holds the quarters of block_n rows that
    // cover them (all block_n rows without clip_boxes; the columns beyond the video are simply not computed)
    const int len = __ldg(p->row_start + m.x + 1) - d.b_row;
    const int quarter = p->block_n / RESCORE_BOXES;
    const int box = p->clip_boxes ? min(RESCORE_BOXES - 1, max(0, (len - 1) / quarter)) : RESCORE_BOXES - 1;
    d.a_hi = &maps->a_hi[mod], d.a_lo = &maps->a_lo[mod], d.b_hi = &maps->b_hi[mod][box], d.b_lo = &maps->b_lo[mod][box];
    d.b_bytes = (box + 1) * quarter * tc::SWIZZLE_BYTES;
    d.b_kb_rows = p->c_kb_rows;
    d.idesc = tc::idesc_f16(BLOCK_M, min(p->block_n, (len + 15) & ~15), p->is_bf16);
    d.tag0 = u, d.tag1 = mod;
    if (++mod == p->n_mod) mod = 0;
    return true;
  }
};

__global__ void __launch_bounds__(192 + 32 * tc::GATHER_WARPS, 1)
vr_rescore_tc_kernel(const __grid_constant__ VrRescoreMaps maps, const __grid_constant__ VrRescoreParams p) {
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const bool gw = p.entry_q && p.gather_warps;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, p.block_n, 3, 4, gw ? 1 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (p.entry_q && !gw) {
      tc::tc_producer_loop_gather(VrRescoreSched(&maps, &p), pipe, lane, p.entry_q, 1, BLOCK_M);
    } else if (lane == 0) {
      tc::tc_producer_loop(VrRescoreSched(&maps, &p), pipe);
    }
  } else if (warp == 1) {
    tc::tc_mma_loop_warp(pipe, tmem_base);
  } else if (warp >= 6) {  // ===================== gather warps 6..9: the listed queries -> A tile =====================
    tc::tc_gather_loop(pipe, threadIdx.x - 192, p.entry_q, BLOCK_M, p.kpad,
                       [&](int u, int mod, int& e0, int& ne, const unsigned short*& hi, const unsigned short*& lo) {
                         const int4 m = __ldg(p.units + u);
                         e0 = m.y, ne = m.z, hi = p.q_hi[mod], lo = p.q_lo[mod];
                       });
  } else {  // ===================== epilogue warps 2..5: one thread = one listed query =====================
    const int row = (warp & 3) * 32 + lane;
    float first = 0.f;
    int u, mod;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, u, mod); ++unit) {
      const int4 m = __ldg(p.units + u);
      const int len = __ldg(p.row_start + m.x + 1) - __ldg(p.row_start + m.x);
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base);
      float cur = MASK_FILL;
      for (int c = 0; c * 32 < len; ++c) {  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + c * 32, r);
        tc::tmem_ld_wait();
        const int n_here = len - c * 32;
        cur = fmaxf(cur, n_here >= 32 ? vr::max32(r) : vr::masked_max32(r, (1u << n_here) - 1u));
      }
      tc::epi_release(pipe, unit);
      if (mod != p.n_mod - 1) {
        first = cur;
      } else if (row < m.z) {
        const float v = mod != 0 ? __fadd_rn(first, cur) : cur;
        p.out[__ldg(p.entry_out + m.y + row)] = __fdiv_rn(v, p.divisor);
      }
    }
  }
  tc::pipe_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------------------------
// operand preparation: fp32 rows -> (optionally L2-normalised) hi/lo 16-bit rows, K padded to a multiple of 64,
// rows regrouped from groups of `gin` to zero-padded groups of `gout` (clips of a video padded to lp).
// One warp per output row.
template <bool BF16>
__device__ __forceinline__ void split16(float x, unsigned short& hi, unsigned short& lo) {
  if (BF16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h), lo = __bfloat16_as_ushort(l);
  } else {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h), lo = __half_as_ushort(l);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x,
                                                         const int* __restrict__ row_index, long long rows_out, int k,
                                                         int kpad, int out_ld, int out_col0, int gin, int gout,
                                                         int normalize,
                                                         unsigned short* __restrict__ hi,
                                                         unsigned short* __restrict__ lo,
                                                         float* __restrict__ hi_err) {
  const int lane = threadIdx.x & 31;
  const long long ro = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (ro >= rows_out) return;
  unsigned short* h = hi + ro * out_ld + out_col0;
  unsigned short* o = lo + ro * out_ld + out_col0;
  long long src;
  if (row_index) {
    src = row_index[ro];
  } else {
    const long long g = ro / gout;
    const int l = (int)(ro - g * gout);
    src = l < gin ? g * gin + l : -1;
  }
  if (src < 0) {
    if (row_index) return;  // gather mode: rows without a source are left untouched (nothing reads them)
    for (int i = lane; i < kpad; i += 32) h[i] = 0, o[i] = 0;
    if (hi_err && lane == 0) hi_err[ro] = 0.f;
    return;
  }
  const float* xr = x + src * k;
  float denom = 1.f;
  if (normalize) {
    float s = 0.f;
    for (int i = lane; i < k; i += 32) s = fmaf(xr[i], xr[i], s);
    denom = fmaxf(sqrtf(warp_sum(s)), 1e-12f);  // F.normalize eps
  }
  float err2 = 0.f;
  for (int i = lane; i < kpad; i += 32) {
    unsigned short a = 0, b = 0;
    if (i < k) {
      const float xv = normalize ? __fdiv_rn(xr[i], denom) : xr[i];
      split16<BF16>(xv, a, b);
      const float d = xv - (BF16 ? __bfloat162float(__ushort_as_bfloat16(a)) : __half2float(__ushort_as_half(a)));
      err2 = fmaf(d, d, err2);
    }
    h[i] = a, o[i] = b;
  }
  if (hi_err) {  // ||x - hi||_2, rounded up: the error bound of products that use the hi half only
    err2 = warp_sum(err2);
    if (lane == 0) hi_err[ro] = __fmul_ru(__fsqrt_ru(err2), 1.0001f);
  }
}

__global__ void mask_bits_kernel(const float* __restrict__ mask, int n_videos, int len, int chunks,
                                 unsigned int* __restrict__ bits) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // one thread per (video, chunk)
  if (i >= (long long)n_videos * chunks) return;
  const int v = (int)(i / chunks), c = (int)(i % chunks);
  unsigned int b = 0;
  for (int j = 0; j < 32; ++j) {
    const int l = c * 32 + j;
    if (l < len && mask[(long long)v * len + l] != 0.f) b |= 1u << j;
  }
  bits[i] = b;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace

int xmlb_make_tmap_2d_u16(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols,
                          unsigned int box_rows, unsigned int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  XMLB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  XMLB_REQUIRE(((uintptr_t)base & 15) == 0 && (cols * 2) % 16 == 0, "tensor map: base/pitch must be 16-byte aligned");
  XMLB_REQUIRE((box_cols == 64 || box_cols == 32) && box_rows >= 1 && box_rows <= 256, "tensor map: bad box");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMLB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XMLB_OK;
}

extern "C" int xmlb_split_rows(const float* x, const int* row_index, long long n_groups, int group_in, int group_out,
                               int k, int kpad, int out_ld, int out_col0, int normalize, int is_bf16,
                               unsigned short* hi, unsigned short* lo, float* hi_err, void* stream) {
  XMLB_REQUIRE(out_ld >= out_col0 + kpad && out_col0 >= 0, "xmlb_split_rows: need out_col0 + kpad <= out_ld");
  XMLB_REQUIRE(x && hi && lo, "xmlb_split_rows: null pointer");
  XMLB_REQUIRE(k >= 1 && kpad >= k && kpad % 64 == 0, "xmlb_split_rows: kpad must be a multiple of 64 and >= k");
  XMLB_REQUIRE(group_in >= 1 && group_out >= group_in, "xmlb_split_rows: need 1 <= group_in <= group_out");
  const long long rows_out = n_groups * group_out;
  if (rows_out == 0) return XMLB_OK;
  XMLB_REQUIRE(rows_out / 8 + 1 < (1ll << 31), "xmlb_split_rows: too many rows");
  const int blocks = ceil_div(rows_out, 8);
  if (is_bf16)
    split_rows_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, row_index, rows_out, k, kpad, out_ld,
                                                                       out_col0, group_in, group_out, normalize, hi,
                                                                       lo, hi_err);
  else
    split_rows_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, row_index, rows_out, k, kpad, out_ld,
                                                                        out_col0, group_in, group_out, normalize, hi,
                                                                        lo, hi_err);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

// Row gather of already split operands: dst_{hi,lo}[r] = src_{hi,lo}[row_index[r]] (16-byte copies, one warp per
// row; rows with a negative index are left untouched).  The queries of a block are split ONCE and then only copied
// into inverted-list order for the grouped kernels (re-scoring, span similarity).
__global__ void __launch_bounds__(256) gather_rows16_kernel(const uint4* __restrict__ src_hi, const uint4* __restrict__ src_lo,
                                                            const int* __restrict__ row_index, long long rows_out,
                                                            int vec_per_row, uint4* __restrict__ dst_hi,
                                                            uint4* __restrict__ dst_lo) {
  const int lane = threadIdx.x & 31;
  const long long ro = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (ro >= rows_out) return;
  const long long src = row_index[ro];
  if (src < 0) return;
  const uint4* sh = src_hi + src * vec_per_row;
  const uint4* sl = src_lo + src * vec_per_row;
  uint4* dh = dst_hi + ro * vec_per_row;
  uint4* dl = dst_lo + ro * vec_per_row;
  // up to 4 x 2 independent 16-byte loads in flight per lane before the first store (a 768-wide row is 3 iterations)
  for (int i0 = lane; i0 < vec_per_row; i0 += 128) {
    uint4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      if (i < vec_per_row) a[u] = __ldg(sh + i), b[u] = __ldg(sl + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      if (i < vec_per_row) dh[i] = a[u], dl[i] = b[u];
    }
  }
}

// Transposing split: x (rows, cols) fp32 -> hi / lo (cols, rpad) 16-bit with hi[c][r] + lo[c][r] ~= x[r][c], columns
// r >= rows zero-filled (rpad multiple of 64).  One pass through a 32 x 33 shared-memory tile instead of a transposing
// copy followed by xmlb_split_rows: the operands of the dW = dY^T . X GEMM of the training step (K = rows).
template <bool BF16>
__global__ void __launch_bounds__(256) split_rows_t_kernel(const float* __restrict__ x, long long rows, int cols,
                                                           long long rpad, unsigned short* __restrict__ hi,
                                                           unsigned short* __restrict__ lo) {
  __shared__ float tile[32][33];
  const long long r0 = blockIdx.y * 32ll;
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const long long r = r0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? __ldg(x + r * cols + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const long long r = r0 + tx;
    if (c < cols && r < rpad) {
      unsigned short a, b;
      split16<BF16>(tile[tx][j], a, b);
      hi[c * rpad + r] = a, lo[c * rpad + r] = b;
    }
  }
}

extern "C" int xmlb_split_rows_t(const float* x, long long rows, int cols, long long rpad, int is_bf16,
                                 unsigned short* hi, unsigned short* lo, void* stream) {
  XMLB_REQUIRE(x && hi && lo, "xmlb_split_rows_t: null pointer");
  XMLB_REQUIRE(rows >= 1 && cols >= 1 && rpad >= rows && rpad % 64 == 0, "xmlb_split_rows_t: rpad must be a multiple of 64, >= rows");
  XMLB_REQUIRE((rpad + 31) / 32 <= 65535, "xmlb_split_rows_t: too many rows");
  const dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)((rpad + 31) / 32));
  if (is_bf16)
    split_rows_t_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, rpad, hi, lo);
  else
    split_rows_t_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, rpad, hi, lo);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_gather_rows16(const unsigned short* src_hi, const unsigned short* src_lo, const int* row_index,
                                  long long rows_out, int kpad, unsigned short* dst_hi, unsigned short* dst_lo,
                                  void* stream) {
  XMLB_REQUIRE(src_hi && src_lo && row_index && dst_hi && dst_lo, "xmlb_gather_rows16: null pointer");
  XMLB_REQUIRE(kpad >= 8 && kpad % 8 == 0, "xmlb_gather_rows16: kpad must be a multiple of 8 (16-byte copies)");
  XMLB_REQUIRE((((uintptr_t)src_hi | (uintptr_t)src_lo | (uintptr_t)dst_hi | (uintptr_t)dst_lo) & 15) == 0,
               "xmlb_gather_rows16: buffers must be 16-byte aligned");
  if (rows_out == 0) return XMLB_OK;
  XMLB_REQUIRE(rows_out / 8 + 1 < (1ll << 31), "xmlb_gather_rows16: too many rows");
  gather_rows16_kernel<<<ceil_div(rows_out, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src_hi), reinterpret_cast<const uint4*>(src_lo), row_index, rows_out, kpad / 8,
      reinterpret_cast<uint4*>(dst_hi), reinterpret_cast<uint4*>(dst_lo));
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_mask_bits(const float* mask, int n_videos, int ctx_len, int lp, unsigned int* bits, void* stream) {
  XMLB_REQUIRE(mask && bits, "xmlb_mask_bits: null pointer");
  XMLB_REQUIRE(lp % 32 == 0 && lp >= ctx_len && ctx_len >= 1, "xmlb_mask_bits: lp must be a multiple of 32, >= ctx_len");
  const long long n = (long long)n_videos * (lp / 32);
  if (n == 0) return XMLB_OK;
  mask_bits_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mask, n_videos, ctx_len, lp / 32, bits);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_vr_scores_tc(const unsigned short* q_hi_a, const unsigned short* q_lo_a,
                                 const unsigned short* q_hi_b, const unsigned short* q_lo_b,
                                 const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                                 const unsigned short* c_hi_b, const unsigned short* c_lo_b,
                                 const unsigned int* mask_bits_a, const unsigned int* mask_bits_b, float* q2c,
                                 int* sched_ws, int n_queries, int n_videos, int lp, int kpad, int is_bf16,
                                 int max_ctas, void* stream) {
  XMLB_REQUIRE(q_hi_a && q_lo_a && c_hi_a && c_lo_a && mask_bits_a && q2c && sched_ws,
               "xmlb_vr_scores_tc: null pointer");
  const bool two = q_hi_b != nullptr;
  XMLB_REQUIRE(!two || (q_lo_b && c_hi_b && c_lo_b && mask_bits_b), "xmlb_vr_scores_tc: incomplete second modality");
  XMLB_REQUIRE(lp >= 32 && lp <= 256 && lp % 32 == 0, "xmlb_vr_scores_tc: lp must be a multiple of 32 in [32, 256]");
  XMLB_REQUIRE(kpad >= 64 && kpad % 64 == 0, "xmlb_vr_scores_tc: kpad must be a multiple of 64");
  XMLB_REQUIRE((long long)n_videos * lp < (1ll << 31), "xmlb_vr_scores_tc: corpus too large for one call");
  if (n_queries == 0 || n_videos == 0) return XMLB_OK;

  VrTcParams p = {};
  p.n_queries = n_queries, p.n_videos = n_videos, p.lp = lp;
  p.vpt = 256 / lp < MAX_VPT ? 256 / lp : MAX_VPT;
  p.block_n = p.vpt * lp;
  p.k_blocks = kpad / BLOCK_K;
  p.n_mod = two ? 2 : 1;
  p.m_tiles = ceil_div(n_queries, BLOCK_M);
  p.n_tiles = ceil_div(n_videos, p.vpt);
  p.mask_bits[0] = mask_bits_a, p.mask_bits[1] = mask_bits_b;
  p.out = q2c;
  p.tile_counter = sched_ws;
  p.divisor = (float)p.n_mod;
  p.idesc = tc::idesc_f16(BLOCK_M, p.block_n, is_bf16 ? 1 : 0);
  p.stages = tc::pipe_stages(p.block_n, 0);
  XMLB_REQUIRE(p.stages >= 2, "xmlb_vr_scores_tc: tile does not fit in shared memory");
  const size_t smem = tc::pipe_smem_bytes(p.block_n, p.stages, 0);

  VrMaps maps;
  const unsigned long long corpus_rows = (unsigned long long)n_videos * lp;
  const unsigned short* qh[2] = {q_hi_a, q_hi_b};
  const unsigned short* ql[2] = {q_lo_a, q_lo_b};
  const unsigned short* ch[2] = {c_hi_a, c_hi_b};
  const unsigned short* cl[2] = {c_lo_a, c_lo_b};
  for (int m = 0; m < p.n_mod; ++m) {
    int rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_hi[m], qh[m], n_queries, kpad, BLOCK_M, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_lo[m], ql[m], n_queries, kpad, BLOCK_M, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.b_hi[m], ch[m], corpus_rows, kpad, p.block_n, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.b_lo[m], cl[m], corpus_rows, kpad, p.block_n, BLOCK_K))) return rc;
  }
  if (!two) {
    maps.a_hi[1] = maps.a_hi[0], maps.a_lo[1] = maps.a_lo[0], maps.b_hi[1] = maps.b_hi[0], maps.b_lo[1] = maps.b_lo[0];
  }

  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = sms;
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const long long total = (long long)p.m_tiles * p.n_tiles;
  if (total < grid) grid = (int)total;
  XMLB_CUDA(cudaMemsetAsync(sched_ws, 0, sizeof(int), (cudaStream_t)stream));
  XMLB_CUDA(cudaFuncSetAttribute(vr_scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vr_scores_tc_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_vr_scores_tc_packed(const unsigned short* q_hi_a, const unsigned short* q_lo_a,
                                        const unsigned short* q_hi_b, const unsigned short* q_lo_b,
                                        const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                                        const unsigned short* c_hi_b, const unsigned short* c_lo_b,
                                        const int* tile_meta, const unsigned int* tile_starts, float* q2c,
                                        int* sched_ws, int n_queries, int n_videos, long long n_packed_rows,
                                        int n_tiles, int hi_only, const int* m_tile_list,
                                        const int* n_m_tiles, int kpad, int is_bf16, int max_ctas, void* stream) {
  XMLB_REQUIRE(q_hi_a && q_lo_a && c_hi_a && c_lo_a && tile_meta && tile_starts && q2c && sched_ws,
               "xmlb_vr_scores_tc_packed: null pointer");
  const bool two = q_hi_b != nullptr;
  XMLB_REQUIRE(!two || (q_lo_b && c_hi_b && c_lo_b), "xmlb_vr_scores_tc_packed: incomplete second modality");
  XMLB_REQUIRE(kpad >= 64 && kpad % 64 == 0, "xmlb_vr_scores_tc_packed: kpad must be a multiple of 64");
  XMLB_REQUIRE(n_packed_rows > 0 && n_packed_rows < (1ll << 31), "xmlb_vr_scores_tc_packed: bad packed row count");
  if (n_queries == 0 || n_tiles == 0) return XMLB_OK;
  VrPackedParams p = {};
  p.n_queries = n_queries, p.n_videos = n_videos, p.n_tiles = n_tiles;
  p.m_tiles = ceil_div(n_queries, BLOCK_M);
  p.k_blocks = kpad / BLOCK_K;
  p.n_mod = two ? 2 : 1;
  p.tile_meta = tile_meta, p.tile_starts = tile_starts;
  p.out = q2c;
  p.tile_counter = sched_ws;
  XMLB_REQUIRE((m_tile_list == nullptr) == (n_m_tiles == nullptr),
               "xmlb_vr_scores_tc_packed: m_tile_list and n_m_tiles go together");
  p.m_tile_list = m_tile_list, p.n_m_tiles = n_m_tiles;
  p.divisor = (float)p.n_mod;
  p.idesc = tc::idesc_f16(BLOCK_M, 256, is_bf16 ? 1 : 0);
  p.terms = hi_only ? 1 : 3;
  p.stages = tc::pipe_stages(256, PACKED_EXTRA_SMEM, p.terms);
#ifdef XMLB_ENABLE_PROBES  // limiter experiments only (python -m tvretrieval_b200.build --probes; tools/vr_filter_probe.py)
  if (const char* e = getenv("XMLB_VR_PROBE")) p.probe = atoi(e);
  if (const char* e = getenv("XMLB_VR_STAGES")) p.stages = atoi(e) < p.stages ? atoi(e) : p.stages;
#endif
  XMLB_REQUIRE(p.stages >= 2, "xmlb_vr_scores_tc_packed: tile does not fit in shared memory");
  const size_t smem = tc::pipe_smem_bytes(256, p.stages, PACKED_EXTRA_SMEM, p.terms);
  XMLB_REQUIRE(((uintptr_t)tile_meta & 15) == 0, "xmlb_vr_scores_tc_packed: tile_meta must be 16-byte aligned");
  VrMaps maps;
  const unsigned short* qh[2] = {q_hi_a, q_hi_b};
  const unsigned short* ql[2] = {q_lo_a, q_lo_b};
  const unsigned short* ch[2] = {c_hi_a, c_hi_b};
  const unsigned short* cl[2] = {c_lo_a, c_lo_b};
  for (int m = 0; m < p.n_mod; ++m) {
    int rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_hi[m], qh[m], n_queries, kpad, BLOCK_M, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_lo[m], ql[m], n_queries, kpad, BLOCK_M, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.b_hi[m], ch[m], n_packed_rows, kpad, 256, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.b_lo[m], cl[m], n_packed_rows, kpad, 256, BLOCK_K))) return rc;
  }
  if (!two) {
    maps.a_hi[1] = maps.a_hi[0], maps.a_lo[1] = maps.a_lo[0], maps.b_hi[1] = maps.b_hi[0], maps.b_lo[1] = maps.b_lo[0];
  }
  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = sms;
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const long long total = (long long)p.m_tiles * p.n_tiles;
  if (total < grid) grid = (int)total;
  XMLB_CUDA(cudaMemsetAsync(sched_ws, 0, sizeof(int), (cudaStream_t)stream));
  XMLB_CUDA(cudaFuncSetAttribute(vr_scores_tc_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vr_scores_tc_packed_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_vr_rescore_tc_kb(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a,
                                     const unsigned short* qg_hi_b, const unsigned short* qg_lo_b,
                                     const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                                     const unsigned short* c_hi_b, const unsigned short* c_lo_b, int c_kblocked,
                                     const int* row_start, const int* units, const int* n_units, int max_units,
                                     const int* entry_out, const int* entry_q, int gather_warps,
                                     long long n_query_rows, float* cand_val, int* sched_ws, long long n_entries,
                                     long long n_packed_rows, int max_len, int kpad, int is_bf16, void* stream) {
  XMLB_REQUIRE(qg_hi_a && qg_lo_a && c_hi_a && c_lo_a && row_start && units && n_units && entry_out && cand_val &&
                   sched_ws, "xmlb_vr_rescore_tc: null pointer");
  const bool two = qg_hi_b != nullptr;
  XMLB_REQUIRE(!two || (qg_lo_b && c_hi_b && c_lo_b), "xmlb_vr_rescore_tc: incomplete second modality");
  XMLB_REQUIRE(kpad >= 64 && kpad % 64 == 0, "xmlb_vr_rescore_tc: kpad must be a multiple of 64");
  XMLB_REQUIRE(max_len >= 1 && max_len <= 256, "xmlb_vr_rescore_tc: a video must have at most 256 clips");
  XMLB_REQUIRE(n_packed_rows > 0 && n_packed_rows < (1ll << 31) && n_entries < (1ll << 31),
               "xmlb_vr_rescore_tc: bad row count");
  XMLB_REQUIRE(((uintptr_t)units & 15) == 0, "xmlb_vr_rescore_tc: units must be 16-byte aligned");
  XMLB_REQUIRE(!entry_q || n_query_rows > 0, "xmlb_vr_rescore_tc: n_query_rows is required with entry_q");
  if (n_entries == 0 || max_units == 0) return XMLB_OK;
  VrRescoreParams p = {};
  p.n_mod = two ? 2 : 1;
  p.k_blocks = kpad / BLOCK_K;
  p.block_n = max_len <= 64 ? 64 : max_len <= 128 ? 128 : 256;
  p.units = reinterpret_cast<const int4*>(units), p.n_units = n_units, p.entry_out = entry_out, p.entry_q = entry_q;
  p.gather_warps = entry_q && gather_warps ? 1 : 0, p.kpad = kpad, p.is_bf16 = is_bf16 ? 1 : 0;
  p.q_hi[0] = qg_hi_a, p.q_lo[0] = qg_lo_a, p.q_hi[1] = two ? qg_hi_b : qg_hi_a, p.q_lo[1] = two ? qg_lo_b : qg_lo_a;
  p.row_start = row_start, p.out = cand_val, p.unit_counter = sched_ws;
  p.divisor = (float)p.n_mod;
  p.idesc = tc::idesc_f16(BLOCK_M, p.block_n, is_bf16 ? 1 : 0);
  p.stages = tc::pipe_stages(p.block_n, 0);
  XMLB_REQUIRE(p.stages >= 2, "xmlb_vr_rescore_tc: tile does not fit in shared memory");
  XMLB_REQUIRE(!c_kblocked || n_packed_rows * (kpad / BLOCK_K) < (1ll << 31),
               "xmlb_vr_rescore_tc: k-blocked corpus too large for 32-bit TMA coordinates");
  p.c_kb_rows = c_kblocked ? (int)n_packed_rows : 0;
  p.clip_boxes = p.gather_warps;  // (the other producers expect whole stages)
  const size_t smem = tc::pipe_smem_bytes(p.block_n, p.stages, 0);
  VrRescoreMaps maps;
  const unsigned short* qh[2] = {qg_hi_a, qg_hi_b};
  const unsigned short* ql[2] = {qg_lo_a, qg_lo_b};
  const unsigned short* ch[2] = {c_hi_a, c_hi_b};
  const unsigned short* cl[2] = {c_lo_a, c_lo_b};
  // gather mode: the A maps address single rows of the (n_query_rows, kpad) query arrays (TMA gather4)
  const unsigned long long a_rows = entry_q ? (unsigned long long)n_query_rows : (unsigned long long)n_entries;
  const unsigned int a_box = entry_q ? 1u : (unsigned int)BLOCK_M;
  for (int m = 0; m < p.n_mod; ++m) {
    int rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_hi[m], qh[m], a_rows, kpad, a_box, BLOCK_K))) return rc;
    if ((rc = xmlb_make_tmap_2d_u16(&maps.a_lo[m], ql[m], a_rows, kpad, a_box, BLOCK_K))) return rc;
    const unsigned long long b_rows = c_kblocked ? (unsigned long long)n_packed_rows * (kpad / BLOCK_K) : n_packed_rows;
    const unsigned long long b_cols = c_kblocked ? BLOCK_K : kpad;
    for (int b = p.clip_boxes ? 0 : RESCORE_BOXES - 1; b < RESCORE_BOXES; ++b) {
      const unsigned int box_rows = (unsigned int)(p.block_n / RESCORE_BOXES * (b + 1));
      if ((rc = xmlb_make_tmap_2d_u16(&maps.b_hi[m][b], ch[m], b_rows, b_cols, box_rows, BLOCK_K))) return rc;
      if ((rc = xmlb_make_tmap_2d_u16(&maps.b_lo[m][b], cl[m], b_rows, b_cols, box_rows, BLOCK_K))) return rc;
    }
  }
  if (!two) {
    maps.a_hi[1] = maps.a_hi[0], maps.a_lo[1] = maps.a_lo[0];
    for (int b = 0; b < RESCORE_BOXES; ++b) maps.b_hi[1][b] = maps.b_hi[0][b], maps.b_lo[1][b] = maps.b_lo[0][b];
  }
  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = max_units < sms ? max_units : sms;
  XMLB_CUDA(cudaMemsetAsync(sched_ws, 0, sizeof(int), (cudaStream_t)stream));
  XMLB_CUDA(cudaFuncSetAttribute(vr_rescore_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vr_rescore_tc_kernel<<<grid, p.gather_warps ? 192 + 32 * tc::GATHER_WARPS : 192, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_vr_rescore_tc_ex(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a,
                                     const unsigned short* qg_hi_b, const unsigned short* qg_lo_b,
                                     const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                                     const unsigned short* c_hi_b, const unsigned short* c_lo_b, const int* row_start,
                                     const int* units, const int* n_units, int max_units, const int* entry_out,
                                     const int* entry_q, int gather_warps, long long n_query_rows, float* cand_val,
                                     int* sched_ws, long long n_entries, long long n_packed_rows, int max_len, int kpad,
                                     int is_bf16, void* stream) {
  return xmlb_vr_rescore_tc_kb(qg_hi_a, qg_lo_a, qg_hi_b, qg_lo_b, c_hi_a, c_lo_a, c_hi_b, c_lo_b, 0, row_start, units,
                               n_units, max_units, entry_out, entry_q, gather_warps, n_query_rows, cand_val, sched_ws,
                               n_entries, n_packed_rows, max_len, kpad, is_bf16, stream);
}

extern "C" int xmlb_vr_rescore_tc(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a,
                                  const unsigned short* qg_hi_b, const unsigned short* qg_lo_b,
                                  const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                                  const unsigned short* c_hi_b, const unsigned short* c_lo_b, const int* row_start,
                                  const int* units, const int* n_units, int max_units, const int* entry_out,
                                  float* cand_val, int* sched_ws, long long n_entries, long long n_packed_rows,
                                  int max_len, int kpad, int is_bf16, void* stream) {
  return xmlb_vr_rescore_tc_ex(qg_hi_a, qg_lo_a, qg_hi_b, qg_lo_b, c_hi_a, c_lo_a, c_hi_b, c_lo_b, row_start, units,
                               n_units, max_units, entry_out, nullptr, 0, 0, cand_val, sched_ws, n_entries,
                               n_packed_rows, max_len, kpad, is_bf16, stream);
}
