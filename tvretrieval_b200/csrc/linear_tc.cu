// nn.Linear (+bias, +ReLU, +residual) on the tcgen05 tensor cores with split-precision operands.
//   replaces the cuBLAS sgemm behind LinearLayer / the Q,K,V projections / BertSelfOutput.dense /
//   {video,sub}_query_linear (reference model_components.py:160-163, 278-280, 314; model_xml.py:459-460)
// out[m][n] = act( sum_k x[m][k] * w[n][k] + bias[n] + residual[m][n] ), x and w given as 16-bit (hi, lo) pairs
// (xmlb_split_rows), three MMAs per product, fp32 accumulation in TMEM -- same mainloop as the video-level score
// kernel (tc_pipeline.cuh).
//
// K-CHUNKED ACCUMULATION.  The fp32 accumulator of tcgen05.mma truncates instead of rounding (measured on B200:
// 2.6e-5 abs at K = 768 for O(1) outputs against 1e-6 for an FMA loop -- a bias that grows with the number of
// accumulator updates, 3 * K / 16).  A unit of the pipeline is therefore one K-CHUNK of an output tile (128 elements
// of K by default = 24 updates): the tensor core sums a chunk in TMEM, eight epilogue warps drain it and add it to
// register accumulators with round-to-nearest fp32 adds (96 columns per thread: two warps share each TMEM lane
// quadrant, one per half of the 192-column tile), while the MMAs of the next chunk fill the other TMEM accumulator.  The result
// is as accurate as an fp32 FMA loop (error ~1e-6 relative at K = 3072), at tensor-core speed.
//
// FUSED OUTPUT FORMATS.  Besides the fp32 row-major output the epilogue can emit what the next kernel wants, so the
// activation never makes an extra round trip through HBM: the 16-bit (hi, lo) split of the output rows (operand of
// the next Linear / of the attention kernel's QK^T), and -- for a trailing block of output columns, the V part of a
// fused QKV projection -- the split TRANSPOSED per sequence, vt[b][c][l] = out[b * seq + l][col0 + c], which is the
// K-major B operand of the attention kernel's P.V product.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_pipeline.cuh"
#include "xmlb200.h"

namespace {

using tc::BLOCK_K;
using tc::BLOCK_M;

constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
// Output tile = 128 rows x 192 columns, 96 columns per epilogue thread: a CTA of 10 warps puts 3 warps on some of the
// SM's four sub-partitions (16 K registers each), which caps a thread at 168 registers -- 96 running sums + 32 freshly
// loaded values + addressing fit, 128 + 32 do not.
constexpr int HALF_N = 96;
constexpr int MAX_BLOCK_N = 2 * HALF_N;

struct LinMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

struct LinParams {
  int rows, out_dim, k_blocks, kb_per_chunk, n_chunks, block_n, m_tiles, n_tiles, stages, relu, is_bf16;
  const float* bias;
  const float* residual;
  float* out;              // fp32 (rows, out_dim), or null
  unsigned short* o_hi;    // split row-major output of columns [0, o_cols): (rows, o_ld) each, or null
  unsigned short* o_lo;
  int o_ld, o_cols;
  unsigned short* t_hi;    // split transposed output of columns [t_col0, out_dim), or null
  unsigned short* t_lo;
  int t_col0, t_seq, t_ld;
  int* tile_counter;  // zeroed before the launch
  unsigned int idesc;
};

struct LinSched {  // output-feature tile fastest: the activation tile is re-read from L2 by its sibling tiles
  const LinMaps* maps;
  const LinParams* p;
  int tile, chunk;
  __device__ LinSched(const LinMaps* m, const LinParams* pp) : maps(m), p(pp), tile(0), chunk(1 << 30) {}
  __device__ bool next(tc::UnitDesc& u) {
    if (chunk >= p->n_chunks) {
      tile = atomicAdd(p->tile_counter, 1);
      chunk = 0;
    }
    if (tile >= p->m_tiles * p->n_tiles) return false;
    u.a_hi = &maps->a_hi, u.a_lo = &maps->a_lo, u.b_hi = &maps->b_hi, u.b_lo = &maps->b_lo;
    u.a_row = (tile / p->n_tiles) * BLOCK_M;
    u.b_row = (tile % p->n_tiles) * p->block_n;
    u.k_block0 = chunk * p->kb_per_chunk;
    u.k_blocks = min(p->kb_per_chunk, p->k_blocks - u.k_block0);
    u.idesc = p->idesc;
    u.tag0 = tile, u.tag1 = chunk;
    ++chunk;
    return true;
  }
};

__device__ __forceinline__ void split16(float x, int is_bf16, unsigned short& hi, unsigned short& lo) {
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

__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const __grid_constant__ LinMaps maps, const __grid_constant__ LinParams p) {
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, p.block_n, 3, EPI_WARPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) tc::tc_producer_loop(LinSched(&maps, &p), pipe);
  } else if (warp == 1) {
    tc::tc_mma_loop_warp(pipe, tmem_base);
  } else {  // ===================== epilogue warps 2..9 =====================
    const int row_in_tile = (warp & 3) * 32 + lane;   // TMEM lane = output row of the tile
    const int half = (warp - 2) >> 2;                 // which half of the accumulator columns
    float acc[HALF_N];
    int t, chunk;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, t, chunk); ++unit) {
      const int n_tile = t % p.n_tiles;
      const int n_begin = n_tile * p.block_n + half * HALF_N;              // first output column of this thread
      const int n_end = min(n_tile * p.block_n + p.block_n, p.out_dim);    // (warp-uniform)
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base) + half * HALF_N;
#pragma unroll
      for (int g = 0; g < HALF_N / 32; ++g) {
        if (n_begin + g * 32 < n_end) {
          uint32_t r[32];
          tc::tmem_ld_32x32(taddr + g * 32, r);
          tc::tmem_ld_wait();
          if (chunk == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[g * 32 + i] = __uint_as_float(r[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[g * 32 + i] = __fadd_rn(acc[g * 32 + i], __uint_as_float(r[i]));
          }
        }
      }
      tc::epi_release(pipe, unit);  // the accumulator is in registers: the MMAs of the chunk after next may start
      if (chunk != p.n_chunks - 1) continue;

      // ---- last chunk of the tile: bias / residual / ReLU, then the requested output formats ----
      const long long row = (long long)(t / p.n_tiles) * BLOCK_M + row_in_tile;
      if (row >= p.rows) continue;
      const long long tb = p.t_hi ? row / p.t_seq : 0;  // sequence of this row, position inside it
      const int tl = p.t_hi ? (int)(row - tb * p.t_seq) : 0;
#pragma unroll
      for (int g = 0; g < HALF_N / 32; ++g) {
        const int n0 = n_begin + g * 32;
        if (n0 >= n_end) break;
        const bool full = n0 + 32 <= n_end && (p.out_dim & 3) == 0;
        float* v = acc + g * 32;  // (g is a compile-time constant after unrolling: still registers)
        if (full) {
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
              v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
            }
          }
          if (p.residual) {
            const float* res = p.residual + row * p.out_dim + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(res + i));
              v[i] += a.x, v[i + 1] += a.y, v[i + 2] += a.z, v[i + 3] += a.w;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (n0 + i < n_end) {
              if (p.bias) v[i] += __ldg(p.bias + n0 + i);
              if (p.residual) v[i] += __ldg(p.residual + row * p.out_dim + n0 + i);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (p.out) {
          float* o = p.out + row * p.out_dim + n0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < n_end) o[i] = v[i];
          }
        }
        if (p.o_hi && n0 < p.o_cols) {  // 16-bit split, row-major (o_cols and o_ld are multiples of 8: 16-byte stores)
          unsigned short* oh = p.o_hi + row * p.o_ld + n0;
          unsigned short* ol = p.o_lo + row * p.o_ld + n0;
          if (full && n0 + 32 <= p.o_cols) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              unsigned short h[8], l[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) split16(v[i + j], p.is_bf16, h[j], l[j]);
              *reinterpret_cast<uint4*>(oh + i) =
                  make_uint4(h[0] | (uint32_t)h[1] << 16, h[2] | (uint32_t)h[3] << 16, h[4] | (uint32_t)h[5] << 16,
                             h[6] | (uint32_t)h[7] << 16);
              *reinterpret_cast<uint4*>(ol + i) =
                  make_uint4(l[0] | (uint32_t)l[1] << 16, l[2] | (uint32_t)l[3] << 16, l[4] | (uint32_t)l[5] << 16,
                             l[6] | (uint32_t)l[7] << 16);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (n0 + i < n_end && n0 + i < p.o_cols) split16(v[i], p.is_bf16, oh[i], ol[i]);
            }
          }
        }
        if (p.t_hi && n0 + 32 > p.t_col0) {  // transposed split: consecutive lanes = consecutive positions l
          const int n_vt = p.out_dim - p.t_col0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = n0 + i - p.t_col0;
            if (c >= 0 && n0 + i < n_end) {
              const long long at = (tb * n_vt + c) * (long long)p.t_ld + tl;
              split16(v[i], p.is_bf16, p.t_hi[at], p.t_lo[at]);
            }
          }
        }
      }
    }
  }
  tc::pipe_teardown(tmem_base);
}

}  // namespace

extern "C" int xmlb_linear_tc_ex(const unsigned short* x_hi, const unsigned short* x_lo, const unsigned short* w_hi,
                                 const unsigned short* w_lo, const float* bias, const float* residual, float* out,
                                 unsigned short* out_hi, unsigned short* out_lo, int out16_ld, int out16_cols,
                                 unsigned short* vt_hi, unsigned short* vt_lo, int vt_col0, int vt_seq, int vt_ld,
                                 int* sched_ws, long long rows, int out_dim, int kpad, int relu, int is_bf16,
                                 int k_chunk, void* stream) {
  XMLB_REQUIRE(x_hi && x_lo && w_hi && w_lo && sched_ws, "xmlb_linear_tc: null pointer");
  XMLB_REQUIRE(out || out_hi || vt_hi, "xmlb_linear_tc: no output requested");
  XMLB_REQUIRE(rows >= 0 && rows < (1ll << 31) - 256 && out_dim >= 1, "xmlb_linear_tc: bad shape");
  XMLB_REQUIRE(kpad >= 64 && kpad % 64 == 0, "xmlb_linear_tc: kpad must be a multiple of 64");
  XMLB_REQUIRE((!out || ((uintptr_t)out & 15) == 0) && (!bias || ((uintptr_t)bias & 15) == 0) &&
                   (!residual || ((uintptr_t)residual & 15) == 0),
               "xmlb_linear_tc: out / bias / residual must be 16-byte aligned");
  XMLB_REQUIRE(k_chunk == 0 || (k_chunk >= BLOCK_K && k_chunk % BLOCK_K == 0),
               "xmlb_linear_tc: k_chunk must be a multiple of 32 (0 = default)");
  if (out_hi) {
    XMLB_REQUIRE(out_lo && out16_cols >= 1 && out16_cols <= out_dim && out16_ld >= out16_cols && out16_ld % 8 == 0 &&
                     out16_cols % 8 == 0 && (((uintptr_t)out_hi | (uintptr_t)out_lo) & 15) == 0,
                 "xmlb_linear_tc: split output needs out_lo, out16_cols <= out_dim, out16_ld >= out16_cols, both "
                 "multiples of 8, 16-byte aligned buffers");
  }
  if (vt_hi) {
    XMLB_REQUIRE(vt_lo && vt_col0 >= 0 && vt_col0 < out_dim && vt_seq >= 1 && vt_ld >= vt_seq && rows % vt_seq == 0,
                 "xmlb_linear_tc: transposed output needs vt_lo, 0 <= vt_col0 < out_dim, vt_ld >= vt_seq >= 1 and "
                 "rows a multiple of vt_seq");
  }
  if (rows == 0) return XMLB_OK;
  LinParams p = {};
  p.rows = (int)rows, p.out_dim = out_dim;
  p.k_blocks = kpad / BLOCK_K;
  p.kb_per_chunk = (k_chunk ? k_chunk : 128) / BLOCK_K;
  p.n_chunks = ceil_div(p.k_blocks, p.kb_per_chunk);
  p.block_n = out_dim >= MAX_BLOCK_N ? MAX_BLOCK_N : (out_dim + 15) / 16 * 16;
  p.m_tiles = ceil_div(rows, BLOCK_M);
  p.n_tiles = ceil_div(out_dim, p.block_n);
  p.relu = relu, p.is_bf16 = is_bf16 ? 1 : 0;
  p.bias = bias, p.residual = residual, p.out = out, p.tile_counter = sched_ws;
  p.o_hi = out_hi, p.o_lo = out_lo, p.o_ld = out16_ld, p.o_cols = out_hi ? out16_cols : 0;
  p.t_hi = vt_hi, p.t_lo = vt_lo, p.t_col0 = vt_hi ? vt_col0 : out_dim, p.t_seq = vt_hi ? vt_seq : 1, p.t_ld = vt_ld;
  p.idesc = tc::idesc_f16(BLOCK_M, p.block_n, is_bf16 ? 1 : 0);
  p.stages = tc::pipe_stages(p.block_n, 0);
  XMLB_REQUIRE(p.stages >= 2, "xmlb_linear_tc: tile does not fit in shared memory");
  const size_t smem = tc::pipe_smem_bytes(p.block_n, p.stages, 0);

  LinMaps maps;
  int rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.a_hi, x_hi, rows, kpad, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.a_lo, x_lo, rows, kpad, BLOCK_M, BLOCK_K))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.b_hi, w_hi, out_dim, kpad, p.block_n, BLOCK_K))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.b_lo, w_lo, out_dim, kpad, p.block_n, BLOCK_K))) return rc;

  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = (long long)p.m_tiles * p.n_tiles;
  const int grid = total < sms ? (int)total : sms;
  XMLB_CUDA(cudaMemsetAsync(sched_ws, 0, sizeof(int), (cudaStream_t)stream));
  XMLB_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_linear_tc(const unsigned short* x_hi, const unsigned short* x_lo, const unsigned short* w_hi,
                              const unsigned short* w_lo, const float* bias, const float* residual, float* out,
                              int* sched_ws, long long rows, int out_dim, int kpad, int relu, int is_bf16,
                              void* stream) {
  XMLB_REQUIRE(out, "xmlb_linear_tc: null pointer");
  return xmlb_linear_tc_ex(x_hi, x_lo, w_hi, w_lo, bias, residual, out, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, 0,
                           0, sched_ws, rows, out_dim, kpad, relu, is_bf16, 0, stream);
}
