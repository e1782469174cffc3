// nn.Linear (+bias, +ReLU, +residual) on the tcgen05 tensor cores with split-precision operands.
//   replaces the cuBLAS sgemm behind LinearLayer / the Q,K,V projections / BertSelfOutput.dense /
//   {video,sub}_query_linear (reference model_components.py:160-163, 278-280, 314; model_xml.py:459-460)
// out[m][n] = act( sum_k x[m][k] * w[n][k] + bias[n] + residual[m][n] ), x and w given as 16-bit (hi, lo) pairs
// (xmlb_split_rows), three MMAs per product, fp32 accumulation in TMEM -- same mainloop as the video-level score
// kernel (tc_pipeline.cuh); the epilogue threads each own one output row and stream it out in 128-byte pieces.
#include "tc_pipeline.cuh"
#include "xmlb200.h"

namespace {

using tc::BLOCK_K;
using tc::BLOCK_M;

struct LinMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

struct LinParams {
  int rows, out_dim, k_blocks, block_n, m_tiles, n_tiles, stages, relu;
  const float* bias;
  const float* residual;
  float* out;
  int* tile_counter;  // zeroed before the launch
  unsigned int idesc;
};

struct LinSched {  // output-feature tile fastest: the activation tile is re-read from L2 by its sibling tiles
  const LinMaps* maps;
  const LinParams* p;
  __device__ LinSched(const LinMaps* m, const LinParams* pp) : maps(m), p(pp) {}
  __device__ bool next(tc::UnitDesc& u) {
    const int tile = atomicAdd(p->tile_counter, 1);
    if (tile >= p->m_tiles * p->n_tiles) return false;
    u.a_hi = &maps->a_hi, u.a_lo = &maps->a_lo, u.b_hi = &maps->b_hi, u.b_lo = &maps->b_lo;
    u.a_row = (tile / p->n_tiles) * BLOCK_M;
    u.b_row = (tile % p->n_tiles) * p->block_n;
    u.k_blocks = p->k_blocks;
    u.idesc = p->idesc;
    u.tag0 = tile, u.tag1 = 0;
    return true;
  }
};

__global__ void __launch_bounds__(192, 1)
linear_tc_kernel(const __grid_constant__ LinMaps maps, const __grid_constant__ LinParams p) {
  extern __shared__ unsigned char smem_raw[];
  tc::Pipe pipe;
  const uint32_t tmem_base = tc::pipe_setup(pipe, smem_raw, p.stages, p.block_n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) tc::tc_producer_loop(LinSched(&maps, &p), pipe);
  } else if (warp == 1) {
    if (lane == 0) tc::tc_mma_loop(pipe, tmem_base);
  } else {  // ===================== epilogue warps 2..5 =====================
    const int row_in_tile = (warp & 3) * 32 + lane;
    const bool vec_ok = (p.out_dim & 3) == 0;
    int t, tag1;
    for (uint32_t unit = 0; tc::epi_next(pipe, unit, t, tag1); ++unit) {
      const int m_tile = t / p.n_tiles, n_tile = t % p.n_tiles;
      const long long row = (long long)m_tile * BLOCK_M + row_in_tile;
      const uint32_t taddr = tc::epi_wait(pipe, unit, tmem_base);
      const int n_begin = n_tile * p.block_n;
      const int n_end = min(n_begin + p.block_n, p.out_dim);
      for (int n0 = n_begin; n0 < n_end; n0 += 32) {  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + (n0 - n_begin), r);
        tc::tmem_ld_wait();
        if (row < p.rows) {
          float* __restrict__ o = p.out + row * p.out_dim + n0;
          const float* __restrict__ res = p.residual ? p.residual + row * p.out_dim + n0 : nullptr;
          if (vec_ok && n0 + 32 <= n_end) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 v = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                     __uint_as_float(r[i + 3]));
              if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
              }
              if (res) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(res + i));
                v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
              }
              if (p.relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
              *reinterpret_cast<float4*>(o + i) = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (n0 + i < n_end) {
                float v = __uint_as_float(r[i]);
                if (p.bias) v += __ldg(p.bias + n0 + i);
                if (res) v += __ldg(res + i);
                if (p.relu) v = fmaxf(v, 0.f);
                o[i] = v;
              }
            }
          }
        }
      }
      tc::epi_release(pipe, unit);
    }
  }
  tc::pipe_teardown(tmem_base);
}

}  // namespace

extern "C" int xmlb_linear_tc(const unsigned short* x_hi, const unsigned short* x_lo, const unsigned short* w_hi,
                              const unsigned short* w_lo, const float* bias, const float* residual, float* out,
                              int* sched_ws, long long rows, int out_dim, int kpad, int relu, int is_bf16,
                              void* stream) {
  XMLB_REQUIRE(x_hi && x_lo && w_hi && w_lo && out && sched_ws, "xmlb_linear_tc: null pointer");
  XMLB_REQUIRE(rows >= 0 && rows < (1ll << 31) - 256 && out_dim >= 1, "xmlb_linear_tc: bad shape");
  XMLB_REQUIRE(kpad >= 64 && kpad % 64 == 0, "xmlb_linear_tc: kpad must be a multiple of 64");
  XMLB_REQUIRE(((uintptr_t)out & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0) &&
                   (!residual || ((uintptr_t)residual & 15) == 0),
               "xmlb_linear_tc: out / bias / residual must be 16-byte aligned");
  if (rows == 0) return XMLB_OK;
  LinParams p = {};
  p.rows = (int)rows, p.out_dim = out_dim;
  p.k_blocks = kpad / BLOCK_K;
  p.block_n = out_dim >= 256 ? 256 : (out_dim + 15) / 16 * 16;
  p.m_tiles = ceil_div(rows, BLOCK_M);
  p.n_tiles = ceil_div(out_dim, p.block_n);
  p.relu = relu, p.bias = bias, p.residual = residual, p.out = out, p.tile_counter = sched_ws;
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
  linear_tc_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
