// Fused multi-head attention core on the tcgen05 tensor cores (context encoders: sequences of up to 256 clips).
//   replaces reference BertSelfAttention.forward after the projections (model_components.py:277-303):
//     out = merge_heads( softmax( Q_h K_h^T / sqrt(dh) + (1 - mask) * -10000 ) V_h )
//   for self attention (one key mask per sequence) and for the cross attention of model_xml.py:357-373 (full
//   (Lq, Lk) mask).  No (batch, heads, Lq, Lk) score tensor ever exists in HBM.
//
// One unit = (sequence b, head h, tile of 128 query rows).  Per unit, inside one persistent CTA (192 threads):
//   warp 0 / lane 0   TMA producer: streams 64-wide K-chunks of Q_h (128 x 64) and K_h (Lk x 64), then 64-key
//                     chunks of V_h^T (dh x 64), all as 16-bit (hi, lo) pairs, through a ring of shared-memory stages
//   warp 1 / lane 0   MMA issuer: S = Q K^T into TMEM columns [0, Lk) (3 MMAs per product: hi*lo + lo*hi + hi*hi,
//                     fp32 accumulate), then O = P V into TMEM columns [256, 256 + dh)
//   warps 2..5        one thread per query row: reads its row of S from TMEM (tcgen05.ld), applies the 1/sqrt(dh)
//                     scale and the additive -10000 mask in fp32 exactly like the reference (including the
//                     fully-masked rows of padded clips), row max, e = exp(x - max), row
//                     sum; writes e as (hi, lo) 16-bit halves straight into the A-operand slot of the stage whose
//                     B slot receives the matching V^T chunk (128B-swizzled K-major layout, the one TMA produces),
//                     so P never leaves the SM; finally reads O from TMEM, divides by the row sum and stores the
//                     context rows as fp32 and / or as the (hi, lo) split the output projection consumes.
// Operands come pre-split from the projection GEMM's epilogue (linear_tc.cu): Q and K row-major (tokens x hidden),
// V transposed per sequence (hidden x tokens), so every tile is a plain 2-D TMA box.
// The kernel is HBM-bound by construction (per unit it reads 3 * L * dh * 4 B and writes L * dh * 4 B for
// 4 * L^2 * dh * 3 tensor-core flops), which is why one accumulator set per CTA suffices: the producer runs ahead
// into the free stages while the softmax of the current unit executes.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include "tc_common.cuh"
#include "xmlb200.h"

namespace {

constexpr int AT_M = 128;            // query rows per unit (TMEM lanes)
constexpr int AT_KC = 64;            // K-chunk: 64 16-bit elements = one 128-byte swizzle row
constexpr int AT_A_BYTES = AT_M * AT_KC * 2;   // one A tile (hi or lo)
constexpr int AT_O_COL = 256;        // TMEM column of the output accumulator
constexpr int AT_MAX_STAGES = 6;

struct AttMaps {
  CUtensorMap q_hi, q_lo, k_hi, k_lo, v_hi, v_lo;
};

struct AttParams {
  int batch, len_q, len_k, hidden, n_heads, dh, n_qt;
  int kbox;        // rows of the K box = len_k rounded up to 16 (MMA N of S)
  int brows;       // rows reserved for the B slot of a stage = max(kbox, dh)
  int stages, n_kc, n_jc;
  int q_col0, k_col0;
  int is_bf16;
  const float* mask;
  long long mask_batch_stride, mask_q_stride;
  float* out;             // fp32 (batch * len_q, hidden) or null
  unsigned short* o_hi;   // split output (batch * len_q, o_ld) or null
  unsigned short* o_lo;
  int o_ld;
  float inv_sqrt_dh;
};

struct AttSmem {
  uint32_t base, bar;
  int stages, stage_bytes, b_lo_off;
  __device__ uint32_t stage(int s) const { return base + s * stage_bytes; }
  __device__ uint32_t full(int s) const { return bar + 8u * s; }
  __device__ uint32_t empty(int s) const { return bar + 8u * (AT_MAX_STAGES + s); }
  __device__ uint32_t pfull(int s) const { return bar + 8u * (2 * AT_MAX_STAGES + s); }
  __device__ uint32_t s_full() const { return bar + 8u * (3 * AT_MAX_STAGES + 0); }
  __device__ uint32_t s_empty() const { return bar + 8u * (3 * AT_MAX_STAGES + 1); }
  __device__ uint32_t o_full() const { return bar + 8u * (3 * AT_MAX_STAGES + 2); }
  __device__ uint32_t o_empty() const { return bar + 8u * (3 * AT_MAX_STAGES + 3); }
  __device__ uint32_t tmem_slot() const { return bar + 8u * (3 * AT_MAX_STAGES + 4); }
};
constexpr int AT_BAR_BYTES = 8 * (3 * AT_MAX_STAGES + 5);

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 8 consecutive probabilities -> one 16-byte group of hi halves and one of lo halves
__device__ __forceinline__ void split8(const float* e, int is_bf16, uint4& hi, uint4& lo) {
  unsigned short h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (is_bf16) {
      const __nv_bfloat16 a = __float2bfloat16_rn(e[j]);
      const __nv_bfloat16 b = __float2bfloat16_rn(e[j] - __bfloat162float(a));
      h[j] = __bfloat16_as_ushort(a), l[j] = __bfloat16_as_ushort(b);
    } else {
      const __half a = __float2half_rn(e[j]);
      const __half b = __float2half_rn(e[j] - __half2float(a));
      h[j] = __half_as_ushort(a), l[j] = __half_as_ushort(b);
    }
  }
  hi = make_uint4(h[0] | (uint32_t)h[1] << 16, h[2] | (uint32_t)h[3] << 16, h[4] | (uint32_t)h[5] << 16,
                  h[6] | (uint32_t)h[7] << 16);
  lo = make_uint4(l[0] | (uint32_t)l[1] << 16, l[2] | (uint32_t)l[3] << 16, l[4] | (uint32_t)l[5] << 16,
                  l[6] | (uint32_t)l[7] << 16);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(192, 1)
attention_tc_kernel(const __grid_constant__ AttMaps maps, const __grid_constant__ AttParams p) {
  extern __shared__ unsigned char smem_raw[];
  AttSmem sm;
  sm.base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.stages = p.stages;
  sm.b_lo_off = 2 * AT_A_BYTES + p.brows * AT_KC * 2;
  sm.stage_bytes = 2 * AT_A_BYTES + 2 * p.brows * AT_KC * 2;
  sm.bar = sm.base + p.stages * sm.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(sm.full(s), 1);
      tc::mbar_init(sm.empty(s), 1);
      tc::mbar_init(sm.pfull(s), 4);
    }
    tc::mbar_init(sm.s_full(), 1);
    tc::mbar_init(sm.s_empty(), 4);
    tc::mbar_init(sm.o_full(), 1);
    tc::mbar_init(sm.o_empty(), 4);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(sm.tmem_slot(), 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot()));

  const int total_units = p.batch * p.n_heads * p.n_qt;
  const int chunks_per_unit = p.n_kc + p.n_jc;
  const uint32_t qk_bytes = 2u * AT_A_BYTES + 2u * p.kbox * AT_KC * 2;
  const uint32_t v_bytes = 2u * p.dh * AT_KC * 2;

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer =====================
      uint32_t n = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int qt = unit % p.n_qt, h = (unit / p.n_qt) % p.n_heads, b = unit / (p.n_qt * p.n_heads);
        const int q_row = b * p.len_q + qt * AT_M, k_row = b * p.len_k, v_row = b * p.hidden + h * p.dh;
        for (int c = 0; c < chunks_per_unit; ++c, ++n) {
          const int s = n % p.stages;
          tc::mbar_wait(sm.empty(s), ((n / p.stages) & 1u) ^ 1u);
          const uint32_t st = sm.stage(s);
          if (c < p.n_kc) {
            tc::mbar_expect_tx(sm.full(s), qk_bytes);
            tc::tma_load_2d(st, &maps.q_hi, sm.full(s), p.q_col0 + h * p.dh + c * AT_KC, q_row);
            tc::tma_load_2d(st + AT_A_BYTES, &maps.q_lo, sm.full(s), p.q_col0 + h * p.dh + c * AT_KC, q_row);
            tc::tma_load_2d(st + 2 * AT_A_BYTES, &maps.k_hi, sm.full(s), p.k_col0 + h * p.dh + c * AT_KC, k_row);
            tc::tma_load_2d(st + sm.b_lo_off, &maps.k_lo, sm.full(s), p.k_col0 + h * p.dh + c * AT_KC, k_row);
          } else {
            tc::mbar_expect_tx(sm.full(s), v_bytes);
            tc::tma_load_2d(st + 2 * AT_A_BYTES, &maps.v_hi, sm.full(s), (c - p.n_kc) * AT_KC, v_row);
            tc::tma_load_2d(st + sm.b_lo_off, &maps.v_lo, sm.full(s), (c - p.n_kc) * AT_KC, v_row);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================== MMA issuer =====================
      const uint32_t idesc_s = tc::idesc_f16(AT_M, p.kbox, p.is_bf16);
      const uint32_t idesc_o = tc::idesc_f16(AT_M, p.dh, p.is_bf16);
      uint32_t n = 0, it = 0, pf_parity = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
        tc::mbar_wait(sm.s_empty(), (it & 1u) ^ 1u);  // the softmax warps have consumed the previous S
        tc::fence_after_sync();
        for (int c = 0; c < p.n_kc; ++c, ++n) {
          const int s = n % p.stages;
          tc::mbar_wait(sm.full(s), (n / p.stages) & 1u);
          tc::fence_after_sync();
          const uint32_t st = sm.stage(s);
          const uint64_t a_hi = tc::smem_desc_kmajor<128>(st), a_lo = tc::smem_desc_kmajor<128>(st + AT_A_BYTES);
          const uint64_t b_hi = tc::smem_desc_kmajor<128>(st + 2 * AT_A_BYTES);
          const uint64_t b_lo = tc::smem_desc_kmajor<128>(st + sm.b_lo_off);
#pragma unroll
          for (int k = 0; k < AT_KC / 16; ++k) {
            const uint64_t off = (uint64_t)(k * 32 >> 4);
            tc::umma_f16(tmem_base, a_hi + off, b_lo + off, idesc_s, (c | k) != 0);
            tc::umma_f16(tmem_base, a_lo + off, b_hi + off, idesc_s, 1u);
            tc::umma_f16(tmem_base, a_hi + off, b_hi + off, idesc_s, 1u);
          }
          tc::umma_commit(sm.empty(s));
        }
        tc::umma_commit(sm.s_full());
        tc::mbar_wait(sm.o_empty(), (it & 1u) ^ 1u);  // the previous unit's O has been read
        tc::fence_after_sync();
        for (int c = 0; c < p.n_jc; ++c, ++n) {
          const int s = n % p.stages;
          tc::mbar_wait(sm.full(s), (n / p.stages) & 1u);      // V^T chunk landed (TMA)
          tc::mbar_wait(sm.pfull(s), (pf_parity >> s) & 1u);   // P chunk written by the softmax warps
          pf_parity ^= 1u << s;
          tc::fence_after_sync();
          const uint32_t st = sm.stage(s);
          const uint64_t a_hi = tc::smem_desc_kmajor<128>(st), a_lo = tc::smem_desc_kmajor<128>(st + AT_A_BYTES);
          const uint64_t b_hi = tc::smem_desc_kmajor<128>(st + 2 * AT_A_BYTES);
          const uint64_t b_lo = tc::smem_desc_kmajor<128>(st + sm.b_lo_off);
          const int ksteps = min(AT_KC / 16, (p.len_k - c * AT_KC + 15) / 16);  // keys beyond len_k carry P = 0
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t off = (uint64_t)(k * 32 >> 4);
            tc::umma_f16(tmem_base + AT_O_COL, a_hi + off, b_lo + off, idesc_o, (c | k) != 0);
            tc::umma_f16(tmem_base + AT_O_COL, a_lo + off, b_hi + off, idesc_o, 1u);
            tc::umma_f16(tmem_base + AT_O_COL, a_hi + off, b_hi + off, idesc_o, 1u);
          }
          tc::umma_commit(sm.empty(s));
        }
        tc::umma_commit(sm.o_full());
      }
    }
  } else {  // ===================== softmax + output warps 2..5: one thread per query row =====================
    const uint32_t quad = (uint32_t)warp & 3u;
    const int r = quad * 32 + lane;                       // row of the tile = TMEM lane
    const uint32_t t_s = tmem_base + ((quad * 32u) << 16);
    const uint32_t t_o = t_s + AT_O_COL;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;  // swizzled K-major A tile
    const uint32_t rx = (uint32_t)(r & 7);
    uint32_t n = 0, it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
      const int qt = unit % p.n_qt, h = (unit / p.n_qt) % p.n_heads, b = unit / (p.n_qt * p.n_heads);
      const int qi = qt * AT_M + r;                       // query position inside the sequence
      const bool row_ok = qi < p.len_q;
      const float* mrow = p.mask + (long long)b * p.mask_batch_stride + (row_ok ? (long long)qi * p.mask_q_stride : 0);
      tc::mbar_wait(sm.s_full(), it & 1u);
      tc::fence_after_sync();
      // x = s / sqrt(dh) + (1 - m) * -10000, the reference's fp32 expression (the division is a multiplication by the
      // reciprocal: exact for head sizes 64 and 256, within 1 ulp otherwise), e = exp(x - max).  Sequences of up to 128
      // keys keep x in registers between the maximum and the exponentials (one pass over TMEM and the mask).
      float mx = -INFINITY, sum = 0.f;
      n += p.n_kc;
      if (p.len_k <= 128) {
        float x[128];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g * 32 < p.len_k) {
            uint32_t v[32];
            tc::tmem_ld_32x32(t_s + g * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float xv = -INFINITY;  // keys beyond the sequence: weight exp(-inf) = 0
              if (g * 32 + i < p.len_k) {
                const float madd = __fmul_rn(__fsub_rn(1.f, __ldg(mrow + g * 32 + i)), ATT_MASK_FILL);
                xv = __fadd_rn(__fmul_rn(__uint_as_float(v[i]), p.inv_sqrt_dh), madd);
              }
              x[g * 32 + i] = xv;
              mx = fmaxf(mx, xv);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < p.n_jc) {
            const int s = n % p.stages;
            tc::mbar_wait(sm.empty(s), ((n / p.stages) & 1u) ^ 1u);  // the MMAs that last read this stage are done
            const uint32_t a_hi = sm.stage(s) + row_off, a_lo = a_hi + AT_A_BYTES;
#pragma unroll
            for (int g = 0; g < 8; ++g) {  // 16-byte groups of 8 keys: chunk index g XOR (row & 7)
              float e[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int col = c * 64 + g * 8 + j;
                e[j] = col < p.len_k ? __expf(x[col] - mx) : 0.f;
                sum += e[j];
              }
              uint4 hi, lo;
              split8(e, p.is_bf16, hi, lo);
              const uint32_t at = (((uint32_t)g) ^ rx) * 16u;
              st_shared_v4(a_hi + at, hi);
              st_shared_v4(a_lo + at, lo);
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(sm.pfull(s));
            ++n;
          }
        }
      } else {
        // ---- pass 1: row maximum ----
        for (int j0 = 0; j0 < p.len_k; j0 += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(t_s + j0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (j0 + i < p.len_k) {
              const float madd = __fmul_rn(__fsub_rn(1.f, __ldg(mrow + j0 + i)), ATT_MASK_FILL);
              mx = fmaxf(mx, __fadd_rn(__fmul_rn(__uint_as_float(v[i]), p.inv_sqrt_dh), madd));
            }
          }
        }
        // ---- pass 2: e = exp(x - max) -> (hi, lo) halves into the A slot of the stage of each 64-key chunk ----
        for (int c = 0; c < p.n_jc; ++c, ++n) {
          const int s = n % p.stages;
          tc::mbar_wait(sm.empty(s), ((n / p.stages) & 1u) ^ 1u);
          const uint32_t a_hi = sm.stage(s) + row_off, a_lo = a_hi + AT_A_BYTES;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int j0 = c * AT_KC + half * 32;
            float e[32];
            if (j0 < p.len_k) {
              uint32_t v[32];
              tc::tmem_ld_32x32(t_s + j0, v);
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float xe = 0.f;
                if (j0 + i < p.len_k) {
                  const float madd = __fmul_rn(__fsub_rn(1.f, __ldg(mrow + j0 + i)), ATT_MASK_FILL);
                  xe = __expf(__fadd_rn(__fmul_rn(__uint_as_float(v[i]), p.inv_sqrt_dh), madd) - mx);
                  sum += xe;
                }
                e[i] = xe;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = 0.f;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 hi, lo;
              split8(e + g * 8, p.is_bf16, hi, lo);
              const uint32_t at = (((uint32_t)(half * 4 + g)) ^ rx) * 16u;
              st_shared_v4(a_hi + at, hi);
              st_shared_v4(a_lo + at, lo);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(sm.pfull(s));
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(sm.s_empty());  // S fully consumed
      // ---- output: O / sum ----
      tc::mbar_wait(sm.o_full(), it & 1u);
      tc::fence_after_sync();
      const float inv = __fdiv_rn(1.f, sum);
      const long long orow = (long long)b * p.len_q + qi;
      for (int d0 = 0; d0 < p.dh; d0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(t_o + d0, v);
        tc::tmem_ld_wait();
        if (row_ok) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __fmul_rn(__uint_as_float(v[i]), inv);
          const int col = h * p.dh + d0;
          if (p.out) {
            float* dst = p.out + orow * p.hidden + col;
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
          }
          if (p.o_hi) {
            unsigned short* dh_ = p.o_hi + orow * p.o_ld + col;
            unsigned short* dl_ = p.o_lo + orow * p.o_ld + col;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 hi, lo;
              split8(o + g * 8, p.is_bf16, hi, lo);
              *reinterpret_cast<uint4*>(dh_ + g * 8) = hi;
              *reinterpret_cast<uint4*>(dl_ + g * 8) = lo;
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(sm.o_empty());
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

extern "C" int xmlb_attention_tc(const unsigned short* q_hi, const unsigned short* q_lo, int q_ld, int q_col0,
                                 const unsigned short* k_hi, const unsigned short* k_lo, int k_ld, int k_col0,
                                 const unsigned short* vt_hi, const unsigned short* vt_lo, int vt_ld,
                                 const float* mask, long long mask_batch_stride, long long mask_q_stride, float* out,
                                 unsigned short* out_hi, unsigned short* out_lo, int out16_ld, int batch, int len_q,
                                 int len_k, int hidden, int n_heads, int is_bf16, void* stream) {
  XMLB_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && mask, "xmlb_attention_tc: null pointer");
  XMLB_REQUIRE(out || (out_hi && out_lo), "xmlb_attention_tc: no output requested");
  XMLB_REQUIRE(n_heads > 0 && hidden % n_heads == 0, "xmlb_attention_tc: hidden %% n_heads != 0");
  const int dh = hidden / n_heads;
  XMLB_REQUIRE(dh % 64 == 0 && dh <= 256, "xmlb_attention_tc: head size must be 64, 128, 192 or 256 (got %d)", dh);
  XMLB_REQUIRE(len_k >= 1 && len_k <= 256 && len_q >= 1, "xmlb_attention_tc: need 1 <= len_k <= 256");
  XMLB_REQUIRE(q_col0 >= 0 && k_col0 >= 0 && q_ld >= q_col0 + hidden && k_ld >= k_col0 + hidden && q_ld % 8 == 0 &&
                   k_ld % 8 == 0 && vt_ld % 8 == 0 && vt_ld >= 64 && vt_ld >= len_k,
               "xmlb_attention_tc: bad operand layout");
  XMLB_REQUIRE(!out || ((uintptr_t)out & 15) == 0, "xmlb_attention_tc: out must be 16-byte aligned");
  XMLB_REQUIRE(!out_hi || (out16_ld >= hidden && out16_ld % 8 == 0 && (((uintptr_t)out_hi | (uintptr_t)out_lo) & 15) == 0),
               "xmlb_attention_tc: bad split output layout");
  XMLB_REQUIRE((long long)batch * len_q < (1ll << 31) - 256 && (long long)batch * hidden < (1ll << 31),
               "xmlb_attention_tc: too many rows");
  if (batch == 0) return XMLB_OK;
  AttParams p = {};
  p.batch = batch, p.len_q = len_q, p.len_k = len_k, p.hidden = hidden, p.n_heads = n_heads, p.dh = dh;
  p.n_qt = ceil_div(len_q, AT_M);
  p.kbox = (len_k + 15) / 16 * 16;
  p.brows = p.kbox > dh ? p.kbox : dh;
  p.n_kc = dh / AT_KC, p.n_jc = ceil_div(len_k, AT_KC);
  p.q_col0 = q_col0, p.k_col0 = k_col0, p.is_bf16 = is_bf16 ? 1 : 0;
  p.mask = mask, p.mask_batch_stride = mask_batch_stride, p.mask_q_stride = mask_q_stride;
  p.out = out, p.o_hi = out_hi, p.o_lo = out_lo, p.o_ld = out16_ld;
  p.inv_sqrt_dh = 1.f / sqrtf((float)dh);
  const int stage_bytes = 2 * AT_A_BYTES + 2 * p.brows * AT_KC * 2;
  int stages = (227 * 1024 - 1024 - AT_BAR_BYTES) / stage_bytes;
  if (stages > AT_MAX_STAGES) stages = AT_MAX_STAGES;
  XMLB_REQUIRE(stages >= 2, "xmlb_attention_tc: tiles do not fit in shared memory");
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + AT_BAR_BYTES;

  AttMaps maps;
  int rc;
  const unsigned long long q_rows = (unsigned long long)batch * len_q, k_rows = (unsigned long long)batch * len_k;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.q_hi, q_hi, q_rows, q_ld, AT_M, AT_KC))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.q_lo, q_lo, q_rows, q_ld, AT_M, AT_KC))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.k_hi, k_hi, k_rows, k_ld, p.kbox, AT_KC))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.k_lo, k_lo, k_rows, k_ld, p.kbox, AT_KC))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.v_hi, vt_hi, (unsigned long long)batch * hidden, vt_ld, dh, AT_KC))) return rc;
  if ((rc = xmlb_make_tmap_2d_u16(&maps.v_lo, vt_lo, (unsigned long long)batch * hidden, vt_ld, dh, AT_KC))) return rc;

  int dev = 0, sms = 0;
  XMLB_CUDA(cudaGetDevice(&dev));
  XMLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = (long long)batch * n_heads * p.n_qt;
  const int grid = total < sms ? (int)total : sms;
  XMLB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(maps, p);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
