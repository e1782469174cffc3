// Epilogue shared by the packed video-retrieval kernels (vr_scores_tc.cu, vr_filter_pair.cu): per query row, the
// masked max over each video's clips of one 128 x 256 fp32 accumulator tile, read from TMEM 32 columns at a time.
#pragma once
#include "tc_common.cuh"

namespace vr {

constexpr int MAX_TILE_VIDEOS = 32;  // packed layout: at most this many videos share a tile
constexpr int PACKED_EXTRA_SMEM = 128 * (MAX_TILE_VIDEOS + 1) * 4;  // float [128][MAX_TILE_VIDEOS + 1]

__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 16]));
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) m[i] = fmaxf(m[i], m[i + w]);
  return m[0];
}
__device__ __forceinline__ float masked_max32(const uint32_t (&r)[32], unsigned int mask) {
  float m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float a = (mask >> i) & 1u ? __uint_as_float(r[i]) : MASK_FILL;
    const float b = (mask >> (i + 16)) & 1u ? __uint_as_float(r[i + 16]) : MASK_FILL;
    m[i] = fmaxf(a, b);
  }
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) m[i] = fmaxf(m[i], m[i + w]);
  return m[0];
}

// One unit = one modality of one tile.  taddr: this warp's lane quadrant of the accumulator; used: columns in use;
// starts: 8 x 32-bit map of the columns where a video starts; out_row: this query's score row at the tile's first
// packed ordinal; my_best: shared-memory row where the first modality's maxima are parked (nothing on the
// accumulator-release path waits on global memory).  release() hands the accumulator back once it is fully read.
template <class Release>
__device__ __forceinline__ void packed_epilogue(uint32_t taddr, int used, const unsigned int* __restrict__ starts,
                                                float* __restrict__ out_row, bool q_ok, int mod, int n_mod,
                                                float divisor, uint32_t my_best, Release release) {
  const bool last = mod == n_mod - 1;
  int j = -1;
  float cur = MASK_FILL;
  auto flush = [&]() {
    if (j < 0) return;
    if (!last) {
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_best + 4u * j), "f"(cur) : "memory");
    } else {
      float v = cur;
      if (mod != 0) {
        float first;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(first) : "r"(my_best + 4u * j) : "memory");
        v = __fadd_rn(first, cur);
      }
      if (q_ok) out_row[j] = __fdiv_rn(v, divisor);
    }
  };
  for (int c = 0; c < 8 && c * 32 < used; ++c) {  // warp-uniform
    uint32_t r[32];
    tc::tmem_ld_32x32(taddr + c * 32, r);
    tc::tmem_ld_wait();
    const int n_here = min(32, used - c * 32);
    const unsigned int valid = n_here == 32 ? 0xffffffffu : (1u << n_here) - 1u;
    unsigned int sb = __ldg(starts + c) & valid;
    if (sb == 0u && n_here == 32) {  // the whole chunk continues the current video: branch-free max tree
      cur = fmaxf(cur, max32(r));
    } else {  // walk the video segments of this chunk (all conditions are uniform over the CTA)
      int pos = 0;
      while (true) {
        const int nxt = sb ? __ffs(sb) - 1 : n_here;  // next video start, or end of the used columns
        if (nxt > pos) {
          const unsigned int seg = (nxt == 32 ? 0xffffffffu : (1u << nxt) - 1u) & ~((1u << pos) - 1u);
          cur = fmaxf(cur, masked_max32(r, seg));
        }
        if (nxt >= n_here) break;
        flush();  // a new video starts at column nxt
        ++j;
        cur = MASK_FILL;
        pos = nxt;
        sb &= sb - 1u;
      }
    }
  }
  release();  // all TMEM reads of this accumulator are done
  flush();
}

}  // namespace vr
