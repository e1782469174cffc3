// Greedy temporal NMS over ranked moment lists, one CTA per query.
//   reference utils/temporal_nms.py:25-74 (temporal_non_maximum_suppression, IoU = intersection / convex hull,
//   strict '>' threshold, at most `max_per_group` survivors per video) wrapped the way
//   baselines/clip_alignment_with_language/inference.py:189-265 wraps it: group the first n ranked predictions by
//   video, NMS per group, stable re-sort of all survivors by score, keep max_after_nms.
// Input lists must already be ranked by score descending (they come from xmlb_span_topk).  Output = indices into
// the input list, ranked (score desc; exact ties: group of earlier first appearance first, then list order --
// the order Python's stable sorted() gives the reference).
#include "common.cuh"
#include "xmlb200.h"

namespace {
constexpr int NMS_T = 256;
constexpr int NMS_MAX = 4096;  // ranked spans per query (12-bit fields of the sort key; 116 KB of shared memory)
constexpr int NMS_BITS = 12;

template <int CAP>  // capacity in spans: 1024 (29 KB, several CTAs per SM) or NMS_MAX
struct NmsSmem {
  int vid[CAP];
  float st[CAP], ed[CAP], score[CAP];
  short leader[CAP];
  short kept_in_group[CAP];
  unsigned char state[CAP];  // 0 = pending, 1 = kept, 2 = suppressed/dropped
  unsigned long long key[CAP];
};

__device__ __forceinline__ unsigned int pos_float_key(float f) { return __float_as_uint(f) | 0x80000000u; }

template <int CAP>
__global__ void __launch_bounds__(NMS_T) temporal_nms_kernel(const int* __restrict__ video_idx,
                                                             const float* __restrict__ st, const float* __restrict__ ed,
                                                             const float* __restrict__ score,
                                                             const int* __restrict__ n_valid, int n_in, double thd,
                                                             int max_per_group, int max_out, int* __restrict__ out_idx,
                                                             int* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem<CAP>& sm = *reinterpret_cast<NmsSmem<CAP>*>(smem_raw);
  const long long q = blockIdx.x;
  const int t = threadIdx.x;
  const int n = n_valid ? min(n_valid[q], n_in) : n_in;
  for (int i = t; i < n; i += NMS_T) {
    sm.vid[i] = video_idx ? video_idx[q * n_in + i] : 0;
    sm.st[i] = st[q * n_in + i];
    sm.ed[i] = ed[q * n_in + i];
    sm.score[i] = score[q * n_in + i];
    sm.state[i] = 0;
    sm.kept_in_group[i] = 0;
  }
  __syncthreads();
  for (int i = t; i < n; i += NMS_T) {
    int l = i;
    const int v = sm.vid[i];
    for (int j = 0; j < i; ++j)
      if (sm.vid[j] == v) {
        l = j;
        break;
      }
    sm.leader[i] = (short)l;
  }
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (sm.state[i] != 0) continue;  // uniform: state[i] is final once every earlier head has been processed
    const int g = sm.leader[i];
    const bool room = sm.kept_in_group[g] < max_per_group;
    __syncthreads();
    if (t == 0) {
      sm.state[i] = room ? 1 : 2;
      if (room) sm.kept_in_group[g] += 1;
    }
    if (room) {
      const double s0 = sm.st[i], e0 = sm.ed[i];
      for (int j = i + 1 + t; j < n; j += NMS_T) {
        if (sm.state[j] == 0 && sm.leader[j] == g) {
          const double s1 = sm.st[j], e1 = sm.ed[j];
          const double inter = fmax(0.0, fmin(e0, e1) - fmax(s0, s1));
          const double hull = fmax(e0, e1) - fmin(s0, s1);
          const double iou = hull == 0.0 ? 0.0 : 1.0 * inter / hull;
          if (iou > thd) sm.state[j] = 2;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  int pow2 = 1;
  while (pow2 < n) pow2 <<= 1;
  for (int i = t; i < pow2; i += NMS_T) {
    unsigned long long k = 0ull;
    if (i < n && sm.state[i] == 1)
      k = ((unsigned long long)pos_float_key(sm.score[i]) << 32) | ((unsigned long long)(NMS_MAX - 1 - sm.leader[i]) << NMS_BITS) |
          (unsigned long long)(NMS_MAX - 1 - i);
    sm.key[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < pow2; i += NMS_T) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool desc = (i & k) == 0;
          const unsigned long long a = sm.key[i], b = sm.key[ixj];
          if (desc ? (a < b) : (a > b)) sm.key[i] = b, sm.key[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
  int cnt = 0;
  for (int i = t; i < max_out; i += NMS_T) {
    const bool ok = i < pow2 && sm.key[i] != 0ull;
    out_idx[q * max_out + i] = ok ? NMS_MAX - 1 - (int)(sm.key[i] & (unsigned long long)(NMS_MAX - 1)) : -1;
    cnt += ok;
  }
  cnt = warp_sum_int(cnt);
  __shared__ int total;
  if (t == 0) total = 0;
  __syncthreads();
  if ((t & 31) == 0) atomicAdd(&total, cnt);
  __syncthreads();
  if (t == 0) out_count[q] = total;
}
}  // namespace

extern "C" int xmlb_temporal_nms(const int* video_idx, const float* st, const float* ed, const float* score,
                                 const int* n_valid, int n_queries, int n_in, double iou_thd, int max_per_group,
                                 int max_out, int* out_idx, int* out_count, void* stream) {
  XMLB_REQUIRE(st && ed && score && out_idx && out_count, "xmlb_temporal_nms: null pointer");
  XMLB_REQUIRE(n_in >= 1 && n_in <= NMS_MAX, "xmlb_temporal_nms: n_in must be in [1, %d]", NMS_MAX);
  XMLB_REQUIRE(max_out >= 1 && max_per_group >= 1, "xmlb_temporal_nms: bad limits");
  if (n_queries == 0) return XMLB_OK;
  if (n_in <= 1024) {
    XMLB_CUDA(cudaFuncSetAttribute(temporal_nms_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(NmsSmem<1024>)));
    temporal_nms_kernel<1024><<<n_queries, NMS_T, sizeof(NmsSmem<1024>), (cudaStream_t)stream>>>(
        video_idx, st, ed, score, n_valid, n_in, iou_thd, max_per_group, max_out, out_idx, out_count);
  } else {
    XMLB_CUDA(cudaFuncSetAttribute(temporal_nms_kernel<NMS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(NmsSmem<NMS_MAX>)));
    temporal_nms_kernel<NMS_MAX><<<n_queries, NMS_T, sizeof(NmsSmem<NMS_MAX>), (cudaStream_t)stream>>>(
        video_idx, st, ed, score, n_valid, n_in, iou_thd, max_per_group, max_out, out_idx, out_count);
  }
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
