// Retrieval metrics on the device: the per-query part of reference standalone_eval/eval.py:83-252
// (eval_by_task_type) for ranked prediction lists that are already device tensors (the output of the search engine),
// so that a per-epoch evaluation during training (reference train.py:181-183 -> inference.py:472-531) needs neither
// the Python prediction lists nor a host loop.
//   first_hit[q][t] = rank of the first prediction of query q that is "correct" at IoU threshold t (INT_MAX if none):
//     VCMR (mode 0): right video and IoU(pred span, ground-truth span) >= thd                    eval.py:145-170
//     SVMR (mode 1): same, but ranks count only the predictions on the ground-truth video        eval.py:216-218
//     VR   (mode 2): right video (thresholds ignored, n_thds must be 1)                          eval.py:232-236
//   IoU = intersection / convex hull in fp32 (eval.py:54-69), 0 for an empty hull.
// R@K is then mean(first_hit < K): a reduction over n_queries ints that the caller does (eval_metrics.py).
// One warp per query; lanes stride over the ranked predictions, chunks of 32 in rank order.
#include "common.cuh"
#include "xmlb200.h"

namespace {
constexpr int MAX_THDS = 8;

__global__ void __launch_bounds__(256) eval_first_hit_kernel(const int* __restrict__ pred_vid,
                                                             const float* __restrict__ pred_st,
                                                             const float* __restrict__ pred_ed,
                                                             const int* __restrict__ n_valid,
                                                             const int* __restrict__ gt_vid,
                                                             const float* __restrict__ gt_st,
                                                             const float* __restrict__ gt_ed, const float* __restrict__ thds,
                                                             int n_thds, int n_queries, int n_pred, int mode,
                                                             int* __restrict__ first_hit) {
  const int lane = threadIdx.x & 31;
  const long long q = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int n = n_valid ? min(n_valid[q], n_pred) : n_pred;
  const int gv = gt_vid[q];
  const float gs = mode == 2 ? 0.f : gt_st[q], ge = mode == 2 ? 0.f : gt_ed[q];
  int best[MAX_THDS];
#pragma unroll
  for (int t = 0; t < MAX_THDS; ++t) best[t] = 0x7fffffff;
  int matched_before = 0;  // SVMR: predictions on the ground-truth video seen in earlier chunks
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool in = i < n;
    const bool vid_ok = in && pred_vid[q * n_pred + i] == gv;
    const unsigned int mbal = __ballot_sync(0xffffffffu, vid_ok);
    const int rank = mode == 1 ? matched_before + __popc(mbal & ((1u << lane) - 1u)) : i;
    float iou = 0.f;
    if (vid_ok && mode != 2) {
      const float ps = pred_st[q * n_pred + i], pe = pred_ed[q * n_pred + i];
      const float inter = fmaxf(0.f, __fsub_rn(fminf(pe, ge), fmaxf(ps, gs)));
      const float hull = __fsub_rn(fmaxf(pe, ge), fminf(ps, gs));
      iou = hull != 0.f ? __fdiv_rn(inter, hull) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < MAX_THDS; ++t) {
      if (t < n_thds) {
        const bool hit = vid_ok && (mode == 2 || iou >= thds[t]);
        const unsigned int hbal = __ballot_sync(0xffffffffu, hit);
        if (hbal && best[t] == 0x7fffffff) {
          const int first_lane = __ffs(hbal) - 1;
          best[t] = __shfl_sync(0xffffffffu, rank, first_lane);
        }
      }
    }
    matched_before += __popc(mbal);
  }
  if (lane == 0)
    for (int t = 0; t < n_thds; ++t) first_hit[q * n_thds + t] = best[t];
}
}  // namespace

extern "C" int xmlb_eval_first_hit(const int* pred_vid, const float* pred_st, const float* pred_ed, const int* n_valid,
                                   const int* gt_vid, const float* gt_st, const float* gt_ed, const float* iou_thds,
                                   int n_thds, int n_queries, int n_pred, int mode, int* first_hit, void* stream) {
  XMLB_REQUIRE(pred_vid && gt_vid && first_hit, "xmlb_eval_first_hit: null pointer");
  XMLB_REQUIRE(mode >= 0 && mode <= 2, "xmlb_eval_first_hit: mode must be 0 (VCMR), 1 (SVMR) or 2 (VR)");
  XMLB_REQUIRE(mode == 2 ? n_thds == 1 : (pred_st && pred_ed && gt_st && gt_ed && iou_thds && n_thds >= 1 && n_thds <= MAX_THDS),
               "xmlb_eval_first_hit: spans and 1..%d thresholds are required for VCMR / SVMR (n_thds = 1 for VR)", MAX_THDS);
  XMLB_REQUIRE(n_pred >= 1, "xmlb_eval_first_hit: n_pred must be >= 1");
  if (n_queries == 0) return XMLB_OK;
  eval_first_hit_kernel<<<ceil_div(n_queries, 8), 256, 0, (cudaStream_t)stream>>>(
      pred_vid, pred_st, pred_ed, n_valid, gt_vid, gt_st, gt_ed, iou_thds, n_thds, n_queries, n_pred, mode, first_hit);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
