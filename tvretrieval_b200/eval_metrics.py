"""Vectorised TVR retrieval metrics: drop-in for `standalone_eval.eval.eval_retrieval` / `eval_by_task_type`
(reference standalone_eval/eval.py:83-276), which the reference's per-epoch evaluation calls on the prediction lists
(baselines/crossmodal_moment_localization/inference.py:489-500).

The reference loops over queries in Python (one small numpy program per query, plus per-query list comprehensions for
SVMR); here every task is a handful of array operations over (n_queries, n_predictions) matrices.  Same inputs, same
output dictionaries (keys, order, rounding), same float32 IoU arithmetic: R@K for K in (1, 5, 10, 100) at the given
temporal IoU thresholds, overall and by description type.  Host-side numpy: the
prediction lists arrive as Python objects, so this is the vectorisation that matters.
"""
from collections import OrderedDict

import numpy as np

TASK_TYPES = OrderedDict([("VCMR", "Video Corpus Moment Retrieval"), ("SVMR", "Single Video Moment Retrieval"),
                          ("VR", "regular Video Retrieval")])
DESC_TYPE2IDX = {"v": 0, "t": 1, "vt": 2}


def get_rounded_percentage(float_number, n_floats=2):
    return round(float_number * 100, n_floats)


def compute_temporal_iou_batch(preds, gt):
    """reference eval.py:54-69: intersection over the HULL of the two spans (not the true union), 0 where empty.
    preds (..., 2), gt broadcastable to it; float32 in, float32 out."""
    inter = np.maximum(0, np.minimum(preds[..., 1], gt[..., 1]) - np.maximum(preds[..., 0], gt[..., 0]))
    hull = np.maximum(preds[..., 1], gt[..., 1]) - np.minimum(preds[..., 0], gt[..., 0])
    return np.divide(inter, hull, out=np.zeros_like(inter), where=hull != 0)


def _prediction_matrix(pred_lists, max_pred):
    """list of [[vid, st, ed, ...], ...] -> (n, P, 3) float32 zero-padded + (n, P) validity."""
    counts = np.asarray([min(len(p), max_pred) for p in pred_lists], dtype=np.int64)
    width = int(counts.max()) if len(counts) else 0
    out = np.zeros((len(pred_lists), width, 3), dtype=np.float32)
    for i, p in enumerate(pred_lists):
        if counts[i]:
            out[i, :counts[i]] = np.asarray([e[:3] for e in p[:counts[i]]], dtype=np.float32)
    return out, np.arange(width)[None, :] < counts[:, None]


def _recall(hit, topks, prefix, select=None):
    """hit (n, P) bool ranked left to right -> {prefix + "r<k>": percentage of rows with a hit in the first k}."""
    out = OrderedDict()
    for k in topks:
        any_k = hit[:, :k].any(axis=1)
        if select is None:
            out["{}r{}".format(prefix, k)] = get_rounded_percentage(np.mean(any_k))
        else:
            with np.errstate(invalid="ignore", divide="ignore"):  # a type without queries gives nan, as the reference
                out["{}r{}".format(prefix, k)] = get_rounded_percentage(
                    1.0 * np.sum(np.logical_and(any_k, select[0])) / select[1])
    return out


def eval_by_task_type(moment_predictions, video2idx, ground_truth, iou_thds=(0.5, 0.7), recall_topks=(1, 5, 10, 100),
                      task_type="SVMR", max_pred_per_query=100, match_number=True, verbose=True, use_desc_type=True):
    """reference eval.py:83-252 -> (metrics, metrics_by_type)."""
    assert task_type in TASK_TYPES, "task_type must be one of {}".format(list(TASK_TYPES.keys()))
    if verbose:
        print("Running evaluation with task_type {}, n results {}; n gt {}".format(task_type, len(moment_predictions),
                                                                                   len(ground_truth)))
    pred_by_id = {e["desc_id"]: e for e in moment_predictions}
    gt_by_id = {e["desc_id"]: e for e in ground_truth}
    if match_number:
        assert set(gt_by_id.keys()) == set(pred_by_id.keys()), "desc_ids in predictions and ground_truth must match"
    items = [(k, g) for k, g in gt_by_id.items() if match_number or k in pred_by_id]
    preds, valid = _prediction_matrix([pred_by_id[k]["predictions"] for k, _ in items], max_pred_per_query)
    # the reference compares the float32 video column with the python int id (eval.py:145)
    gt_vid = np.asarray([video2idx[g["vid_name"]] for _, g in items], dtype=np.float64)
    vid_matched = (preds[..., 0].astype(np.float64) == gt_vid[:, None]) & valid

    # ground-truth spans: one [st, ed] per query (TVR), or >= 4 annotations of which >= 2 must overlap (DiDeMo)
    n_ts = np.asarray([len(g["ts"]) if len(g["ts"]) >= 4 else 1 for _, g in items], dtype=np.int64)
    ts = np.zeros((len(items), int(n_ts.max()) if len(items) else 1, 2), dtype=np.float32)
    for i, (_, g) in enumerate(items):
        ts[i, :n_ts[i]] = np.asarray(g["ts"], dtype=np.float32).reshape(-1, 2)[:n_ts[i]]
    ts_valid = np.arange(ts.shape[1])[None, :] < n_ts[:, None]
    iou = compute_temporal_iou_batch(preds[:, :, None, 1:3], ts[:, None, :, :])      # (n, P, T)
    iou = iou * vid_matched[:, :, None].astype(np.float32)   # wrong-video predictions score 0 (eval.py:157,169)
    need = np.where(n_ts >= 4, 2, 1)[:, None]
    corrects = [(((iou >= thd) & ts_valid[:, None, :]).sum(-1) >= need) & valid for thd in iou_thds]

    desc_types = np.asarray([DESC_TYPE2IDX[g["type"]] for _, g in items]) if use_desc_type else None
    metrics, metrics_by_type = OrderedDict(), OrderedDict()

    def by_type(hits_per_prefix):
        for desc_type, t in DESC_TYPE2IDX.items():
            sel = desc_types == t
            for prefix, hit in hits_per_prefix:
                metrics_by_type.update(_recall(hit, recall_topks, "{}-{}".format(desc_type, prefix), (sel, np.sum(sel))))

    if task_type == "VCMR":
        hits = [("{}-".format(thd), c) for thd, c in zip(iou_thds, corrects)]
    elif task_type == "SVMR":
        # only the predictions on the ground-truth video count, in their own ranking (eval.py:216-218):
        # rank among the matched predictions = running count of matches
        rank = np.cumsum(vid_matched, axis=1) - 1
        hits = []
        for thd, c in zip(iou_thds, corrects):
            compact = np.zeros_like(c)
            rows, cols = np.nonzero(vid_matched)
            compact[rows, rank[rows, cols]] = c[rows, cols]
            hits.append(("{}-".format(thd), compact))
    else:  # VR
        hits = [("", vid_matched)]
    for prefix, hit in hits:
        metrics.update(_recall(hit, recall_topks, prefix))
    if use_desc_type:
        by_type(hits)
        metrics_by_type["desc_type_ratio"] = "v {} t {} vt {}".format(
            *[get_rounded_percentage(1.0 * np.sum(desc_types == DESC_TYPE2IDX[k]) / len(desc_types))
              for k in ["v", "t", "vt"]])
    return metrics, metrics_by_type


def eval_retrieval(submission, ground_truth, iou_thds=(0.5, 0.7), verbose=True, match_number=True,
                   use_desc_type=True):
    """reference eval.py:255-276 -> OrderedDict {task: metrics, ..., task + "_by_type": metrics_by_type, ...}."""
    video2idx = submission["video2idx"]
    submitted = [k for k in TASK_TYPES if k in submission]
    if verbose:
        print("Evaluating for task {}".format(submitted))
    raw = {}
    for task in submitted:
        raw[task], raw[task + "_by_type"] = eval_by_task_type(
            submission[task], video2idx, ground_truth, iou_thds=iou_thds, recall_topks=(1, 5, 10, 100),
            task_type=task, max_pred_per_query=100, match_number=match_number, verbose=verbose,
            use_desc_type=use_desc_type)
    out = OrderedDict((task, raw[task]) for task in submitted)
    if use_desc_type:
        for task in submitted:
            out[task + "_by_type"] = raw[task + "_by_type"]
    return out


# ---------------------------------------------------------------------------------------------------------------
# device path: metrics straight from the search engine's device tensors (no prediction lists, no host loop)
# ---------------------------------------------------------------------------------------------------------------
def eval_search_result_device(result, gt_video_pos, gt_ts, ctx_len, clip_length=1.5, desc_types=None,
                              iou_thds=(0.5, 0.7), recall_topks=(1, 5, 10, 100), max_pred_per_query=100, tasks=None):
    """eval_retrieval for an engine.SearchResult that still lives on the GPU: the same metric dictionaries (keys, order,
    rounding) the reference evaluator gives for the submission built from that result, computed by xmlb_eval_first_hit.
    gt_video_pos (Nq,) int: corpus position (index into video_metas) of each query's ground-truth video; gt_ts
    (Nq, 2) float seconds; desc_types (Nq,) ints 0 / 1 / 2 = "v" / "t" / "vt" (None: no by-type metrics).
    -> OrderedDict like eval_retrieval's (tasks present in `result`, or the given subset)."""
    import torch
    from . import _lib
    dev = result.top_video_idx.device if result.top_video_idx is not None else result.svmr_flat_idx.device
    gt_vid = torch.as_tensor(gt_video_pos, device=dev).to(torch.int32).contiguous()
    ts = torch.as_tensor(gt_ts, device=dev, dtype=torch.float32)
    gs, ge = ts[:, 0].contiguous(), ts[:, 1].contiguous()
    thds = torch.tensor(list(iou_thds), device=dev, dtype=torch.float32)
    nq = len(gt_vid)
    stream = torch.cuda.current_stream().cuda_stream
    types = None if desc_types is None else np.asarray(desc_types)

    def first_hit(mode, vid, st=None, ed=None):
        n_thd = 1 if mode == 2 else len(iou_thds)
        vid = vid[:, :max_pred_per_query].to(torch.int32).contiguous()
        st = None if st is None else st[:, :max_pred_per_query].contiguous()
        ed = None if ed is None else ed[:, :max_pred_per_query].contiguous()
        out = torch.empty(nq, n_thd, device=dev, dtype=torch.int32)
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        rc = _lib.lib().xmlb_eval_first_hit(p(vid), p(st), p(ed), None, p(gt_vid), p(gs), p(ge), p(thds), n_thd, nq,
                                            vid.shape[1], mode, p(out), stream)
        _lib.check(rc, "xmlb_eval_first_hit")
        return out.cpu().numpy()

    def metrics_of(fh, prefixes):
        overall, by_type = OrderedDict(), OrderedDict()
        for j, prefix in enumerate(prefixes):
            for k in recall_topks:
                overall["{}r{}".format(prefix, k)] = get_rounded_percentage(np.mean(fh[:, j] < k))
        if types is not None:
            for name, t in DESC_TYPE2IDX.items():
                sel = types == t
                for j, prefix in enumerate(prefixes):
                    for k in recall_topks:
                        with np.errstate(invalid="ignore", divide="ignore"):
                            by_type["{}-{}r{}".format(name, prefix, k)] = get_rounded_percentage(
                                1.0 * np.sum((fh[:, j] < k) & sel) / np.sum(sel))
            by_type["desc_type_ratio"] = "v {} t {} vt {}".format(
                *[get_rounded_percentage(1.0 * np.sum(types == DESC_TYPE2IDX[k]) / len(types)) for k in ["v", "t", "vt"]])
        return overall, by_type

    raw = OrderedDict()
    span_prefixes = ["{}-".format(t) for t in iou_thds]
    cells = ctx_len * ctx_len
    want = lambda t: tasks is None or t in tasks  # noqa: E731
    if result.span_flat_idx is not None and want("VCMR"):
        flat = result.span_flat_idx.long()
        rank, rem = flat // cells, flat % cells
        vid = torch.gather(result.top_video_idx.long(), 1, rank)
        st = (rem // ctx_len).float() * clip_length
        ed = (rem % ctx_len).float() * clip_length + clip_length
        raw["VCMR"] = metrics_of(first_hit(0, vid, st, ed), span_prefixes)
    if result.svmr_flat_idx is not None and want("SVMR"):
        flat = result.svmr_flat_idx.long()
        st = (flat // ctx_len).float() * clip_length
        ed = ((flat % ctx_len) + 1).float() * clip_length
        vid = gt_vid.view(-1, 1).expand_as(flat)  # SVMR predictions are made on the ground-truth video
        raw["SVMR"] = metrics_of(first_hit(1, vid, st, ed), span_prefixes)
    if result.top_video_idx is not None and want("VR"):
        raw["VR"] = metrics_of(first_hit(2, result.top_video_idx), [""])
    out = OrderedDict((t, raw[t][0]) for t in TASK_TYPES if t in raw)
    if types is not None:
        for t in TASK_TYPES:
            if t in raw:
                out[t + "_by_type"] = raw[t][1]
    return out
