"""Drop-in drivers with the signatures of reference baselines/crossmodal_moment_localization/inference.py:
`compute_context_info` (:32), `compute_query2ctx_info` (:252), `compute_query2ctx_info_svmr_only` (:107),
`generate_min_max_length_mask` (:170), `get_svmr_res_from_st_ed_probs` (:195), `get_eval_res` (:448), plus the
post-processing helpers it imports (`get_submission_top_n`, `post_processing_vcmr_nms`,
`post_processing_svmr_nms`; reference baselines/clip_alignment_with_language/inference.py:503,228,247).

Same inputs (model, duck-typed dataset, `opt`) and same outputs (lists of
{"desc_id", "desc", "predictions": [[video_idx, st_sec, ed_sec, score], ...]}), but the tensor work runs on the
xmlb200 kernels through `engine.VCMRSearcher`, and the host section is vectorised.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from . import ops
from .engine import CorpusIndex, VCMRSearcher


# ------------------------------------------------------------------------------------------------ collate
def pad_sequences_1d(sequences, dtype=torch.float32, device=torch.device("cpu"), fixed_length=None):
    """reference utils/tensor_utils.py:5-53 (torch branch): zero-pad to the batch max, float {0,1} mask."""
    if isinstance(sequences[0], (list, np.ndarray)):
        sequences = [torch.as_tensor(s, dtype=dtype, device=device) for s in sequences]
    lengths = [len(s) for s in sequences]
    width = fixed_length if fixed_length is not None else max(lengths)
    padded = torch.zeros((len(sequences), width) + tuple(sequences[0].shape[1:]), dtype=dtype, device=device)
    mask = torch.zeros((len(sequences), width), dtype=torch.float32, device=device)
    for i, (s, n) in enumerate(zip(sequences, lengths)):
        padded[i, :n] = s
        mask[i, :n] = 1
    return padded, mask


def start_end_collate(batch):
    """reference start_end_dataset.py:346-359."""
    metas = [e["meta"] for e in batch]
    out = {}
    for k in batch[0]["model_inputs"].keys():
        if "feat" in k:
            out[k] = pad_sequences_1d([e["model_inputs"][k] for e in batch], dtype=torch.float32)
    if "st_ed_indices" in batch[0]["model_inputs"]:
        out["st_ed_indices"] = torch.stack([e["model_inputs"]["st_ed_indices"] for e in batch], dim=0)
    return metas, out


def prepare_batch_inputs(batched_model_inputs, device, non_blocking=False):
    """reference start_end_dataset.py:362-370: (`x_feat`: (padded, mask)) -> `x_feat`, `x_mask` on device."""
    out = {}
    for k, v in batched_model_inputs.items():
        if "feat" in k:
            out[k] = v[0].to(device, non_blocking=non_blocking)
            out[k.replace("feat", "mask")] = v[1].to(device, non_blocking=non_blocking)
        else:
            out[k] = v.to(device, non_blocking=non_blocking)
    return out


def _loader(dataset, batch_size, opt):
    return DataLoader(dataset, collate_fn=start_end_collate, batch_size=batch_size,
                      num_workers=getattr(opt, "num_workers", 0), shuffle=False,
                      pin_memory=getattr(opt, "pin_memory", False))


# ------------------------------------------------------------------------------------------------ corpus
def cat_tensor(tensor_list):
    """reference inference.py:71-87: zero-pad each batch to the global max length, concatenate."""
    if len(tensor_list) == 0:
        return None
    if tensor_list[0].dim() not in (2, 3):
        raise ValueError("Only support 2/3 dimensional tensors")
    width = max(t.shape[1] for t in tensor_list)
    total = sum(t.shape[0] for t in tensor_list)
    out = tensor_list[0].new_zeros((total, width) + tuple(tensor_list[0].shape[2:]))
    row = 0
    for t in tensor_list:
        out[row:row + t.shape[0], :t.shape[1]] = t
        row += t.shape[0]
    return out


def compute_context_info(model, eval_dataset, opt):
    """reference inference.py:32-97.  Batches of `opt.eval_context_bsz` videos in dataset order, each padded to
    its own max length (the padded rows are part of the function)."""
    model.eval()
    eval_dataset.set_data_mode("context")
    metas, acc = [], {k: [] for k in ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2",
                                      "sub_mask")}
    with torch.no_grad():
        for batch_metas, batch in _loader(eval_dataset, opt.eval_context_bsz, opt):
            metas.extend(batch_metas)
            x = prepare_batch_inputs(batch, device=opt.device, non_blocking=getattr(opt, "pin_memory", False))
            v1, v2, s1, s2 = model.encode_context(x["video_feat"], x["video_mask"], x["sub_feat"], x["sub_mask"])
            if "video" in opt.ctx_mode:
                acc["video_feat1"].append(v1), acc["video_feat2"].append(v2), acc["video_mask"].append(x["video_mask"])
            if "sub" in opt.ctx_mode:
                acc["sub_feat1"].append(s1), acc["sub_feat2"].append(s2), acc["sub_mask"].append(x["sub_mask"])
    return dict(video_metas=metas, **{k: cat_tensor(v) for k, v in acc.items()})


# ------------------------------------------------------------------------------------------------ queries
def generate_min_max_length_mask(array_shape, min_l, max_l):
    """reference inference.py:170-192: (1, ..., 1, L, L) float mask, 1 where min_l <= n - m < max_l."""
    length = array_shape[-1]
    d = np.arange(length)[None, :] - np.arange(length)[:, None]
    band = ((d >= min_l) & (d < max_l)).astype(np.float32)
    return band.reshape((1,) * (len(array_shape) - 2) + band.shape)


def _gather_queries(eval_dataset, opt, with_gt):
    """Runs the dataset in query mode through the same collate as the reference and returns the metas plus
    all queries padded to one (Nq, Lq, Dq) pinned host tensor (+ mask)."""
    eval_dataset.set_data_mode("query")
    eval_dataset.load_gt_vid_name_for_query(with_gt)
    metas, feats, masks = [], [], []
    for batch_metas, batch in _loader(eval_dataset, opt.eval_query_bsz, opt):
        metas.extend(batch_metas)
        feats.append(batch["query_feat"][0])
        masks.append(batch["query_feat"][1])
        if getattr(opt, "debug", False):
            break
    feat, mask = cat_tensor(feats), cat_tensor(masks)
    if torch.cuda.is_available():
        feat, mask = feat.pin_memory(), mask.pin_memory()
    return metas, feat, mask


def _searcher(model, opt, ctx_info, max_before_nms, max_n_videos):
    cache = ctx_info.get("_xmlb_index")
    if cache is None:
        cache = CorpusIndex.from_ctx_info(ctx_info)
        ctx_info["_xmlb_index"] = cache
    return VCMRSearcher(model, cache, q2c_alpha=opt.q2c_alpha, min_pred_l=opt.min_pred_l, max_pred_l=opt.max_pred_l,
                        max_n_videos=max_n_videos, max_before_nms=max_before_nms)


def _rows_to_predictions(video_idx, st_sec, ed_sec, score):
    """(..., K) arrays -> nested lists of [int, float, float, float] rows like the reference's float() boxing
    (inference.py:436-438); the boxing runs in numpy's C loops (object arrays), not in a Python loop per row."""
    rows = np.empty(np.shape(video_idx) + (4,), dtype=object)
    rows[..., 0] = np.asarray(video_idx, dtype=np.int64).astype(object)
    rows[..., 1] = np.asarray(st_sec, dtype=np.float64).astype(object)  # float32 -> double, exactly like float()
    rows[..., 2] = np.asarray(ed_sec, dtype=np.float64).astype(object)
    rows[..., 3] = np.asarray(score, dtype=np.float64).astype(object)
    return rows.tolist()


def _svmr_predictions(flat_idx, score, query_metas, video2idx, ctx_len, clip_length):
    """Host section of get_svmr_res_from_st_ed_probs, reference inference.py:225-240: ed index += 1, x clip."""
    st_i, ed_i = np.divmod(flat_idx.astype(np.int64), ctx_len)
    st_sec = st_i.astype(np.float32) * clip_length
    ed_sec = (ed_i + 1).astype(np.float32) * clip_length
    res = []
    for i, m in enumerate(query_metas):
        vid = np.full(flat_idx.shape[1], video2idx[m["vid_name"]], dtype=np.int64)
        res.append(dict(desc_id=m["desc_id"], desc=m["desc"],
                        predictions=_rows_to_predictions(vid, st_sec[i], ed_sec[i], score[i])))
    return res


def get_svmr_res_from_st_ed_probs(svmr_gt_st_probs, svmr_gt_ed_probs, query_metas, video2idx, clip_length,
                                  min_pred_l, max_pred_l, max_before_nms, device="cuda"):
    """reference inference.py:195-241, numpy (Nq, L) probabilities in; the outer product / band mask / per-query
    full argsort is replaced by the band-limited top-k kernel."""
    st = torch.as_tensor(svmr_gt_st_probs, dtype=torch.float32, device=device).unsqueeze(1)
    ed = torch.as_tensor(svmr_gt_ed_probs, dtype=torch.float32, device=device).unsqueeze(1)
    idx, score = ops.span_topk(st, ed, None, min_pred_l, max_pred_l, max_before_nms, tie_desc=True)
    return _svmr_predictions(idx.cpu().numpy(), score.cpu().numpy(), query_metas, video2idx, st.shape[-1],
                             clip_length)


def load_external_vr_res2(external_vr_res_path, top_n_vr_videos=5):
    """reference inference.py:244-249: desc_id -> top retrieved [video_idx, 0, 0, score] rows of a VR submission."""
    import json
    with open(external_vr_res_path) as fh:
        external_vr_res = json.load(fh)
    external_vr_res = get_submission_top_n(external_vr_res, top_n=top_n_vr_videos)["VR"]
    return {e["desc_id"]: e["predictions"] for e in external_vr_res}


def compute_query2ctx_info(model, eval_dataset, opt, ctx_info, max_before_nms=1000, max_n_videos=100,
                           tasks=("SVMR",)):
    """reference inference.py:252-445 -> {"SVMR"/"VCMR"/"VR": [...]} (keys with empty results dropped)."""
    is_svmr, is_vr, is_vcmr = "SVMR" in tasks, "VR" in tasks, "VCMR" in tasks
    video2idx = eval_dataset.video2idx
    video_metas = ctx_info["video_metas"]
    model.eval()
    query_metas, qfeat, qmask = _gather_queries(eval_dataset, opt, is_svmr)
    external = None
    if getattr(opt, "external_inference_vr_res_path", None) is not None:
        # reference inference.py:264-273,349-355: the top videos of every query come from another system's VR
        # submission (dataset video ids -> positions in video_metas; raw scores, exponentiated by the search)
        query2video = load_external_vr_res2(opt.external_inference_vr_res_path, top_n_vr_videos=max_n_videos)
        video_idx2meta_idx = {video2idx[m["vid_name"]]: i for i, m in enumerate(video_metas)}
        rows = [query2video[m["desc_id"]] for m in query_metas]
        external = (torch.tensor([[video_idx2meta_idx[e[0]] for e in r] for r in rows], dtype=torch.int32),
                    torch.tensor([[e[3] for e in r] for r in rows], dtype=torch.float32))
    searcher = _searcher(model, opt, ctx_info, max_before_nms, max_n_videos)
    gt = None
    if is_svmr:
        meta_idx = {m["vid_name"]: i for i, m in enumerate(video_metas)}
        gt = torch.tensor([meta_idx[m["vid_name"]] for m in query_metas], dtype=torch.int32)
    run_tasks = tuple(t for t in ("VCMR", "VR", "SVMR") if t in tasks)
    with torch.no_grad():
        out = searcher.search_host(qfeat, qmask, gt, run_tasks, external_topk=external)

    return host_section(out, query_metas, video_metas, video2idx, searcher.index.ctx_len, opt.clip_length, tasks)


def host_section(out, query_metas, video_metas, video2idx, ctx_len, clip_length, tasks):
    """Host section of compute_query2ctx_info (reference inference.py:391-445), vectorised: the numpy result arrays
    of VCMRSearcher.search_host -> {"SVMR"/"VCMR"/"VR": [{desc_id, desc, predictions}]} in the reference's format
    (flat index -> (rank, st, ed); rank -> position in video_metas -> dataset video id; clip indices -> seconds)."""
    is_svmr, is_vr, is_vcmr = "SVMR" in tasks, "VR" in tasks, "VCMR" in tasks
    meta_to_video_idx = np.asarray([video2idx[m["vid_name"]] for m in video_metas], dtype=np.int64)
    clip = clip_length
    res = {}
    if is_svmr:
        res["SVMR"] = _svmr_predictions(out["svmr_flat_idx"], out["svmr_score"], query_metas, video2idx, ctx_len, clip)
    if is_vr:
        top_idx, top_sc = out["top_video_idx"][:, :100], out["top_video_score"][:, :100]
        vr_rows = np.empty(top_idx.shape + (4,), dtype=object)
        vr_rows[..., 0] = meta_to_video_idx[top_idx].astype(object)
        vr_rows[..., 1] = vr_rows[..., 2] = 0
        vr_rows[..., 3] = top_sc.astype(np.float64).astype(object)
        vr_rows = vr_rows.tolist()
        res["VR"] = [dict(desc_id=m["desc_id"], desc=m["desc"], predictions=vr_rows[i])
                     for i, m in enumerate(query_metas)]
    if is_vcmr:
        flat = out["span_flat_idx"].astype(np.int64)
        # the reference unravels with (max_n_videos, opt.max_ctx_l, opt.max_ctx_l) (inference.py:424-425), which
        # silently assumes the corpus tensors are max_ctx_l wide; the true width is used here
        rank, rem = np.divmod(flat, ctx_len * ctx_len)
        st_i, ed_i = np.divmod(rem, ctx_len)
        meta = np.take_along_axis(out["top_video_idx"].astype(np.int64), rank, axis=1)
        vid = meta_to_video_idx[meta]
        st_sec = st_i.astype(np.float32) * clip
        ed_sec = ed_i.astype(np.float32) * clip + clip
        vcmr_rows = _rows_to_predictions(vid, st_sec, ed_sec, out["span_score"])
        res["VCMR"] = [dict(desc_id=m["desc_id"], desc=m["desc"], predictions=vcmr_rows[i])
                       for i, m in enumerate(query_metas)]
    return {k: v for k, v in res.items() if len(v) != 0}


def compute_query2ctx_info_svmr_only(model, eval_dataset, opt, ctx_info, max_before_nms=1000, max_n_videos=200,
                                     tasks=("SVMR",)):
    """reference inference.py:107-167: start/end distributions on each query's ground-truth video only."""
    model.eval()
    video2idx = eval_dataset.video2idx
    video_metas = ctx_info["video_metas"]
    query_metas, qfeat, qmask = _gather_queries(eval_dataset, opt, True)
    searcher = _searcher(model, opt, ctx_info, max_before_nms, max_n_videos)
    meta_idx = {m["vid_name"]: i for i, m in enumerate(video_metas)}
    gt = torch.tensor([meta_idx[m["vid_name"]] for m in query_metas], dtype=torch.int32)
    with torch.no_grad():
        out = searcher.search_host(qfeat, qmask, gt, ("SVMR",))
    return dict(SVMR=_svmr_predictions(out["svmr_flat_idx"], out["svmr_score"], query_metas, video2idx,
                                       searcher.index.ctx_len, opt.clip_length))


def get_eval_res(model, eval_dataset, opt, tasks, max_after_nms):
    """reference inference.py:448-464."""
    context_info = compute_context_info(model, eval_dataset, opt)
    if "VCMR" in tasks or "VR" in tasks:
        eval_res = compute_query2ctx_info(model, eval_dataset, opt, context_info,
                                          max_before_nms=opt.max_before_nms, max_n_videos=opt.max_vcmr_video,
                                          tasks=tasks)
    else:
        eval_res = compute_query2ctx_info_svmr_only(model, eval_dataset, opt, context_info,
                                                    max_before_nms=opt.max_before_nms, max_n_videos=max_after_nms,
                                                    tasks=tasks)
    eval_res["video2idx"] = eval_dataset.video2idx
    return eval_res


def eval_metrics_device(model, eval_dataset, opt, tasks=("SVMR", "VCMR", "VR"), ctx_info=None, max_after_nms=100):
    """Per-epoch validation metrics without prediction lists: the retrieval of get_eval_res on the device and the
    metrics of standalone_eval.eval_retrieval (reference inference.py:489-500 on top of eval.py:83-276) computed from
    the engine's device tensors by xmlb_eval_first_hit -- the numbers eval_epoch writes to *_metrics.json, at the cost
    of the search itself (the list building and the Python evaluator take ~1 s per 10 K queries).  Ground truth comes
    from eval_dataset.query_data (vid_name, ts, type), like in the reference.  No NMS, top-100 predictions."""
    from .eval_metrics import DESC_TYPE2IDX, eval_search_result_device
    model.eval()
    if ctx_info is None:
        ctx_info = compute_context_info(model, eval_dataset, opt)
    query_metas, qfeat, qmask = _gather_queries(eval_dataset, opt, True)
    searcher = _searcher(model, opt, ctx_info, opt.max_before_nms, opt.max_vcmr_video)
    meta_idx = {m["vid_name"]: i for i, m in enumerate(ctx_info["video_metas"])}
    gt_by_id = {q["desc_id"]: q for q in eval_dataset.query_data}
    gts = [gt_by_id[m["desc_id"]] for m in query_metas]
    gt = torch.tensor([meta_idx[g["vid_name"]] for g in gts], dtype=torch.int32)
    run_tasks = tuple(t for t in ("VCMR", "VR", "SVMR") if t in tasks)
    dev = searcher.index.device
    with torch.no_grad():
        res = searcher.search(qfeat.to(dev), qmask.to(dev), gt.to(dev), run_tasks)
    use_types = all("type" in g for g in gts)
    return eval_search_result_device(res, gt, [g["ts"] for g in gts], searcher.index.ctx_len, opt.clip_length,
                                     desc_types=[DESC_TYPE2IDX[g["type"]] for g in gts] if use_types else None,
                                     max_pred_per_query=min(100, max_after_nms), tasks=run_tasks)


# ------------------------------------------------------------------------------------------------ post-processing
def get_submission_top_n(submission, top_n=100):
    """reference baselines/clip_alignment_with_language/inference.py:503-515."""
    out = dict(video2idx=submission["video2idx"])
    for k, lst in submission.items():
        if k != "video2idx":
            for e in lst:
                e["predictions"] = e["predictions"][:top_n]
            out[k] = lst
    return out


NMS_MAX_PREDICTIONS = 4096  # xmlb_temporal_nms works on up to this many ranked predictions per query


def _nms_lists(res_list, nms_thd, max_before_nms, max_after_nms, per_video, device="cuda"):
    n_in = min(max_before_nms, max(len(e["predictions"]) for e in res_list))
    if n_in > NMS_MAX_PREDICTIONS:
        raise ValueError("temporal NMS handles at most %d predictions per query (got max_before_nms=%d with lists of "
                         "up to %d); lower max_before_nms" % (NMS_MAX_PREDICTIONS, max_before_nms, n_in))
    nq = len(res_list)
    arr = np.zeros((nq, n_in, 4), dtype=np.float64)
    cnt = np.zeros(nq, dtype=np.int32)
    for i, e in enumerate(res_list):
        p = np.asarray(e["predictions"][:n_in], dtype=np.float64).reshape(-1, 4)
        arr[i, :len(p)] = p
        cnt[i] = len(p)
    t = torch.from_numpy(arr).to(device)
    vid = t[..., 0].to(torch.int32) if per_video else None
    kept, kcnt = ops.temporal_nms(t[..., 1].float(), t[..., 2].float(), t[..., 3].float(), nms_thd, max_after_nms,
                                  video_idx=vid, n_valid=torch.from_numpy(cnt).to(device))
    kept, kcnt = kept.cpu().numpy(), kcnt.cpu().numpy()
    for i, e in enumerate(res_list):
        e["predictions"] = [e["predictions"][j] for j in kept[i, :kcnt[i]].tolist()]
    return res_list


def post_processing_vcmr_nms(vcmr_res, nms_thd=0.6, max_before_nms=1000, max_after_nms=100):
    """reference baselines/clip_alignment_with_language/inference.py:228-244 (per-video NMS, re-rank)."""
    return _nms_lists(vcmr_res, nms_thd, max_before_nms, max_after_nms, per_video=True)


def post_processing_svmr_nms(svmr_res, nms_thd=0.6, max_before_nms=1000, max_after_nms=100):
    """reference baselines/clip_alignment_with_language/inference.py:247-265."""
    return _nms_lists(svmr_res, nms_thd, max_before_nms, max_after_nms, per_video=False)


# ------------------------------------------------------------------------------------------------ per-epoch evaluation
POST_PROCESSING_MMS_FUNC = {"SVMR": post_processing_svmr_nms, "VCMR": post_processing_vcmr_nms}


def _save_json(data, path, pretty=False):
    import json
    with open(path, "w") as fh:
        if pretty:
            json.dump(data, fh, indent=4, sort_keys=False)
        else:
            json.dump(data, fh)


def eval_epoch(model, eval_dataset, opt, save_submission_filename, tasks=("SVMR",), max_after_nms=100):
    """reference inference.py:472-531: run the retrieval drivers, write the top-`max_after_nms` submission (and, on the
    validation split, its metrics) under opt.results_dir; with opt.nms_thd != -1 repeat after temporal NMS.
    -> (metrics, metrics_nms, written file paths).  opt fields read here: results_dir, eval_split_name, debug,
    dset_name, nms_thd, max_before_nms (plus those of the drivers)."""
    import os
    from .eval_metrics import eval_retrieval
    model.eval()
    raw = get_eval_res(model, eval_dataset, opt, tasks, max_after_nms=max_after_nms)
    iou_thds = (0.5, 0.7)
    submission_path = os.path.join(opt.results_dir, save_submission_filename)
    submission = get_submission_top_n(raw, top_n=max_after_nms)
    _save_json(submission, submission_path)
    metrics = metrics_nms = None
    paths = [submission_path]
    on_val = opt.eval_split_name == "val"  # test_public has no ground truth
    if on_val:
        metrics = eval_retrieval(submission, eval_dataset.query_data, iou_thds=iou_thds, match_number=not opt.debug,
                                 verbose=opt.debug, use_desc_type=opt.dset_name == "tvr")
        metrics_path = submission_path.replace(".json", "_metrics.json")
        _save_json(metrics, metrics_path, pretty=True)
        paths.append(metrics_path)
    if opt.nms_thd != -1:
        after = dict(video2idx=raw["video2idx"])
        for task, nms in POST_PROCESSING_MMS_FUNC.items():
            if task in raw:
                after[task] = nms(raw[task], nms_thd=opt.nms_thd, max_before_nms=opt.max_before_nms,
                                  max_after_nms=max_after_nms)
        nms_path = submission_path.replace(".json", "_nms_thd_{}.json".format(opt.nms_thd))
        _save_json(after, nms_path)
        if on_val:
            metrics_nms = eval_retrieval(after, eval_dataset.query_data, iou_thds=iou_thds,
                                         match_number=not opt.debug, verbose=opt.debug)
            nms_metrics_path = nms_path.replace(".json", "_metrics.json")
            _save_json(metrics_nms, nms_metrics_path, pretty=True)
            paths += [nms_path, nms_metrics_path]
        else:
            paths = [nms_path]
    return metrics, metrics_nms, paths
