"""Generate the golden vectors that pin `oracle/xml_oracle.py` to the REAL reference.

Run in the build container only (it imports jayleicn/TVRetrieval from /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

It executes the unmodified reference model + drivers (model_xml.py, inference.py, temporal_nms.py, ...)
on small seeded synthetic datasets and writes `tests/golden/<case>.npz` holding the weights, the
inputs and every reference output the tests compare against.  Shims: `easydict`, `h5py.File`
(tests/golden/_shims) and `numpy.int` (reference inference.py:289,293 uses the removed alias).
"""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("XML_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_shims"), REFERENCE, REPO]
np.int = int  # noqa

import torch  # noqa: E402
from easydict import EasyDict  # noqa: E402
from baselines.crossmodal_moment_localization.model_xml import XML, xml_base_config  # noqa: E402
from baselines.crossmodal_moment_localization import inference as ref_inf  # noqa: E402
from baselines.clip_alignment_with_language.inference import (  # noqa: E402
    post_processing_vcmr_nms, post_processing_svmr_nms, get_submission_top_n)
from utils.temporal_nms import temporal_non_maximum_suppression  # noqa: E402

from tvretrieval_b200.synthetic import SyntheticEvalDataset  # noqa: E402

CASES = {
    # BASELINE.json configs[0] analogue: video-only, SVMR-only driver (reference inference.py:107-167)
    "video_only_svmr": dict(
        cfg=dict(ctx_mode="video", merge_two_stream=False, cross_att=False, hidden_size=64, n_heads=4,
                 visual_input_size=48, query_input_size=40, sub_input_size=40, max_ctx_l=32, max_desc_l=12),
        data=dict(n_videos=10, n_queries=10, seed=11), ctx_bsz=4, q_bsz=4, max_n_videos=5,
        max_before_nms=40, tasks=("SVMR",)),
    # the north-star path: video_sub, cross attention, merged ConvSE, full VCMR/SVMR/VR driver
    "video_sub_vcmr": dict(
        cfg=dict(ctx_mode="video_sub", merge_two_stream=True, cross_att=True, hidden_size=64, n_heads=4,
                 visual_input_size=48, query_input_size=40, sub_input_size=40, max_ctx_l=32, max_desc_l=12),
        data=dict(n_videos=24, n_queries=12, seed=12), ctx_bsz=10, q_bsz=5, max_n_videos=8,
        max_before_nms=50, tasks=("VCMR", "SVMR", "VR")),
    # two streams, no cross attention, per-stream ConvSE averaged (reference model_xml.py:580-585)
    "video_sub_nocross": dict(
        cfg=dict(ctx_mode="video_sub", merge_two_stream=False, cross_att=False, hidden_size=32, n_heads=2,
                 visual_input_size=24, query_input_size=20, sub_input_size=28, max_ctx_l=24, max_desc_l=8),
        data=dict(n_videos=14, n_queries=9, seed=13), ctx_bsz=6, q_bsz=4, max_n_videos=6,
        max_before_nms=30, tasks=("VCMR", "SVMR", "VR")),
}


def make_opt(case, cfg):
    return EasyDict(eval_context_bsz=case["ctx_bsz"], eval_query_bsz=case["q_bsz"], num_workers=0,
                    pin_memory=False, device=torch.device("cpu"), ctx_mode=cfg["ctx_mode"],
                    external_inference_vr_res_path=None, q2c_alpha=20.0, min_pred_l=2, max_pred_l=16,
                    max_ctx_l=cfg["max_ctx_l"], clip_length=1.5, debug=False)


def preds_to_array(res_list):
    """list of {desc_id, desc, predictions=[[vid, st, ed, score]...]} -> (Nq, K, 4) float64 (+ counts)."""
    k = max(len(e["predictions"]) for e in res_list)
    out = np.zeros((len(res_list), k, 4), dtype=np.float64)
    cnt = np.zeros(len(res_list), dtype=np.int64)
    for i, e in enumerate(res_list):
        p = np.asarray(e["predictions"], dtype=np.float64).reshape(-1, 4)
        out[i, :len(p)] = p
        cnt[i] = len(p)
    return out, cnt


def run_case(name, case):
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(case["cfg"])
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    ds = SyntheticEvalDataset(max_ctx_l=cfg.max_ctx_l, max_desc_l=cfg.max_desc_l,
                              video_dim=cfg.visual_input_size, sub_dim=cfg.sub_input_size,
                              query_dim=cfg.query_input_size, ctx_mode=cfg.ctx_mode, min_ctx_l=3, **case["data"])
    opt = make_opt(case, cfg)
    out = {"cfg_json": json.dumps({k: v for k, v in cfg.items()}),
           "case_json": json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in case.items()})}
    for k, v in model.state_dict().items():
        out["w/" + k] = v.numpy()
    out["ctx_lens"] = np.asarray(ds.ctx_lens)
    out["video2idx"] = np.asarray([ds.video2idx[v["vid_name"]] for v in ds.video_data])
    out["query_gt_meta_idx"] = np.asarray([int(q["vid_name"].split("_")[1]) for q in ds.query_data])
    for i in range(len(ds.video_data)):
        if ds.use_video:
            out["video_feat/%d" % i] = ds.video_feats[i].numpy()
        if ds.use_sub:
            out["sub_feat/%d" % i] = ds.sub_feats[i].numpy()
    for i, q in enumerate(ds.query_feats):
        out["query_feat/%d" % i] = q.numpy()

    with torch.no_grad():
        ctx = ref_inf.compute_context_info(model, ds, opt)
        for k in ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2", "sub_mask"):
            if ctx[k] is not None:
                out["ctx/" + k] = ctx[k].numpy()

        # raw model outputs for all queries in one batch, cross=True (reference model_xml.py:553-586)
        from baselines.crossmodal_moment_localization.start_end_dataset import start_end_collate, \
            prepare_batch_inputs
        ds.set_data_mode("query")
        ds.load_gt_vid_name_for_query(True)
        batch = start_end_collate([ds[i] for i in range(len(ds))])
        inputs = prepare_batch_inputs(batch[1], device=opt.device)
        out["query_feat_padded"] = inputs["query_feat"].numpy()
        out["query_mask"] = inputs["query_mask"].numpy()
        vq, sq = model.encode_query(inputs["query_feat"], inputs["query_mask"])
        out["video_query"], out["sub_query"] = vq.numpy(), sq.numpy()
        q2c, st, ed = model.get_pred_from_raw_query(
            inputs["query_feat"], inputs["query_mask"], ctx["video_feat1"], ctx["video_feat2"], ctx["video_mask"],
            ctx["sub_feat1"], ctx["sub_feat2"], ctx["sub_mask"], cross=True)
        out["cross/q2c"], out["cross/st"], out["cross/ed"] = q2c.numpy(), st.numpy(), ed.numpy()

        # cross=False (in-batch) outputs on each query's GT video, as the SVMR-only driver feeds them
        gt = torch.as_tensor(out["query_gt_meta_idx"])
        pick = lambda t: None if t is None else t[gt]  # noqa: E731
        q2c, st, ed = model.get_pred_from_raw_query(
            inputs["query_feat"], inputs["query_mask"], pick(ctx["video_feat1"]), pick(ctx["video_feat2"]),
            pick(ctx["video_mask"]), pick(ctx["sub_feat1"]), pick(ctx["sub_feat2"]), pick(ctx["sub_mask"]),
            cross=False)
        out["inbatch/q2c"], out["inbatch/st"], out["inbatch/ed"] = q2c.numpy(), st.numpy(), ed.numpy()

        if "VCMR" in case["tasks"]:
            res = ref_inf.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=case["max_before_nms"],
                                                 max_n_videos=case["max_n_videos"], tasks=case["tasks"])
        else:
            res = ref_inf.compute_query2ctx_info_svmr_only(model, ds, opt, ctx,
                                                           max_before_nms=case["max_before_nms"],
                                                           max_n_videos=case["max_n_videos"], tasks=case["tasks"])
        for task, lst in res.items():
            arr, cnt = preds_to_array(lst)
            out["res/%s" % task], out["res/%s_count" % task] = arr, cnt
            out["res/%s_desc_id" % task] = np.asarray([e["desc_id"] for e in lst])

        # optional NMS post-processing (reference inference.py:507-515), thd 0.5
        nms_funcs = {"SVMR": post_processing_svmr_nms, "VCMR": post_processing_vcmr_nms}
        for task, fn in nms_funcs.items():
            if task in res:
                lst = fn(copy.deepcopy(res[task]), nms_thd=0.5, max_before_nms=case["max_before_nms"],
                         max_after_nms=20)
                arr, cnt = preds_to_array(lst)
                out["nms/%s" % task], out["nms/%s_count" % task] = arr, cnt
        top = get_submission_top_n(dict(video2idx=ds.video2idx, **copy.deepcopy(res)), top_n=7)
        for task in res:
            out["top7/%s" % task] = preds_to_array(top[task])[0]

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def nms_vectors():
    """Known-answer vectors for utils/temporal_nms.py:25-74 on random span lists (incl. edge cases)."""
    rng = np.random.RandomState(7)
    out = {}
    n_cases = 0
    for n in (1, 2, 3, 17, 60, 150):
        for thd in (0.3, 0.5, 0.7):
            st = np.floor(rng.rand(n) * 40) * 1.5
            ln = (np.floor(rng.rand(n) * 14) + 3) * 1.5
            sc = np.round(rng.rand(n), 3)  # rounded -> some exact score ties
            preds = [[float(a), float(a + b), float(c)] for a, b, c in zip(st, ln, sc)]
            kept = temporal_non_maximum_suppression(copy.deepcopy(preds), thd, max_after_nms=100)
            out["in/%d" % n_cases] = np.asarray(preds, dtype=np.float64).reshape(-1, 3)
            out["thd/%d" % n_cases] = np.float64(thd)
            out["out/%d" % n_cases] = np.asarray(kept, dtype=np.float64).reshape(-1, 3)
            n_cases += 1
    out["n_cases"] = np.int64(n_cases)
    path = os.path.join(HERE, "temporal_nms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


TRAIN_VARIANTS = {
    # reference defaults (hinge, all in-batch negatives)
    "plain": dict(),
    # the second training phase: hard negatives from the top of the ranking, lse ranking loss, span loss re-weighted
    "hard": dict(use_hard_negative=True, hard_pool_size=3, ranking_loss_type="lse", lw_st_ed=0.01, lw_neg_q=0.7),
}
TRAIN_SEED = 77


def train_batch(case, cfg):
    """One training batch built from the case's synthetic dataset: item i = (query i, its ground-truth video),
    padded like start_end_collate (start_end_dataset.py:346-359); st/ed labels are seeded random clip indices."""
    ds = SyntheticEvalDataset(max_ctx_l=cfg.max_ctx_l, max_desc_l=cfg.max_desc_l,
                              video_dim=cfg.visual_input_size, sub_dim=cfg.sub_input_size,
                              query_dim=cfg.query_input_size, ctx_mode=cfg.ctx_mode, min_ctx_l=3, **case["data"])
    from utils.tensor_utils import pad_sequences_1d
    gt = [int(q["vid_name"].split("_")[1]) for q in ds.query_data]
    inputs = {}
    inputs["query_feat"], inputs["query_mask"] = pad_sequences_1d(list(ds.query_feats), dtype=torch.float32)
    for name, used, feats in (("video", ds.use_video, ds.video_feats), ("sub", ds.use_sub, ds.sub_feats)):
        if used:
            inputs[name + "_feat"], inputs[name + "_mask"] = pad_sequences_1d([feats[v] for v in gt],
                                                                               dtype=torch.float32)
        else:
            inputs[name + "_feat"] = inputs[name + "_mask"] = None
    inputs["tef_feat"] = inputs["tef_mask"] = None
    rng = np.random.RandomState(5)
    lens = np.asarray([ds.ctx_lens[v] for v in gt])
    st = (rng.rand(len(gt)) * (lens - 1)).astype(np.int64)
    ed = np.minimum(lens - 1, st + 1 + (rng.rand(len(gt)) * 6).astype(np.int64))
    inputs["st_ed_indices"] = torch.from_numpy(np.stack([st, ed], 1))
    return inputs, np.asarray(gt)


def train_vectors():
    """XML.forward + losses + backward (reference model_xml.py:212-251,588-637) with dropout off (eval mode):
    loss, the four reported floats and the gradient of every parameter."""
    out = {"seed": np.int64(TRAIN_SEED), "variants_json": json.dumps(TRAIN_VARIANTS)}
    for name, case in CASES.items():
        for variant, override in TRAIN_VARIANTS.items():
            cfg = copy.deepcopy(xml_base_config)
            cfg.update(case["cfg"])
            cfg.update(override)
            torch.manual_seed(2018)
            model = XML(cfg).eval()
            inputs, gt = train_batch(case, cfg)
            torch.manual_seed(TRAIN_SEED)
            loss, parts = model(**inputs)
            loss.backward()
            key = "%s/%s/" % (name, variant)
            out[key + "loss"] = np.float64(loss.item())
            out[key + "parts"] = np.asarray([parts[k] for k in ("loss_st_ed", "loss_neg_ctx", "loss_neg_q",
                                                                "loss_overall")], dtype=np.float64)
            out[key + "st_ed_indices"] = inputs["st_ed_indices"].numpy()
            out[key + "gt"] = gt
            for pname, prm in model.named_parameters():
                out[key + "g/" + pname] = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).numpy()
            print(key, "loss %.6f" % loss.item(), parts)
    path = os.path.join(HERE, "train_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def external_vr_vectors():
    """compute_query2ctx_info with --external_inference_vr_res_path (reference inference.py:264-273,349-355) on the
    video_sub_vcmr case: the VCMR video lists come from another system's VR submission instead of the model's own
    video retrieval.  The "external" submission is synthetic: per query, a seeded random set of videos with seeded
    cosine-like scores in descending order."""
    import tempfile
    name = "video_sub_vcmr"
    case = CASES[name]
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(case["cfg"])
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    ds = SyntheticEvalDataset(max_ctx_l=cfg.max_ctx_l, max_desc_l=cfg.max_desc_l,
                              video_dim=cfg.visual_input_size, sub_dim=cfg.sub_input_size,
                              query_dim=cfg.query_input_size, ctx_mode=cfg.ctx_mode, min_ctx_l=3, **case["data"])
    opt = make_opt(case, cfg)
    rng = np.random.RandomState(3)
    k = case["max_n_videos"]
    video_ids = [ds.video2idx[v["vid_name"]] for v in ds.video_data]
    vr = []
    for q in ds.query_data:
        vids = rng.permutation(video_ids)[:k + 2]  # two more than needed: get_submission_top_n truncates
        scores = np.sort(rng.uniform(-0.2, 0.6, size=k + 2))[::-1]
        vr.append(dict(desc_id=q["desc_id"], desc=q["desc"],
                       predictions=[[int(v), 0, 0, float(s)] for v, s in zip(vids, scores)]))
    submission = dict(video2idx=ds.video2idx, VR=vr)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
        json.dump(submission, fh)
    opt.external_inference_vr_res_path = fh.name
    with torch.no_grad():
        ctx = ref_inf.compute_context_info(model, ds, opt)
        res = ref_inf.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=case["max_before_nms"],
                                             max_n_videos=k, tasks=("VCMR", "VR"))
    os.unlink(fh.name)
    out = {"case": name, "submission_json": json.dumps(submission)}
    for task, lst in res.items():
        out["res/%s" % task] = preds_to_array(lst)[0]
    path = os.path.join(HERE, "external_vr.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def visualization_vectors():
    """XML.get_visualization_data (reference model_xml.py:253-289) on the video_sub_vcmr training batch."""
    name = "video_sub_vcmr"
    case = CASES[name]
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(case["cfg"])
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    inputs, _ = train_batch(case, cfg)
    with torch.no_grad():
        data = model.get_visualization_data(**inputs)
    out = {"case": name, "n": np.int64(len(data))}
    for i, d in enumerate(data):
        for k, v in d.items():
            out["%d/%s" % (i, k)] = np.asarray(v)
    path = os.path.join(HERE, "visualization.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def eval_metric_vectors():
    """standalone_eval.eval.eval_retrieval (reference standalone_eval/eval.py:255-276) on a synthetic submission:
    variable numbers of predictions per query (some beyond the 100 that are evaluated), predictions engineered to hit
    the IoU thresholds from both sides, TVR-style single ground-truth spans and DiDeMo-style multi-annotation ones."""
    from standalone_eval.eval import eval_retrieval
    rng = np.random.RandomState(21)
    n_videos, n_desc = 30, 80
    video2idx = {"vid_%d" % i: 100 + 3 * i for i in range(n_videos)}
    names = list(video2idx)
    gts, sub = [], dict(video2idx=video2idx, VCMR=[], SVMR=[], VR=[])
    for d in range(n_desc):
        vid = names[rng.randint(n_videos)]
        st = float(np.round(rng.uniform(0, 60), 1))
        ed = st + float(np.round(rng.uniform(3, 20), 1))
        multi = d % 9 == 4
        ts = [[st + float(rng.choice([-3, 0, 1.5])), ed + float(rng.choice([-1.5, 0, 3]))] for _ in range(4 + d % 2)] \
            if multi else [st, ed]
        gts.append(dict(desc_id=1000 + d, desc="q%d" % d, type=["v", "t", "vt"][rng.randint(3)], vid_name=vid, ts=ts))

        def spans(n):
            out = []
            for _ in range(n):
                kind = rng.randint(4)
                if kind == 0:    # near the ground truth: IoU around the thresholds
                    a = st + rng.choice([-4.5, -3, -1.5, 0, 1.5, 3])
                    b = ed + rng.choice([-4.5, -3, -1.5, 0, 1.5, 3])
                    b = max(b, a + 1.5)
                else:
                    a = float(np.floor(rng.uniform(0, 80) / 1.5) * 1.5)
                    b = a + 1.5 * rng.randint(2, 17)
                out.append((float(a), float(b)))
            return out
        n_vcmr = int(rng.choice([1, 3, 40, 100, 130])) if d % 7 else 100  # (the reference cannot take 0)
        sub["VCMR"].append(dict(desc_id=1000 + d, desc="q%d" % d, predictions=[
            [video2idx[vid] if rng.rand() < 0.3 else video2idx[names[rng.randint(n_videos)]], a, b, float(rng.rand())]
            for a, b in spans(n_vcmr)]))
        sub["SVMR"].append(dict(desc_id=1000 + d, desc="q%d" % d, predictions=[
            [video2idx[vid], a, b, float(rng.rand())] for a, b in spans(int(rng.choice([1, 12, 100, 120])))]))
        order = rng.permutation(n_videos)[:int(rng.choice([5, 30]))]
        sub["VR"].append(dict(desc_id=1000 + d, desc="q%d" % d, predictions=[
            [video2idx[names[v]], 0, 0, float(rng.rand())] for v in order]))
    out = dict(submission=sub, ground_truth=gts, expected={})
    for use_type in (True, False):
        for thds in ((0.5, 0.7), (0.3,)):
            res = eval_retrieval(copy.deepcopy(sub), copy.deepcopy(gts), iou_thds=thds, verbose=False,
                                 use_desc_type=use_type)
            out["expected"]["%s/%s" % (use_type, ",".join(map(str, thds)))] = \
                [[task, list(m.items())] for task, m in res.items()]
    path = os.path.join(HERE, "eval_metrics.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


ADAM_SHAPES = {"enc.weight": (24, 16), "enc.bias": (24,), "enc.LayerNorm.weight": (24,), "big.weight": (130, 70),
               "conv.weight": (1, 1, 5)}
ADAM_HYPER = dict(lr=1e-2, warmup=0.25, t_total=8, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                  max_grad_norm=1.0)
ADAM_STEPS = 5


def bert_adam_vectors():
    """BertAdam.step (reference optimization.py:273-338) for a few steps on seeded parameters / gradients with the
    two parameter groups of train.py:151-156 (weight decay 0.01 / 0.0) and the warmup_linear schedule.  Gradients
    of some tensors exceed max_grad_norm (clipped), others do not; "conv.weight" gets no gradient at step 2."""
    from baselines.crossmodal_moment_localization.optimization import BertAdam
    g = torch.Generator().manual_seed(99)
    params = {k: torch.nn.Parameter(torch.randn(*shp, generator=g) * 0.5) for k, shp in ADAM_SHAPES.items()}
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in params.items() if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in params.items() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = BertAdam(groups, **ADAM_HYPER)
    out = {"hyper_json": json.dumps(ADAM_HYPER), "shapes_json": json.dumps(ADAM_SHAPES), "n_steps": np.int64(ADAM_STEPS)}
    for k, p in params.items():
        out["p0/" + k] = p.detach().numpy().copy()
    for step in range(ADAM_STEPS):
        for k, p in params.items():
            scale = {"enc.weight": 0.01, "big.weight": 0.05}.get(k, 1.0)  # small-norm and large-norm gradients
            grad = torch.randn(*p.shape, generator=g) * scale
            if k == "conv.weight" and step == 2:
                p.grad = None
                out["has_grad/%d/%s" % (step, k)] = np.int64(0)
                continue
            out["has_grad/%d/%s" % (step, k)] = np.int64(1)
            out["g/%d/%s" % (step, k)] = grad.numpy().copy()
            p.grad = grad.clone()
        opt.step()
        for k, p in params.items():
            out["p/%d/%s" % (step, k)] = p.detach().numpy().copy()
            if p.grad is not None:
                out["g_after/%d/%s" % (step, k)] = p.grad.numpy().copy()
        out["lr/%d" % step] = np.asarray(opt.get_lr(), dtype=np.float64)
    for k, p in params.items():
        out["m/" + k] = opt.state[p]["next_m"].numpy().copy()
        out["v/" + k] = opt.state[p]["next_v"].numpy().copy()
    path = os.path.join(HERE, "bert_adam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(4)
    if "--adam-only" in sys.argv:
        bert_adam_vectors()
        sys.exit(0)
    if "--eval-only" in sys.argv:
        np.bool = bool  # standalone_eval/eval.py uses the removed alias (numpy < 1.24)
        eval_metric_vectors()
        sys.exit(0)
    if "--visualization-only" in sys.argv:
        visualization_vectors()
        sys.exit(0)
    if "--external-vr-only" in sys.argv:
        external_vr_vectors()
        sys.exit(0)
    if "--train-only" not in sys.argv:
        for case_name, case_def in CASES.items():
            run_case(case_name, case_def)
        nms_vectors()
    train_vectors()
    bert_adam_vectors()
    external_vr_vectors()
    visualization_vectors()
    np.bool = bool  # standalone_eval/eval.py uses the removed alias (numpy < 1.24)
    eval_metric_vectors()
