"""GPU parity of the drop-in XML model and drivers, through the C ABI, against
(a) the committed golden vectors produced by the real reference (tests/golden/*.npz) and
(b) the CPU oracle on larger TVR-shaped inputs (H=768, L=128, resnet_i3d features).
Floating point: north_star tolerance 1e-3 relative (the asserts below use tighter bounds where they hold);
ranked indices: exact wherever the oracle's own neighbouring scores differ by more than 5e-5 relative."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import xml_oracle as O
from tests.golden_io import CASE_NAMES, GoldenCase

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(got, want, rtol=1e-4, atol=1e-5):
    got = got.detach().cpu() if torch.is_tensor(got) else torch.as_tensor(got)
    want = want.detach().cpu() if torch.is_tensor(want) else torch.as_tensor(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    torch.testing.assert_close(got.double(), want.double(), rtol=rtol, atol=atol)


def build_model(cfg, weights):
    from tvretrieval_b200.model_xml import XML, AttrDict
    model = XML(AttrDict(cfg))
    model.load_state_dict(weights)
    return model.to(DEV).eval()


class Opt:
    def __init__(self, case, cfg):
        self.eval_context_bsz, self.eval_query_bsz = case["ctx_bsz"], case["q_bsz"]
        self.num_workers, self.pin_memory, self.device = 0, False, torch.device(DEV)
        self.ctx_mode, self.external_inference_vr_res_path = cfg["ctx_mode"], None
        self.q2c_alpha, self.min_pred_l, self.max_pred_l = 20.0, 2, 16
        self.max_ctx_l, self.clip_length, self.debug = cfg["max_ctx_l"], 1.5, False
        self.max_before_nms, self.max_vcmr_video = case["max_before_nms"], case["max_n_videos"]


@pytest.fixture(scope="module", params=CASE_NAMES)
def golden(request):
    g = GoldenCase(request.param)
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    ds = SyntheticEvalDataset(max_ctx_l=g.cfg["max_ctx_l"], max_desc_l=g.cfg["max_desc_l"],
                              video_dim=g.cfg["visual_input_size"], sub_dim=g.cfg["sub_input_size"],
                              query_dim=g.cfg["query_input_size"], ctx_mode=g.cfg["ctx_mode"], min_ctx_l=3,
                              **g.case["data"])
    # the dataset is regenerated from its seed; make sure it is the one the goldens were made from
    assert ds.ctx_lens == g.ctx_lens
    for i in range(g.n_queries):
        assert torch.equal(ds.query_feats[i], g.query_feats[i])
    g.ds = ds
    g.model = build_model(g.cfg, g.weights)
    g.opt = Opt(g.case, g.cfg)
    return g


def test_golden_context_encoding(golden):
    from tvretrieval_b200 import inference as I
    ctx = I.compute_context_info(golden.model, golden.ds, golden.opt)
    assert [m["vid_name"] for m in ctx["video_metas"]] == [v["vid_name"] for v in golden.ds.video_data]
    for k, want in golden.ctx().items():
        if want is None:
            assert ctx[k] is None
        else:
            close(ctx[k], want, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("cross", [True, False])
def test_golden_pred_from_raw_query(golden, cross):
    ctx = {k: (None if v is None else v.to(DEV)) for k, v in golden.ctx().items()}
    if not cross:
        gt = torch.as_tensor(golden.query_gt_meta_idx, device=DEV)
        ctx = {k: (None if v is None else v[gt].contiguous()) for k, v in ctx.items()}
    with torch.no_grad():
        q2c, st, ed = golden.model.get_pred_from_raw_query(
            golden.query_feat.to(DEV), golden.query_mask.to(DEV), ctx["video_feat1"], ctx["video_feat2"],
            ctx["video_mask"], ctx["sub_feat1"], ctx["sub_feat2"], ctx["sub_mask"], cross=cross)
    tag = "cross" if cross else "inbatch"
    close(q2c, golden.t(tag + "/q2c"), rtol=1e-4, atol=2e-6)
    want_st, want_ed = golden.t(tag + "/st"), golden.t(tag + "/ed")
    close(st, want_st, rtol=1e-4, atol=5e-5), close(ed, want_ed, rtol=1e-4, atol=5e-5)
    assert torch.equal(st.cpu() == -1e10, want_st == -1e10)


def preds_array(lst):
    k = max(len(e["predictions"]) for e in lst)
    out = np.zeros((len(lst), k, 4))
    for i, e in enumerate(lst):
        out[i, :len(e["predictions"])] = np.asarray(e["predictions"], dtype=np.float64).reshape(-1, 4)
    return out


def assert_ranked_equal(got, ref, score_rtol=1e-4, tie_rtol=5e-5):
    """Rows [video_idx, st, ed, score] must match exactly in the positive-score region, except where two
    neighbouring reference scores are closer than `tie_rtol` relative: two fp32 implementations with different
    summation orders differ by ~1e-6 relative on the final scores (exp(20*q2c) amplifies), and the tensor-core
    query encoder adds ~1e-5 (its fp32 accumulator truncates), so gaps of that size are ties no two implementations
    agree on.  Returns the number of such swapped rows."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    np.testing.assert_allclose(got[..., 3], ref[..., 3], rtol=score_rtol, atol=1e-12)
    pos = ref[..., 3] > 0
    bad = (got[..., :3] != ref[..., :3]).any(-1) & pos
    for q, r in zip(*np.nonzero(bad)):
        s = ref[q, :, 3]
        lo, hi = max(r - 1, 0), min(r + 1, len(s) - 1)
        gap = min(abs(s[r] - s[lo]) if lo != r else np.inf, abs(s[r] - s[hi]) if hi != r else np.inf)
        assert gap <= tie_rtol * abs(s[r]), "rank mismatch at query %d rank %d (gap %g)" % (q, r, gap)
    return int(bad.sum())


def test_golden_driver(golden):
    from tvretrieval_b200 import inference as I
    c = golden.case
    ctx = I.compute_context_info(golden.model, golden.ds, golden.opt)
    if "VCMR" in c["tasks"]:
        res = I.compute_query2ctx_info(golden.model, golden.ds, golden.opt, ctx, max_before_nms=c["max_before_nms"],
                                       max_n_videos=c["max_n_videos"], tasks=tuple(c["tasks"]))
    else:
        res = I.compute_query2ctx_info_svmr_only(golden.model, golden.ds, golden.opt, ctx,
                                                 max_before_nms=c["max_before_nms"], max_n_videos=c["max_n_videos"])
    assert sorted(res.keys()) == sorted(c["tasks"])
    for task in c["tasks"]:
        assert [e["desc_id"] for e in res[task]] == golden.z["res/%s_desc_id" % task].tolist()
        assert all(isinstance(e["predictions"][0][0], int) and isinstance(e["predictions"][0][3], float)
                   for e in res[task])
        n_swapped = assert_ranked_equal(preds_array(res[task]), golden.z["res/" + task])
        assert n_swapped == 0
    # post-processing on the reference's own ranked lists: top-n truncation and NMS (thd 0.5)
    for task, fn in (("VCMR", I.post_processing_vcmr_nms), ("SVMR", I.post_processing_svmr_nms)):
        if task not in c["tasks"]:
            continue
        ref_in, ref_out, cnt = golden.z["res/" + task], golden.z["nms/" + task], golden.z["nms/%s_count" % task]
        lst = [dict(desc_id=0, desc="", predictions=[[int(p[0]), p[1], p[2], p[3]] for p in row.tolist()])
               for row in ref_in]
        out = fn(copy.deepcopy(lst), nms_thd=0.5, max_before_nms=c["max_before_nms"], max_after_nms=20)
        for q, e in enumerate(out):
            assert len(e["predictions"]) == cnt[q]
            assert np.array_equal(np.asarray(e["predictions"], dtype=np.float64).reshape(-1, 4), ref_out[q, :cnt[q]]), (task, q)
        top = I.get_submission_top_n(dict(video2idx={}, **{task: copy.deepcopy(lst)}), top_n=7)
        assert np.array_equal(preds_array(top[task]), golden.z["top7/" + task])


def test_search_with_packed_query_encoder_equals_padded():
    """The engine switches to XML.encode_query_packed for large query blocks; forced on here, the search must return
    the ranks of the padded encoder (scores to rounding: the fused attention sums in a different order)."""
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.inference import cat_tensor
    g = GoldenCase("video_sub_vcmr")
    model = build_model(g.cfg, g.weights)
    with torch.no_grad():
        acc = {k: [] for k in ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2", "sub_mask")}
        for b in g.context_batches():
            b = {k: v.to(DEV) for k, v in b.items()}
            v1, v2, s1, s2 = model.encode_context(b["video_feat"], b["video_mask"], b["sub_feat"], b["sub_mask"])
            for k, v in zip(acc, (v1, v2, b["video_mask"], s1, s2, b["sub_mask"])):
                acc[k].append(v)
        index = CorpusIndex.from_ctx_info({k: cat_tensor(v) for k, v in acc.items()})
    kw = dict(max_n_videos=g.case["max_n_videos"], max_before_nms=g.case["max_before_nms"])
    qf, qm = g.query_feat.to(DEV), g.query_mask.to(DEV)
    gt = torch.as_tensor(g.query_gt_meta_idx, dtype=torch.int32, device=DEV)
    # `packed` also runs the two-pass retrieval, in host mode with its filter pass pipelined over 3 uploaded pieces
    padded, packed = VCMRSearcher(model, index, **kw), VCMRSearcher(model, index, two_pass=True, encode_chunk=5, **kw)
    padded.packed_queries, packed.packed_min_queries = False, 0
    want = padded.search(qf, qm, gt, tasks=("VCMR", "VR", "SVMR"))
    for got in (packed.search(qf, qm, gt, tasks=("VCMR", "VR", "SVMR")),
                packed.search(g.query_feat.pin_memory(), g.query_mask.pin_memory(), gt.cpu(),
                              tasks=("VCMR", "VR", "SVMR"), host=True)):
        for name in ("top_video_idx", "span_flat_idx", "svmr_flat_idx"):
            assert torch.equal(getattr(got, name), getattr(want, name)), name
        for name in ("top_video_score", "span_score", "svmr_score"):
            torch.testing.assert_close(getattr(got, name), getattr(want, name), rtol=2e-5, atol=1e-12)


def test_eval_epoch_writes_submission_and_metrics(tmp_path):
    """eval_epoch (reference inference.py:472-531): submission + metrics files before and after NMS; the metrics are
    those of the reference evaluator on the written submission."""
    import json
    from tvretrieval_b200 import inference as I
    from tvretrieval_b200.eval_metrics import eval_retrieval
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    g = GoldenCase("video_sub_vcmr")
    ds = SyntheticEvalDataset(max_ctx_l=g.cfg["max_ctx_l"], max_desc_l=g.cfg["max_desc_l"],
                              video_dim=g.cfg["visual_input_size"], sub_dim=g.cfg["sub_input_size"],
                              query_dim=g.cfg["query_input_size"], ctx_mode=g.cfg["ctx_mode"], min_ctx_l=3,
                              **g.case["data"])
    model = build_model(g.cfg, g.weights)
    opt = Opt(g.case, g.cfg)
    opt.results_dir, opt.eval_split_name, opt.dset_name, opt.nms_thd = str(tmp_path), "val", "tvr", 0.5
    metrics, metrics_nms, paths = I.eval_epoch(model, ds, opt, "sub.json", tasks=("VCMR", "SVMR", "VR"),
                                               max_after_nms=20)
    assert [os.path.basename(p) for p in paths] == ["sub.json", "sub_metrics.json", "sub_nms_thd_0.5.json",
                                                    "sub_nms_thd_0.5_metrics.json"]
    sub = json.load(open(paths[0]))
    assert set(sub) == {"video2idx", "VCMR", "SVMR", "VR"} and all(len(e["predictions"]) <= 20 for e in sub["VCMR"])
    again = eval_retrieval(sub, ds.query_data, verbose=False)
    assert json.load(open(paths[1])) == json.loads(json.dumps(again)) == json.loads(json.dumps(metrics))
    assert set(metrics_nms) == {"VCMR", "SVMR", "VCMR_by_type", "SVMR_by_type"}
    # the device path (no lists, no host evaluator) returns the metrics eval_epoch wrote
    dev_metrics = I.eval_metrics_device(model, ds, opt, tasks=("VCMR", "SVMR", "VR"), max_after_nms=20)
    assert json.dumps(dev_metrics) == json.dumps(metrics)
    assert metrics["SVMR"]["0.5-r100"] >= metrics["SVMR"]["0.5-r1"]
    after = json.load(open(paths[2]))
    assert set(after) == {"video2idx", "VCMR", "SVMR"}


def test_device_metrics_equal_the_evaluator_on_the_submission():
    """eval_metrics.eval_search_result_device (xmlb_eval_first_hit on the engine's device tensors) returns exactly the
    dictionaries the evaluator computes from the submission built out of the same result (reference
    standalone_eval/eval.py:83-276 semantics: R@K at IoU 0.5 / 0.7, overall and by description type)."""
    import json
    from tvretrieval_b200 import inference as I
    from tvretrieval_b200.eval_metrics import DESC_TYPE2IDX, eval_retrieval, eval_search_result_device
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    g = GoldenCase("video_sub_vcmr")
    ds = SyntheticEvalDataset(max_ctx_l=g.cfg["max_ctx_l"], max_desc_l=g.cfg["max_desc_l"],
                              video_dim=g.cfg["visual_input_size"], sub_dim=g.cfg["sub_input_size"],
                              query_dim=g.cfg["query_input_size"], ctx_mode=g.cfg["ctx_mode"], min_ctx_l=3,
                              **g.case["data"])
    for i, q in enumerate(ds.query_data):  # all three description types, ground-truth spans of varying length
        q["type"] = ("v", "t", "vt")[i % 3]
        q["ts"] = [1.5 * (i % 4), 1.5 * (i % 4) + 3.0 + 1.5 * (i % 3)]
    model = build_model(g.cfg, g.weights)
    opt = Opt(g.case, g.cfg)
    ctx = I.compute_context_info(model, ds, opt)
    k_vid, k_span = g.case["max_n_videos"], g.case["max_before_nms"]
    searcher = I._searcher(model, opt, ctx, k_span, k_vid)
    metas = ctx["video_metas"]
    pos = {m["vid_name"]: i for i, m in enumerate(metas)}
    gt_pos = torch.tensor([pos[q["vid_name"]] for q in ds.query_data], dtype=torch.int32)
    qf, qm = GoldenCase.pad(ds.query_feats)
    res = searcher.search(qf.to(DEV), qm.to(DEV), gt_pos.to(DEV), tasks=("VCMR", "VR", "SVMR"))
    out = {s: getattr(res, s).cpu().numpy() for s in res.__slots__ if getattr(res, s) is not None}
    query_metas = [dict(desc_id=q["desc_id"], desc=q["desc"], vid_name=q["vid_name"]) for q in ds.query_data]
    sub = I.host_section(out, query_metas, metas, ds.video2idx, searcher.index.ctx_len, 1.5, ("VCMR", "VR", "SVMR"))
    sub["video2idx"] = ds.video2idx
    want = eval_retrieval(sub, ds.query_data, verbose=False)
    got = eval_search_result_device(res, gt_pos, [q["ts"] for q in ds.query_data], searcher.index.ctx_len, 1.5,
                                    desc_types=[DESC_TYPE2IDX[q["type"]] for q in ds.query_data])
    assert list(got) == list(want)
    diff = {(t, k): (got[t].get(k), want[t][k]) for t in want for k in want[t] if got[t].get(k) != want[t][k]}
    assert not diff, diff
    assert json.dumps(got, sort_keys=False) == json.dumps(want, sort_keys=False)
    assert want["SVMR"]["0.5-r100"] > 0 and want["VR"]["r100"] > 0  # (the case is not degenerate)


def test_golden_visualization_data():
    """XML.get_visualization_data / return_modular_att / return_similaity (reference model_xml.py:253-289,410-416,
    498-500) against the reference's own output."""
    from tests.golden_io import VisualizationCase
    vc = VisualizationCase()
    tc = vc.train
    model = build_model(tc.cfg, tc.weights)
    inputs = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in tc.inputs.items()}
    got = model.get_visualization_data(**inputs)
    assert len(got) == len(vc.items) and set(got[0]) == set(vc.items[0])
    for g, w in zip(got, vc.items):
        assert np.array_equal(g["st_ed_indices"], w["st_ed_indices"])
    vc.check(got, rtol=1e-4, atol=1e-5)


def test_golden_external_vr_lists(tmp_path):
    """--external_inference_vr_res_path (reference inference.py:264-273,349-355): the VCMR video lists come from
    another system's VR submission; compared with the reference's own output for the same submission file
    (tests/golden/external_vr.npz)."""
    import json
    import os
    from tests.golden_io import GOLDEN_DIR
    from tvretrieval_b200 import inference as I
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    z = np.load(os.path.join(GOLDEN_DIR, "external_vr.npz"))
    g = GoldenCase(str(z["case"]))
    ds = SyntheticEvalDataset(max_ctx_l=g.cfg["max_ctx_l"], max_desc_l=g.cfg["max_desc_l"],
                              video_dim=g.cfg["visual_input_size"], sub_dim=g.cfg["sub_input_size"],
                              query_dim=g.cfg["query_input_size"], ctx_mode=g.cfg["ctx_mode"], min_ctx_l=3,
                              **g.case["data"])
    model = build_model(g.cfg, g.weights)
    opt = Opt(g.case, g.cfg)
    path = tmp_path / "external_vr.json"
    path.write_text(str(z["submission_json"]))
    opt.external_inference_vr_res_path = str(path)
    ctx = I.compute_context_info(model, ds, opt)
    res = I.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=g.case["max_before_nms"],
                                   max_n_videos=g.case["max_n_videos"], tasks=("VCMR", "VR"))
    sub = json.loads(str(z["submission_json"]))
    k = g.case["max_n_videos"]
    for e, want in zip(res["VR"], sub["VR"]):  # VR output = the external lists, scores exponentiated
        assert [p[0] for p in e["predictions"]] == [p[0] for p in want["predictions"][:k]]
    close(preds_array(res["VR"])[..., 3], z["res/VR"][..., 3], rtol=1e-6, atol=0)
    assert assert_ranked_equal(preds_array(res["VCMR"]), z["res/VCMR"]) == 0


# ---------------------------------------------------------------------------------------------------------
# TVR-shaped comparison against the oracle (BASELINE.json configs[1] dims, fewer videos so the CPU finishes fast)
# ---------------------------------------------------------------------------------------------------------
def tvr_case(ctx_mode, n_videos, n_queries, hidden, max_ctx_l, video_dim, seed):
    from tvretrieval_b200.model_xml import XML, xml_base_config
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(hidden_size=hidden, max_ctx_l=max_ctx_l, max_desc_l=30, visual_input_size=video_dim, ctx_mode=ctx_mode)
    if ctx_mode != "video_sub":
        cfg.update(merge_two_stream=False, cross_att=False)
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    ds = SyntheticEvalDataset(n_videos, n_queries, max_ctx_l, 30, video_dim, 768, 768, ctx_mode=ctx_mode, seed=seed,
                              video_split=2048 if video_dim == 3072 else None)
    weights = {k: v.clone() for k, v in model.state_dict().items()}
    return cfg, model.to(DEV), weights, ds


def oracle_batches(ds, bsz):
    for lo in range(0, len(ds.video_data), bsz):
        b = {}
        if ds.use_video:
            b["video_feat"], b["video_mask"] = GoldenCase.pad(ds.video_feats[lo:lo + bsz])
        if ds.use_sub:
            b["sub_feat"], b["sub_mask"] = GoldenCase.pad(ds.sub_feats[lo:lo + bsz])
        yield b


# Accuracy of the product's query path against the exact (float64) evaluation of the reference arithmetic, as a
# relative deviation of the final VCMR / VR scores: the pooled query vectors come from split-precision tensor-core
# GEMMs with K-chunked fp32 accumulation (~2e-6), q2c from the split-precision corpus contraction (5e-7 abs, x20 through
# exp(20 q2c) = 1e-5), the span probabilities from fp32 softmax of fp32-accurate logits (~1e-5).  Lists must be ordered,
# complete and valued like the float64 lists to within RANK_TAU; the reference's own fp32 arithmetic, measured by the
# same yardstick in the same test, is only determined to ~1e-5 as well (see tests/rank_check.py).
RANK_TAU = 1e-4


def test_tvr_shape_video_sub_vcmr():
    """video_sub, resnet_i3d (Dv=3072), H=768, L=128: full driver; corpus encoding vs the fp32 oracle (values), the
    query path vs the reference arithmetic in float64 on the SAME encoded corpus (ranks)."""
    from tests import rank_check as R
    from tvretrieval_b200 import inference as I
    n_videos, n_queries, k_vid, k_span = 150, 24, 100, 200
    cfg, model, weights, ds = tvr_case("video_sub", n_videos, n_queries, 768, 128, 3072, seed=1234)
    case = dict(ctx_bsz=64, q_bsz=10, max_before_nms=k_span, max_n_videos=k_vid)
    opt = Opt(case, cfg)
    ctx = I.compute_context_info(model, ds, opt)
    with torch.no_grad():
        octx = O.context_info(cfg, weights, oracle_batches(ds, case["ctx_bsz"]))
    for k in ("video_feat1", "video_feat2", "sub_feat1", "sub_feat2"):
        close(ctx[k], octx[k], rtol=1e-3, atol=1e-4)
    res = I.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=k_span, max_n_videos=k_vid,
                                   tasks=("VCMR", "SVMR", "VR"))
    qf, qm = GoldenCase.pad(ds.query_feats)
    v2i = np.asarray([ds.video2idx[v["vid_name"]] for v in ds.video_data])

    # (1) ranks: the engine's device-side lists against float64 on the product's own encoded corpus
    searcher = I._searcher(model, opt, ctx, k_span, k_vid)
    raw = searcher.search(qf.to(DEV), qm.to(DEV), tasks=("VCMR", "VR"))
    vr64, st64, ed64 = R.fp64_scores(cfg, weights, ctx, qf, qm, device=DEV)
    stats = R.check_search_result(vr64, st64, ed64, raw.top_video_idx, raw.top_video_score, raw.span_flat_idx,
                                  raw.span_score, 128)
    # the reference's own fp32 arithmetic (CPU oracle) on the same corpus, by the same yardstick
    cctx = {k: ctx[k].cpu() for k in R.CTX_KEYS}
    with torch.no_grad():
        o32 = O.query_batch_tensor_section(cfg, weights, cctx, qf, qm, q2c_alpha=20.0, max_n_videos=k_vid,
                                           max_before_nms=k_span, min_pred_l=2, max_pred_l=16)
    stats32 = R.check_search_result(vr64, st64, ed64, o32["top_video_idx"], o32["top_video_score"],
                                    o32["span_flat_idx"], o32["span_score"], 128)
    print("\nrank parity vs float64 -- kernels: %s\n                         fp32 oracle (%d threads): %s"
          % (R.fmt(stats), torch.get_num_threads(), R.fmt(stats32)))
    R.assert_within(stats, RANK_TAU, "kernels")
    R.assert_within(stats32, RANK_TAU, "fp32 oracle")

    # (2) the host section decodes exactly the device lists (reference inference.py:391-445)
    want_vcmr = O.decode_vcmr(raw.span_flat_idx.cpu().numpy(), raw.span_score.cpu().numpy(),
                              raw.top_video_idx.cpu().numpy(), v2i, k_vid, 128, 1.5)
    assert np.array_equal(preds_array(res["VCMR"]), want_vcmr)
    want_vr = np.zeros((n_queries, k_vid, 4))
    want_vr[..., 0] = v2i[raw.top_video_idx.cpu().numpy()]
    want_vr[..., 3] = raw.top_video_score.cpu().numpy()
    assert np.array_equal(preds_array(res["VR"]), want_vr)

    # (3) end to end against the oracle run on ITS OWN encoding of the corpus: values (the two corpus encodings differ
    # by fp32 rounding, so near-tied ranks legitimately differ; ranks are covered by (1))
    with torch.no_grad():
        o = O.query_batch_tensor_section(cfg, weights, octx, qf, qm, q2c_alpha=20.0, max_n_videos=k_vid,
                                         max_before_nms=k_span, min_pred_l=2, max_pred_l=16)
    np.testing.assert_allclose(raw.span_score.cpu().numpy(), o["span_score"].numpy(), rtol=1e-3)
    np.testing.assert_allclose(raw.top_video_score.cpu().numpy(), o["top_video_score"].numpy(), rtol=1e-3)
    gt = np.asarray([int(q["vid_name"].split("_")[1]) for q in ds.query_data])
    rows = np.arange(n_queries)
    sv = O.svmr_from_probs(o["st_prob"].numpy()[rows, gt], o["ed_prob"].numpy()[rows, gt], 1.5, 2, 16, k_span)
    want_svmr = np.concatenate([np.broadcast_to(v2i[gt][:, None, None], (n_queries, k_span, 1)), sv], axis=-1)
    assert_ranked_equal(preds_array(res["SVMR"]), want_svmr.astype(np.float64), score_rtol=1e-3, tie_rtol=2e-4)


def test_bench_scale_parity():
    """BASELINE configs[2] on one GPU, the configuration bench.py times (21,793 videos x 10,000 queries, L<=128,
    H=768): the two-pass search of the full query block, checked on 64 sampled queries against (a) the reference
    arithmetic in float64 over the whole corpus (top-100 videos, top-200 moments), (b) the one-pass exact kernel
    (bit-equal), (c) the candidate-superset property of the filter pass (no overflow, exact top-100 inside)."""
    import bench
    from tests import rank_check as R
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.synthetic import corpus_lengths, synthetic_queries
    args = bench.parse_args([])
    cfg = bench.model_config(args)
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    weights = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    ctx, _ = bench.encode_corpus_shard(model, args, lens, 0, args.n_videos, torch.device(DEV))
    index = CorpusIndex.from_ctx_info(ctx, precision=args.precision)
    qf_cpu, qm_cpu = synthetic_queries(args.n_queries, 30, 768)
    qf, qm = qf_cpu.to(DEV), qm_cpu.to(DEV)
    two = VCMRSearcher(model, index)
    assert two.two_pass
    two.debug = {}
    full = two.search(qf, qm)                      # the benchmarked call: one block of 10,000 queries
    cand = two.debug["cand"]
    assert int(cand.n_flagged) == 0                # no candidate list overflowed
    n_cand = (cand.col >= 0).sum(1)
    sample = torch.arange(0, args.n_queries, args.n_queries // 64, device=DEV)[:64]
    # (a) float64 reference over the whole corpus
    vr64, st64, ed64 = R.fp64_scores(dict(cfg), weights, ctx, qf_cpu[sample.cpu()], qm_cpu[sample.cpu()], device=DEV,
                                     chunk=512)
    stats = R.check_search_result(vr64, st64, ed64, full.top_video_idx[sample], full.top_video_score[sample],
                                  full.span_flat_idx[sample], full.span_score[sample], args.max_ctx_l)
    print("\nbench-scale rank parity vs float64 (64 of %d queries, %d videos): %s; candidates per query %d..%d"
          % (args.n_queries, args.n_videos, R.fmt(stats), int(n_cand.min()), int(n_cand.max())))
    R.assert_within(stats, RANK_TAU, "bench scale")
    del vr64, st64, ed64
    # (b) one-pass exact kernel over all pairs == two-pass, on the sampled queries as a block of their own (the small
    # block runs its 64-row query_linear on the exact-fp32 SIMT kernel instead of the tensor cores, so its span
    # scores differ from the full block's in the last bits; the video lists do not depend on that layer)
    one, two_s = VCMRSearcher(model, index, two_pass=False), VCMRSearcher(model, index, two_pass=True)
    one.packed_min_queries = two_s.packed_min_queries = 0
    want, got = one.search(qf[sample], qm[sample]), two_s.search(qf[sample], qm[sample])
    for name in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name
    for name in ("top_video_idx", "top_video_score"):
        assert torch.equal(getattr(full, name)[sample], getattr(want, name)), name
    torch.testing.assert_close(full.span_score[sample], want.span_score, rtol=1e-4, atol=0)
    # (c) the exact top-100 of every sampled query lies inside its candidate list
    ids = cand.ids[sample].long()
    hit = (ids.unsqueeze(1) == want.top_video_idx.long().unsqueeze(2)).any(2)
    assert bool(hit.all())


def test_tvr_shape_video_only_svmr():
    """BASELINE.json configs[0]: video-only resnet (Dv=2048), 10 videos x 10 queries, L=32, H=768, SVMR-only."""
    from tvretrieval_b200 import inference as I
    cfg, model, weights, ds = tvr_case("video", 10, 10, 768, 32, 2048, seed=1234)
    case = dict(ctx_bsz=200, q_bsz=100, max_before_nms=200, max_n_videos=100)
    opt = Opt(case, cfg)
    ctx = I.compute_context_info(model, ds, opt)
    res = I.compute_query2ctx_info_svmr_only(model, ds, opt, ctx, max_before_nms=200)
    with torch.no_grad():
        octx = O.context_info(cfg, weights, oracle_batches(ds, 200))
        gt = torch.as_tensor([int(q["vid_name"].split("_")[1]) for q in ds.query_data])
        qf, qm = GoldenCase.pad(ds.query_feats)
        _, st, ed = O.pred_from_raw_query(cfg, weights, qf, qm, octx["video_feat1"][gt], octx["video_feat2"][gt],
                                          octx["video_mask"][gt], None, None, None, cross=False)
    sv = O.svmr_from_probs(torch.softmax(st, -1).numpy(), torch.softmax(ed, -1).numpy(), 1.5, 2, 16, 200)
    v2i = np.asarray([ds.video2idx[v["vid_name"]] for v in ds.video_data])
    want = np.concatenate([np.broadcast_to(v2i[gt.numpy()][:, None, None], (10, 200, 1)), sv], axis=-1)
    got = preds_array(res["SVMR"])
    assert_ranked_equal(got, want.astype(np.float64), score_rtol=1e-3)
    # zero-score tail (fewer than 200 in-band cells for short videos): same canonical order as the oracle
    tail = want[..., 3] == 0
    assert tail.any() and np.array_equal(got[..., :3][tail], want[..., :3].astype(np.float64)[tail])
    assert (got[..., 3][tail] == 0).all()


@pytest.mark.parametrize("precision,max_cand", [("f16x3", None), ("f16x3", 100), ("bf16x3", None)])
def test_two_pass_search_equals_one_pass(precision, max_cand):
    """VCMRSearcher(two_pass=True) -- hi-only filter over the corpus, exact re-scoring of the candidates, in-kernel
    fallback for overflowed rows -- returns bit-identical results to the one-pass split-precision search."""
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths, synthetic_queries
    cfg, model, weights, ds = tvr_case("video_sub", 4, 4, 256, 64, 3072, seed=3)
    n_videos, nq = 1300, 300
    lens = corpus_lengths(n_videos, 64, seed=11)
    with torch.no_grad():
        video, sub, mask = corpus_batch(lens, 0, n_videos, 3072, 768, DEV, seed=11, video_split=2048)
        v1, v2, s1, s2 = model.encode_context(video, mask, sub, mask)
        index = CorpusIndex(v1, v2, mask, s1, s2, mask, precision=precision)
        qf, qm = synthetic_queries(nq, 30, 768, seed=12)
        qf, qm = qf.to(DEV), qm.to(DEV)
        kw = dict(max_n_videos=100, max_before_nms=200, query_chunk=256)
        want = VCMRSearcher(model, index, two_pass=False, **kw).search(qf, qm)
        two = VCMRSearcher(model, index, two_pass=True, max_candidates=max_cand, **kw)
        assert two.two_pass
        got = two.search(qf, qm)
    for name in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name
    assert VCMRSearcher(model, index, **kw).two_pass == (precision == "f16x3")  # automatic choice
    if max_cand is None:
        # the optional operand layouts (k-blocked corpus copy for the re-scoring kernel, shared-memory image of
        # f2cat fetched by plain bulk copies) and the host-buffer entry point in several pieces per block
        # (per-piece filter passes behind the uploads): the same bits
        os.environ["XMLB_F2_IMAGE"] = "1"
        try:
            alt = CorpusIndex(v1, v2, mask, s1, s2, mask, precision=precision, rescore_kblocked=True)
        finally:
            del os.environ["XMLB_F2_IMAGE"]
        assert alt.f2cat[0].dim() == 4 and alt.video_tc_kb[0].dim() == 3 and index.video_tc_kb is None
        searcher = VCMRSearcher(model, alt, two_pass=True, encode_chunk=64, **kw)
        searcher.min_piece = 32
        assert len(searcher._piece_bounds(256, True)) >= 4
        host = searcher.search_host(qf.cpu().pin_memory(), qm.cpu().pin_memory())
        for name in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score"):
            assert (torch.from_numpy(host[name]) == getattr(want, name).cpu()).all(), name


@pytest.mark.parametrize("precision", ["f16x3", "f32"])
def test_index_grows_and_round_trips_through_a_file(precision, tmp_path):
    """CorpusIndex.add_videos (incremental add) and save / load: searching the grown or reloaded index returns
    exactly what an index built from all videos at once returns."""
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths, synthetic_queries
    cfg, model, weights, ds = tvr_case("video_sub", 4, 4, 128, 64, 3072, seed=3)
    n_videos, n_first, nq = 700, 450, 90
    lens = corpus_lengths(n_videos, 64, seed=21)
    lens[5] = 64
    with torch.no_grad():
        video, sub, mask = corpus_batch(lens, 0, n_videos, 3072, 768, DEV, seed=21, video_split=2048)
        feats = model.encode_context(video, mask, sub, mask)
        parts = lambda lo, hi, w: (feats[0][lo:hi, :w], feats[1][lo:hi, :w], mask[lo:hi, :w], feats[2][lo:hi, :w],  # noqa: E731
                                   feats[3][lo:hi, :w], mask[lo:hi, :w])
        full = CorpusIndex(*parts(0, n_videos, 64), precision=precision)
        grown = CorpusIndex(*parts(0, n_first, 64), precision=precision)
        w = int(lens[n_first:].max())  # the added batch is narrower than the index, as a later context batch can be
        grown.add_videos(*[t.contiguous() for t in parts(n_first, n_videos, w)])
        assert grown.n_videos == n_videos
        path = grown.save(str(tmp_path / "corpus.xmlb"))
        loaded = CorpusIndex.load(path, DEV)
        assert loaded.nbytes() == grown.nbytes()
        qf, qm = synthetic_queries(nq, 30, 768, seed=22)
        qf, qm = qf.to(DEV), qm.to(DEV)
        gt = torch.randint(0, n_videos, (nq,), generator=torch.Generator().manual_seed(2)).to(torch.int32).to(DEV)
        kw = dict(max_n_videos=100, max_before_nms=200)
        tasks = ("VCMR", "VR", "SVMR")
        want = VCMRSearcher(model, full, **kw).search(qf, qm, gt, tasks)
        for index in (grown, loaded):
            got = VCMRSearcher(model, index, **kw).search(qf, qm, gt, tasks)
            for name in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score", "svmr_flat_idx",
                         "svmr_score"):
                assert torch.equal(getattr(got, name), getattr(want, name)), name
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.xmlb"
        bad.write_bytes(b"not an index")
        CorpusIndex.load(str(bad), DEV)


def test_normalized_corpus_cache_is_not_served_to_a_new_corpus():
    """get_pred_from_raw_query caches the L2-normalised corpus per tensor object: a NEW corpus of the same shape that
    the allocator places at the freed address of the old one (kernel-written tensors all have _version 0) must not
    be scored against the old normalisation."""
    cfg, model, weights, ds = tvr_case("video_sub", 4, 4, 64, 16, 32, seed=3)
    g = torch.Generator().manual_seed(0)
    qf = torch.randn(5, 6, 768, generator=g).to(DEV)
    qm = torch.ones(5, 6, device=DEV)
    mask = torch.ones(7, 16, device=DEV)
    with torch.no_grad():
        outs = []
        for seed in (1, 2):
            gg = torch.Generator().manual_seed(seed)
            v1, s1 = (torch.randn(7, 16, 64, generator=gg).to(DEV) for _ in range(2))
            v2, s2 = (torch.randn(7, 16, 64, generator=gg).to(DEV) for _ in range(2))
            ptr = v1.data_ptr()
            q2c, _, _ = model.get_pred_from_raw_query(qf, qm, v1, v2, mask, s1, s2, mask, cross=True)
            want = O.pred_from_raw_query(dict(cfg), weights, qf.cpu(), qm.cpu(), v1.cpu(), v2.cpu(), mask.cpu(),
                                         s1.cpu(), s2.cpu(), mask.cpu(), cross=True)[0]
            close(q2c, want, rtol=1e-4, atol=2e-6)
            outs.append((ptr, q2c))
            del v1, s1, v2, s2  # the next iteration's tensors typically re-use these addresses
    assert not torch.equal(outs[0][1], outs[1][1])


def test_model_rejects_cpu_tensors():
    from tvretrieval_b200._lib import XmlbError
    cfg, model, weights, ds = tvr_case("video", 4, 4, 64, 16, 32, seed=3)
    with pytest.raises(XmlbError):
        model.cpu().encode_query(torch.randn(2, 5, 768), torch.ones(2, 5))
