"""Pin oracle/xml_oracle.py against outputs of the REAL reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import xml_oracle as O
from tests.golden_io import CASE_NAMES, GOLDEN_DIR, TRAIN_VARIANTS, GoldenCase, TrainCase

RTOL, ATOL = 1e-5, 1e-6  # same torch CPU kernels on both sides -> essentially bit-equal


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    torch.testing.assert_close(a.double(), b.double(), rtol=rtol, atol=atol)


@pytest.fixture(scope="module", params=CASE_NAMES)
def case(request):
    return GoldenCase(request.param)


def test_context_encoding(case):
    with torch.no_grad():
        ctx = O.context_info(case.cfg, case.weights, case.context_batches())
    ref = case.ctx()
    for k, v in ref.items():
        if v is None:
            assert ctx[k] is None
        else:
            close(ctx[k], v)


def test_encode_query(case):
    with torch.no_grad():
        vq, sq = O.encode_query(case.cfg, case.weights, case.query_feat, case.query_mask)
    close(vq, case.t("video_query"))
    close(sq, case.t("sub_query"))


@pytest.mark.parametrize("cross", [True, False])
def test_pred_from_raw_query(case, cross):
    ctx = case.ctx()
    if not cross:
        gt = torch.as_tensor(case.query_gt_meta_idx)
        ctx = {k: (None if v is None else v[gt]) for k, v in ctx.items()}
    with torch.no_grad():
        q2c, st, ed = O.pred_from_raw_query(case.cfg, case.weights, case.query_feat, case.query_mask,
                                            ctx["video_feat1"], ctx["video_feat2"], ctx["video_mask"],
                                            ctx["sub_feat1"], ctx["sub_feat2"], ctx["sub_mask"], cross=cross)
    tag = "cross" if cross else "inbatch"
    close(q2c, case.t(tag + "/q2c"))
    close(st, case.t(tag + "/st"))
    close(ed, case.t(tag + "/ed"))


def run_driver(case, canonical):
    """Oracle version of compute_query2ctx_info (tensor + host sections) on the golden ctx."""
    ctx = case.ctx()
    c = case.case
    outs = []
    with torch.no_grad():
        for lo, (qf, qm) in case.query_batches():
            outs.append(O.query_batch_tensor_section(
                case.cfg, case.weights, ctx, qf, qm, q2c_alpha=20.0, max_n_videos=c["max_n_videos"],
                max_before_nms=c["max_before_nms"], min_pred_l=2, max_pred_l=16, canonical_ties=canonical))
    cat = lambda k: torch.cat([o[k] for o in outs]).numpy()  # noqa: E731
    return {k: cat(k) for k in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score")}, outs


@pytest.mark.parametrize("canonical", [True, False])
def test_vcmr_vr_driver(case, canonical):
    if "VCMR" not in case.case["tasks"]:
        pytest.skip("SVMR-only case")
    r, _ = run_driver(case, canonical)
    c = case.case
    # VR: [video_idx, 0, 0, score]
    ref_vr = case.z["res/VR"]
    assert np.array_equal(case.video2idx[r["top_video_idx"]], ref_vr[..., 0].astype(np.int64))
    close(r["top_video_score"], ref_vr[..., 3].astype(np.float32), rtol=1e-6, atol=0)
    # VCMR rows
    dec = O.decode_vcmr(r["span_flat_idx"], r["span_score"], r["top_video_idx"], case.video2idx,
                        c["max_n_videos"], case.cfg["max_ctx_l"], 1.5)
    ref = case.z["res/VCMR"]
    pos = ref[..., 3] > 0  # tie order among exact zeros is unspecified in the reference
    assert pos.sum() > 0.9 * pos.size
    assert np.array_equal(dec[..., :3][pos], ref[..., :3][pos])
    close(dec[..., 3], ref[..., 3], rtol=1e-6, atol=0)
    assert np.array_equal(dec[:, :7][pos[:, :7]], case.z["top7/VCMR"][pos[:, :7]])


def test_svmr(case):
    c = case.case
    ref = case.z["res/SVMR"]
    if "VCMR" in c["tasks"]:
        _, outs = run_driver(case, True)
        st = torch.cat([o["st_prob"] for o in outs]).numpy()
        ed = torch.cat([o["ed_prob"] for o in outs]).numpy()
        rows = np.arange(len(st))
        st, ed = st[rows, case.query_gt_meta_idx], ed[rows, case.query_gt_meta_idx]
    else:  # SVMR-only driver: cross=False on the GT videos, softmax (reference inference.py:139-157)
        st = torch.softmax(case.t("inbatch/st"), -1).numpy()
        ed = torch.softmax(case.t("inbatch/ed"), -1).numpy()
    got = O.svmr_from_probs(st, ed, 1.5, 2, 16, c["max_before_nms"])
    pos = ref[..., 3] > 0
    assert np.array_equal(got[..., :2][pos].astype(np.float64), ref[..., 1:3][pos])
    close(got[..., 2], ref[..., 3].astype(np.float32), rtol=1e-6, atol=0)
    assert np.array_equal(ref[..., 0], np.broadcast_to(case.video2idx[case.query_gt_meta_idx][:, None],
                                                       ref[..., 0].shape))


def test_temporal_nms_known_answers():
    z = np.load(GOLDEN_DIR + "/temporal_nms.npz")
    for i in range(int(z["n_cases"])):
        got = O.temporal_nms(z["in/%d" % i].tolist(), float(z["thd/%d" % i]))
        assert np.array_equal(np.asarray(got, dtype=np.float64).reshape(-1, 3), z["out/%d" % i]), i


def test_vcmr_nms(case):
    if "VCMR" not in case.case["tasks"]:
        pytest.skip("SVMR-only case")
    c = case.case
    ref_in, ref_out, cnt = case.z["res/VCMR"], case.z["nms/VCMR"], case.z["nms/VCMR_count"]
    for q in range(len(ref_in)):
        preds = [[int(p[0]), p[1], p[2], p[3]] for p in ref_in[q].tolist()]
        got = O.vcmr_nms(preds, 0.5, c["max_before_nms"], 20)
        assert len(got) == cnt[q]
        assert np.array_equal(np.asarray(got, dtype=np.float64), ref_out[q, :cnt[q]])


@pytest.mark.parametrize("variant", TRAIN_VARIANTS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_train_forward_loss_and_gradients(name, variant):
    """O.train_forward (XML.forward + losses, model_xml.py:212-251,588-637) against the reference's loss, reported
    floats and the gradient of every parameter (same negative sampling: torch.randint under the same seed)."""
    tc = TrainCase(name, variant)
    w = {k: v.clone().requires_grad_(True) for k, v in tc.weights.items()}
    i = tc.inputs
    torch.manual_seed(tc.seed)
    loss, parts = O.train_forward(tc.cfg, w, i["query_feat"], i["query_mask"], i["video_feat"], i["video_mask"],
                                  i["sub_feat"], i["sub_mask"], i["st_ed_indices"])
    assert abs(loss.item() - tc.loss) <= 1e-5 * abs(tc.loss)
    for k, v in tc.parts.items():
        assert abs(parts[k] - v) <= 1e-5 * max(1e-3, abs(v)), k
    loss.backward()
    assert set(tc.grads) == set(w)
    for k, g in tc.grads.items():
        got = w[k].grad if w[k].grad is not None else torch.zeros_like(w[k])
        close(got, g, rtol=1e-4, atol=1e-7)


def test_bert_adam_steps():
    """O.bert_adam_step (optimization.py:273-338) against the reference optimizer: parameters and clipped gradients
    after every step, final moments, scheduled learning rates."""
    from tests.golden_io import AdamCase
    ac = AdamCase()
    h = ac.hyper
    params, states = ac.initial(), {k: {} for k in ac.names}
    for step in range(ac.n_steps):
        for k in ac.names:
            g = ac.grad(step, k)
            if g is None:
                continue
            O.bert_adam_step(params[k], g, states[k], lr=h["lr"], weight_decay=ac.weight_decay(k),
                             schedule=h["schedule"], warmup=h["warmup"], t_total=h["t_total"], b1=h["b1"], b2=h["b2"],
                             e=h["e"], max_grad_norm=h["max_grad_norm"])
            close(g, ac.t("g_after/%d/%s" % (step, k)), rtol=1e-5, atol=1e-8)
        for k in ac.names:
            close(params[k], ac.t("p/%d/%s" % (step, k)), rtol=1e-5, atol=1e-7)  # 1-ulp differences of the update
    for k in ac.names:
        close(states[k]["next_m"], ac.t("m/" + k), rtol=1e-5, atol=1e-8)
        close(states[k]["next_v"], ac.t("v/" + k), rtol=1e-5, atol=1e-10)
    # get_lr() after the last step: lr * multiplier(step count) per parameter, in parameter-group order
    order = [k for k in ac.names if ac.weight_decay(k) > 0] + [k for k in ac.names if ac.weight_decay(k) == 0]
    want = [h["lr"] * O.lr_multiplier(h["schedule"], h["warmup"], h["t_total"], states[k]["step"]) for k in order]
    np.testing.assert_allclose(ac.z["lr/%d" % (ac.n_steps - 1)], want, rtol=1e-12)


def test_visualization_data():
    """O.visualization_data (XML.get_visualization_data, model_xml.py:253-289) against the reference's output."""
    from tests.golden_io import VisualizationCase
    vc = VisualizationCase()
    tc = vc.train
    i = tc.inputs
    with torch.no_grad():
        got = O.visualization_data(tc.cfg, tc.weights, i["query_feat"], i["query_mask"], i["video_feat"],
                                   i["video_mask"], i["sub_feat"], i["sub_mask"])
    vc.check(got, RTOL, ATOL)
