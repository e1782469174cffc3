"""The drop-in claim of INTEGRATION.md section A, tested against the reference's OWN unmodified code (vendored into
baseline/_ref by tools/vendor_reference.py): its inference drivers run with this repo's XML class / instances on the
GPU and must reproduce what they produced with the reference's XML on the CPU (the committed golden vectors).
The CPU-only test checks that the vendored copy is the code that made those vectors."""
import numpy as np
import pytest
import torch

from tests import reference_loader as RL
from tests.golden_io import GoldenCase

needs_ref = pytest.mark.skipif(not RL.available(), reason="baseline/_ref not vendored (tools/vendor_reference.py)")


def make_opt(ref, g, device):
    c, cfg = g.case, g.cfg
    return ref.EasyDict(eval_context_bsz=c["ctx_bsz"], eval_query_bsz=c["q_bsz"], num_workers=0, pin_memory=False,
                        device=torch.device(device), ctx_mode=cfg["ctx_mode"], external_inference_vr_res_path=None,
                        q2c_alpha=20.0, min_pred_l=2, max_pred_l=16, max_ctx_l=cfg["max_ctx_l"], clip_length=1.5,
                        debug=False, device_ids=[0])


def dataset(g):
    from tvretrieval_b200.synthetic import SyntheticEvalDataset
    return SyntheticEvalDataset(max_ctx_l=g.cfg["max_ctx_l"], max_desc_l=g.cfg["max_desc_l"],
                                video_dim=g.cfg["visual_input_size"], sub_dim=g.cfg["sub_input_size"],
                                query_dim=g.cfg["query_input_size"], ctx_mode=g.cfg["ctx_mode"], min_ctx_l=3,
                                **g.case["data"])


def preds_array(lst):
    k = max(len(e["predictions"]) for e in lst)
    out = np.zeros((len(lst), k, 4))
    for i, e in enumerate(lst):
        out[i, :len(e["predictions"])] = np.asarray(e["predictions"], dtype=np.float64).reshape(-1, 4)
    return out


def run_reference_drivers(ref, model, g, device):
    ds, opt, c = dataset(g), make_opt(ref, g, device), g.case
    with torch.no_grad():
        ctx = ref.inference.compute_context_info(model, ds, opt)
        return ref.inference.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=c["max_before_nms"],
                                                    max_n_videos=c["max_n_videos"], tasks=tuple(c["tasks"]))


def positive_rows_equal(got, want, rtol):
    """[video_idx, st, ed, score] rows; the reference's order among exact ties (the structural zeros) is unspecified
    (SURVEY.md Appendix B-7), so only the strictly positive prefix is compared row by row."""
    assert got.shape == want.shape
    np.testing.assert_allclose(got[..., 3], want[..., 3], rtol=rtol, atol=1e-12)
    pos = want[..., 3] > 0
    assert np.array_equal(got[..., :3][pos], want[..., :3][pos])


@needs_ref
def test_vendored_reference_reproduces_the_goldens_on_cpu():
    ref = RL.load()
    g = GoldenCase("video_sub_vcmr")
    model = ref.XML(ref.EasyDict(g.cfg))
    model.load_state_dict(g.weights)
    res = run_reference_drivers(ref, model.eval(), g, "cpu")
    for task in g.case["tasks"]:
        positive_rows_equal(preds_array(res[task]), g.z["res/" + task], rtol=1e-6)


@needs_ref
@pytest.mark.gpu
def test_reference_drivers_run_on_the_dropin_model():
    """reference inference.compute_context_info + compute_query2ctx_info, unmodified, driving tvretrieval_b200's XML on
    cuda: encode_context and get_pred_from_raw_query(cross=True) are served by the kernels, everything else
    (exp / softmax / topk / gather / einsum / band mask / full sort / host lists) is the reference's own torch code."""
    from tvretrieval_b200.model_xml import XML, AttrDict
    ref = RL.load()
    g = GoldenCase("video_sub_vcmr")
    model = XML(AttrDict(g.cfg))
    model.load_state_dict(g.weights)
    res = run_reference_drivers(ref, model.to("cuda").eval(), g, "cuda")
    for task in g.case["tasks"]:
        positive_rows_equal(preds_array(res[task]), g.z["res/" + task], rtol=1e-4)


@needs_ref
@pytest.mark.gpu
def test_reference_setup_model_with_swapped_class(tmp_path, monkeypatch):
    """INTEGRATION.md section A literally: the one-line swap `XML = tvretrieval_b200.model_xml.XML` inside the
    reference's inference module; its own setup_model() then loads a reference-format checkpoint into the drop-in
    class, and its own eval path runs on it."""
    import tvretrieval_b200.model_xml as ours
    ref = RL.load()
    g = GoldenCase("video_sub_vcmr")
    ref_model = ref.XML(ref.EasyDict(g.cfg))
    ref_model.load_state_dict(g.weights)
    ckpt = tmp_path / "model.ckpt"
    torch.save(dict(model=ref_model.state_dict(), model_cfg=dict(g.cfg), epoch=3), str(ckpt))  # train.py:219-223
    monkeypatch.setattr(ref.inference, "XML", lambda cfg: ours.XML(ours.AttrDict(cfg)))
    opt = make_opt(ref, g, "cuda")
    opt.ckpt_filepath = str(ckpt)
    model = ref.inference.setup_model(opt)
    assert isinstance(model, ours.XML) and next(model.parameters()).is_cuda
    res = run_reference_drivers(ref, model.eval(), g, "cuda")
    positive_rows_equal(preds_array(res["VCMR"]), g.z["res/VCMR"], rtol=1e-4)
