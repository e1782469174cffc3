"""GPU parity of the training step (BASELINE config #4), through the C ABI:
  * XML.forward loss, reported parts and the gradient of EVERY parameter against the golden vectors produced by the
    real reference (tests/golden/train_step.npz), same negative sampling under the same torch seed;
  * BertAdam.step (one fused multi-tensor launch pair) against the reference optimizer's golden run
    (tests/golden/bert_adam.npz) and against the oracle on multi-chunk tensors;
  * the dropout kernel: keep rate, scaling, determinism in (seed, index), forward/backward mask identity;
  * a TVR-shaped (H=768, L=128, bsz=32) step against the oracle, and that a few optimizer steps reduce the loss.
Tolerance: gradients rtol 2e-3 of each tensor's largest entry (north_star: 1e-3 relative fp32 on outputs; gradients
accumulate a few more roundings), losses rtol 1e-4."""
import numpy as np
import pytest
import torch

from oracle import xml_oracle as O
from tests.golden_io import CASE_NAMES, TRAIN_VARIANTS, AdamCase, TrainCase

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build_model(cfg, weights):
    from tvretrieval_b200.model_xml import XML, AttrDict
    model = XML(AttrDict(cfg))
    model.load_state_dict(weights)
    return model.to(DEV)


def to_dev(inputs):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in inputs.items()}


def assert_grads_close(model, want, rtol=2e-3, loose=(), loose_rtol=1e-2):
    """`loose`: substrings of parameter names checked at `loose_rtol` instead (see the TVR-shaped test)."""
    got = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    assert set(got) == set(want)
    largest = max(w.abs().max().item() for w in want.values())
    bad = []
    for k, w in want.items():
        g = got[k].detach().cpu().double()
        w = w.double()
        # per-tensor scale, floored for tensors whose gradient is ~0 next to the others (e.g. attention query biases
        # when the span loss weight is 0.01): there only rounding noise is left on both sides
        scale = max(w.abs().max().item(), 1e-4 * largest)
        err = (g - w).abs().max().item()
        if err > (loose_rtol if any(x in k for x in loose) else rtol) * scale:
            bad.append("%s: max |grad diff| %.3e vs largest entry %.3e" % (k, err, scale))
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("variant", TRAIN_VARIANTS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_golden_train_step(name, variant):
    tc = TrainCase(name, variant)
    model = build_model(tc.cfg, tc.weights).eval()  # the goldens were made with dropout off
    torch.manual_seed(tc.seed)
    loss, parts = model(**to_dev(tc.inputs))
    assert loss.dim() == 0 and loss.requires_grad
    assert abs(loss.item() - tc.loss) <= 1e-4 * abs(tc.loss)
    for k, v in tc.parts.items():
        assert isinstance(parts[k], float)
        assert abs(parts[k] - v) <= 1e-4 * max(1e-3, abs(v)), k
    loss.backward()
    assert_grads_close(model, tc.grads)


def test_train_mode_dropout_step_runs_and_is_seeded():
    """train() mode: dropout masks come from torch's default generator, so the same seed gives the same loss and
    gradients, a different seed different ones; every parameter still gets a finite gradient."""
    tc = TrainCase("video_sub_vcmr", "plain")
    model = build_model(tc.cfg, tc.weights).train()
    inputs = to_dev(tc.inputs)

    def run(seed):
        model.zero_grad(set_to_none=True)
        torch.manual_seed(seed)
        loss, _ = model(**inputs)
        loss.backward()
        return loss.item(), {k: p.grad.clone() for k, p in model.named_parameters()}
    l1, g1 = run(5)
    l2, g2 = run(5)
    l3, _ = run(6)
    assert abs(l1 - l2) <= 1e-6 * abs(l1) and abs(l1 - l3) > 1e-6 * abs(l1)
    for k in g1:
        assert torch.isfinite(g1[k]).all(), k
        # (torch's conv1d weight gradient in the recomputed backward accumulates with atomics: allow its jitter)
        torch.testing.assert_close(g1[k], g2[k], rtol=1e-4, atol=1e-7, msg=k)
    model.eval()
    torch.manual_seed(tc.seed)
    assert abs(model(**inputs)[0].item() - tc.loss) <= 1e-4 * abs(tc.loss)  # eval mode: dropout is the identity


def test_dropout_kernel():
    from tvretrieval_b200 import ops
    x = torch.randn(1 << 20, device=DEV) + 3.0
    for p in (0.1, 0.5):
        y = ops.dropout(x, p, seed=1234)
        kept = y != 0
        assert abs(kept.float().mean().item() - (1 - p)) < 3e-3
        torch.testing.assert_close(y[kept], x[kept] / (1 - p), rtol=1e-6, atol=0)
        assert torch.equal(y, ops.dropout(x, p, seed=1234))
        assert not torch.equal(kept, ops.dropout(x, p, seed=1235) != 0)
        # the mask is a function of the GLOBAL element index: a chunk computed with index0 equals the slice
        assert torch.equal(ops.dropout(x[1000:5003].clone(), p, seed=1234, index0=1000), y[1000:5003])
        mask = ops.dropout_mask(x.shape, p, 1234, x.device)
        torch.testing.assert_close(y, x * mask, rtol=1e-6, atol=0)
    # neighbouring elements are not correlated: P(keep i and keep i+1) = (1 - p)^2
    k = (ops.dropout_mask((1 << 20,), 0.5, 7, x.device) != 0).float()
    assert abs((k[:-1] * k[1:]).mean().item() - 0.25) < 3e-3
    # backward applies the same mask
    xg = x.clone().requires_grad_(True)
    ops.dropout(xg, 0.5, seed=99).sum().backward()
    torch.testing.assert_close(xg.grad, ops.dropout_mask(x.shape, 0.5, 99, x.device))
    assert ops.dropout(x, 0.0, seed=1) is x


def test_attention_dropout_matches_mask():
    """xmlb_attention_train: softmax(QK^T) * mask(seed) . V with the same mask the backward recomputation uses."""
    from tvretrieval_b200 import autograd, ops
    g = torch.Generator().manual_seed(3)
    n, lq, lk, hid, heads = 5, 7, 9, 32, 4
    q, k, v = (torch.randn(n, l, hid, generator=g).to(DEV) for l in (lq, lk, lk))
    mask = (torch.rand(n, 1, lk, generator=g) > 0.2).float().to(DEV)
    got = ops.attention(q, k, v, mask, heads, dropout_p=0.3, seed=77)
    want = autograd.t_attention(q, k, v, mask, heads, dropout_p=0.3, seed=77)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    # chunked batches (max_batch) see the same global mask
    got2 = ops.attention(q, k, v, mask, heads, max_batch=2, dropout_p=0.3, seed=77)
    torch.testing.assert_close(got2, got, rtol=0, atol=0)


def adam_groups(ac, params):
    return [{"params": [params[k] for k in ac.names if ac.weight_decay(k) > 0], "weight_decay": 0.01},
            {"params": [params[k] for k in ac.names if ac.weight_decay(k) == 0], "weight_decay": 0.0}]


def test_bert_adam_golden():
    from tvretrieval_b200.optimization import BertAdam
    ac = AdamCase()
    params = {k: torch.nn.Parameter(v.to(DEV)) for k, v in ac.initial().items()}
    opt = BertAdam(adam_groups(ac, params), **ac.hyper)
    assert opt.get_lr() == [0]
    for step in range(ac.n_steps):
        for k in ac.names:
            g = ac.grad(step, k)
            params[k].grad = None if g is None else g.to(DEV)
        opt.step()
        for k in ac.names:
            torch.testing.assert_close(params[k].detach().cpu(), ac.t("p/%d/%s" % (step, k)), rtol=1e-5, atol=2e-7)
            if params[k].grad is not None:  # clipped in place, like clip_grad_norm_
                torch.testing.assert_close(params[k].grad.cpu(), ac.t("g_after/%d/%s" % (step, k)), rtol=1e-5,
                                           atol=1e-8)
        np.testing.assert_allclose(opt.get_lr(), ac.z["lr/%d" % step], rtol=1e-12)
    for k in ac.names:
        st = opt.state[params[k]]
        torch.testing.assert_close(st["next_m"].cpu(), ac.t("m/" + k), rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(st["next_v"].cpu(), ac.t("v/" + k), rtol=1e-5, atol=1e-10)


@pytest.mark.parametrize("max_grad_norm", [1.0, -1.0])
def test_bert_adam_large_tensors_vs_oracle(max_grad_norm):
    """Many chunks per tensor (norm reduction across CTAs), odd sizes, unaligned views."""
    from tvretrieval_b200.optimization import BertAdam
    g = torch.Generator().manual_seed(8)
    shapes = [(768, 3072), (768,), (100003,), (1, 1, 5), (8192,), (8193,)]
    hyper = dict(lr=3e-3, warmup=0.1, t_total=20, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 max_grad_norm=max_grad_norm)
    cpu = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
    params = [torch.nn.Parameter(c.clone().to(DEV)) for c in cpu]
    opt = BertAdam(params, weight_decay=0.01, **hyper)
    states = [{} for _ in shapes]
    for step in range(3):
        for i, (c, p) in enumerate(zip(cpu, params)):
            grad = torch.randn(*c.shape, generator=g) * (0.001 if i % 2 else 0.05)
            p.grad = grad.clone().to(DEV)
            O.bert_adam_step(c, grad, states[i], lr=hyper["lr"], weight_decay=0.01, schedule="warmup_linear",
                             warmup=0.1, t_total=20, b1=0.9, b2=0.999, e=1e-6, max_grad_norm=max_grad_norm)
        opt.step()
        for c, p in zip(cpu, params):
            torch.testing.assert_close(p.detach().cpu(), c, rtol=1e-5, atol=1e-7)


def test_tvr_shaped_step_vs_oracle_and_training_reduces_loss():
    """H=768, Dv=3072, L<=128, bsz=32 (config #4 dims at a batch the CPU oracle differentiates in seconds)."""
    from tvretrieval_b200.model_xml import XML, AttrDict, xml_base_config
    from tvretrieval_b200.optimization import BertAdam
    cfg = dict(xml_base_config, hidden_size=768, visual_input_size=3072, max_ctx_l=128, max_desc_l=30,
               use_hard_negative=True, hard_pool_size=20, lw_st_ed=0.01, drop=0.0, input_drop=0.0)
    torch.manual_seed(2018)
    model = XML(AttrDict(cfg)).to(DEV).train()
    weights = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    n = 32

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(16, 129, (n,), generator=g)
        lens[0] = 128
        qlens = torch.randint(5, 31, (n,), generator=g)
        video_mask = (torch.arange(128)[None] < lens[:, None]).float()
        query_mask = (torch.arange(30)[None] < qlens[:, None]).float()
        unit = lambda t: t / (t.norm(dim=-1, keepdim=True) + 1e-5)  # noqa: E731
        video = unit(torch.randn(n, 128, 3072, generator=g)) * video_mask[..., None]
        sub = unit(torch.randn(n, 128, 768, generator=g)) * video_mask[..., None]
        query = unit(torch.randn(n, 30, 768, generator=g)) * query_mask[..., None]
        st = (torch.rand(n, generator=g) * (lens - 1)).long()
        ed = torch.minimum(lens - 1, st + 3)
        return dict(query_feat=query, query_mask=query_mask, video_feat=video, video_mask=video_mask, sub_feat=sub,
                    sub_mask=video_mask, tef_feat=None, tef_mask=None, st_ed_indices=torch.stack([st, ed], 1))

    inputs = batch(1234)
    query, query_mask, video, video_mask, sub = (inputs[k] for k in ("query_feat", "query_mask", "video_feat",
                                                                     "video_mask", "sub_feat"))
    dev_inputs = to_dev(inputs)

    # The sampled negative is "the video at rank r of the sorted row" (model_xml.py:617-622).  Random-init scores
    # cluster in [0.03, 0.15] (smallest in-batch gap ~1e-6) while two fp32 implementations of the encoders differ by
    # ~1e-5 there, so a few neighbouring ranks are swapped between them.  Gradient parity is only defined when both
    # sample the SAME negatives: pick the sampling seed for which they do.
    # The oracle is evaluated in float64: measured on the B200 box, torch's fp32 CPU kernels are themselves off by
    # 2-17 % (against float64) on the four video_input_proj gradients of this batch, while the kernels here are
    # within 2e-6 of float64 on every tensor.
    f64 = lambda t: t.double()  # noqa: E731
    w = {k: v.double().requires_grad_(True) for k, v in weights.items()}
    query64, query_mask64, video64, video_mask64, sub64 = map(f64, (query, query_mask, video, video_mask, sub))
    with torch.no_grad():
        v1, v2, s1, s2 = O.encode_context(cfg, w, video64, video_mask64, sub64, video_mask64)
        q2c_cpu = O.pred_from_raw_query(cfg, w, query64, query_mask64, v1, v2, video_mask64, s1, s2, video_mask64,
                                        cross=False)[0].float()
        d = dev_inputs
        ctx = model.encode_context(d["video_feat"], d["video_mask"], d["sub_feat"], d["sub_mask"])
        q2c_gpu = model.get_pred_from_raw_query(d["query_feat"], d["query_mask"], ctx[0], ctx[1], d["video_mask"],
                                                ctx[2], ctx[3], d["sub_mask"], cross=False)[0].cpu()
    torch.testing.assert_close(q2c_gpu, q2c_cpu, rtol=1e-4, atol=2e-5)

    def sampled(q2c, seed):
        torch.manual_seed(seed)
        out = []
        for m in (q2c, q2c.t()):
            masked = m.clone()
            masked.fill_diagonal_(999)
            order = torch.sort(masked, descending=True, dim=1)[1]
            pick = torch.randint(1, min(1 + cfg["hard_pool_size"], n), size=(n,))
            out.append(order[torch.arange(n), pick])
        return torch.stack(out)

    sample_seed = next(s for s in range(11, 111) if torch.equal(sampled(q2c_gpu, s), sampled(q2c_cpu, s)))
    torch.manual_seed(sample_seed)
    o_loss, o_parts = O.train_forward(cfg, w, query64, query_mask64, video64, video_mask64, sub64, video_mask64,
                                      inputs["st_ed_indices"])
    o_loss.backward()
    want = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in w.items()}
    assert model.train_precision == "f16x3"
    for prec in ("f32", "f16x3"):
        model.train_precision = prec
        model.zero_grad(set_to_none=True)
        torch.manual_seed(sample_seed)
        loss, parts = model(**dev_inputs)
        loss.backward()
        assert abs(loss.item() - o_loss.item()) <= 1e-4 * abs(o_loss.item())
        for k in o_parts:
            assert abs(parts[k] - o_parts[k]) <= 1e-4 * max(1e-3, abs(o_parts[k])), k
        # Every gradient within 1e-3 of float64 (relative to the tensor's largest entry).  Exception for the default
        # tensor-core mode: the input projections, whose ReLU masks flip for the few pre-activations that lie within
        # the forward rounding error of zero -- a property of ANY fp32-accurate forward pass (measured on this very
        # batch: torch's own CPU fp32 is off by 17 % on video_input_proj, the tensor-core mode by 0.4 % on
        # sub_input_proj, the exact-fp32 kernels by 0.01 %; profiles/r02_train_grad_accuracy.txt).
        assert_grads_close(model, want, rtol=1e-3, loose=("input_proj",) if prec == "f16x3" else ())
    # a few fused optimizer steps on the same batch reduce the loss
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    named = list(model.named_parameters())
    opt = BertAdam([{"params": [p for k, p in named if not any(nd in k for nd in no_decay)], "weight_decay": 0.01},
                    {"params": [p for k, p in named if any(nd in k for nd in no_decay)], "weight_decay": 0.0}],
                   lr=1e-4, warmup=0.01, t_total=100, schedule="warmup_linear")
    first = None
    for it in range(6):
        opt.zero_grad()
        torch.manual_seed(11)
        loss, _ = model(**dev_inputs)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < first
