#!/usr/bin/env python
"""Accuracy of the training-step gradients at TVR dims (H=768, Dv=3072, L<=128, bsz=32): the kernels (exact-fp32
linears) and torch's own fp32 CPU evaluation of the same arithmetic, both against a float64 evaluation.
Test infrastructure (it imports the oracle): run on the GPU box, output kept in profiles/r01_train_grad_accuracy.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xml_oracle as O
from tvretrieval_b200.model_xml import XML, AttrDict, xml_base_config
cfg = dict(xml_base_config, hidden_size=768, visual_input_size=3072, max_ctx_l=128, max_desc_l=30,
           use_hard_negative=True, hard_pool_size=20, lw_st_ed=0.01, drop=0.0, input_drop=0.0)
torch.manual_seed(2018)
model = XML(AttrDict(cfg)).to('cuda').train()
weights = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
n = 32
g = torch.Generator().manual_seed(1234)
lens = torch.randint(16, 129, (n,), generator=g); lens[0] = 128
qlens = torch.randint(5, 31, (n,), generator=g)
video_mask = (torch.arange(128)[None] < lens[:, None]).float()
query_mask = (torch.arange(30)[None] < qlens[:, None]).float()
unit = lambda t: t / (t.norm(dim=-1, keepdim=True) + 1e-5)
video = unit(torch.randn(n, 128, 3072, generator=g)) * video_mask[..., None]
sub = unit(torch.randn(n, 128, 768, generator=g)) * video_mask[..., None]
query = unit(torch.randn(n, 30, 768, generator=g)) * query_mask[..., None]
st = (torch.rand(n, generator=g) * (lens - 1)).long(); ed = torch.minimum(lens - 1, st + 3)
sted = torch.stack([st, ed], 1)
def run_oracle(dev, dt):
    w = {k: v.to(dev, dt).requires_grad_(True) for k, v in weights.items()}
    torch.manual_seed(11)
    c = lambda t: t.to(dev, dt)
    loss, parts = O.train_forward(cfg, w, c(query), c(query_mask), c(video), c(video_mask), c(sub), c(video_mask), sted.to(dev))
    loss.backward()
    return loss.item(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)).double().cpu() for k, v in w.items()}
l64, g64 = run_oracle('cuda', torch.float64)
l32c, g32c = run_oracle('cpu', torch.float32)
inputs = dict(query_feat=query.cuda(), query_mask=query_mask.cuda(), video_feat=video.cuda(), video_mask=video_mask.cuda(),
              sub_feat=sub.cuda(), sub_mask=video_mask.cuda(), tef_feat=None, tef_mask=None, st_ed_indices=sted.cuda())
grads, losses = {}, {}
for prec in ("f32", "f16x3"):
    model.train_precision = prec
    model.zero_grad(set_to_none=True)
    torch.manual_seed(11)
    loss, _ = model(**inputs)
    loss.backward()
    grads[prec] = {k: p.grad.double().cpu() for k, p in model.named_parameters()}
    losses[prec] = loss.item()
print("loss: float64 %.8f  torch CPU fp32 %.8f  kernels f32 %.8f  kernels f16x3 %.8f" % (l64, l32c, losses["f32"], losses["f16x3"]))
worst = {"cpu": 0.0, "f32": 0.0, "f16x3": 0.0}
for k in g64:
    s = g64[k].abs().max().item()
    e = lambda d: (d[k] - g64[k]).abs().max().item() / max(s, 1e-12)
    errs = (e(g32c), e(grads["f32"]), e(grads["f16x3"]))
    for name, v in zip(worst, errs):
        worst[name] = max(worst[name], v)
    print("%-48s max |grad| %.2e   max err / max |grad|:  torch CPU fp32 %.1e   kernels f32 %.1e   kernels f16x3 %.1e" % ((k, s) + errs))
print("worst over all tensors: torch CPU fp32 %.1e   kernels f32 %.1e   kernels f16x3 %.1e" % (worst["cpu"], worst["f32"], worst["f16x3"]))
# forward + backward time at the reference batch size (128), CUDA events, 5 repeats after 2 warm-ups
import time
rep = lambda t: t.repeat(4, *([1] * (t.dim() - 1)))
big = {k: (rep(v) if torch.is_tensor(v) else v) for k, v in inputs.items()}
for prec in ("f32", "f16x3"):
    model.train_precision = prec
    ts = []
    for i in range(7):
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss, _ = model(**big); loss.backward()
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("forward + backward, bsz 128, train_precision=%s: %.1f ms" % (prec, 1e3 * sorted(ts[2:])[len(ts[2:]) // 2]))
