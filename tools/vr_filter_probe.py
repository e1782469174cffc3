#!/usr/bin/env python
"""Limiter experiments on the hi-only (filter) pass of the packed tcgen05 VR kernel at the bench shape.
XMLB_VR_PROBE bit 0: epilogue only waits/releases; bit 1: producer loads no B tiles; XMLB_VR_STAGES: ring depth.
Needs a library built with the probe knobs compiled in:  python -m tvretrieval_b200.build --force --probes
(and a plain `python -m tvretrieval_b200.build --force` afterwards)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tvretrieval_b200 import ops  # noqa: E402
from tvretrieval_b200.engine import CorpusPacking  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nv, nq, L, H = 21793, 10000, 128, 768
lens = torch.randint(16, L + 1, (nv,), generator=g, device=dev)
mask = (torch.arange(L, device=dev)[None] < lens[:, None]).float()
pk = CorpusPacking(mask)


def halves(rows):
    out = []
    for _ in range(2):
        x = torch.randn(rows, H, device=dev, generator=g)
        out.append(ops.split_rows(x, normalize=True))
        del x
    return out


c, q = halves(pk.n_rows), halves(nq)
flops = 2.0 * H * pk.n_rows * nq * 2
out = torch.empty(nq, nv, device=dev)


def run(tag, hi_only=True, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    ms = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.vr_scores_tc_packed(q[0], c[0], pk, nv, q_b=q[1], c_b=c[1], ordinal=True, hi_only=hi_only, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    for k in env:
        del os.environ[k]
    t = sorted(ms[2:])[1]
    print("%-44s %8.2f ms  %7.1f TFLOP/s (1 MMA per product)" % (tag, t, flops / t / 1e9), flush=True)



ops.FILTER_ON_CTA_PAIRS = False
run("hi-only single-CTA baseline")
run("hi-only, epilogue = wait+release", XMLB_VR_PROBE=1)
run("hi-only, no B loads", XMLB_VR_PROBE=2)
run("hi-only, no B loads, no epilogue", XMLB_VR_PROBE=3)
for s in (6, 4, 3, 2):
    run("hi-only, %d stages" % s, XMLB_VR_STAGES=s)
run("3-term baseline (3 MMAs per product)", hi_only=False)
run("3-term, epilogue = wait+release", hi_only=False, XMLB_VR_PROBE=1)

# CTA-pair kernel (tcgen05 cta_group::2) against the single-CTA hi-only kernel: same values, less operand traffic
ops.FILTER_ON_CTA_PAIRS = False
ref = ops.vr_scores_tc_packed(q[0], c[0], pk, nv, q_b=q[1], c_b=c[1], ordinal=True, hi_only=True).clone()
ops.FILTER_ON_CTA_PAIRS = True
got = ops.vr_scores_tc_packed(q[0], c[0], pk, nv, q_b=q[1], c_b=c[1], ordinal=True, hi_only=True)
torch.cuda.synchronize()
d = (got - ref).abs()
print("pair vs single-CTA hi-only: max |diff| %.3e, identical %.6f" % (d.max().item(), (got == ref).float().mean().item()))
run("hi-only on CTA pairs (cta_group::2)")
