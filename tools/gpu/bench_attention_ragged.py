#!/usr/bin/env python
"""Times xmlb_attention_ragged_qkv alone on the bench's query set (10 K queries, TVR lengths, 4 heads of 192)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tvretrieval_b200 import ops  # noqa: E402
from tvretrieval_b200.synthetic import synthetic_queries  # noqa: E402


def main(n=10000, hid=768, heads=4):
    dev = torch.device("cuda", 0)
    _, qm = synthetic_queries(n, 30, 16)
    lens = (qm != 0).sum(1).to(torch.int32)
    cu = torch.zeros(n + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    t_rows = int(cu[-1])
    torch.manual_seed(0)
    qkv = torch.randn(t_rows, 3 * hid, device=dev)
    cu = cu.to(dev)
    out = ops.attention_ragged_qkv(qkv, cu, int(lens.max()), heads)
    # reference: per-sequence softmax attention in float64 on a sample of the sequences
    worst = 0.0
    for s in range(0, n, 997):
        lo, hi = int(cu[s]), int(cu[s + 1])
        x = qkv[lo:hi].double().view(hi - lo, 3, heads, hid // heads)
        q, k, v = x[:, 0].transpose(0, 1), x[:, 1].transpose(0, 1), x[:, 2].transpose(0, 1)
        p = torch.softmax(q @ k.transpose(1, 2) / (hid // heads) ** 0.5, -1)
        ref = (p @ v).transpose(0, 1).reshape(hi - lo, hid)
        worst = max(worst, float((out[lo:hi].double() - ref).abs().max()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        ops.attention_ragged_qkv(qkv, cu, int(lens.max()), heads)
    b.record()
    torch.cuda.synchronize()
    print("attention_ragged_qkv: %d tokens of %d queries, %.3f ms per call, max abs error vs float64 %.2e"
          % (t_rows, n, a.elapsed_time(b) / 20, worst))


if __name__ == "__main__":
    main()
