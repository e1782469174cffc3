#!/usr/bin/env python
"""One training step (bsz 128, TVR dims) bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def main():
    args = bench.parse_args(["--workload", "train"] + sys.argv[1:])
    from tvretrieval_b200.model_xml import XML
    dev = torch.device("cuda", 0)
    cfg = bench.model_config(args)
    cfg.update(use_hard_negative=True, hard_pool_size=20, lw_st_ed=0.01)
    torch.manual_seed(2018)
    model = XML(cfg).to(dev).train()
    n, L = args.train_bsz, args.max_ctx_l
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(16, L + 1, (n,), generator=g)
    qlens = torch.randint(5, 31, (n,), generator=g)
    vm = (torch.arange(L)[None] < lens[:, None]).float()
    qm = (torch.arange(30)[None] < qlens[:, None]).float()
    st = (torch.rand(n, generator=g) * (lens - 1)).long()
    d = dict(query_feat=torch.randn(n, 30, 768, generator=g) * qm[..., None], query_mask=qm,
             video_feat=torch.randn(n, L, args.video_dim, generator=g) * vm[..., None], video_mask=vm,
             sub_feat=torch.randn(n, L, 768, generator=g) * vm[..., None], sub_mask=vm, tef_feat=None, tef_mask=None,
             st_ed_indices=torch.stack([st, torch.minimum(lens - 1, st + 3)], 1))
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}
    for _ in range(3):
        model.zero_grad()
        loss, _ = model(**d)
        loss.backward()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.zero_grad()
    loss, _ = model(**d)
    loss.backward()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
