#!/usr/bin/env python
"""Diagnostic (not a pytest file): repeats the TVR-shaped driver comparison of tests/test_gpu_model.py in one process
and prints hashes of the kernels' outputs and of the CPU oracle's, to tell which side varies from run to run."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xml_oracle as O  # noqa: E402
from tests.test_gpu_model import Opt, oracle_batches, preds_array, tvr_case  # noqa: E402
from tests.golden_io import GoldenCase  # noqa: E402
from tvretrieval_b200 import inference as I  # noqa: E402

h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]  # noqa: E731
cfg, model, weights, ds = tvr_case("video_sub", 150, 24, 768, 128, 3072, seed=1234)
case = dict(ctx_bsz=64, q_bsz=10, max_before_nms=200, max_n_videos=100)
opt = Opt(case, cfg)
for it in range(4):
    ctx = I.compute_context_info(model, ds, opt)
    res = I.compute_query2ctx_info(model, ds, opt, ctx, max_before_nms=200, max_n_videos=100, tasks=("VCMR", "VR"))
    with torch.no_grad():
        octx = O.context_info(cfg, weights, oracle_batches(ds, 64))
        qf, qm = GoldenCase.pad(ds.query_feats)
        o = O.query_batch_tensor_section(cfg, weights, octx, qf, qm, q2c_alpha=20.0, max_n_videos=100,
                                         max_before_nms=200, min_pred_l=2, max_pred_l=16)
    print(it, "gpu ctx", h(ctx["video_feat1"].cpu().numpy()), h(ctx["video_feat2"].cpu().numpy()),
          "vcmr", h(preds_array(res["VCMR"])), "vr", h(preds_array(res["VR"])),
          "| oracle ctx", h(octx["video_feat1"].numpy()), "span", h(o["span_score"].numpy()),
          h(o["span_flat_idx"].numpy()), flush=True)
