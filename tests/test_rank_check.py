"""CPU checks of the float64 rank discriminator (tests/rank_check.py): the reference's own fp32 arithmetic (the
oracle) must pass it, and a list with a real ranking error must fail it."""
import copy

import numpy as np
import pytest
import torch

from oracle import xml_oracle as O
from tests import rank_check as R


def small_case(n_videos=40, n_queries=6, hidden=64, ctx_l=24, seed=5):
    cfg = dict(merge_two_stream=True, cross_att=True, span_predictor_type="conv", encoder_type="transformer",
               visual_input_size=96, query_input_size=48, sub_input_size=48, hidden_size=hidden, conv_kernel_size=5,
               max_ctx_l=ctx_l, max_desc_l=8, n_heads=4, ctx_mode="video_sub", no_modular=False,
               initializer_range=0.02, stack_conv_predictor_conv_kernel_sizes=-1, conv_stride=1,
               input_drop=0.1, drop=0.1)
    w = O.init_weights(cfg, seed=seed)
    gen = torch.Generator().manual_seed(seed)
    lens = torch.randint(6, ctx_l + 1, (n_videos,), generator=gen)
    lens[0] = ctx_l
    mask = (torch.arange(ctx_l)[None] < lens[:, None]).float()
    video = torch.randn(n_videos, ctx_l, 96, generator=gen) * mask.unsqueeze(2)
    sub = torch.randn(n_videos, ctx_l, 48, generator=gen) * mask.unsqueeze(2)
    with torch.no_grad():
        ctx = O.context_info(cfg, w, [dict(video_feat=video, video_mask=mask, sub_feat=sub, sub_mask=mask)])
    qlen = torch.randint(3, 9, (n_queries,), generator=gen)
    qm = (torch.arange(8)[None] < qlen[:, None]).float()
    qf = torch.randn(n_queries, 8, 48, generator=gen) * qm.unsqueeze(2)
    return cfg, w, ctx, qf, qm


def test_fp32_oracle_is_consistent_with_fp64_order():
    cfg, w, ctx, qf, qm = small_case()
    k, m, L = 10, 50, cfg["max_ctx_l"]
    with torch.no_grad():
        o = O.query_batch_tensor_section(cfg, w, ctx, qf, qm, q2c_alpha=20.0, max_n_videos=k, max_before_nms=m,
                                         min_pred_l=2, max_pred_l=16)
    vr, st, ed = R.fp64_scores(cfg, w, ctx, qf, qm, chunk=16)  # chunked == unchunked arithmetic per video
    stats = R.check_search_result(vr, st, ed, o["top_video_idx"], o["top_video_score"], o["span_flat_idx"],
                                  o["span_score"], L)
    R.assert_within(stats, 2e-5, "fp32 oracle")
    # a corrupted list (two far-apart ranks exchanged) must be detected
    bad = copy.deepcopy(o["span_flat_idx"])
    bad[:, [0, m - 1]] = bad[:, [m - 1, 0]]
    stats = R.check_search_result(vr, st, ed, o["top_video_idx"], o["top_video_score"], bad, o["span_score"], L)
    assert stats["vcmr"]["inversion"] > 1e-2
    with pytest.raises(AssertionError):
        R.assert_within(stats, 2e-5)
    # a list that drops the best video must be detected as well
    worse = o["top_video_idx"].clone()
    order = torch.sort(vr, dim=1, descending=True)[1]
    worse[:, 0] = order[:, k + 3]
    stats = R.check_search_result(vr, st, ed, worse, o["top_video_score"], None, None, L)
    assert stats["vr"]["dropped_above_kept"] > 0 and stats["vr"]["kept_below_kth"] > 0


def test_list_stats_exact_list_is_zero():
    s = np.sort(np.random.RandomState(0).rand(4, 30))[:, ::-1]
    st = R.list_stats(s[:, :10], s, 10)
    assert st["inversion"] == 0 and st["kept_below_kth"] == 0 and st["positions_off"] == 0
