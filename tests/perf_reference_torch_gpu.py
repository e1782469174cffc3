#!/usr/bin/env python
"""Measurement script (not a pytest file): the reference's arithmetic run THROUGH PYTORCH ON THE SAME B200 -- the
"reference single-GPU PyTorch path" of the north star -- next to the kernels, on the bench workload.

The reference itself cannot travel to the GPU box; its restatement in plain torch ops (oracle/xml_oracle.py, pinned
to the real reference by tests/golden) is executed on cuda:0 instead: same op sequence as reference
inference.py:302-389 (dense (Q, Nv, L) logits, softmax, gather, (Q,100,L,L) span tensor, full sort) in batches of
Q = 50 queries (the reference's eval_query_bsz), fp32 with TF32 off (the parity setting) and on.
Also times one training step (config #4: bsz 128, L <= 128, hard negatives) both ways.

    python tests/perf_reference_torch_gpu.py [--n-videos 21793] [--batches 3]      -> profiles/r01_torch_gpu_reference.txt
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from oracle import xml_oracle as O  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-videos", type=int, default=21793)
    ap.add_argument("--n-queries", type=int, default=10000)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--train-bsz", type=int, default=128)
    a = ap.parse_args()
    sys.argv = sys.argv[:1]
    args = bench.parse_args()  # the bench defaults (dims, precision, context batch size)
    args.n_videos, args.n_queries = a.n_videos, a.n_queries
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.optimization import BertAdam
    from tvretrieval_b200.synthetic import corpus_lengths, synthetic_queries
    dev = torch.device("cuda:0")
    cfg = bench.model_config(args)
    torch.manual_seed(2018)
    model = XML(cfg).eval().to(dev)
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}  # on the GPU: the oracle runs there
    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    ctx, _ = bench.encode_corpus_shard(model, args, lens, 0, args.n_videos, dev)
    qf_cpu, qm_cpu = synthetic_queries(args.n_queries, 30, 768)
    qf, qm = qf_cpu.to(dev), qm_cpu.to(dev)
    print("# %s, torch %s; %d videos x %d queries, L<=%d, H=%d" % (torch.cuda.get_device_name(0), torch.__version__,
                                                                  args.n_videos, args.n_queries, args.max_ctx_l,
                                                                  args.hidden))
    # ---- query path ----
    index = CorpusIndex.from_ctx_info(ctx, precision=args.precision)
    searcher = VCMRSearcher(model, index)
    with torch.no_grad():
        ms_ours = timed(lambda: searcher.search(qf, qm), 3)
    print("kernels (VCMRSearcher.search, all %d queries):  %9.2f ms  -> %10.1f queries/s"
          % (args.n_queries, ms_ours, args.n_queries / ms_ours * 1e3))
    del index, searcher
    torch.cuda.empty_cache()
    q = 50
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        i = [0]

        def ref_batch():
            lo = (i[0] % a.batches) * q
            i[0] += 1
            O.query_batch_tensor_section(cfg, weights, ctx, qf[lo:lo + q], qm[lo:lo + q], q2c_alpha=20.0,
                                         max_n_videos=100, max_before_nms=200, min_pred_l=2, max_pred_l=16,
                                         canonical_ties=False)
        with torch.no_grad():
            ms = timed(ref_batch, a.batches)
        print("torch ops on the GPU, reference op sequence, Q=50, TF32 %-3s: %9.2f ms per batch -> %10.1f queries/s "
              "(x%.0f slower than the kernels)" % ("on" if tf32 else "off", ms, q / ms * 1e3,
                                                   (args.n_queries / ms_ours) / (q / ms)))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    del ctx
    torch.cuda.empty_cache()
    # ---- training step (config #4) ----
    n = a.train_bsz
    g = torch.Generator().manual_seed(1234)
    L = args.max_ctx_l
    ln = torch.randint(L // 8, L + 1, (n,), generator=g)
    ln[0] = L
    ql = torch.randint(5, 31, (n,), generator=g)
    vmask = (torch.arange(L)[None] < ln[:, None]).float()
    qmask = (torch.arange(30)[None] < ql[:, None]).float()
    unit = lambda t: t / (t.norm(dim=-1, keepdim=True) + 1e-5)  # noqa: E731
    inputs = dict(query_feat=unit(torch.randn(n, 30, 768, generator=g)) * qmask[..., None], query_mask=qmask,
                  video_feat=unit(torch.randn(n, L, args.video_dim, generator=g)) * vmask[..., None], video_mask=vmask,
                  sub_feat=unit(torch.randn(n, L, 768, generator=g)) * vmask[..., None], sub_mask=vmask,
                  tef_feat=None, tef_mask=None)
    st = (torch.rand(n, generator=g) * (ln - 1)).long()
    inputs["st_ed_indices"] = torch.stack([st, torch.minimum(ln - 1, st + 3)], 1)
    inputs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inputs.items()}
    tcfg = dict(cfg, use_hard_negative=True, hard_pool_size=20, lw_st_ed=0.01)
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")

    def groups(named):
        return [{"params": [p for k, p in named if not any(nd in k for nd in no_decay)], "weight_decay": 0.01},
                {"params": [p for k, p in named if any(nd in k for nd in no_decay)], "weight_decay": 0.0}]

    for prec in ("f32", "f16x3"):
        torch.manual_seed(2018)
        m = XML(type(cfg)(tcfg)).to(dev).train()
        m.train_precision = prec
        opt = BertAdam(groups(list(m.named_parameters())), lr=1e-4, warmup=0.01, t_total=1000)

        def fwd_bwd():
            opt.zero_grad()
            loss, _ = m(**inputs)
            loss.backward()
        ms = timed(fwd_bwd, 5)
        ms_opt = timed(opt.step, 5)
        print("training step bsz=%d, kernels (train_precision=%s, dropout 0.1): forward+backward %8.2f ms, fused "
              "BertAdam.step %6.3f ms" % (n, prec, ms, ms_opt))
    w = {k: v.detach().clone().requires_grad_(True) for k, v in weights.items()}

    def ref_fwd_bwd():
        for v in w.values():
            v.grad = None
        loss, _ = O.train_forward(tcfg, w, inputs["query_feat"], inputs["query_mask"], inputs["video_feat"],
                                  inputs["video_mask"], inputs["sub_feat"], inputs["sub_mask"], inputs["st_ed_indices"])
        loss.backward()
    ms = timed(ref_fwd_bwd, 5)
    print("training step bsz=%d, torch ops on the GPU (reference op sequence, fp32, no dropout): forward+backward "
          "%8.2f ms" % (n, ms))


if __name__ == "__main__":
    main()
