#!/usr/bin/env python
"""Times the small kernels of the query path on the bench workload (21.8K videos x 10K queries), one kernel at a time
with CUDA events, 10 repeats each, inputs = what the search itself feeds them.  Prints one line per kernel."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def counters(reset=True):
    from tvretrieval_b200 import _lib
    out = (ctypes.c_longlong * 4)()
    _lib.lib().xmlb_debug_counters(out, int(reset))
    return list(out)


def main():
    args = bench.parse_args(sys.argv[1:])
    from tvretrieval_b200 import ops
    from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.synthetic import corpus_lengths, synthetic_queries
    dev = torch.device("cuda", 0)
    torch.manual_seed(2018)
    model = XML(bench.model_config(args)).eval().to(dev)
    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    ctx, _ = bench.encode_corpus_shard(model, args, lens, 0, args.n_videos, dev)
    index = CorpusIndex.from_ctx_info(ctx, precision=args.precision)
    del ctx
    qf, qm = synthetic_queries(args.n_queries, 30, 768)
    qf, qm = qf.to(dev), qm.to(dev)
    s = VCMRSearcher(model, index)
    with torch.no_grad():
        res = s.search(qf, qm)
        lens_q = (qm != 0).sum(1).to(torch.int64).cpu()
        vq, sq = s._encode_pieces(s._device_pieces(qf, qm, s._piece_bounds(len(qf), False)), lens_q, None, width=30,
                                  bounds=s._piece_bounds(len(qf), False))
        lists = s.span_lists(res.top_video_idx)
        st, ed = s.span_probs(vq, sq, lists)
        nq = len(qf)
        st, ed = st.view(nq, 100, index.ctx_len), ed.view(nq, 100, index.ctx_len)
        counters()
        t = timed(lambda: ops.span_topk(st, ed, res.top_video_score, 2, 16, 200))
        c = counters()
        print("span_topk (all 100 slots valid): %.3f ms; overflowed rows %d of %d, mean survivors %.0f"
              % (t, c[0], c[2], c[1] / max(1, c[2])))
        for frac in (2, 8):
            valid = (torch.arange(100, device=dev)[None] % frac == (torch.arange(nq, device=dev)[:, None] % frac))
            valid = valid.to(torch.uint8).contiguous()
            counters()
            t = timed(lambda: ops.span_topk(st, ed, res.top_video_score, 2, 16, 200, slot_valid=valid, zero_fill=False))
            c = counters()
            print("span_topk (1/%d of the slots valid, as on %d GPUs): %.3f ms; overflowed rows %d of %d, mean survivors %.0f"
                  % (frac, frac, t, c[0], c[2], c[1] / max(1, c[2])))
        for mode in (("warps",) if index.f2cat is not None and index.f2cat[0].dim() >= 3 else ("copy", "warps")):
            ops.GATHER = mode
            print("span_probs [%s]: %.3f ms" % (mode, timed(lambda: s.span_probs(vq, sq, lists))))
            print("top_videos [%s] (filter + select + rescore + topk): %.3f ms" % (mode, timed(lambda: s.top_videos(vq, sq, 100), 3)))
        print("pair lists: %.3f ms" % timed(lambda: s.span_lists(res.top_video_idx)))
        print("top_videos (filter + select + rescore + topk): %.3f ms" % timed(lambda: s.top_videos(vq, sq, 100), 3))
        bounds = s._piece_bounds(len(qf), False)
        print("encode_query (packed, %d pieces): %.3f ms"
              % (len(bounds), timed(lambda: s._encode_pieces(s._device_pieces(qf, qm, bounds), lens_q, None, width=30,
                                                              bounds=bounds), 5)))


if __name__ == "__main__":
    main()
