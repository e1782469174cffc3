#!/usr/bin/env python
"""Encodes a few context batches of the bench corpus (200 videos each, TVR dims) and brackets the last ones with
cudaProfilerStart/Stop, for `ncu --profile-from-start off`; prints videos/s without the profiler."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def main():
    args = bench.parse_args(sys.argv[1:])
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths
    dev = torch.device("cuda", 0)
    torch.manual_seed(2018)
    model = XML(bench.model_config(args)).eval().to(dev)
    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    batches = [corpus_batch(lens, b, args.ctx_bsz, args.video_dim, 768, dev) for b in range(6)]
    with torch.no_grad():
        for v, s, m in batches[:2]:
            model.encode_context(v, m, s, m)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for v, s, m in batches:
            model.encode_context(v, m, s, m)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("encode_context: %.1f videos/s (%d videos, %.2f ms per batch of %d)"
              % (6 * args.ctx_bsz / dt, 6 * args.ctx_bsz, 1e3 * dt / 6, args.ctx_bsz))
        torch.cuda.profiler.start()
        v, s, m = batches[2]
        model.encode_context(v, m, s, m)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
