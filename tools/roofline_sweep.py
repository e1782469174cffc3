#!/usr/bin/env python
"""Roofline sweep of the corpus query x clip contraction (BASELINE.json configs[4]): the filter pass of the video
retrieval (vr_filter_pair_kernel, tcgen05 cta_group::2, 1 MMA per product) and the one-pass exact kernel
(vr_scores_tc_packed_kernel, 3 MMAs per product) over
    queries per pass x hidden size x clips per video,
on a synthetic packed corpus (random unit vectors; the kernels' time does not depend on the values).  Prints one JSON
line per point: algorithmic TFLOP/s (2 * H * S per query and modality) against the measured bf16 peaks of
MEASURED_PEAKS.json.  One GPU; timing by CUDA events, 3 warm-ups, inputs (>= 1 GB of corpus) larger than L2.

    python tools/roofline_sweep.py [--videos 8192] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--hidden", type=int, nargs="*", default=[256, 768, 1024])
    ap.add_argument("--ctx-l", type=int, nargs="*", default=[128, 256])
    ap.add_argument("--queries", type=int, nargs="*", default=[256, 1024, 4096, 16384])
    args = ap.parse_args()
    from tvretrieval_b200 import ops
    from tvretrieval_b200.engine import CorpusPacking
    from tvretrieval_b200.synthetic import corpus_lengths
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    burst, sustained = peaks.get("bf16_tflops", 1590.0), peaks.get("bf16_tflops_sustained", 1400.0)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(7)
    for length in args.ctx_l:
        lens = corpus_lengths(args.videos, length, seed=3).to(dev)
        mask = (torch.arange(length, device=dev)[None] < lens[:, None]).float()
        packing = CorpusPacking(mask)
        s_valid = int(lens.sum())
        for hid in args.hidden:
            kpad = (hid + 63) // 64 * 64
            corpus = []
            for _ in range(2):  # two modalities
                x = torch.randn(packing.n_rows, hid, generator=gen, device=dev)
                corpus.append(ops.split_rows(x, kpad=kpad, normalize=True))
                del x
            for nq in args.queries:
                qs = [ops.split_rows(torch.randn(nq, hid, generator=gen, device=dev), kpad=kpad, normalize=True)
                      for _ in range(2)]
                out = torch.empty(nq, args.videos, device=dev)
                for name, hi_only, mma in (("vr_filter_pair_kernel (hi halves, 1 MMA per product)", True, 1),
                                           ("vr_scores_tc_packed_kernel (hi/lo split, 3 MMAs per product)", False, 3)):
                    def run():
                        ops.vr_scores_tc_packed(qs[0], corpus[0], packing, args.videos, q_b=qs[1], c_b=corpus[1],
                                                ordinal=True, hi_only=hi_only, out=out)
                    for _ in range(3):
                        run()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(args.reps):
                        run()
                    b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / args.reps
                    flops = 2.0 * 2 * hid * s_valid * nq
                    tf = flops / (ms / 1e3) / 1e12
                    corpus_gb = 2 * packing.n_rows * kpad * 2 * (1 if hi_only else 2) / 1e9
                    print(json.dumps({"kernel": name, "queries_per_pass": nq, "hidden": hid, "max_ctx_l": length,
                                      "videos": args.videos, "valid_clips": s_valid, "ms": round(ms, 4),
                                      "algorithmic_tflops": round(tf, 1), "executed_tflops": round(tf * mma, 1),
                                      "frac_of_bf16_burst_peak_executed": round(tf * mma / burst, 3),
                                      "frac_of_bf16_sustained_peak_executed": round(tf * mma / sustained, 3),
                                      "corpus_operand_gb": round(corpus_gb, 2),
                                      "hbm_bound_ms_at_%d_gbs" % int(peaks.get("hbm_gbs", 6545)):
                                          round(corpus_gb / peaks.get("hbm_gbs", 6545) * 1e3, 3)}))
                    sys.stdout.flush()
            del corpus
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
