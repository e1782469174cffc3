#!/usr/bin/env python
"""Benchmark of the XML full-corpus VCMR query path (BASELINE.json metric: queries/sec, 21.8K-video shape).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of ALL queries (default 10,000) over the whole encoded corpus (default 21,793 videos,
L<=128 clips, H=768, video_sub / resnet_i3d, synthetic features, random-init weights): query encoders ->
video-level scores -> top-100 videos -> span distributions of the selected pairs -> top-200 moments.
`value` times it with the raw query features already resident in HBM (CUDA events, max over ranks);
`e2e` times the same through `VCMRSearcher.search_host` from pinned HOST buffers to host numpy results.
The corpus (34 GB encoded) is far larger than L2, so no explicit L2 flush is needed between steps.
`--impl reference` times the CPU oracle port of the reference's own query path (torch CPU ops, all host
threads) on a bounded sample and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "queries/sec full-corpus VCMR (21.8K-video shape)"


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="search", choices=["search", "train"],
                    help="search = full-corpus VCMR queries/s (BASELINE metric, default); train = the training step of "
                         "BASELINE configs[3] (bsz 128 per GPU, hard negatives, VR + SVMR losses, BertAdam), "
                         "data-parallel over the GPUs")
    ap.add_argument("--train-bsz", type=int, default=128)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"],
                    help="ours = the CUDA kernels; reference = the reference's own code on the host CPU cores; "
                         "torch_gpu = the reference's own code through PyTorch on the same GPU")
    ap.add_argument("--n-videos", type=int, default=21793)
    ap.add_argument("--n-queries", type=int, default=10000)
    ap.add_argument("--max-ctx-l", type=int, default=128)
    ap.add_argument("--hidden", type=int, default=768)
    ap.add_argument("--video-dim", type=int, default=3072)
    ap.add_argument("--ctx-bsz", type=int, default=200)
    ap.add_argument("--query-chunk", type=int, default=16384)
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "bf16x3", "f32"],
                    help="video-level score kernel: tcgen05 split-precision (f16x3 / bf16x3) or exact-fp32 SIMT")
    ap.add_argument("--one-pass", action="store_true",
                    help="video retrieval with the one-pass split-precision kernel over all pairs (no filter pass)")
    ap.add_argument("--padded-corpus", action="store_true",
                    help="tensor-core VR kernel on the padded (Nv x L) corpus instead of the packed valid clips")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--device-queries", action="store_true",
                    help="generate the synthetic queries on the device, block by block (stress configs whose raw queries "
                         "would not fit in pinned host memory on every rank); implies --no-e2e")
    ap.add_argument("--transport", default=None, choices=["peer", "collective"],
                    help="list exchanges of the sharded search: peer-memory stores from the ranking kernels (default with "
                         "NCCL) or torch.distributed collectives")
    ap.add_argument("--no-parity", action="store_true", help="skip the float64 parity check after the timed region")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip timing the reference through PyTorch on the same GPU (gpu_reference in the JSON line)")
    ap.add_argument("--cuda-profiler", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    return ap.parse_args(argv)


def model_config(args):
    from tvretrieval_b200.model_xml import xml_base_config
    import copy
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(hidden_size=args.hidden, max_ctx_l=args.max_ctx_l, max_desc_l=30, visual_input_size=args.video_dim,
               query_input_size=768, sub_input_size=768, ctx_mode="video_sub")
    return cfg


def workload_config(args, n_gpus):
    return {"workload": "XML video_sub resnet_i3d full-corpus VCMR+VR: %d videos x %d queries, L<=%d, H=%d, Dv=%d, "
                        "Lq<=30, top-100 videos, top-200 moments" % (args.n_videos, args.n_queries, args.max_ctx_l,
                                                                     args.hidden, args.video_dim),
            "n_videos": args.n_videos, "n_queries": args.n_queries, "max_ctx_l": args.max_ctx_l,
            "hidden": args.hidden, "sharding": "videos/%d" % n_gpus, "l2": "inputs larger than L2 (no flush)",
            "q2c_alpha": 20, "min_pred_l": 2, "max_pred_l": 16, "max_vcmr_video": 100, "max_before_nms": 200}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the run by a thread polling NVML every 5 ms (an
    `nvidia-smi -lms` child needs longer to start than a short multi-GPU run lasts)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(v.strip().isdigit() for v in vis.split(",")) and gpu_index < len(vis.split(",")):
            gpu_index = int(vis.split(",")[gpu_index])
        self.gpu_index, self.thread, self.stop_flag, self.samples, self.err = gpu_index, None, False, [], None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), smax,
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(get_reasons(h))))
                time.sleep(0.005)
        except Exception as exc:  # reported in the JSON line, never fatal
            self.err = "%s: %s" % (type(exc).__name__, exc)

    def start(self):
        import threading
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def mark(self):
        """Samples taken before this call (model warm-up) are dropped if enough remain afterwards."""
        self.mark_at = len(self.samples)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=5)
        samples = self.samples
        if getattr(self, "mark_at", 0) and len(samples) - self.mark_at >= 3:
            samples = samples[self.mark_at:]
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        sm = sorted(x[0] for x in samples)
        bits = 0
        for x in samples:
            bits |= x[3]
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(samples[0][1]),
                "reasons": sorted(name for bit, name in self.REASONS if bits & bit),
                "power_w_max": max(x[2] for x in samples), "samples": len(samples)}


# ------------------------------------------------------------------------------------------------ ours
def encode_corpus_shard(model, args, lens, vid_lo, vid_hi, device):
    """Encodes videos [vid_lo, vid_hi) with exactly the batching of compute_context_info (global batches of
    ctx_bsz videos in corpus order, each padded to its own max length)."""
    from tvretrieval_b200.synthetic import corpus_batch
    n, L, H = vid_hi - vid_lo, args.max_ctx_l, args.hidden
    out = {k: torch.zeros(n, L, H, device=device) for k in ("video_feat1", "video_feat2", "sub_feat1", "sub_feat2")}
    mask = torch.zeros(n, L, device=device)
    bsz = args.ctx_bsz
    t0 = time.perf_counter()
    events = []
    with torch.no_grad():
        for b in range(vid_lo // bsz, (vid_hi - 1) // bsz + 1):
            video, sub, m = corpus_batch(lens, b, bsz, args.video_dim, 768, device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            v1, v2, s1, s2 = model.encode_context(video, m, sub, m)
            e1.record()
            events.append((e0, e1))
            g_lo = b * bsz
            lo, hi = max(g_lo, vid_lo), min(g_lo + len(m), vid_hi)
            sl, dl = slice(lo - g_lo, hi - g_lo), slice(lo - vid_lo, hi - vid_lo)
            w = m.shape[1]
            for k, t in zip(("video_feat1", "video_feat2", "sub_feat1", "sub_feat2"), (v1, v2, s1, s2)):
                out[k][dl, :w] = t[sl]
            mask[dl, :w] = m[sl]
    torch.cuda.synchronize()
    out["video_mask"] = out["sub_mask"] = mask
    # (wall seconds incl. generating the synthetic features, seconds inside XML.encode_context by CUDA events)
    return out, (time.perf_counter() - t0, sum(a.elapsed_time(b) for a, b in events) / 1e3)


# ------------------------------------------------------------------------------------------------ reference legs
def _reference_opt(ref_ns, cfg, device, bsz, max_ctx_l):
    return ref_ns.EasyDict(eval_query_bsz=bsz, eval_context_bsz=200, num_workers=0, pin_memory=False, device=device,
                           ctx_mode=cfg["ctx_mode"], external_inference_vr_res_path=None, q2c_alpha=20.0,
                           min_pred_l=2, max_pred_l=16, max_ctx_l=max_ctx_l, clip_length=1.5, debug=False)


def reference_query_path(ref_ns, model, cfg, ctx, qf, qm, bsz, device, sync=None):
    """One run of the reference's UNMODIFIED compute_query2ctx_info(tasks=VCMR+VR) (baseline/_ref, reference
    inference.py:252-445) over the given queries in batches of `bsz` against the encoded corpus `ctx` (tensors on
    `device`).  -> (seconds in total, seconds in the tensor section :302-389, result dict).  The split comes from
    replacing the module's tqdm progress bar by a timer (tests/reference_loader.SectionTimer)."""
    from tests import reference_loader as RL
    nv, L = ctx["video_mask"].shape
    ds = RL.QueryDataset(qf, qm, nv, L)
    info = dict(video_metas=[{"vid_name": "vid_%05d" % i} for i in range(nv)], **{k: ctx[k] for k in CTX_KEYS})
    timer = RL.SectionTimer(sync)
    ref_ns.inference.tqdm = timer
    opt = _reference_opt(ref_ns, cfg, device, bsz, L)
    if sync:
        sync()
    t0 = time.perf_counter()
    with torch.no_grad():
        res = ref_ns.inference.compute_query2ctx_info(model, ds, opt, info, max_before_nms=200,
                                                      max_n_videos=min(100, nv), tasks=("VCMR", "VR"))
    if sync:
        sync()
    return time.perf_counter() - t0, timer.seconds.get("Computing q embedding", float("nan")), res


CTX_KEYS = ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2", "sub_mask")


def reference_model(cfg, weights, device):
    """The reference's own XML (baseline/_ref) with the given weights, or None when baseline/_ref is not vendored."""
    from tests import reference_loader as RL
    if not RL.available():
        return None, None
    ref_ns = RL.load()
    model = ref_ns.XML(ref_ns.EasyDict(dict(cfg)))
    model.load_state_dict(weights)
    return ref_ns, model.to(device).eval()


def extrapolate_seconds(n_a, t_a, n_b, t_b, n_total):
    """Seconds per pass at n_total videos from two timed corpus sizes: t(n) = fixed + per_video * n (the fixed part --
    query encoding, loader, host lists -- must not be scaled with the corpus, or the CPU arm would look slower than it
    is).  -> (seconds, description)."""
    if n_b > n_a and t_b > t_a:
        per_video = (t_b - t_a) / (n_b - n_a)
        fixed = max(0.0, t_a - per_video * n_a)
    else:
        per_video, fixed = t_b / n_b, 0.0
    return fixed + per_video * n_total, ("extrapolated to %d videos as fixed %.3f s + %.3e s/video (from passes over %d "
                                         "and %d videos)" % (n_total, fixed, per_video, n_a, n_b))


def cpu_baseline(args, cfg, weights, ctx_dev, query_feat, query_mask):
    """The reference's query path timed on the host cores, on a bounded sample: Q=50 queries (the reference's
    eval_query_bsz) against the first Nv_s videos of the same encoded corpus; Nv_s is calibrated so the sample
    takes about --cpu-seconds.  queries/sec is extrapolated linearly in Nv_s / Nv.
    kind "reference" = the unmodified reference from baseline/_ref; "port" = oracle/xml_oracle.py when the
    vendored copy is absent."""
    n_thr = os.cpu_count() or 1
    torch.set_num_threads(n_thr)
    nq = min(50, len(query_feat))
    qf, qm = query_feat[:nq].cpu(), query_mask[:nq].cpu()
    n_total = ctx_dev["video_feat1"].shape[0]
    ref_ns, ref_model = reference_model(cfg, weights, "cpu")
    kind = "reference" if ref_model is not None else "port"

    def run(n_sub, reps):
        ctx = {k: ctx_dev[k][:n_sub].cpu() for k in CTX_KEYS}
        best = float("inf")
        for _ in range(reps):
            if ref_model is not None:
                t, _, _ = reference_query_path(ref_ns, ref_model, cfg, ctx, qf, qm, nq, torch.device("cpu"))
            else:
                from oracle import xml_oracle as O
                t0 = time.perf_counter()
                with torch.no_grad():
                    O.query_batch_tensor_section(dict(cfg), weights, ctx, qf, qm, q2c_alpha=20.0,
                                                 max_n_videos=min(100, n_sub), max_before_nms=200, min_pred_l=2,
                                                 max_pred_l=16, canonical_ties=False)
                t = time.perf_counter() - t0
            best = min(best, t)
        return best

    n_cal = min(n_total, 400)
    run(n_cal, 1)  # warm-up (thread pools, allocator)
    t_cal = run(n_cal, 1)
    n_sub = int(min(n_total, max(n_cal, n_cal * (args.cpu_seconds / 3.0) / max(t_cal, 1e-3))))
    t = run(n_sub, 2)
    t_full, fit = extrapolate_seconds(n_cal, t_cal, n_sub, t, n_total)
    what = ("the unmodified reference compute_query2ctx_info (baseline/_ref, model + driver incl. full sort and host "
            "lists), torch CPU fp32" if kind == "reference" else
            "oracle port of the reference query path incl. full sort, torch CPU fp32")
    return {"value": nq / t_full, "unit": "queries/s", "cores": n_thr, "kind": kind,
            "sample": "%d queries x first %d of %d videos (%.2f s per pass, best of 2), %s; %s"
                      % (nq, n_sub, n_total, t, what, fit)}


def gpu_reference(args, cfg, weights, ctx, qf_cpu, qm_cpu, device):
    """The reference single-GPU PyTorch path on the SAME B200 and the SAME encoded corpus: the unmodified reference
    model + compute_query2ctx_info (baseline/_ref) on cuda, fp32 with TF32 off and on, at the reference's query batch
    (50) and at a large one.  queries/s over the whole call (tensor section + host lists) and over the tensor section
    alone (reference inference.py:302-389)."""
    ref_ns, model = reference_model(cfg, weights, device)
    if model is None:
        return {"unavailable": "baseline/_ref not vendored (tools/vendor_reference.py)"}
    out = {"api": "baseline/_ref baselines.crossmodal_moment_localization.inference.compute_query2ctx_info "
                  "(unmodified), reference XML on cuda, tasks VCMR+VR, top-100 videos, top-200 moments",
           "runs": []}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for tf32, bsz, nq in ((False, 50, 250), (True, 50, 250), (True, 400, 800), (False, 400, 800)):
            nq = min(nq, len(qf_cpu))
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
            run = {"tf32": tf32, "query_batch": bsz, "n_queries": nq}
            try:
                reference_query_path(ref_ns, model, cfg, ctx, qf_cpu[:min(bsz, nq)], qm_cpu[:min(bsz, nq)], bsz,
                                     device, torch.cuda.synchronize)  # warm-up: one batch
                t, t_tensor, _ = reference_query_path(ref_ns, model, cfg, ctx, qf_cpu[:nq], qm_cpu[:nq], bsz, device,
                                                      torch.cuda.synchronize)
                run.update(queries_per_s=nq / t, queries_per_s_tensor_section=nq / t_tensor, seconds=t)
            except torch.cuda.OutOfMemoryError:
                run["error"] = "out of memory"
            torch.cuda.empty_cache()
            out["runs"].append(run)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    ok = [r for r in out["runs"] if "queries_per_s" in r]
    if ok:
        out["best_queries_per_s"] = max(r["queries_per_s"] for r in ok)
        out["best_queries_per_s_tensor_section"] = max(r["queries_per_s_tensor_section"] for r in ok)
    return out


def parity_check(args, cfg, weights, ctx, res, qf_cpu, qm_cpu, device, n_sample=32, tau=1e-4):
    """After the timed region: the lists the timed search returned for `n_sample` of its queries against the reference
    arithmetic evaluated in float64 over the whole corpus (tests/rank_check.py; oracle/ used as the checker only)."""
    from tests import rank_check as R
    nq = len(qf_cpu)
    sample = torch.arange(0, nq, max(1, nq // n_sample))[:n_sample]
    vr64, st64, ed64 = R.fp64_scores(dict(cfg), weights, ctx, qf_cpu[sample], qm_cpu[sample], device=device, chunk=512)
    sd = sample.to(device)
    stats = R.check_search_result(vr64, st64, ed64, res.top_video_idx[sd], res.top_video_score[sd],
                                  res.span_flat_idx[sd], res.span_score[sd], args.max_ctx_l)
    worst = max(v for st_ in stats.values() for k, v in st_.items() if k != "positions_off")
    return {"parity_checked": bool(worst <= tau),
            "parity": {"against": "reference arithmetic in float64 (oracle/xml_oracle.py) over all %d videos"
                                  % ctx["video_mask"].shape[0], "n_queries": int(len(sample)), "tau": tau,
                       "worst_relative_deviation": worst, "stats": stats}}


def host_postprocess_ms(out, n_videos, ctx_len):
    """a18 (reference inference.py:391-445): building the VR + VCMR prediction lists of ALL queries from the numpy
    arrays search_host returned (vectorised host section of tvretrieval_b200.inference), milliseconds."""
    from tvretrieval_b200.inference import host_section
    nq = len(out["top_video_idx"])
    video_metas = [{"vid_name": "vid_%05d" % i} for i in range(n_videos)]
    video2idx = {m["vid_name"]: i for i, m in enumerate(video_metas)}
    metas = [dict(desc_id=i, desc="q%d" % i) for i in range(nq)]
    t0 = time.perf_counter()
    res = host_section(out, metas, video_metas, video2idx, ctx_len, 1.5, ("VCMR", "VR"))
    dt = time.perf_counter() - t0
    assert len(res["VCMR"]) == nq and len(res["VCMR"][0]["predictions"]) == out["span_flat_idx"].shape[1]
    return 1e3 * dt


def ncu_traffic(args, world, searcher):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed
    ncu --set full capture (profiles/ncu_traffic.json); only valid for the shape and mode it was captured at."""
    try:
        t = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    same = (world == 1 and searcher.two_pass and args.precision == t.get("precision")
            and [args.n_videos, args.n_queries, args.max_ctx_l, args.hidden] == t.get("shape"))
    return t.get("dram_bytes_per_launch") if same else None


def run_ours(args):
    from tvretrieval_b200 import _lib
    from tvretrieval_b200.engine import CorpusIndex, PhaseTimer, VCMRSearcher, host_memory_near
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.synthetic import corpus_lengths, synthetic_queries

    if args.device_queries:
        args.no_e2e = args.no_parity = args.no_gpu_reference = args.no_cpu_baseline = True
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run for N>1" % (args.gpus, world))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    cfg = model_config(args)
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    weights_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(device)

    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    vid_lo = rank * args.n_videos // world
    vid_hi = (rank + 1) * args.n_videos // world
    ctx, t_enc = encode_corpus_shard(model, args, lens, vid_lo, vid_hi, device)
    index = CorpusIndex.from_ctx_info(ctx, vid_lo=vid_lo, precision=args.precision, packed=not args.padded_corpus)
    keep_ctx = ctx if (rank == 0 and world == 1 and not (args.no_cpu_baseline and args.no_parity
                                                          and args.no_gpu_reference)) else None
    if keep_ctx is None:
        del ctx
    torch.cuda.empty_cache()

    if args.device_queries:
        gen = torch.Generator(device=device).manual_seed(4321)
        qlen = torch.randint(5, 31, (args.n_queries,), generator=gen, device=device)
        qm = (torch.arange(30, device=device)[None] < qlen[:, None]).float()
        qf = torch.empty(args.n_queries, 30, 768, device=device)
        for lo in range(0, args.n_queries, 8192):
            x = torch.randn(min(8192, args.n_queries - lo), 30, 768, generator=gen, device=device)
            qf[lo:lo + len(x)] = x / (x.norm(dim=-1, keepdim=True) + 1e-5) * qm[lo:lo + len(x)].unsqueeze(2)
        qf_cpu = qm_cpu = qf_pin = qm_pin = None
    else:
        qf_cpu, qm_cpu = synthetic_queries(args.n_queries, 30, 768)
        with host_memory_near(device):  # the pinned upload buffers on the GPU's own NUMA node
            qf_pin, qm_pin = qf_cpu.pin_memory(), qm_cpu.pin_memory()
        qf, qm = qf_pin.to(device), qm_pin.to(device)

    two_pass = False if args.one_pass else None  # None: automatic (on for the packed f16x3 index)
    if world == 1:
        searcher = VCMRSearcher(model, index, query_chunk=args.query_chunk, two_pass=two_pass)
    else:
        from tvretrieval_b200.sharding import ShardedSearcher
        searcher = ShardedSearcher(model, index, n_videos_total=args.n_videos, query_chunk=args.query_chunk,
                                   two_pass=two_pass, transport=args.transport)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ---- device-resident timing ----
        sampler = ClockSampler(local_rank)
        sampler.start()  # begins before the warm-up so that very short runs still get samples under load
        for _ in range(args.warmup):
            searcher.search(qf, qm)
        torch.cuda.synchronize()
        sampler.mark()
        timer = PhaseTimer()
        searcher.timer = timer
        barrier()
        launches0 = _lib.launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.cuda_profiler:
            torch.cuda.profiler.start()
        start.record()
        for _ in range(args.steps):
            res = searcher.search(qf, qm)
        end.record()
        if args.cuda_profiler:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        barrier()
        clocks = sampler.stop()
        launches = _lib.launch_count() - launches0
        ms_total = max_over_ranks(start.elapsed_time(end))
        phases = timer.totals_ms()
        searcher.timer = None
        assert res.span_score.shape == (args.n_queries, 200) and bool((res.span_score[:, 0] > 0).all())

        # ---- end to end from pinned host buffers ----
        e2e = None
        if not args.no_e2e:
            rr = 0 if world > 1 else None  # sharded: the (identical) result is copied to the host of rank 0 only
            for _ in range(max(1, args.warmup // 2)):
                searcher.search_host(qf_pin, qm_pin, result_rank=rr)
            e2e_timer = PhaseTimer()
            searcher.timer = e2e_timer
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out = searcher.search_host(qf_pin, qm_pin, result_rank=rr)
            barrier()
            t_e2e = max_over_ranks(time.perf_counter() - t0)
            searcher.timer = None
            d2h = sum(v.nbytes for v in out.values()) if out is not None else 0
            e2e = {"value": args.n_queries * args.steps / t_e2e, "unit": "queries/s",
                   "h2d_bytes_per_step": qf_pin.numel() * 4 + qm_pin.numel() * 4, "d2h_bytes_per_step": d2h,
                   "h2d_note": "summed over ranks: each rank uploads only its slice of the queries",
                   "ms_per_step": 1e3 * t_e2e / args.steps,
                   # device time of the phases of the host-buffer call (per-piece filter passes summed); what is
                   # left of ms_per_step is exposed upload, result download and host / launch gaps
                   "phases_ms_per_step": {k: v / args.steps for k, v in e2e_timer.totals_ms().items()},
                   "api": "tvretrieval_b200.engine.VCMRSearcher.search_host (the call compute_query2ctx_info makes)"}

    ms_per_step = ms_total / args.steps
    value = args.n_queries / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel: the corpus query x clip contraction (video-level scores) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    n_local = vid_hi - vid_lo
    s_pad = n_local * args.max_ctx_l
    s_valid = int(lens[vid_lo:vid_hi].sum())
    packed = index.packing is not None
    s_count = s_valid if packed else s_pad  # clips the kernel actually contracts against
    flops_per_step = 2.0 * 2 * args.hidden * s_count * args.n_queries  # 2 modalities x 2*H*S per query
    n_calls = timer.counts().get("vr_scores", 1)
    vr_ms_per_step = phases.get("vr_scores", 0.0) / args.steps
    achieved = flops_per_step / (vr_ms_per_step / 1e3) / 1e12 if vr_ms_per_step > 0 else 0.0
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tc_mode = args.precision != "f32"
    mma_per_product = 1 if (searcher.two_pass or not tc_mode) else 3
    if searcher.two_pass:
        kernel_name = ("vr_filter_pair_kernel, filter pass of the two-pass search (tcgen05 cta_group::2 kind::f16 on CTA "
                       "pairs, M=256 x N=256 tiles, hi halves only: 1 MMA per product over ALL pairs, fp32 TMEM "
                       "accumulate; the candidates it leaves are re-scored exactly by vr_rescore_tc_kernel, phases "
                       "vr_select/vr_rescore)")
    elif tc_mode:
        kernel_name = ("vr_scores_tc_packed_kernel (tcgen05 kind::f16, %s split: 3 MMAs per product, fp32 TMEM "
                       "accumulate)" % args.precision)
    else:
        kernel_name = "gemm_simt_kernel (fp32 SIMT, EPI_VRMAX)"
    roofline = {"bound": "tensor", "kernel": kernel_name,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "mma_flops_executed_per_algorithmic": mma_per_product,
                "frac_of_peak_executed": mma_per_product * achieved / peak,
                "traffic": ncu_traffic(args, world, searcher), "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                "flops_counted": ("algorithmic 2*H*S per modality per query, S = %d valid clips (packed corpus, tile fill "
                                  "%.3f); the padded corpus would be S_pad = %d" % (s_valid, index.packing.fill, s_pad))
                if packed else "algorithmic 2*H*S_pad per modality per query, S_pad = n_videos*L = %d padded clips" % s_pad,
                "ms_per_step_in_kernel": vr_ms_per_step, "launches_per_step": n_calls / args.steps,
                "share_of_step": vr_ms_per_step / ms_per_step}

    line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32 (corpus contraction: %s)" % ("16-bit hi/lo split tensor-core products, fp32 accumulate"
                                                       if tc_mode else "fp32 FMA"),
            "data": "synthetic", "config": dict(workload_config(args, world), precision=args.precision,
                                                video_retrieval="two-pass" if searcher.two_pass else "one-pass",
                                                exchange=getattr(searcher, "transport", "none (one GPU)")),
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
            "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "corpus_encode": {"videos_per_s": n_local / t_enc[1], "seconds": t_enc[1],
                              "seconds_incl_synthetic_feature_generation": t_enc[0],
                              "precision": model.context_precision, "index_gb": index.nbytes() / 1e9}}
    if e2e is not None:
        line["e2e"] = e2e
    if e2e is not None and rank == 0:
        line["host_postprocess_ms"] = host_postprocess_ms(out, args.n_videos, args.max_ctx_l)
    if rank == 0 and world == 1 and not args.no_parity:
        line.update(parity_check(args, cfg, weights_cpu, keep_ctx, res, qf_cpu, qm_cpu, device))
    else:
        line["parity_checked"] = False
        line["parity"] = {"skipped": "run at N=1 (sharded results are bit-equal to N=1: tests/test_gpu_sharded.py)"
                          if world > 1 else "--no-parity"}
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        del searcher, index
        torch.cuda.empty_cache()
        line["gpu_reference"] = gpu_reference(args, cfg, weights_cpu, keep_ctx, qf_cpu, qm_cpu, device)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, cfg, weights_cpu, keep_ctx, qf_cpu, qm_cpu)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arms
def _synthetic_encoded_corpus(n, L, H, gen):
    """Random stand-in for an encoded corpus (the query path's cost does not depend on the values)."""
    lens = torch.randint(L // 8, L + 1, (n,), generator=gen)
    lens[0] = L
    mask = (torch.arange(L)[None] < lens[:, None]).float()
    ctx = {"video_mask": mask, "sub_mask": mask}
    for k in ("video_feat1", "video_feat2", "sub_feat1", "sub_feat2"):
        ctx[k] = torch.randn(n, L, H, generator=gen)
    return ctx


def run_reference(args):
    """CPU arm: the reference's own query path on the host cores -- the UNMODIFIED reference model + driver from
    baseline/_ref (kind "reference"), or the oracle port when the vendored copy is absent (kind "port").  None of this
    repo's kernels, models or engine are on this path.  Each step = 50 queries (reference eval_query_bsz) against a
    bounded sample of the corpus; queries/s scaled linearly to the full corpus."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_thr = os.cpu_count() or 1
    torch.set_num_threads(n_thr)
    cfg = dict(model_config(args))
    from tvretrieval_b200.synthetic import synthetic_queries
    nq = min(50, args.n_queries)
    qf, qm = synthetic_queries(nq, 30, 768)
    L, H = args.max_ctx_l, args.hidden
    gen = torch.Generator().manual_seed(4321)
    from tests import reference_loader as RL
    if RL.available():
        torch.manual_seed(2018)
        ref_ns = RL.load()
        model = ref_ns.XML(ref_ns.EasyDict(cfg)).eval()
        kind = "reference"

        def one_pass(ctx):
            return reference_query_path(ref_ns, model, cfg, ctx, qf, qm, nq, torch.device("cpu"))[0]
    else:
        from oracle import xml_oracle as O
        w = O.init_weights(cfg)
        kind = "port"

        def one_pass(ctx):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.query_batch_tensor_section(cfg, w, ctx, qf, qm, q2c_alpha=20.0,
                                             max_n_videos=min(100, len(ctx["video_mask"])), max_before_nms=200,
                                             min_pred_l=2, max_pred_l=16, canonical_ties=False)
            return time.perf_counter() - t0

    n_cal = min(args.n_videos, 400)
    cal = _synthetic_encoded_corpus(n_cal, L, H, gen)
    one_pass(cal)
    t_cal = one_pass(cal)
    budget = 150.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
    n_sub = int(min(args.n_videos, max(n_cal, n_cal * min(budget, args.cpu_seconds) / 3.0 / max(t_cal, 1e-3))))
    ctx = _synthetic_encoded_corpus(n_sub, L, H, gen)
    for _ in range(args.warmup):
        one_pass(ctx)
    times = [one_pass(ctx) for _ in range(args.steps)]
    t = sum(times) / len(times)
    t_full, fit = extrapolate_seconds(n_cal, t_cal, n_sub, t, args.n_videos)
    value = nq / t_full
    what = ("the UNMODIFIED reference (baseline/_ref): XML.get_pred_from_raw_query(cross=True) inside "
            "inference.compute_query2ctx_info incl. full sort and host lists" if kind == "reference" else
            "oracle port of get_pred_from_raw_query(cross=True) + exp/softmax/topk/gather/einsum/band mask/full sort")
    sample = ("each step = %d queries (reference eval_query_bsz) x %d of %d videos, synthetic encoded corpus, torch CPU "
              "fp32, %s; %s" % (nq, n_sub, args.n_videos, what, fit))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": n_thr, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_torch_gpu(args):
    """The "reference single-GPU PyTorch path" of north_star on cuda:0: the UNMODIFIED reference (baseline/_ref) encodes
    the synthetic corpus with its own XML.encode_context and answers the queries with its own compute_query2ctx_info;
    no kernel, model or engine of this repo is on this path (tvretrieval_b200.synthetic only generates the inputs)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from tests import reference_loader as RL
    if not RL.available():
        print(json.dumps({"impl": "torch_gpu", "unavailable": "baseline/_ref not vendored"}))
        return
    from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths, synthetic_queries
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    cfg = dict(model_config(args))
    torch.manual_seed(2018)
    ref_ns = RL.load()
    model = ref_ns.XML(ref_ns.EasyDict(cfg)).to(device).eval()
    lens = corpus_lengths(args.n_videos, args.max_ctx_l)
    L, H = args.max_ctx_l, args.hidden
    ctx = {k: torch.zeros(args.n_videos, L, H, device=device) for k in ("video_feat1", "video_feat2", "sub_feat1",
                                                                          "sub_feat2")}
    mask = torch.zeros(args.n_videos, L, device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in range((args.n_videos + args.ctx_bsz - 1) // args.ctx_bsz):
            video, sub, m = corpus_batch(lens, b, args.ctx_bsz, args.video_dim, 768, device)
            outs = model.encode_context(video, m, sub, m)
            lo, w = b * args.ctx_bsz, m.shape[1]
            for k, t in zip(("video_feat1", "video_feat2", "sub_feat1", "sub_feat2"), outs):
                ctx[k][lo:lo + len(m), :w] = t
            mask[lo:lo + len(m), :w] = m
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t0
    ctx["video_mask"] = ctx["sub_mask"] = mask
    qf, qm = synthetic_queries(args.n_queries, 30, 768)
    weights = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    del model
    g = gpu_reference(args, cfg, weights, ctx, qf, qm, device)
    best = g.get("best_queries_per_s", 0.0)
    line = {"impl": "torch_gpu", "metric": METRIC, "value": best, "unit": "queries/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, 1), "gpu_reference": g,
            "corpus_encode": {"videos_per_s": args.n_videos / t_enc, "seconds": t_enc}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ training step
def run_train(args):
    """BASELINE configs[3]: XML.forward + backward + BertAdam.step on a synthetic batch of `--train-bsz` items per GPU
    (video_sub, H=768, Dv=3072, L<=128, in-batch hard negatives, lw_st_ed=0.01), data-parallel: every rank its own
    batch, gradients averaged with one NCCL all-reduce (sharding.all_reduce_gradients) before the fused optimizer
    step.  `value` = training samples/s over all ranks; `gpu_reference` = the unmodified reference model (baseline/_ref)
    with torch eager fp32 + the reference's BertAdam on the same GPU, same batch (N=1 only)."""
    from tvretrieval_b200 import _lib
    from tvretrieval_b200.model_xml import XML
    from tvretrieval_b200.optimization import BertAdam
    from tvretrieval_b200.sharding import all_reduce_gradients
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    cfg = model_config(args)
    cfg.update(use_hard_negative=True, hard_pool_size=20, lw_st_ed=0.01, lw_neg_q=1, lw_neg_ctx=1, margin=0.1)
    n, L = args.train_bsz, args.max_ctx_l

    def batch(dev, seed):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(16, L + 1, (n,), generator=g)
        lens[0] = L
        qlens = torch.randint(5, 31, (n,), generator=g)
        vm = (torch.arange(L)[None] < lens[:, None]).float()
        qm = (torch.arange(30)[None] < qlens[:, None]).float()
        unit = lambda t: t / (t.norm(dim=-1, keepdim=True) + 1e-5)  # noqa: E731
        st = (torch.rand(n, generator=g) * (lens - 1)).long()
        d = dict(query_feat=unit(torch.randn(n, 30, 768, generator=g)) * qm[..., None], query_mask=qm,
                 video_feat=unit(torch.randn(n, L, args.video_dim, generator=g)) * vm[..., None], video_mask=vm,
                 sub_feat=unit(torch.randn(n, L, 768, generator=g)) * vm[..., None], sub_mask=vm, tef_feat=None,
                 tef_mask=None, st_ed_indices=torch.stack([st, torch.minimum(lens - 1, st + 3)], 1))
        return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}

    def optimizer(model, cls):
        no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
        named = list(model.named_parameters())
        return cls([{"params": [p for k, p in named if not any(nd in k for nd in no_decay)], "weight_decay": 0.01},
                    {"params": [p for k, p in named if any(nd in k for nd in no_decay)], "weight_decay": 0.0}],
                   lr=1e-4, warmup=0.01, t_total=1000, schedule="warmup_linear")

    torch.manual_seed(2018)
    model = XML(cfg).to(device).train()
    weights = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = optimizer(model, BertAdam)
    inputs = batch(device, 1234 + rank)
    phases = {"forward_backward": 0.0, "all_reduce": 0.0, "optimizer": 0.0}

    def step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad()
        ev[0].record()
        loss, _ = model(**inputs)
        loss.backward()
        ev[1].record()
        if world > 1:
            all_reduce_gradients(model.parameters())
        ev[2].record()
        opt.step()
        ev[3].record()
        return ev, loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        step(False)
    torch.cuda.synchronize()
    sampler.mark()
    barrier()
    launches0 = _lib.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    evs = [step(True) for _ in range(args.steps)]
    end.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms_total = start.elapsed_time(end)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    for ev, _ in evs:
        phases["forward_backward"] += ev[0].elapsed_time(ev[1]) / args.steps
        phases["all_reduce"] += ev[1].elapsed_time(ev[2]) / args.steps
        phases["optimizer"] += ev[2].elapsed_time(ev[3]) / args.steps
    ms_per_step = ms_total / args.steps
    n_params = sum(p.numel() for p in model.parameters())
    line = {"metric": "training samples/sec, XML video_sub step (configs[3])", "value": world * n / (ms_per_step / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (GEMMs: 16-bit hi/lo split tensor-core products, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "XML.forward + backward + BertAdam.step, video_sub resnet_i3d, bsz %d per GPU, L<=%d, "
                                   "H=%d, Dv=%d, hard negatives (pool 20), lw_st_ed 0.01; data parallel, gradients averaged "
                                   "by one NCCL all-reduce over a flat %.0f MB buffer"
                                   % (n, L, args.hidden, args.video_dim, n_params * 4 / 1e6),
                       "train_precision": model.train_precision, "parameters": n_params},
            "clocks": clocks, "gpu_launches": launches, "phases_ms_per_step": phases, "loss_last": float(evs[-1][1])}
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        ref_ns, ref_model = reference_model(cfg, weights, device)
        if ref_model is not None:
            from baselines.crossmodal_moment_localization.optimization import BertAdam as RefAdam
            ref_model.train()
            ropt = optimizer(ref_model, RefAdam)
            times = []
            for i in range(3 + args.steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ropt.zero_grad()
                rloss, _ = ref_model(**inputs)
                rloss.backward()
                ropt.step()
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            t_ref = sum(times[3:]) / args.steps
            line["gpu_reference"] = {"api": "baseline/_ref XML.forward + backward + reference BertAdam.step, torch eager "
                                            "fp32 on the same GPU and batch", "ms_per_step": 1e3 * t_ref,
                                     "samples_per_s": n / t_ref}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.workload == "train":
        run_train(a)
    elif a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch_gpu":
        run_torch_gpu(a)
    else:
        run_ours(a)
