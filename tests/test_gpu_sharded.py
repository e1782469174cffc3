"""Multi-GPU equivalence: the video-sharded search (tvretrieval_b200/sharding.py) must return exactly the
single-GPU result.  With >= 2 GPUs the two ranks use NCCL (one GPU each); on a single-GPU box both ranks share
cuda:0 and the two all-gathers go through gloo (staged through the host) -- same kernels, same merge logic."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, use_nccl, precision, transport):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if use_nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
    try:
        from tvretrieval_b200.engine import CorpusIndex, VCMRSearcher
        from tvretrieval_b200.model_xml import XML, xml_base_config
        from tvretrieval_b200.sharding import ShardedSearcher, ShardPlan
        from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths, synthetic_queries
        n_videos, nq, L, H, k_vid, k_span = 61, 37, 32, 64, 16, 40
        cfg = copy.deepcopy(xml_base_config)
        cfg.update(hidden_size=H, max_ctx_l=L, max_desc_l=12, visual_input_size=48, query_input_size=40,
                   sub_input_size=40, n_heads=4)
        torch.manual_seed(7)
        model = XML(cfg).to(dev).eval()
        lens = corpus_lengths(n_videos, L, seed=5)
        lens[3] = 2  # a very short video: fewer in-band cells than k_span when it is selected
        with torch.no_grad():
            video, sub, mask = corpus_batch(lens, 0, n_videos, 48, 40, dev, seed=5, video_split=32)
            v1, v2, s1, s2 = model.encode_context(video, mask, sub, mask)
            full = CorpusIndex(v1, v2, mask, s1, s2, mask, precision=precision)
            lo, hi = ShardPlan(n_videos, world, rank).video_range()
            shard = CorpusIndex(v1[lo:hi].contiguous(), v2[lo:hi].contiguous(), mask[lo:hi].contiguous(),
                                s1[lo:hi].contiguous(), s2[lo:hi].contiguous(), mask[lo:hi].contiguous(), vid_lo=lo,
                                precision=precision)
            qf, qm = synthetic_queries(nq, 12, 40, seed=9)
            qf, qm = qf.to(dev), qm.to(dev)
            gt = torch.randint(0, n_videos, (nq,), generator=torch.Generator().manual_seed(1)).to(torch.int32).to(dev)
            kw = dict(max_n_videos=k_vid, max_before_nms=k_span, query_chunk=16)
            tasks = ("VCMR", "VR", "SVMR")
            want = VCMRSearcher(model, full, two_pass=False, **kw).search(qf, qm, gt, tasks)
            # the shards additionally use the two-pass (filter + exact re-score) video retrieval
            kw["two_pass"] = precision != "f32"
            sharded = ShardedSearcher(model, shard, n_videos_total=n_videos, transport=transport, **kw)
            assert sharded.transport == (transport or ("peer" if use_nccl else "collective"))
            got = sharded.search(qf, qm, gt, tasks)
            again = sharded.search(qf, qm, gt, tasks)  # workspace re-use across calls
            for name in ("top_video_idx", "span_flat_idx", "span_score", "svmr_flat_idx"):
                assert torch.equal(getattr(got, name), getattr(again, name)), name
        for name in ("top_video_idx", "top_video_score", "span_flat_idx", "span_score", "svmr_flat_idx", "svmr_score"):
            a, b = getattr(got, name), getattr(want, name)
            assert torch.equal(a, b), "rank %d: %s differs from the single-GPU result" % (rank, name)
        assert (want.span_score[:, 0] > 0).all()
        # host-buffer entry point: every rank uploads only its query slice
        host = sharded.search_host(qf.cpu(), qm.cpu(), gt.cpu(), tasks)
        for name in ("top_video_idx", "span_flat_idx", "span_score", "svmr_flat_idx"):
            assert (torch.from_numpy(host[name]) == getattr(want, name).cpu()).all(), name
        # ... in several pieces per block: every rank encodes its share of EACH piece, the shares are exchanged and
        # the filter pass runs piece by piece (with 3 ranks the last rank's share of a 4-query piece is empty)
        sharded.min_piece, sharded.encode_chunk = 4, 8
        assert len(sharded._piece_bounds(16, True)) == 4
        for _ in range(2):  # twice: the exchange workspace is re-used across calls
            host = sharded.search_host(qf.cpu(), qm.cpu(), gt.cpu(), tasks)
            for name in ("top_video_idx", "span_flat_idx", "span_score", "svmr_flat_idx"):
                assert (torch.from_numpy(host[name]) == getattr(want, name).cpu()).all(), name + " (pieces)"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", [None, "collective"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("precision", ["f32", "f16x3"])
def test_sharded_search_equals_single_gpu(precision, world, transport):
    """transport None = peer-memory exchange when the ranks have a GPU each (NCCL), else the collective path."""
    use_nccl = torch.cuda.device_count() >= world
    if transport == "collective" and not use_nccl:
        pytest.skip("same as transport=None on a single-GPU box")
    port = 29500 + (os.getpid() * 7 + len(precision) + 13 * world + (5 if transport else 0)) % 2000
    mp.spawn(_worker, args=(world, port, use_nccl, precision, transport), nprocs=world, join=True)
