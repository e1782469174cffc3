"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: partition arithmetic, the two all-gathers and the
global-index bookkeeping of tvretrieval_b200/sharding.py.  The kernels themselves are covered by the gpu tests
(test_gpu_sharded.py runs the full sharded search against the single-GPU result)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_videos, nq, k):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tvretrieval_b200.sharding import ShardPlan, all_gather_cat, all_gather_rows, local_slot_mask
        plan = ShardPlan(n_videos, world, rank)
        lo, hi = plan.video_range()
        # every rank scores its own videos with the same global function; local top-k -> gather -> merge
        g = torch.Generator().manual_seed(0)
        scores = torch.rand(nq, n_videos, generator=g)
        scores[:, 5] = scores[:, n_videos - 2]  # cross-shard exact tie
        local = scores[:, lo:hi]
        val, idx = torch.sort(local, dim=1, descending=True, stable=True)
        val, idx = val[:, :k].contiguous(), (idx[:, :k] + lo).to(torch.int32).contiguous()
        g_val, g_idx = all_gather_cat(val), all_gather_cat(idx)
        assert g_val.shape == (nq, world * k)
        for r in range(world):
            r_lo, r_hi = plan.video_range(r)
            blk = g_idx[:, r * k:(r + 1) * k]
            assert ((blk >= r_lo) & (blk < r_hi)).all()
        assert torch.equal(g_val, torch.gather(scores, 1, g_idx.long()))
        # canonical merge (score desc, global id asc) of the gathered candidates == global top-k
        key = g_val.double() * 1e12 - g_idx.double()
        order = torch.sort(key, dim=1, descending=True)[1][:, :k]
        merged = torch.gather(g_idx, 1, order).long()
        want = torch.sort(scores.double() * 1e12 - torch.arange(n_videos).double(), dim=1, descending=True)[1][:, :k]
        assert torch.equal(merged, want)
        mine = local_slot_mask(merged, lo, hi)
        assert torch.equal(mine.bool(), plan.owner_of(merged) == rank)
        counts = all_gather_rows(mine.sum(1, keepdim=True).to(torch.int64))
        assert (counts.view(world, nq).sum(0) == k).all()  # every selected video has exactly one owner
        # owner exchange: each query's per-rank lists meet on the rank that owns the query; gathering the owners'
        # slices restores what a plain all-gather delivers
        from tvretrieval_b200.sharding import exchange_to_owners, gather_from_owners
        o_val, o_idx = exchange_to_owners(val, idx)
        q_lo, q_hi, per = plan.query_range(nq)
        assert o_val.shape == (per, world * k) and o_idx.dtype == torch.int32
        assert torch.equal(o_val[:q_hi - q_lo], g_val[q_lo:q_hi]) and torch.equal(o_idx[:q_hi - q_lo], g_idx[q_lo:q_hi])
        assert torch.equal(gather_from_owners(o_val, nq), g_val) and torch.equal(gather_from_owners(o_idx, nq), g_idx)
        # query split: rank-major concatenation restores query order
        q_lo, q_hi, per = plan.query_range(nq)
        packed = torch.full((per, 3), -1.0)
        packed[:q_hi - q_lo] = torch.arange(q_lo, q_hi, dtype=torch.float32)[:, None]
        allq = all_gather_rows(packed)[:nq]
        assert torch.equal(allq[:, 0], torch.arange(nq, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos,nq,k", [(37, 11, 7), (16, 5, 8)])
def test_two_rank_gather_and_merge(n_videos, nq, k):
    port = 29500 + (os.getpid() + n_videos) % 2000
    mp.spawn(_worker, args=(2, port, n_videos, nq, k), nprocs=2, join=True)


def test_shard_plan_covers_corpus():
    from tvretrieval_b200.sharding import ShardPlan
    for n, g in ((21793, 8), (21793, 4), (10, 3), (100000, 8)):
        edges = [ShardPlan(n, g, r).video_range() for r in range(g)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(g - 1))
        ids = torch.arange(n)
        owner = ShardPlan(n, g, 0).owner_of(ids)
        for r, (lo, hi) in enumerate(edges):
            assert (owner[lo:hi] == r).all()
        qs = [ShardPlan(n, g, r).query_range(1003) for r in range(g)]
        assert qs[0][0] == 0 and max(q[1] for q in qs) == 1003


def _ddp_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tvretrieval_b200.sharding import all_reduce_gradients
        g = torch.Generator().manual_seed(3)
        shapes = [(5, 4), (7,), (1, 1, 5), (3,)]
        params = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes]
        grads = [[torch.randn(*s, generator=g) for s in shapes] for _ in range(world)]  # same on every rank
        for i, p in enumerate(params):
            p.grad = grads[rank][i].clone()
        params[3].grad = None if rank == 1 else params[3].grad  # a parameter unused on one rank
        frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
        all_reduce_gradients(params + [frozen])
        for i, p in enumerate(params):
            want = sum(grads[r][i] for r in range(world) if not (i == 3 and r == 1)) / world
            assert torch.allclose(p.grad, want, atol=1e-7), i
        assert frozen.grad is None
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_average():
    mp.spawn(_ddp_worker, args=(2, 29500 + (os.getpid() + 977) % 2000), nprocs=2, join=True)
