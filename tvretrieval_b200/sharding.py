"""Video-sharded multi-GPU search (one process per GPU, torch.distributed over NCCL / NVLink).

The reference has no distributed inference path; this is new design.  The corpus is
partitioned by video into contiguous ranges (rank r holds videos [r*Nv/G, (r+1)*Nv/G)), everything up to the
per-(query, video) start/end distributions is independent per video, and the two places where videos couple --
top-k videos over the whole corpus and the top-k moments over the selected videos -- are resolved exactly with
small exchanges of ranked (score, id) lists.  Every list exchange has the same shape:
all-to-all so that each query's lists meet on the rank that OWNS the query (queries are split evenly), merge
there, all-gather the merged lists -- 1 / G of the traffic and of the merge work of a plain all-gather.

  1. query encoding is split by query across ranks, the pooled (Nq, H) query vectors are all-gathered;
  2. each rank runs the filter pass on its shard; the corpus-wide K-th largest approximate score is found from the
     ranks' local top-K values (exchange #1), so each rank re-scores exactly only ITS part of the global
     candidate set; exchange #2 merges the exact (score, global video id) lists into the global top-K
     (ranked by score desc, video id asc), known to every rank;
  3. each rank computes span distributions and its local top-M moments only for the selected videos it owns, with
     flat indices expressed in GLOBAL rank coordinates; exchange #3 merges (score, flat index).

The result on every rank is identical to the single-GPU result (same arithmetic per cell, same canonical ranking).
"""
import torch
import torch.distributed as dist

from . import ops
from .engine import SearchResult, VCMRSearcher

NEG = -3.0e38


class ShardPlan:
    """Pure index arithmetic of the partition (testable without a GPU)."""

    def __init__(self, n_videos_total, world_size, rank):
        assert 0 <= rank < world_size and n_videos_total >= world_size
        self.n_videos_total, self.world_size, self.rank = n_videos_total, world_size, rank

    def video_range(self, rank=None):
        r = self.rank if rank is None else rank
        return r * self.n_videos_total // self.world_size, (r + 1) * self.n_videos_total // self.world_size

    def query_range(self, n_queries, rank=None):
        """Equal-size query slices (padded so that all_gather sees equal shapes): -> (lo, hi, per_rank)."""
        r = self.rank if rank is None else rank
        per = (n_queries + self.world_size - 1) // self.world_size
        return min(r * per, n_queries), min((r + 1) * per, n_queries), per

    def owner_of(self, video_ids):
        """rank owning each global video id (LongTensor in, LongTensor out)."""
        bounds = torch.tensor([self.video_range(r)[1] for r in range(self.world_size)], device=video_ids.device)
        return torch.bucketize(video_ids, bounds, right=True)


def _gather(out, t, group):
    """all_gather_into_tensor; gloo cannot gather CUDA tensors, so they are staged through the host there (only
    used by the single-GPU equivalence tests -- production runs use NCCL)."""
    if t.is_cuda and dist.get_backend(group) == "gloo":
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(host, t.cpu(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, t, group=group)


def all_gather_cat(t, group=None):
    """(n, k) on every rank -> (n, world * k): rank r's block occupies columns [r*k, (r+1)*k).  Works with NCCL
    (CUDA tensors) and gloo (CPU tensors)."""
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
    _gather(out, t, group)
    return out.view(world, t.shape[0], t.shape[1]).permute(1, 0, 2).reshape(t.shape[0], world * t.shape[1]).contiguous()


def all_gather_rows(t, group=None):
    """(n, ...) on every rank -> (world * n, ...), rank-major."""
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _gather(out, t, group)
    return out


def _all_to_all(out, inp, group):
    """all_to_all_single with equal splits (gloo: CUDA tensors are staged through the host, as in _gather)."""
    if inp.is_cuda and dist.get_backend(group) == "gloo":
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_to_all_single(host, inp.cpu(), group=group)
        out.copy_(host)
    else:
        dist.all_to_all_single(out, inp, group=group)


def exchange_to_owners(val, idx, group=None):
    """Per-rank candidate lists (val fp32, idx int32) of shape (n, kl) for ALL n queries -> the lists of every rank
    for the queries THIS rank owns (rank r owns rows [r * per, (r + 1) * per), per = ceil(n / world)):
    (per, world * kl) val and idx, rank p's block in columns [p * kl, (p + 1) * kl).  One all-to-all of 8-byte
    (score, id) pairs: each rank receives 1 / world of what an all-gather would deliver."""
    world = dist.get_world_size(group)
    n, kl = val.shape
    per = (n + world - 1) // world
    packed = torch.zeros(world * per, kl, 2, dtype=torch.int32, device=val.device)
    packed[:n, :, 0] = val.contiguous().view(torch.int32)
    packed[:n, :, 1] = idx
    out = torch.empty_like(packed)
    _all_to_all(out, packed, group)
    out = out.view(world, per, kl, 2).permute(1, 0, 2, 3).reshape(per, world * kl, 2)
    return out[..., 0].contiguous().view(torch.float32), out[..., 1].contiguous()


def gather_from_owners(t, n, group=None):
    """(per, ...) results of the queries this rank owns -> (n, ...) for all queries on every rank."""
    return all_gather_rows(t, group)[:n]


def merge_ranked_lists(val, idx, k, group=None, tie_desc=False, finish=None):
    """Global top-k of per-rank ranked candidate lists (val, idx) of shape (n, k_loc), ranked (val desc, idx
    asc|desc); missing entries are marked by idx < 0 and come back as (-1, 0).  Each query is merged on the rank
    that owns it (exchange_to_owners), `finish(idx, val)` post-processes the owner's slice, and the merged lists
    are all-gathered: every rank returns the full (n, k) result."""
    n = len(val)
    val = torch.where(idx < 0, torch.full_like(val, -1.0), val)  # missing entries rank below every candidate
    o_val, o_idx = exchange_to_owners(val, idx, group)
    # the selection kernel wants distinct ids: give the missing entries distinct negative ones
    cols = torch.arange(o_idx.shape[1], device=o_idx.device, dtype=o_idx.dtype)
    o_idx = torch.where(o_idx < 0, -1 - cols, o_idx)
    m_idx, m_val = ops.topk_rows(o_val, k, ids=o_idx, tie_desc=tie_desc)
    missing = m_idx < 0
    m_idx = torch.where(missing, torch.full_like(m_idx, -1), m_idx).contiguous()
    m_val = torch.where(missing, torch.zeros_like(m_val), m_val).contiguous()
    if finish is not None:
        m_idx, m_val = finish(m_idx, m_val)
    both = gather_from_owners(torch.stack([m_val.view(torch.int32), m_idx], dim=2), n, group)
    return both[..., 1].contiguous(), both[..., 0].contiguous().view(torch.float32)


def local_slot_mask(top_ids, vid_lo, vid_hi):
    """1 where a selected (global) video id belongs to this rank's shard."""
    return ((top_ids >= vid_lo) & (top_ids < vid_hi)).to(torch.uint8)


class ShardedSearcher(VCMRSearcher):
    """Same `search()` contract as VCMRSearcher; `index` holds only this rank's videos (index.vid_lo set)."""

    def __init__(self, model, index, n_videos_total, group=None, **kw):
        super().__init__(model, index, **kw)
        self.group = group
        self.plan = ShardPlan(n_videos_total, dist.get_world_size(group), dist.get_rank(group))
        lo, hi = self.plan.video_range()
        assert index.vid_lo == lo and index.n_videos == hi - lo, "index does not hold this rank's shard"
        # The candidate filter compares scores of different shards against one corpus-wide threshold, so its error
        # bound must hold on every shard: use the largest corpus-side bound (tc_err, a host float per modality).
        tc_err = getattr(index, "tc_err", None)
        if tc_err:
            names = sorted(tc_err)
            t = torch.tensor([tc_err[n] for n in names], dtype=torch.float32)
            if dist.get_backend(group) != "gloo":
                t = t.to(index.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            index.tc_err = {n: float(v) for n, v in zip(names, t.cpu())}

    def _my_slice(self, n):
        lo, hi, _ = self.plan.query_range(n)
        return lo, hi

    def _global_kth(self, approx, k):
        """The k-th largest APPROXIMATE score over all shards: local top-k values, one all-gather of (Nq, k) floats
        per rank, k-th largest of the union.  With it every rank keeps only the candidates that can reach the
        corpus-wide top-k (~(k + margin) / G per rank) instead of the candidates of its own local top-k (~k per
        rank), so the exact re-scoring work scales with 1 / G."""
        with self._phase("gather_kth"):
            k_loc = min(k, approx.shape[1])
            idx, val = ops.topk_rows(approx, k_loc)
            if k_loc < k:
                val = torch.cat([val, val.new_full((len(val), k - k_loc), NEG)], 1)
                idx = torch.cat([idx, idx.new_zeros((len(idx), k - k_loc))], 1)
            o_val, _ = exchange_to_owners(val, idx, self.group)  # merged on the rank that owns the query
            _, top = ops.topk_rows(o_val, k)
            return gather_from_owners(top[:, k - 1].contiguous(), len(approx), self.group).contiguous()

    def _gather_encoded(self, video_query, sub_query, n):
        """Each rank encoded its slice of the block; the pooled vectors are all-gathered (rank-major = query order)."""
        lo, hi, per = self.plan.query_range(n)
        hid = self.model.config.hidden_size
        packed = torch.zeros(per, 2, hid, device=self.index.device)
        if hi > lo:
            packed[:hi - lo, 0], packed[:hi - lo, 1] = video_query, sub_query
        with self._phase("gather_queries"):
            allq = all_gather_rows(packed, self.group)[:n]
        return allq[:, 0].contiguous(), allq[:, 1].contiguous()

    def _search_encoded(self, video_query, sub_query, gt_video_idx, tasks):
        ix = self.index
        res = SearchResult()
        vid_lo, vid_hi = ix.vid_lo, ix.vid_lo + ix.n_videos
        nq = len(video_query)
        k = self.max_n_videos
        if ("VR" in tasks or "VCMR" in tasks) and self._external is not None:
            res.top_video_idx, res.top_video_score = self._external  # the same lists on every rank: nothing to merge
        elif "VR" in tasks or "VCMR" in tasks:
            k_loc = min(k, ix.n_videos)
            idx, val = self.top_videos(video_query, sub_query, k_loc, k_global=k)  # global ids, exp(alpha * score)
            if k_loc < k:  # shard smaller than k: pad so every rank contributes k columns
                idx = torch.cat([idx, idx.new_full((nq, k - k_loc), 2 ** 31 - 1)], 1)
                val = torch.cat([val, val.new_full((nq, k - k_loc), NEG)], 1)
            with self._phase("merge_videos"):
                o_val, o_idx = exchange_to_owners(val, idx, self.group)
                m_idx, m_val = ops.topk_rows(o_val, k, ids=o_idx)
                both = gather_from_owners(torch.stack([m_val.view(torch.int32), m_idx], dim=2), nq, self.group)
                res.top_video_idx = both[..., 1].contiguous()
                res.top_video_score = both[..., 0].contiguous().view(torch.float32)
        if "VCMR" in tasks:
            m = self.max_before_nms
            with self._phase("pair_lists"):
                valid = local_slot_mask(res.top_video_idx, vid_lo, vid_hi)
                lists = self.span_lists(res.top_video_idx, slot_valid=valid)
            with self._phase("span_probs"):
                st, ed = self.span_probs(video_query, sub_query, lists)
            st, ed = st.view(nq, k, ix.ctx_len), ed.view(nq, k, ix.ctx_len)
            with self._phase("span_topk"):
                idx, val = ops.span_topk(st, ed, res.top_video_score, self.min_pred_l, self.max_pred_l, m,
                                         slot_valid=valid, zero_fill=False)
            with self._phase("merge_spans"):
                total = k * ix.ctx_len * ix.ctx_len
                res.span_flat_idx, res.span_score = merge_ranked_lists(
                    val, idx, m, self.group, finish=lambda i, v: ops.span_zero_fill(i, v, total))
        if "SVMR" in tasks:
            assert gt_video_idx is not None, "SVMR needs the ground-truth video of every query"
            m = self.max_before_nms
            gt = gt_video_idx.view(nq, 1).to(torch.int32)
            valid = local_slot_mask(gt, vid_lo, vid_hi)
            lists = self.span_lists(gt, slot_valid=valid)
            st, ed = self.span_probs(video_query, sub_query, lists)
            idx, val = ops.span_topk(st.view(nq, 1, ix.ctx_len), ed.view(nq, 1, ix.ctx_len), None, self.min_pred_l,
                                     self.max_pred_l, m, slot_valid=valid, tie_desc=True, zero_fill=False)
            # exactly one rank owns each query's ground-truth video: the others contribute only (-1, 0) rows
            cells = ix.ctx_len * ix.ctx_len
            res.svmr_flat_idx, res.svmr_score = merge_ranked_lists(
                val, idx, m, self.group, tie_desc=True,
                finish=lambda i, v: ops.span_zero_fill(i, v, cells, tie_desc=True))
        return res


# ------------------------------------------------------------------------------------------------ data-parallel training
def all_reduce_gradients(parameters, group=None):
    """Data-parallel training step (BASELINE config #4 on several GPUs; SURVEY.md section 8e "Training"): every rank
    ran XML.forward / backward on its own mini-batch (in-batch negatives stay per rank, like the reference under
    DataParallel); the gradients of all parameters are averaged with ONE all-reduce over a flat buffer (81 MB for the
    20.2 M-parameter video_sub model) and written back in place, after which every rank applies the same
    BertAdam.step().  Parameters without a gradient on this rank contribute zeros."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return
    world = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
