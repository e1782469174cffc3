"""Video-sharded multi-GPU search (one process per GPU, torch.distributed over NCCL / NVLink).

The reference has no distributed inference path; this is new design.  The corpus is
partitioned by video into contiguous ranges (rank r holds videos [r*Nv/G, (r+1)*Nv/G)), everything up to the
per-(query, video) start/end distributions is independent per video, and the two places where videos couple --
top-k videos over the whole corpus and the top-k moments over the selected videos -- are resolved exactly with
two small all-gathers:

  1. query encoding is split by query across ranks, the pooled (Nq, H) query vectors are all-gathered;
  2. each rank scores its shard and keeps its local top-K videos; all-gather #1 of (score, global video id);
     every rank merges to the same global top-K (ranked by score desc, video id asc);
  3. each rank computes span distributions and its local top-M moments only for the selected videos it owns, with
     flat indices expressed in GLOBAL rank coordinates; all-gather #2 of (score, flat index); merge.

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


def merge_ranked_lists(val, idx, k, group=None, tie_desc=False):
    """All-gathers per-rank ranked candidate lists (val, idx) of shape (n, k_loc) -- missing entries marked by
    idx < 0 -- and returns the global top-k by (val desc, idx asc|desc); missing entries come back as (-1, 0)."""
    val = torch.where(idx < 0, torch.full_like(val, -1.0), val)  # missing entries rank below every candidate
    g_val, g_idx = all_gather_cat(val, group), all_gather_cat(idx, group)
    # the selection kernel wants distinct ids: give the missing entries distinct negative ones
    cols = torch.arange(g_idx.shape[1], device=g_idx.device, dtype=g_idx.dtype)
    g_idx = torch.where(g_idx < 0, -1 - cols, g_idx)
    idx, val = ops.topk_rows(g_val, k, ids=g_idx, tie_desc=tie_desc)
    missing = idx < 0
    return torch.where(missing, torch.full_like(idx, -1), idx).contiguous(), \
        torch.where(missing, torch.zeros_like(val), val).contiguous()


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

    def _my_slice(self, n):
        lo, hi, _ = self.plan.query_range(n)
        return lo, hi

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
        if "VR" in tasks or "VCMR" in tasks:
            k = self.max_n_videos
            k_loc = min(k, ix.n_videos)
            idx, val = self.top_videos(video_query, sub_query, k_loc)  # global ids, exp(alpha * score)
            if k_loc < k:  # shard smaller than k: pad so every rank contributes k columns
                idx = torch.cat([idx, idx.new_full((nq, k - k_loc), 2 ** 31 - 1)], 1)
                val = torch.cat([val, val.new_full((nq, k - k_loc), NEG)], 1)
            with self._phase("merge_videos"):
                g_val, g_idx = all_gather_cat(val, self.group), all_gather_cat(idx, self.group)
                res.top_video_idx, res.top_video_score = ops.topk_rows(g_val, k, ids=g_idx)
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
                idx, val = merge_ranked_lists(val, idx, m, self.group)
                res.span_flat_idx, res.span_score = ops.span_zero_fill(idx, val, k * ix.ctx_len * ix.ctx_len)
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
            idx, val = merge_ranked_lists(val, idx, m, self.group, tie_desc=True)
            res.svmr_flat_idx, res.svmr_score = ops.span_zero_fill(idx, val, ix.ctx_len * ix.ctx_len, tie_desc=True)
        return res
