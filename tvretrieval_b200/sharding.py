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


class SymmWorkspace:
    """Peer-mapped ("symmetric") workspace of the sharded search: the same allocation on every GPU of the node,
    each rank holding the addresses of all of them (torch.distributed._symmetric_memory: CUDA VMM + NVLink / NVSwitch
    mappings).  The ranking kernels store their lists directly into the other ranks' workspaces (ops.PeerSpec) and a
    device-side barrier on the signal pads separates producers from consumers -- the query path then contains no
    NCCL collective, no pack / unpack pass and no host synchronisation."""

    def __init__(self, group, device, nbytes):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.world, self.rank = self.hdl.world_size, self.hdl.rank
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        assert self.base[self.rank] == self.buf.data_ptr()
        self.nbytes, self.cursor = int(nbytes), 0

    def alloc(self, nbytes):
        off = self.cursor
        self.cursor = (off + int(nbytes) + 255) // 256 * 256
        assert self.cursor <= self.nbytes, "symmetric workspace too small"
        return off

    def ptrs(self, off):
        return [b + off for b in self.base]

    def view(self, off, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        return self.buf[off:off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(shape)

    def barrier(self):
        """All ranks' earlier stores (to any rank) are visible to all ranks' later kernels; stream-ordered."""
        self.hdl.barrier(channel=0)


class PeerExchange:
    """The four coupling points of the sharded search over a SymmWorkspace (see ShardedSearcher)."""

    def __init__(self, group, device, plan, max_queries, hidden, k_videos, k_spans):
        self.world, self.rank = plan.world_size, plan.rank
        per = (max_queries + self.world - 1) // self.world
        g, f = self.world, 4
        sizes = dict(q_all=g * per * 2 * hidden * f, kth_own=g * per * k_videos * f, kth_all=g * per * f,
                     vid_own_i=g * per * k_videos * f, vid_own_v=g * per * k_videos * f,
                     vid_all_i=g * per * k_videos * f, vid_all_v=g * per * k_videos * f,
                     span_own_i=g * per * k_spans * f, span_own_v=g * per * k_spans * f,
                     span_all_i=g * per * k_spans * f, span_all_v=g * per * k_spans * f)
        self.ws = SymmWorkspace(group, device, sum((v + 255) // 256 * 256 for v in sizes.values()) + 4096)
        self.off = {k: self.ws.alloc(v) for k, v in sizes.items()}
        self.max_per, self.hidden, self.k_videos, self.k_spans = per, hidden, k_videos, k_spans

    def per(self, n):
        per = (n + self.world - 1) // self.world
        assert per <= self.max_per, "query block larger than the symmetric workspace was sized for"
        return per

    def owned(self, n):
        per = self.per(n)
        return max(0, min(per, n - self.rank * per))

    def to_owner(self, name_i, name_v, per):
        return ops.PeerSpec(self.ws.ptrs(self.off[name_i]) if name_i else None,
                            self.ws.ptrs(self.off[name_v]) if name_v else None, 1, per, self.rank)

    def to_all(self, name_i, name_v, per):
        return ops.PeerSpec(self.ws.ptrs(self.off[name_i]) if name_i else None,
                            self.ws.ptrs(self.off[name_v]) if name_v else None, 2, per, self.rank)

    def region(self, name, shape, dtype):
        return self.ws.view(self.off[name], shape, dtype)

    # ---- coupling point 0: the pooled query vectors of every rank's slice -> all ranks
    def gather_queries(self, packed, n):
        """packed (per, 2, H) fp32: this rank's slice (zero rows beyond it) -> (n, 2, H) of all ranks."""
        per = self.per(n)
        assert packed.shape == (per, 2, self.hidden)
        off = self.off["q_all"] + self.rank * per * 2 * self.hidden * 4
        ops.peer_copy(packed.contiguous(), self.ws.ptrs(off))
        self.ws.barrier()
        return self.region("q_all", (self.world * per, 2, self.hidden), torch.float32)[:n]

    # ---- coupling point 1: corpus-wide k-th largest approximate score
    def global_kth(self, approx, k):
        n, per, kv = len(approx), self.per(len(approx)), self.k_videos
        assert k == kv
        k_loc = min(k, approx.shape[1])
        ops.topk_rows(approx, k_loc, pad=(k, 0, NEG) if k_loc < k else None, want_idx=False,
                      peer=self.to_owner(None, "kth_own", per))
        self.ws.barrier()
        own = self.region("kth_own", (self.world * per * k,), torch.float32)
        ops.topk_rows(own, k, n_rows=self.owned(n), segments=(k, per * k, self.world * k), out_slice=(k - 1, 1),
                      peer=self.to_all(None, "kth_all", per))
        self.ws.barrier()
        return self.region("kth_all", (self.world * per,), torch.float32)[:n]

    # ---- coupling point 2: exact per-rank video lists -> global top-k on every rank
    def merge_videos(self, idx, val, k):
        n, per = len(idx), self.per(len(idx))
        k_loc = idx.shape[1]
        # (already ranked lists: the k_loc-entry "selection" is the identity; it is the kernel that stores to peers)
        ops.topk_rows(val, k_loc, ids=idx, pad=(k, 2 ** 31 - 1, NEG) if k_loc < k else None,
                      peer=self.to_owner("vid_own_i", "vid_own_v", per))
        self.ws.barrier()
        ops.topk_rows(self.region("vid_own_v", (self.world * per * k,), torch.float32), k,
                      ids=self.region("vid_own_i", (self.world * per * k,), torch.int32), n_rows=self.owned(n),
                      segments=(k, per * k, self.world * k), peer=self.to_all("vid_all_i", "vid_all_v", per))
        self.ws.barrier()
        return (self.region("vid_all_i", (self.world * per, k), torch.int32)[:n].clone(),
                self.region("vid_all_v", (self.world * per, k), torch.float32)[:n].clone())

    # ---- coupling point 3: per-rank moment lists -> global top-m on every rank
    def span_peer(self, n):
        return self.to_owner("span_own_i", "span_own_v", self.per(n))

    def merge_spans(self, n, m, total_cells, tie_desc=False):
        """The per-rank lists were stored by span_topk(peer=span_peer(n)); -> (flat idx, score) (n, m) on all ranks."""
        per = self.per(n)
        assert m == self.k_spans
        self.ws.barrier()
        owned = self.owned(n)
        dev = self.ws.buf.device
        m_idx = torch.empty(max(owned, 1), m, device=dev, dtype=torch.int32)
        m_val = torch.empty(max(owned, 1), m, device=dev, dtype=torch.float32)
        if owned:
            ops.topk_rows(self.region("span_own_v", (self.world * per * m,), torch.float32), m,
                          ids=self.region("span_own_i", (self.world * per * m,), torch.int32), n_rows=owned,
                          segments=(m, per * m, self.world * m), missing_neg=True, tie_desc=tie_desc,
                          out=(m_idx[:owned], m_val[:owned]))
            ops.span_zero_fill(m_idx[:owned], m_val[:owned], total_cells, tie_desc=tie_desc,
                               peer=self.to_all("span_all_i", "span_all_v", per))
        self.ws.barrier()
        return (self.region("span_all_i", (self.world * per, m), torch.int32)[:n].clone(),
                self.region("span_all_v", (self.world * per, m), torch.float32)[:n].clone())


def local_slot_mask(top_ids, vid_lo, vid_hi):
    """1 where a selected (global) video id belongs to this rank's shard."""
    return ((top_ids >= vid_lo) & (top_ids < vid_hi)).to(torch.uint8)


class ShardedSearcher(VCMRSearcher):
    """Same `search()` contract as VCMRSearcher; `index` holds only this rank's videos (index.vid_lo set)."""

    def __init__(self, model, index, n_videos_total, group=None, transport=None, **kw):
        """transport: "peer" = the ranking kernels store their lists straight into the other GPUs' symmetric
        workspaces over NVLink (PeerExchange; needs NCCL + torch symmetric memory, one GPU per rank), "collective" =
        all_to_all / all_gather through torch.distributed (any backend; used by the single-GPU gloo tests).
        None = "peer" when available."""
        super().__init__(model, index, **kw)
        self.group = group
        self.plan = ShardPlan(n_videos_total, dist.get_world_size(group), dist.get_rank(group))
        self.peer = None
        auto = transport is None
        if auto:
            transport = "peer" if (dist.get_backend(group) == "nccl" and self.plan.world_size <= 8) else "collective"
        if transport == "peer":
            try:
                self.peer = PeerExchange(group, index.device, self.plan, self.query_chunk, model.config.hidden_size,
                                         self.max_n_videos, self.max_before_nms)
            except Exception as exc:  # no peer access between the GPUs / symmetric memory not available
                if not auto:
                    raise
                import sys
                print("xmlb200: symmetric memory unavailable (%s: %s); list exchanges use torch.distributed collectives"
                      % (type(exc).__name__, exc), file=sys.stderr)
                ok = torch.zeros(1, device=index.device)
            else:
                ok = torch.ones(1, device=index.device)
            if auto:  # every rank must use the same transport
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if float(ok) == 0:
                    self.peer, transport = None, "collective"
        self.transport = transport
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

    def _n_ranks(self):
        return self.plan.world_size

    def _encode_block_pipelined(self, query_feat_cpu, query_mask_cpu, n):
        """Host-buffer search, encoding + filter pass of one block of n queries: the block is cut into GLOBAL pieces
        (the geometric ramp of _piece_bounds) and every rank uploads and encodes its 1 / world share of EACH piece;
        the pooled vectors of a piece are shared (peer stores into every rank's symmetric workspace + one device-side
        barrier, or one all-gather) and the filter pass over this rank's shard runs on the piece while the shares of
        the next piece are still uploading.  With one contiguous slice per rank (search() from device-resident
        queries) the filter cannot start before every rank has uploaded and encoded everything: end to end the
        uploads of the 8 ranks -- which share the host's PCIe / memory bandwidth -- were fully exposed.
        -> pooled (video_query, sub_query) of all n queries; self._prefilter holds the finished filter pass."""
        from .engine import _Prefilter
        world, rank = self.plan.world_size, self.plan.rank
        dev, hid = self.index.device, self.model.config.hidden_size
        pieces = self._piece_bounds(n, True)
        shares, mine = [], []
        for g_lo, g_hi in pieces:
            q = -(-(g_hi - g_lo) // world)
            lo = min(g_hi, g_lo + rank * q)
            shares.append(q), mine.append((lo, min(g_hi, lo + q)))
        width = query_mask_cpu.shape[1]
        lens = None
        if (self.packed_queries and n // world >= self.packed_min_queries and width <= self.model.PACKED_MAX_LEN):
            lens = (query_mask_cpu != 0).sum(1).to(torch.int64)
        pre = _Prefilter(self, n)
        if self.peer is not None:
            per = self.peer.per(n)
            all_q = self.peer.region("q_all", (world * per, 2, hid), torch.float32)[:n]
        else:
            all_q = torch.empty(n, 2, hid, device=dev)
        done = [0]

        def on_piece(_lo, _hi, a, b):
            j = done[0]
            done[0] += 1
            (g_lo, g_hi), (lo, hi) = pieces[j], mine[j]
            with self._phase("gather_queries"):
                if self.peer is not None:
                    if hi > lo:
                        off = self.peer.off["q_all"] + lo * 2 * hid * 4
                        ops.peer_copy(torch.stack([a, b], 1).contiguous(), self.peer.ws.ptrs(off))
                    self.peer.ws.barrier()
                else:
                    share = torch.zeros(shares[j], 2, hid, device=dev)
                    if hi > lo:
                        share[:hi - lo, 0], share[:hi - lo, 1] = a, b
                    all_q[g_lo:g_hi] = all_gather_rows(share, self.group)[:g_hi - g_lo]
            pre.run(g_lo, g_hi, all_q[g_lo:g_hi, 0].contiguous(), all_q[g_lo:g_hi, 1].contiguous())

        self._encode_pieces(self._host_pieces(query_feat_cpu, query_mask_cpu, mine), lens, on_piece, width=width,
                            bounds=mine)
        self._prefilter = pre
        return all_q[:, 0].contiguous(), all_q[:, 1].contiguous()

    def _global_kth(self, approx, k):
        """The k-th largest APPROXIMATE score over all shards: local top-k values, one all-gather of (Nq, k) floats
        per rank, k-th largest of the union.  With it every rank keeps only the candidates that can reach the
        corpus-wide top-k (~(k + margin) / G per rank) instead of the candidates of its own local top-k (~k per
        rank), so the exact re-scoring work scales with 1 / G."""
        if self.peer is not None:
            with self._phase("gather_kth"):
                return self.peer.global_kth(approx, k)
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
            allq = self.peer.gather_queries(packed, n) if self.peer is not None else \
                all_gather_rows(packed, self.group)[:n]
        return allq[:, 0].contiguous(), allq[:, 1].contiguous()

    def _merge_spans(self, st, ed, video_score, valid, nq, n_slots, m, tie_desc=False):
        """Local top-m moments of this rank's selected videos -> global top-m on every rank (+ zero fill)."""
        ix = self.index
        total = n_slots * ix.ctx_len * ix.ctx_len
        st, ed = st.view(nq, n_slots, ix.ctx_len), ed.view(nq, n_slots, ix.ctx_len)
        if self.peer is not None:
            with self._phase("span_topk"):  # the lists go straight to the ranks that own the queries
                ops.span_topk(st, ed, video_score, self.min_pred_l, self.max_pred_l, m, slot_valid=valid,
                              tie_desc=tie_desc, zero_fill=False, peer=self.peer.span_peer(nq))
            with self._phase("merge_spans"):
                return self.peer.merge_spans(nq, m, total, tie_desc=tie_desc)
        with self._phase("span_topk"):
            idx, val = ops.span_topk(st, ed, video_score, self.min_pred_l, self.max_pred_l, m, slot_valid=valid,
                                     tie_desc=tie_desc, zero_fill=False)
        with self._phase("merge_spans"):
            return merge_ranked_lists(val, idx, m, self.group, tie_desc=tie_desc,
                                      finish=lambda i, v: ops.span_zero_fill(i, v, total, tie_desc=tie_desc))

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
            with self._phase("merge_videos"):
                if self.peer is not None:
                    res.top_video_idx, res.top_video_score = self.peer.merge_videos(idx, val, k)
                else:
                    if k_loc < k:  # shard smaller than k: pad so every rank contributes k columns
                        idx = torch.cat([idx, idx.new_full((nq, k - k_loc), 2 ** 31 - 1)], 1)
                        val = torch.cat([val, val.new_full((nq, k - k_loc), NEG)], 1)
                    o_val, o_idx = exchange_to_owners(val, idx, self.group)
                    m_idx, m_val = ops.topk_rows(o_val, k, ids=o_idx)
                    both = gather_from_owners(torch.stack([m_val.view(torch.int32), m_idx], dim=2), nq, self.group)
                    res.top_video_idx = both[..., 1].contiguous()
                    res.top_video_score = both[..., 0].contiguous().view(torch.float32)
        if "VCMR" in tasks:
            with self._phase("pair_lists"):
                valid = local_slot_mask(res.top_video_idx, vid_lo, vid_hi)
                lists = self.span_lists(res.top_video_idx, slot_valid=valid)
            with self._phase("span_probs"):
                st, ed = self.span_probs(video_query, sub_query, lists)
            res.span_flat_idx, res.span_score = self._merge_spans(st, ed, res.top_video_score, valid, nq, k,
                                                                  self.max_before_nms)
        if "SVMR" in tasks:
            assert gt_video_idx is not None, "SVMR needs the ground-truth video of every query"
            gt = gt_video_idx.view(nq, 1).to(torch.int32)
            valid = local_slot_mask(gt, vid_lo, vid_hi)
            lists = self.span_lists(gt, slot_valid=valid)
            st, ed = self.span_probs(video_query, sub_query, lists)
            # exactly one rank owns each query's ground-truth video: the others contribute only (-1, 0) rows
            res.svmr_flat_idx, res.svmr_score = self._merge_spans(st, ed, None, valid, nq, 1, self.max_before_nms,
                                                                  tie_desc=True)
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
