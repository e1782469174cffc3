"""Device-resident VCMR / VR / SVMR search engine: the query-side hot path kept on the GPU end to end.

One `search()` call = the tensor section of reference `compute_query2ctx_info`
(baselines/crossmodal_moment_localization/inference.py:302-389) for a batch of raw queries, restructured as
  phase 1  query encoders -> video-level scores for the whole corpus -> exact top-k videos   (K1-K6, K10)
  phase 2  similarity curves + ConvSE + softmax ONLY for the selected (query, video) pairs,
           grouped per video through inverted lists so each video is read once per 32 queries  (K7-K9, K11)
  phase 3  band-limited span scores + exact top-k moments                                    (K12, K13)
instead of materialising (Nq, Nv, L) logits and sorting 100*L*L cells per query.  Results are identical to the
reference's (same arithmetic per cell; ranking = score desc, index asc).
"""
import contextlib
import os

import numpy as np
import torch

from . import ops


@contextlib.contextmanager
def host_memory_near(device):
    """Context in which the calling thread runs on the CPUs next to `device` (NVML's ideal CPU affinity of the GPU),
    so that host buffers allocated AND first touched inside it -- the pinned query / result buffers of
    VCMRSearcher.search_host -- live on the GPU's own NUMA node: an upload from the other socket's memory runs at
    about half the PCIe rate, and the ranks of a multi-GPU search otherwise all pull from whichever node their
    processes happened to start on.  The previous affinity is restored on exit; without NVML this does nothing."""
    old = None
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        try:
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
        except Exception:
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        old = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
    except Exception:
        old = None
    try:
        yield
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except OSError:
                pass


class PhaseTimer:
    """CUDA-event timers around the phases of a search (events are recorded on the stream the kernels are
    launched on).  bench.py uses it to report the dominant kernel's duration inside the timed region."""

    def __init__(self):
        self.spans = {}

    @contextlib.contextmanager
    def phase(self, name):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        yield
        end.record()
        self.spans.setdefault(name, []).append((start, end))

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: sum(s.elapsed_time(e) for s, e in v) for k, v in self.spans.items()}

    def counts(self):
        return {k: len(v) for k, v in self.spans.items()}

    def reset(self):
        self.spans = {}


class CorpusPacking:
    """Ragged layout of the corpus operand of the tensor-core VR kernel: only valid clips are stored; whole videos
    are packed (longest first) into tiles of <= 256 consecutive rows, so a tile never splits a video."""
    TILE = 256
    MAX_VIDEOS = 32  # per tile (shared-memory scratch of the kernel epilogue)

    def __init__(self, mask):
        valid = (mask != 0).cpu().numpy()
        n_videos, length = valid.shape
        assert length <= self.TILE, "a video must fit in one tile"
        lens = valid.sum(1)
        # best-fit decreasing bin packing: open a tile with the longest remaining video, then keep adding the longest
        # video that still fits (videos of equal length are taken in corpus order)
        by_len = [[] for _ in range(self.TILE + 1)]
        for v in range(n_videos - 1, -1, -1):
            if lens[v] > 0:
                by_len[lens[v]].append(v)  # pop() yields ascending video index
        cnt = np.asarray([len(b) for b in by_len])
        order, tile_of, metas = [], [], []  # metas: row_start, first ordinal, used columns
        row = 0
        while cnt.any():
            t_row, t_ord, rem = row, len(order), self.TILE
            while rem > 0 and len(order) - t_ord < self.MAX_VIDEOS:
                fits = np.nonzero(cnt[:rem + 1])[0]
                if len(fits) == 0:
                    break
                n = int(fits[-1])
                order.append(by_len[n].pop())
                tile_of.append(len(metas))
                cnt[n] -= 1
                rem -= n
                row += n
            metas.append((t_row, t_ord, self.TILE - rem, len(order) - t_ord))
        order = np.asarray(order, dtype=np.int64)
        tile_of = np.asarray(tile_of, dtype=np.int64)
        olens = lens[order]
        row_start = np.concatenate([[0], np.cumsum(olens)]).astype(np.int64)  # packed row of each video's first clip
        meta = np.asarray(metas, dtype=np.int32).reshape(-1, 4)
        starts = np.zeros((len(metas), 8), dtype=np.uint32)
        col = row_start[:-1] - meta[tile_of, 0]
        np.bitwise_or.at(starts, (tile_of, col // 32), (np.uint32(1) << (col % 32).astype(np.uint32)))
        ii, ll = np.nonzero(valid[order])
        dev = mask.device
        self.n_rows, self.n_tiles = int(row_start[-1]), len(metas)
        self.src_rows = torch.from_numpy((order[ii] * length + ll).astype(np.int32)).to(dev)
        self.tile_meta = torch.from_numpy(meta).to(dev)
        self.tile_starts = torch.from_numpy(starts.view(np.int32)).to(dev)
        self.order = torch.from_numpy(order.astype(np.int32)).to(dev)
        self.row_start = torch.from_numpy(row_start.astype(np.int32)).to(dev)  # (n_packed + 1,)
        self.max_len = int(olens.max()) if len(olens) else 1
        empty = np.nonzero(lens == 0)[0]
        self.n_packed = len(order)
        # ordinal -> video id for every column of the score matrix (videos without valid clips come last)
        self.order_full = torch.from_numpy(np.concatenate([order, empty]).astype(np.int32)).to(dev)
        self.empty = torch.from_numpy(empty.astype(np.int32)).to(dev)
        self.fill = self.n_rows / max(1, self.n_tiles * self.TILE)

    STATE = ("tile_meta", "tile_starts", "order", "row_start", "empty")  # device tensors that define the layout

    @classmethod
    def from_state(cls, state, max_len):
        """Rebuild from saved tables (CorpusIndex.load) -- src_rows is only needed while packing new features."""
        self = cls.__new__(cls)
        for k in cls.STATE:
            setattr(self, k, state[k])
        self.src_rows = None
        self._refresh(max_len)
        return self

    def _refresh(self, max_len):
        self.n_tiles, self.n_packed = len(self.tile_meta), len(self.order)
        self.n_rows = int(self.row_start[-1]) if len(self.row_start) else 0
        self.max_len = int(max_len)
        self.order_full = torch.cat([self.order, self.empty])
        self.fill = self.n_rows / max(1, self.n_tiles * self.TILE)

    def append(self, other, n_videos_before):
        """Append the packing of newly added videos (built on their own masks): their tiles, rows and ordinals
        follow the existing ones; video ids are shifted by the number of videos already in the index."""
        meta = other.tile_meta.clone()
        meta[:, 0] += self.n_rows
        meta[:, 1] += self.n_packed
        self.tile_meta = torch.cat([self.tile_meta, meta])
        self.tile_starts = torch.cat([self.tile_starts, other.tile_starts])
        self.row_start = torch.cat([self.row_start, other.row_start[1:] + self.n_rows])
        self.order = torch.cat([self.order, other.order + n_videos_before])
        self.empty = torch.cat([self.empty, other.empty + n_videos_before])
        self._refresh(max(self.max_len, other.max_len))


class CorpusIndex:
    """Encoded corpus resident in HBM.  Built once from the output of `compute_context_info`
    (reference inference.py:89-97); feat1 is L2-normalised at build time (the reference re-normalises it for
    every query batch, model_xml.py:447)."""

    PRECISIONS = ("f16x3", "bf16x3", "f32")

    @torch.no_grad()
    def __init__(self, video_feat1=None, video_feat2=None, video_mask=None, sub_feat1=None, sub_feat2=None,
                 sub_mask=None, vid_lo=0, precision="f16x3", packed=True, merged_spans=True, rescore_kblocked=False):
        """precision selects the video-level-score kernel: "f16x3" / "bf16x3" = tcgen05 tensor cores with
        hi/lo-split operands (3 MMAs per product, fp32-accurate), "f32" = exact-fp32 SIMT kernel.
        packed=True stores only the valid clips for the tensor-core kernel (CorpusPacking).
        merged_spans=True additionally keeps [feat2_video | feat2_sub] as hi/lo halves for the tensor-core
        similarity-curve kernel of the merged two-stream model (needs both modalities, L <= 256).
        rescore_kblocked=True keeps a second, K-BLOCKED copy of the packed corpus halves for the exact re-scoring
        kernel of the two-pass search (+1x the packed operand, 9.6 GB at the bench shape): it streams every
        candidate video's clips as one contiguous run per k-step instead of 64-byte pieces 2 Kpad bytes apart.
        Off by default: bit-equal but measured no faster (3.57 vs 3.59 ms per 10 K queries) -- unlike the similarity
        kernel, the re-scoring kernel waits on the L2 round trips of its query gather, not on the corpus stream."""
        assert precision in self.PRECISIONS, precision
        ref = video_feat1 if video_feat1 is not None else sub_feat1
        self.n_videos, self.ctx_len, self.hidden = ref.shape
        self.device = ref.device
        self.vid_lo = vid_lo  # global id of the first video (multi-GPU shards)
        self.precision = precision
        self.video_feat1n = self.sub_feat1n = None
        self.video_tc = self.sub_tc = self.video_bits = self.sub_bits = self.packing = self.f2cat = None
        self.video_tc_kb = self.sub_tc_kb = None
        if video_feat1 is not None and sub_feat1 is not None and not torch.equal(video_mask, sub_mask):
            packed = False  # the packed layout shares one packing between the modalities
        if precision == "f32":
            self.video_feat1n = ops.l2norm_rows(video_feat1) if video_feat1 is not None else None
            self.sub_feat1n = ops.l2norm_rows(sub_feat1) if sub_feat1 is not None else None
        else:
            # corpus operand of the tensor-core kernel: normalised, hi/lo split, each video's clips padded to a
            # multiple of 32 rows, hidden padded to a multiple of 64 (zero fill)
            self.lp = (self.ctx_len + 31) // 32 * 32
            self.kpad = (self.hidden + 63) // 64 * 64
            bf16 = precision == "bf16x3"
            if packed:
                self.packing = CorpusPacking(video_mask if video_mask is not None else sub_mask)
                # tc_err[x] = max over clips of ||c - c_hi||_2: error bound of the hi-only filter pass (two-pass search)
                self.tc_err = {}
                for name, feat in (("video", video_feat1), ("sub", sub_feat1)):
                    if feat is not None:
                        hi, lo, err = ops.split_rows(feat, kpad=self.kpad, normalize=True, bf16=bf16,
                                                     row_index=self.packing.src_rows, hi_err=True)
                        setattr(self, name + "_tc", (hi, lo))
                        if rescore_kblocked:
                            setattr(self, name + "_tc_kb", (ops.kblock_rows(hi), ops.kblock_rows(lo)))
                        self.tc_err[name] = float(err.max()) if err.numel() else 0.0
            else:
                if video_feat1 is not None:
                    self.video_tc = ops.split_rows(video_feat1, self.ctx_len, self.lp, self.kpad, normalize=True,
                                                   bf16=bf16)
                    self.video_bits = ops.mask_bits(video_mask, self.lp)
                if sub_feat1 is not None:
                    self.sub_tc = ops.split_rows(sub_feat1, self.ctx_len, self.lp, self.kpad, normalize=True,
                                                 bf16=bf16)
                    self.sub_bits = ops.mask_bits(sub_mask, self.lp)
        if (precision != "f32" and merged_spans and video_feat2 is not None and sub_feat2 is not None
                and self.ctx_len <= 256):
            rows = self.n_videos * self.ctx_len
            self.f2cat = (torch.empty(rows, 2 * self.kpad, device=self.device, dtype=torch.int16),
                          torch.empty(rows, 2 * self.kpad, device=self.device, dtype=torch.int16))
            bf16 = precision == "bf16x3"
            ops.split_rows(video_feat2, kpad=self.kpad, bf16=bf16, out=self.f2cat, out_col0=0)
            ops.split_rows(sub_feat2, kpad=self.kpad, bf16=bf16, out=self.f2cat, out_col0=self.kpad)
            # stored K-BLOCKED, (2 Kpad / 32, rows, 32): the 64-byte pieces a k-step of the similarity kernel reads
            # from consecutive clips are contiguous in HBM (measured: span-probability phase 4.96 -> 3.67 ms; as
            # separate pieces 2 * 2 Kpad bytes apart every one of them opened its own DRAM page)
            hi, lo = self.f2cat
            self.f2cat = None
            # XMLB_F2_IMAGE=1: additionally pre-swizzled into the shared-memory image of the tiles, fetched by plain
            # bulk copies instead of tensor boxes (bit-equal; measured no faster: 3.58-3.61 vs 3.64 ms, so off)
            image = self.ctx_len % 8 == 0 and ops.GATHER == "warps" and os.environ.get("XMLB_F2_IMAGE", "0") == "1"
            hi = ops.kblock_rows(hi, swizzle=image)
            self.f2cat = (hi, ops.kblock_rows(lo, swizzle=image))
        self.video_feat2 = video_feat2.contiguous() if video_feat2 is not None else None
        self.sub_feat2 = sub_feat2.contiguous() if sub_feat2 is not None else None
        self.video_mask = video_mask.contiguous() if video_mask is not None else None
        self.sub_mask = sub_mask.contiguous() if sub_mask is not None else None

    @classmethod
    def from_ctx_info(cls, ctx_info, vid_lo=0, precision="f16x3", packed=True, merged_spans=True,
                      rescore_kblocked=False):
        return cls(ctx_info.get("video_feat1"), ctx_info.get("video_feat2"), ctx_info.get("video_mask"),
                   ctx_info.get("sub_feat1"), ctx_info.get("sub_feat2"), ctx_info.get("sub_mask"), vid_lo=vid_lo,
                   precision=precision, packed=packed, merged_spans=merged_spans, rescore_kblocked=rescore_kblocked)

    # ---- incremental growth and persistence (the reference re-encodes and keeps the
    # corpus in Python lists of tensors, inference.py:32-97, and its profiling scenario is "1K videos are added",
    # baselines/profiling/profile_main.py:1-4) -------------------------------------------------------------------
    TENSORS = ("video_feat1n", "sub_feat1n", "video_feat2", "sub_feat2", "video_mask", "sub_mask", "video_bits",
               "sub_bits")
    PAIRS = ("video_tc", "sub_tc", "f2cat", "video_tc_kb", "sub_tc_kb")

    @torch.no_grad()
    def add_videos(self, video_feat1=None, video_feat2=None, video_mask=None, sub_feat1=None, sub_feat2=None,
                   sub_mask=None):
        """Append newly encoded videos (outputs of XML.encode_context, padded to this index's clip width or
        narrower).  They get the next video ids; existing operand rows are not re-packed.  Searching the grown index
        gives the same result as an index built from all videos at once."""
        def widen(t):  # zero-pad the clip axis like cat_tensor does (reference inference.py:71-87)
            if t is None or t.shape[1] == self.ctx_len:
                return t
            assert t.shape[1] < self.ctx_len, "new videos are wider than the index (max_ctx_l)"
            pad = [0, 0] * (t.dim() - 2) + [0, self.ctx_len - t.shape[1]]
            return torch.nn.functional.pad(t, pad)
        args = [widen(t) for t in (video_feat1, video_feat2, video_mask, sub_feat1, sub_feat2, sub_mask)]
        new = CorpusIndex(*args, vid_lo=self.vid_lo + self.n_videos, precision=self.precision,
                          packed=self.packing is not None, merged_spans=self.f2cat is not None,
                          rescore_kblocked=(self.video_tc_kb or self.sub_tc_kb) is not None)
        assert (new.packing is None) == (self.packing is None) and (new.f2cat is None) == (self.f2cat is None)
        for name in self.TENSORS:
            a, b = getattr(self, name), getattr(new, name)
            assert (a is None) == (b is None), "modalities of the new videos differ from the index: " + name
            if a is not None:
                setattr(self, name, torch.cat([a, b]))
        for name in self.PAIRS:
            a, b = getattr(self, name), getattr(new, name)
            if a is not None:
                dim = 1 if a[0].dim() >= 3 else 0  # f2cat is (k-blocks, rows, 32) or (k-blocks, rows, 4, 8)
                setattr(self, name, (torch.cat([a[0], b[0]], dim), torch.cat([a[1], b[1]], dim)))
        if self.packing is not None:
            self.packing.append(new.packing, self.n_videos)
            self.tc_err = {k: max(v, new.tc_err[k]) for k, v in self.tc_err.items()}
        self.n_videos += new.n_videos
        return self

    def save(self, path):
        """One file: JSON header (shapes, dtypes, offsets) + 4 KiB-aligned raw little-endian tensors, written and
        read in bounded chunks so that a 44 GB index never needs a host copy of itself."""
        import json
        items = {}
        for name in self.TENSORS:
            if getattr(self, name) is not None:
                items[name] = getattr(self, name)
        for name in self.PAIRS:
            if getattr(self, name) is not None:
                items[name + ".hi"], items[name + ".lo"] = getattr(self, name)
        if self.packing is not None:
            for k in CorpusPacking.STATE:
                items["packing." + k] = getattr(self.packing, k)
        header = dict(format="xmlb200-index", version=1, n_videos=self.n_videos, ctx_len=self.ctx_len,
                      hidden=self.hidden, vid_lo=self.vid_lo, precision=self.precision,
                      lp=getattr(self, "lp", None), kpad=getattr(self, "kpad", None),
                      tc_err=getattr(self, "tc_err", None),
                      max_len=self.packing.max_len if self.packing is not None else None, tensors={})
        off = 0
        for k, t in items.items():
            header["tensors"][k] = dict(dtype=str(t.dtype).replace("torch.", ""), shape=list(t.shape), offset=off)
            off += (t.numel() * t.element_size() + 4095) // 4096 * 4096
        blob = json.dumps(header).encode()
        head_len = (len(blob) + 16 + 4095) // 4096 * 4096
        with open(path, "wb") as fh:
            fh.write(b"XMLB200I" + len(blob).to_bytes(8, "little") + blob)
            fh.write(b"\0" * (head_len - 16 - len(blob)))
            for k, t in items.items():
                flat = t.contiguous().view(-1)
                step = max(1, (256 << 20) // max(1, t.element_size()))
                for lo in range(0, flat.numel(), step):
                    fh.write(flat[lo:lo + step].cpu().numpy().tobytes())
                pad = -(flat.numel() * t.element_size()) % 4096
                fh.write(b"\0" * pad)
        return path

    @classmethod
    def load(cls, path, device="cuda"):
        import json
        with open(path, "rb") as fh:
            magic = fh.read(8)
            if magic != b"XMLB200I":
                raise ValueError("%s is not an xmlb200 index file" % path)
            n = int.from_bytes(fh.read(8), "little")
            header = json.loads(fh.read(n))
        if header.get("version") != 1:
            raise ValueError("unsupported index file version %r" % header.get("version"))
        base = (n + 16 + 4095) // 4096 * 4096
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        tensors = {}
        for k, d in header["tensors"].items():
            dtype = getattr(torch, d["dtype"])
            t = torch.empty(d["shape"], dtype=dtype, device=device)
            flat = t.view(-1).view(torch.uint8)
            nbytes, step = flat.numel(), 256 << 20
            for lo in range(0, nbytes, step):
                hi = min(nbytes, lo + step)
                chunk = torch.from_numpy(np.array(mm[base + d["offset"] + lo:base + d["offset"] + hi]))
                flat[lo:hi].copy_(chunk)
            tensors[k] = t
        self = cls.__new__(cls)
        self.n_videos, self.ctx_len, self.hidden = header["n_videos"], header["ctx_len"], header["hidden"]
        self.vid_lo, self.precision, self.device = header["vid_lo"], header["precision"], torch.device(device)
        if header["lp"] is not None:
            self.lp, self.kpad = header["lp"], header["kpad"]
        for name in cls.TENSORS:
            setattr(self, name, tensors.get(name))
        for name in cls.PAIRS:
            setattr(self, name, (tensors[name + ".hi"], tensors[name + ".lo"]) if name + ".hi" in tensors else None)
        self.packing = None
        if "packing.order" in tensors:
            self.packing = CorpusPacking.from_state({k: tensors["packing." + k] for k in CorpusPacking.STATE},
                                                    header["max_len"])
            self.tc_err = header["tc_err"]
        return self

    def nbytes(self):
        tensors = [self.video_feat1n, self.sub_feat1n, self.video_feat2, self.sub_feat2, self.video_mask,
                   self.sub_mask, self.video_bits, self.sub_bits]
        for pair in (self.video_tc, self.sub_tc, self.f2cat, self.video_tc_kb, self.sub_tc_kb):
            tensors.extend(pair or ())
        return sum(t.numel() * t.element_size() for t in tensors if t is not None)


class _Prefilter:
    """Pass 1 of the two-pass video retrieval for one block of queries: hi/lo splits of the normalised pooled
    queries (+ the hi-only error bounds) and the approximate (Nq, Nv) scores, filled for row ranges [lo, hi)."""

    def __init__(self, searcher, n):
        ix, m = searcher.index, searcher.model
        self.s, self.bf16 = searcher, ix.precision == "bf16x3"
        dev = ix.device
        self.corpus = [c for c, on in ((ix.video_tc, m.use_video), (ix.sub_tc, m.use_sub)) if on]
        self.which = [i for i, on in enumerate((m.use_video, m.use_sub)) if on]
        self.split = [(torch.empty(n, ix.kpad, device=dev, dtype=torch.int16),
                       torch.empty(n, ix.kpad, device=dev, dtype=torch.int16),
                       torch.empty(n, device=dev, dtype=torch.float32)) for _ in self.corpus]
        self.approx = torch.empty(n, ix.n_videos, device=dev, dtype=torch.float32)

    def run(self, lo, hi, video_query, sub_query):
        ix = self.s.index
        qs = [(video_query, sub_query)[i] for i in self.which]
        with self.s._phase("vr_scores"):
            for q, (h, l, e) in zip(qs, self.split):
                ops.split_rows(q, kpad=ix.kpad, normalize=True, bf16=self.bf16, hi_err=True, out=(h[lo:hi], l[lo:hi]),
                               err_out=e[lo:hi])
            parts = [(h[lo:hi], l[lo:hi]) for h, l, _ in self.split]
            ops.vr_scores_tc_packed(parts[0], self.corpus[0], ix.packing, ix.n_videos,
                                    q_b=parts[1] if len(parts) == 2 else None,
                                    c_b=self.corpus[1] if len(parts) == 2 else None, bf16=self.bf16, ordinal=True,
                                    hi_only=True, out=self.approx[lo:hi])


class SearchResult:
    """Device tensors for one batch of queries."""
    __slots__ = ("top_video_idx", "top_video_score", "span_flat_idx", "span_score", "svmr_flat_idx", "svmr_score")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


class VCMRSearcher:
    # accumulation-order slack of the two-pass error bound: the fp32 TMEM accumulator truncates, <= 1 ulp(1) per
    # update, 3 * K/16 updates for the exact kernel + K/16 for the filter (K = 768: 192 * 2^-23 = 2.3e-5)
    TWO_PASS_SLACK = 4e-5

    def __init__(self, model, index, q2c_alpha=20.0, min_pred_l=2, max_pred_l=16, max_n_videos=100,
                 max_before_nms=200, query_chunk=16384, two_pass=None, max_candidates=None, encode_chunk=2048):
        """two_pass: find the top videos with a hi-only (1 MMA per product) filter pass over the corpus followed by
        exact re-scoring of the few survivors; same result as the one-pass kernel at a third of the tensor-core
        work.  None = automatic (packed f16x3 index with at least 4 * max_n_videos videos)."""
        self.model, self.index = model, index
        auto = index.packing is not None and index.precision == "f16x3" and index.n_videos >= 4 * max_n_videos
        assert not two_pass or (index.packing is not None and index.precision != "f32"), \
            "the two-pass search needs the packed tensor-core index"
        self.two_pass = auto if two_pass is None else bool(two_pass)
        self.max_candidates = int(max_candidates or max(256, 2 * max_n_videos))
        self.q2c_alpha = float(q2c_alpha)
        self.min_pred_l, self.max_pred_l = int(min_pred_l), int(max_pred_l)
        self.max_n_videos, self.max_before_nms = int(max_n_videos), int(max_before_nms)
        self.query_chunk = int(query_chunk)    # queries searched together (one pass over the corpus per block)
        self.encode_chunk = int(encode_chunk)  # queries uploaded / encoded per piece inside a block
        self.timer = None  # set to a PhaseTimer to time the phases
        self._external = None  # (video positions, exp-scores) of the current block when external lists are given
        self._prefilter = None  # pass-1 state of the current block when it was filled piece by piece
        # tests / bench parity leg: set to a dict to receive the two-pass candidate tables of the last block
        # ("cand": ops.Candidates after exact re-scoring)
        self.debug = None

    def _phase(self, name):
        return self.timer.phase(name) if self.timer is not None else contextlib.nullcontext()

    # ---- phases (also used one by one by the sharded searcher) -------------------------------------
    def score_ids(self):
        """Column -> (global) video id table of video_scores(ordinal=True), or None when columns are video ids."""
        ix = self.index
        if ix.packing is None:
            return None
        return ix.packing.order_full + ix.vid_lo if ix.vid_lo else ix.packing.order_full

    def video_scores(self, video_query, sub_query, ordinal=False):
        """(Nq, Nv) video-level scores.  ordinal=True leaves the columns in the kernel's native order (see
        score_ids) instead of re-ordering them by video id -- what the top-k that follows wants."""
        ix, m = self.index, self.model
        if ix.precision != "f32":
            bf16 = ix.precision == "bf16x3"
            ops_q = [ops.split_rows(q, kpad=ix.kpad, normalize=True, bf16=bf16) if used else None
                     for q, used in ((video_query, m.use_video), (sub_query, m.use_sub))]
            streams = [(q, c, b) for q, c, b in ((ops_q[0], ix.video_tc, ix.video_bits),
                                                 (ops_q[1], ix.sub_tc, ix.sub_bits)) if q is not None]
            a, b = streams[0], (streams[1] if len(streams) == 2 else (None, None, None))
            if ix.packing is not None:
                return ops.vr_scores_tc_packed(a[0], a[1], ix.packing, ix.n_videos, q_b=b[0], c_b=b[1], bf16=bf16,
                                               ordinal=ordinal)
            return ops.vr_scores_tc(a[0], a[1], a[2], ix.n_videos, ix.lp, q_b=b[0], c_b=b[1], bits_b=b[2], bf16=bf16)
        return ops.vr_scores_f32(
            ops.l2norm_rows(video_query) if m.use_video else None, ops.l2norm_rows(sub_query) if m.use_sub else None,
            ix.video_feat1n if m.use_video else None, ix.sub_feat1n if m.use_sub else None,
            ix.video_mask if m.use_video else None, ix.sub_mask if m.use_sub else None)

    def _global_kth(self, approx, k):
        """k-th largest approximate score of each query over the WHOLE corpus when this index is one shard of it
        (ShardedSearcher overrides this); None = the index is the whole corpus."""
        return None

    def top_videos(self, video_query, sub_query, k, k_global=None):
        """Exact top-k videos of this index for every query -> (global video ids int32 (Nq, k), exp(alpha * score)),
        ranked by (score desc, id asc).  Reference: inference.py:317,347-348 on top of model_xml.py:446-452,572-574.
        k_global (sharded search): the k of the final, corpus-wide selection -- the candidate filter then keeps only
        this shard's part of the global candidate set, and entries that did not make it come back as (INT_MAX, 0)."""
        ix, m = self.index, self.model
        if not (self.two_pass and ix.n_videos >= k):
            with self._phase("vr_scores"):
                q2c = self.video_scores(video_query, sub_query, ordinal=True)
            with self._phase("topk_videos"):
                ids = self.score_ids()
                idx, val = ops.topk_rows(q2c, k, alpha=self.q2c_alpha, apply_exp=True, ids=ids)
                if ids is None and ix.vid_lo:
                    idx = idx + ix.vid_lo
            return idx, val
        bf16 = ix.precision == "bf16x3"
        pk = ix.packing
        used = [(q, c, ix.tc_err[name]) for q, c, name, on in ((video_query, ix.video_tc, "video", m.use_video),
                                                               (sub_query, ix.sub_tc, "sub", m.use_sub)) if on]
        pre, self._prefilter = self._prefilter, None
        if pre is None:  # pass 1 over the whole block now (otherwise it ran piece by piece behind the uploads)
            pre = _Prefilter(self, len(video_query))
            pre.run(0, len(video_query), video_query, sub_query)
        split, approx = pre.split, pre.approx
        qa, ca = split[0][:2], used[0][1]
        qb, cb = (split[1][:2], used[1][1]) if len(used) == 2 else (None, None)
        with self._phase("vr_select"):
            # |approx - exact| <= mean over modalities of (||q - q_hi|| * ||c|| + ||q_hi|| * ||c - c_hi||) + slack
            scale = 1.001 / len(used)
            const = scale * sum(e for _, _, e in used) + self.TWO_PASS_SLACK
            cand = ops.select_candidates(approx, k, split[0][2], split[1][2] if len(used) == 2 else None, scale,
                                         const, self.max_candidates, ids=self.score_ids(),
                                         row_kth=self._global_kth(approx, k_global or k))
        with self._phase("vr_rescore"):
            kb = [c for c, on in ((ix.video_tc_kb, m.use_video), (ix.sub_tc_kb, m.use_sub)) if on]
            ca_r, cb_r = (kb[0], kb[1] if len(used) == 2 else None) if all(c is not None for c in kb) else (ca, cb)
            ops.vr_rescore_tc(used[0][0], ca_r, pk, cand, ix.kpad, q_fp32_b=used[1][0] if len(used) == 2 else None,
                              c_b=cb_r, bf16=bf16, q_split_a=qa, q_split_b=qb)
        if self.debug is not None:
            self.debug["cand"] = cand
        with self._phase("topk_videos"):
            idx, val = ops.topk_rows(cand.val, k, alpha=self.q2c_alpha, apply_exp=True, ids=cand.ids)
        with self._phase("vr_fallback"):  # rows whose candidate list overflowed (normally none: both launches idle)
            ops.vr_scores_tc_packed(qa, ca, pk, ix.n_videos, q_b=qb, c_b=cb, bf16=bf16, ordinal=True, out=approx,
                                    m_tiles=(cand.flagged_groups, cand.n_flagged))
            ops.topk_rows(approx, k, alpha=self.q2c_alpha, apply_exp=True, ids=self.score_ids(),
                          row_flags=cand.row_flags, out=(idx, val))
        return idx, val

    def use_span_tc(self):
        m = self.model
        return self.index.f2cat is not None and m.config.merge_two_stream and m.use_video and m.use_sub

    def span_lists(self, top_idx, slot_valid=None):
        """Inverted (video -> queries) lists of the selected pairs; the chunk size is the N tile of the tensor-core
        similarity kernel (larger when many queries share a video), 32 for the SIMT kernel."""
        ix = self.index
        chunk = 32
        if self.use_span_tc():
            avg = top_idx.numel() / max(1, ix.n_videos)
            chunk = 128 if avg >= 96 else 64 if avg >= 40 else 32
        lists = ops.build_pair_lists(top_idx, ix.n_videos, vid_lo=ix.vid_lo, slot_valid=slot_valid, chunk=chunk)
        # selected videos of a single-GPU search all live in this index: every output row will be written
        lists.complete = slot_valid is None and ix.vid_lo == 0 and not hasattr(self, "plan")
        return lists

    def span_probs(self, video_query, sub_query, lists):
        """softmax-normalised start/end distributions for the listed (query, video) pairs -> (rows, L) x 2."""
        ix = self.index
        if self.use_span_tc():
            m = self.model
            qv = ops.linear(video_query, m.video_query_linear.weight, m.video_query_linear.bias)
            qs = ops.linear(sub_query, m.sub_query_linear.weight, m.sub_query_linear.bias)
            if ix.kpad != ix.hidden:
                qv = torch.nn.functional.pad(qv, (0, ix.kpad - ix.hidden))
                qs = torch.nn.functional.pad(qs, (0, ix.kpad - ix.hidden))
            ksize = m.merged_st_predictor.weight.numel()
            rows = getattr(ix, "_span_clip_rows", None)  # (mask it was made from, ksize, rows)
            if rows is None or rows[0] is not ix.video_mask or rows[1] != ksize:
                rows = ix._span_clip_rows = (ix.video_mask, ksize, ops.span_clip_rows(ix.video_mask, ksize))
            return ops.span_probs_tc(ix.f2cat, torch.cat([qv, qs], dim=1), lists, ix.video_mask,
                                     m.merged_st_predictor.weight, m.merged_ed_predictor.weight, ix.ctx_len,
                                     softmax=True, bf16=ix.precision == "bf16x3", clip_rows=rows[2])
        args = self.model.span_streams(video_query, sub_query, ix.video_feat2, ix.sub_feat2, ix.video_mask,
                                       ix.sub_mask)
        return ops.span_logits(softmax=True, lists=lists, **args)

    # ---- search: encode (pipelined in small chunks) -> one block of queries against the whole index ---------
    def _my_slice(self, n):
        """Part [lo, hi) of a block of n queries that THIS process encodes (all of it on one GPU)."""
        return 0, n

    def _n_ranks(self):
        """Processes that search the corpus together (ShardedSearcher: the shards)."""
        return 1

    def _gather_encoded(self, video_query, sub_query, n):
        """Pooled query vectors of the whole block from the locally encoded slice (identity on one GPU)."""
        return video_query, sub_query

    # Encode only the valid query tokens (XML.encode_query_packed): same pooled vectors, ~40 % fewer rows through the
    # encoder for TVR's query lengths.  False = the padded reference layout (XML.encode_query).
    packed_queries = os.environ.get("XMLB_PACKED_QUERIES", "1") != "0"
    # ... for blocks of at least this many queries on this GPU: the layout tables are built on the host (one D2H of
    # the lengths when the masks live on the device), which costs more than it saves on small slices (measured at
    # 8 GPUs, 1,250 queries per rank: 2.1 ms packed vs 1.6 ms padded)
    packed_min_queries = 1024

    min_piece = 256  # (tests lower it to exercise the multi-piece paths on tiny blocks)

    def _piece_bounds(self, n, host):
        """[lo, hi) of the pieces a slice of n queries is uploaded / encoded in.  From host buffers the first piece is
        small (its upload is the only one nothing overlaps) and a short slice (one rank's share of a sharded search)
        is still cut into a few pieces so that uploads and encoding overlap."""
        step = self.encode_chunk
        if not host:  # device-resident queries: nothing to overlap, few large pieces = few launches
            step = max(step, 8192)
            return [(lo, min(n, lo + step)) for lo in range(0, n, step)]
        step = min(step, max(self.min_piece, (n + 3) // 4))
        # geometric ramp: a piece's upload (~1.7 us per query over PCIe) hides behind the encoding + filter pass of
        # the piece before it (~4.5 us per query) as long as it is at most ~2.5x as large; only the first upload is
        # exposed, so it is small (with a first piece of step / 4 the GPU idled ~2 ms per 10 K-query block waiting
        # for the first two uploads)
        size = 256 if step >= 1024 else max(min(128, step), step // 4)
        cuts, lo = [], 0
        while lo < n:
            hi = min(n, lo + size)
            if n - hi < size // 2:  # no tiny last piece
                hi = n
            cuts.append((lo, hi))
            lo, size = hi, min(step, 2 * size)
        return cuts

    def _encode_pieces(self, pieces, lens_cpu=None, on_piece=None, tables_first=False, width=None, bounds=None,
                       lens_dev=None):
        """pieces: iterable of (query_feat, query_mask) device tensors -> pooled (video_query, sub_query).
        lens_cpu: host int tensor, valid tokens of every query in piece order (None: padded encoding); `width` is
        the padded token count of the pieces, lens_dev the same lengths on the device when they already are there.
        on_piece(lo, hi, video_query, sub_query): called after each piece is encoded (pipelined filter pass).
        The packed-layout tables of ALL pieces are made (on the device, from the lengths) before `pieces` is
        touched: when it streams the features from the host, a small H2D copy issued after its bulk uploads would
        wait behind all of them."""
        hid = self.model.config.hidden_size
        vq, sq = [], []
        with self._phase("encode_query"):
            tables = None
            if lens_cpu is not None:
                tables = self.model.packed_query_tables_device(lens_cpu.numpy(), width, self.index.device, bounds,
                                                               lens_dev=lens_dev)
            off = 0
            for i, (qf, qm) in enumerate(pieces):
                if len(qf) == 0:  # (a rank's empty share of a piece of a sharded search)
                    a = b = torch.zeros(0, hid, device=self.index.device)
                elif tables is not None:
                    a, b = self.model.encode_query_packed(qf, tables=tables[i])
                else:
                    a, b = self.model.encode_query(qf, qm)
                vq.append(a), sq.append(b)
                if on_piece is not None:
                    on_piece(off, off + len(a), a, b)
                off += len(a)
            if not vq:
                z = torch.zeros(0, hid, device=self.index.device)
                return z, z
            return (vq[0], sq[0]) if len(vq) == 1 else (torch.cat(vq), torch.cat(sq))

    # Host-buffer searches run the filter pass of the video retrieval per uploaded piece (5 launches over the corpus
    # instead of 1: ~20 GB more L2->HBM traffic, hidden under the tensor-core work) so that only the first piece's
    # upload is exposed instead of the whole 0.9 GB.
    pipelined_filter = True

    def _device_pieces(self, query_feat, query_mask, bounds):
        for lo, hi in bounds:
            yield query_feat[lo:hi], query_mask[lo:hi]

    def _host_pieces(self, query_feat_cpu, query_mask_cpu, bounds):
        """Uploads the pieces on a side stream, so the copy of piece i+1 overlaps the encoding of piece i."""
        dev = self.index.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        copy = self._copy_stream
        copy.wait_stream(main)
        staged = []
        for lo, hi in bounds:
            with torch.cuda.stream(copy):
                qf = query_feat_cpu[lo:hi].to(dev, non_blocking=True)
                qm = query_mask_cpu[lo:hi].to(dev, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
            staged.append((qf, qm, ready))
        for qf, qm, ready in staged:
            main.wait_event(ready)
            qf.record_stream(main), qm.record_stream(main)
            yield qf, qm

    @torch.no_grad()  # inference only: the kernels must never record an autograd graph here
    def search(self, query_feat, query_mask, gt_video_idx=None, tasks=("VCMR", "VR"), host=False,
               external_topk=None):
        """query_feat (Nq, Lq, Dq), query_mask (Nq, Lq) on the index's device (host=True: pinned host tensors, the
        upload is part of the call).  gt_video_idx (Nq,) int: corpus position of each query's ground-truth video,
        needed for SVMR.  external_topk = (video positions int (Nq, max_n_videos), raw scores (Nq, max_n_videos)):
        the video lists of another retrieval system replace this model's own video retrieval (reference
        inference.py:349-355, --external_inference_vr_res_path); their scores enter as exp(q2c_alpha * score).  Queries are processed in blocks of `query_chunk`; inside a block the raw features are
        uploaded / encoded in pieces of `encode_chunk`, then the whole block is searched at once (each pass over
        the corpus is shared by all queries of the block)."""
        outs = []
        dev = self.index.device
        for b_lo in range(0, len(query_feat), self.query_chunk):
            n = min(self.query_chunk, len(query_feat) - b_lo)
            lo, hi = self._my_slice(n)
            qf, qm = query_feat[b_lo + lo:b_lo + hi], query_mask[b_lo + lo:b_lo + hi]
            lens = lens_dev = None
            if (self.packed_queries and hi - lo >= self.packed_min_queries
                    and qm.shape[1] <= self.model.PACKED_MAX_LEN):
                # valid tokens per query (masks are prefix masks); one small D2H when the masks live on the device
                lens = (qm != 0).sum(1)
                if lens.is_cuda:
                    lens_dev = lens
                lens = lens.to(torch.int64).cpu()
            on_piece = None
            self._prefilter = None
            pipelined = (host and self.pipelined_filter and external_topk is None and self.two_pass
                         and ("VR" in tasks or "VCMR" in tasks))
            if pipelined and self._n_ranks() > 1:  # (the same decision on every rank)
                # sharded search from host buffers: pieces of the block are encoded by all ranks together, shared
                # and filtered while the next piece uploads (ShardedSearcher._encode_block_pipelined)
                video_query, sub_query = self._encode_block_pipelined(query_feat[b_lo:b_lo + n],
                                                                      query_mask[b_lo:b_lo + n], n)
            else:
                if pipelined and self.index.n_videos >= self.max_n_videos:
                    # host buffers: the filter pass of piece i runs while piece i+1 is still being uploaded
                    self._prefilter = _Prefilter(self, n)
                    on_piece = self._prefilter.run
                bounds = self._piece_bounds(hi - lo, host)
                video_query, sub_query = self._encode_pieces(self._host_pieces(qf, qm, bounds) if host
                                                             else self._device_pieces(qf, qm, bounds), lens, on_piece,
                                                             tables_first=host, width=qm.shape[1], bounds=bounds,
                                                             lens_dev=lens_dev)
                video_query, sub_query = self._gather_encoded(video_query, sub_query, n)
            gt = None if gt_video_idx is None else gt_video_idx[b_lo:b_lo + n].to(dev, non_blocking=True)
            self._external = None
            if external_topk is not None:
                ext_idx = external_topk[0][b_lo:b_lo + n].to(dev).to(torch.int32).contiguous()
                ext_val = torch.exp(self.q2c_alpha * external_topk[1][b_lo:b_lo + n].to(dev).float()).contiguous()
                assert ext_idx.shape == (n, self.max_n_videos) == ext_val.shape, \
                    "external video lists must hold max_n_videos entries per query"
                self._external = (ext_idx, ext_val)
            outs.append(self._search_encoded(video_query, sub_query, gt, tasks))
            self._external = None
        if len(outs) == 1:
            return outs[0]
        res = SearchResult()
        for s in SearchResult.__slots__:
            if getattr(outs[0], s) is not None:
                setattr(res, s, torch.cat([getattr(o, s) for o in outs]))
        return res

    def _search_encoded(self, video_query, sub_query, gt_video_idx, tasks):
        ix = self.index
        res = SearchResult()
        nq = len(video_query)
        if ("VR" in tasks or "VCMR" in tasks) and self._external is not None:
            res.top_video_idx, res.top_video_score = self._external
        elif "VR" in tasks or "VCMR" in tasks:
            res.top_video_idx, res.top_video_score = self.top_videos(video_query, sub_query, self.max_n_videos)
        if "VCMR" in tasks:
            with self._phase("pair_lists"):
                lists = self.span_lists(res.top_video_idx)
            with self._phase("span_probs"):
                st, ed = self.span_probs(video_query, sub_query, lists)
            st = st.view(nq, self.max_n_videos, ix.ctx_len)
            ed = ed.view(nq, self.max_n_videos, ix.ctx_len)
            with self._phase("span_topk"):
                res.span_flat_idx, res.span_score = ops.span_topk(st, ed, res.top_video_score, self.min_pred_l,
                                                                  self.max_pred_l, self.max_before_nms)
        if "SVMR" in tasks:
            assert gt_video_idx is not None, "SVMR needs the ground-truth video of every query"
            lists = self.span_lists(gt_video_idx.view(nq, 1))
            st, ed = self.span_probs(video_query, sub_query, lists)
            res.svmr_flat_idx, res.svmr_score = ops.span_topk(
                st.view(nq, 1, ix.ctx_len), ed.view(nq, 1, ix.ctx_len), None, self.min_pred_l, self.max_pred_l,
                self.max_before_nms, tie_desc=True)
        return res

    # ---- host-buffer entry point (the e2e path bench.py times) --------------------------------------
    @torch.no_grad()
    def search_host(self, query_feat_cpu, query_mask_cpu, gt_video_idx_cpu=None, tasks=("VCMR", "VR"),
                    external_topk=None, result_rank=None):
        """Pinned host buffers in, numpy arrays out; H2D / D2H copies are part of the call.  The arrays are views of
        pinned staging buffers that the NEXT search_host call of this searcher overwrites.  result_rank (sharded
        search): only that rank copies the (identical) result to its host, the others return None."""
        res = self.search(query_feat_cpu, query_mask_cpu, gt_video_idx_cpu, tasks, host=True,
                          external_topk=external_topk)
        if result_rank is not None and getattr(self, "plan", None) is not None and self.plan.rank != result_rank:
            torch.cuda.current_stream(self.index.device).synchronize()
            return None
        stage = self.__dict__.setdefault("_host_out", {})
        out = {}
        for s in SearchResult.__slots__:
            t = getattr(res, s)
            if t is None:
                continue
            buf = stage.get(s)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = stage[s] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            out[s] = buf
        torch.cuda.current_stream(self.index.device).synchronize()
        return {s: b.numpy() for s, b in out.items()}
