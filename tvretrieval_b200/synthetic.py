"""Deterministic synthetic TVR-shaped inputs.

`SyntheticEvalDataset` implements the duck-typed dataset protocol that the reference drivers
(`compute_context_info` / `compute_query2ctx_info`, reference inference.py:32,252) expect from
`StartEndEvalDataset` (reference start_end_dataset.py:171-343): `set_data_mode`,
`load_gt_vid_name_for_query`, `__len__`, `__getitem__ -> {"meta", "model_inputs"}`, and the
attributes `video2idx`, `max_ctx_len`, `query_data`.  The h5 readers are replaced by seeded
tensors of the same layout (unit-norm rows; `resnet_i3d` = unit-norm 2048 block || unit-norm
1024 block, reference utils/video_feature/normalize_and_concat.py:24-29).
"""
import torch


def _unit_rows(x, eps=1e-5):
    """l2_normalize_np_array semantics (reference utils/basic_utils.py:82-84): eps added to the norm."""
    return x / (x.norm(dim=-1, keepdim=True) + eps)


def make_video_feat(gen, n_clips, video_dim, split=None, device="cpu"):
    x = torch.randn(n_clips, video_dim, generator=gen, device=device)
    if split is None or split <= 0 or split >= video_dim:
        return _unit_rows(x)
    return torch.cat([_unit_rows(x[:, :split]), _unit_rows(x[:, split:])], dim=1)


def make_text_feat(gen, n_rows, dim, device="cpu"):
    return _unit_rows(torch.randn(n_rows, dim, generator=gen, device=device))


class SyntheticEvalDataset(torch.utils.data.Dataset):
    def __init__(self, n_videos, n_queries, max_ctx_l, max_desc_l, video_dim, sub_dim, query_dim,
                 ctx_mode="video_sub", seed=1234, video_split=None, min_ctx_l=None, clip_length=1.5):
        self.ctx_mode = ctx_mode
        self.use_video = "video" in ctx_mode
        self.use_sub = "sub" in ctx_mode
        self.max_ctx_len = max_ctx_l
        self.max_desc_len = max_desc_l
        self.clip_length = clip_length
        self.data_mode = "context"
        self.load_gt_video = False
        gen = torch.Generator().manual_seed(seed)
        lo = max(1, max_ctx_l // 8) if min_ctx_l is None else min_ctx_l
        ctx_lens = torch.randint(lo, max_ctx_l + 1, (n_videos,), generator=gen)
        ctx_lens[n_videos // 2] = max_ctx_l  # padded corpus width must equal max_ctx_l (hazard B-5)
        q_lens = torch.randint(min(5, max_desc_l), max_desc_l + 1, (n_queries,), generator=gen)
        self.ctx_lens = ctx_lens.tolist()
        self.video_feats, self.sub_feats = [], []
        for n in self.ctx_lens:
            self.video_feats.append(make_video_feat(gen, n, video_dim, video_split) if self.use_video else None)
            self.sub_feats.append(make_text_feat(gen, n, sub_dim) if self.use_sub else None)
        self.query_feats = [make_text_feat(gen, int(n), query_dim) for n in q_lens]
        # vid_name -> dataset video idx: deliberately NOT the meta position (hazard B-10)
        perm = torch.randperm(n_videos, generator=gen).tolist()
        self.video_data = [{"vid_name": "vid_%05d" % i, "duration": float(self.ctx_lens[i] * clip_length)}
                           for i in range(n_videos)]
        self.video2idx = {"vid_%05d" % i: 1000 + perm[i] for i in range(n_videos)}
        gt = torch.randint(0, n_videos, (n_queries,), generator=gen).tolist()
        self.query_data = []
        for i in range(n_queries):
            dur = self.video_data[gt[i]]["duration"]
            self.query_data.append({"desc_id": 90000 + i, "desc": "synthetic query %d" % i,
                                    "vid_name": self.video_data[gt[i]]["vid_name"], "duration": dur,
                                    "ts": [0.0, min(dur, 3 * clip_length)], "type": "v"})

    def set_data_mode(self, data_mode):
        assert data_mode in ("context", "query")
        self.data_mode = data_mode

    def load_gt_vid_name_for_query(self, load_gt_video):
        self.load_gt_video = load_gt_video

    def __len__(self):
        return len(self.video_data) if self.data_mode == "context" else len(self.query_data)

    def __getitem__(self, index):
        if self.data_mode == "context":
            raw = self.video_data[index]
            filler = torch.zeros((2, 2))  # reference start_end_dataset.py:314,323,333
            return dict(meta=dict(vid_name=raw["vid_name"], duration=raw["duration"]),
                        model_inputs=dict(
                            video_feat=self.video_feats[index] if self.use_video else filler,
                            sub_feat=self.sub_feats[index] if self.use_sub else filler,
                            tef_feat=filler))
        raw = self.query_data[index]
        return dict(meta=dict(desc_id=raw["desc_id"], desc=raw["desc"],
                              vid_name=raw["vid_name"] if self.load_gt_video else None),
                    model_inputs=dict(query_feat=self.query_feats[index]))


# ---------------------------------------------------------------------------------------------------
# Large synthetic corpora generated batch by batch ON THE DEVICE (bench.py): the raw features of the
# 21.8K-video shape are 34 GB and are never materialised at once.
# ---------------------------------------------------------------------------------------------------
def corpus_lengths(n_videos, max_ctx_l, seed=1234):
    gen = torch.Generator().manual_seed(seed)
    lens = torch.randint(max(1, max_ctx_l // 8), max_ctx_l + 1, (n_videos,), generator=gen)
    lens[n_videos // 2] = max_ctx_l
    return lens


def corpus_batch(lens, batch_idx, bsz, video_dim, sub_dim, device, seed=1234, video_split=2048):
    """Raw features of global context batch `batch_idx` (videos [batch_idx*bsz, ...)), padded to the batch max
    like start_end_collate does.  Deterministic in (seed, batch_idx) regardless of which rank generates it."""
    lo = batch_idx * bsz
    bl = lens[lo:lo + bsz].to(device)
    n, width = len(bl), int(bl.max())
    gen = torch.Generator(device=device).manual_seed(seed * 1000003 + batch_idx)
    mask = (torch.arange(width, device=device)[None] < bl[:, None]).float()
    video = torch.randn(n, width, video_dim, generator=gen, device=device)
    if video_split and 0 < video_split < video_dim:
        video = torch.cat([_unit_rows(video[..., :video_split]), _unit_rows(video[..., video_split:])], dim=-1)
    else:
        video = _unit_rows(video)
    sub = _unit_rows(torch.randn(n, width, sub_dim, generator=gen, device=device))
    return video * mask.unsqueeze(2), sub * mask.unsqueeze(2), mask


def synthetic_queries(n_queries, max_desc_l, query_dim, seed=4321):
    """(Nq, max_desc_l, Dq) unit-norm rows, lengths ~ randint(5, max_desc_l + 1), zero padded; CPU tensors."""
    gen = torch.Generator().manual_seed(seed)
    lens = torch.randint(min(5, max_desc_l), max_desc_l + 1, (n_queries,), generator=gen)
    mask = (torch.arange(max_desc_l)[None] < lens[:, None]).float()
    feat = _unit_rows(torch.randn(n_queries, max_desc_l, query_dim, generator=gen)) * mask.unsqueeze(2)
    return feat, mask
