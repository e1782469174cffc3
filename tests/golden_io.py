"""Load the committed golden fixtures (tests/golden/*.npz, produced by make_golden.py from the real
reference) as torch tensors + rebuilt context/query batches."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = ("video_only_svmr", "video_sub_vcmr", "video_sub_nocross")


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.cfg = json.loads(str(z["cfg_json"]))
        self.case = json.loads(str(z["case_json"]))
        self.weights = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
        self.ctx_lens = z["ctx_lens"].tolist()
        self.n_videos = len(self.ctx_lens)
        self.use_video = "video" in self.cfg["ctx_mode"]
        self.use_sub = "sub" in self.cfg["ctx_mode"]
        self.video_feats = [torch.from_numpy(z["video_feat/%d" % i]) for i in range(self.n_videos)] \
            if self.use_video else None
        self.sub_feats = [torch.from_numpy(z["sub_feat/%d" % i]) for i in range(self.n_videos)] \
            if self.use_sub else None
        self.n_queries = len(z["query_gt_meta_idx"])
        self.query_feats = [torch.from_numpy(z["query_feat/%d" % i]) for i in range(self.n_queries)]
        self.query_feat = torch.from_numpy(z["query_feat_padded"])
        self.query_mask = torch.from_numpy(z["query_mask"])
        self.video2idx = z["video2idx"]
        self.query_gt_meta_idx = z["query_gt_meta_idx"]

    def t(self, key):
        return torch.from_numpy(self.z[key]) if key in self.z.files else None

    def ctx(self):
        return {k: self.t("ctx/" + k) for k in
                ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2", "sub_mask")}

    @staticmethod
    def pad(seqs):
        width = max(len(s) for s in seqs)
        feat = torch.zeros(len(seqs), width, seqs[0].shape[1])
        mask = torch.zeros(len(seqs), width)
        for i, s in enumerate(seqs):
            feat[i, :len(s)] = s
            mask[i, :len(s)] = 1
        return feat, mask

    def context_batches(self):
        """Same batching as the reference DataLoader(shuffle=False, batch_size=eval_context_bsz) +
        start_end_collate (pad to the batch max)."""
        bsz = self.case["ctx_bsz"]
        for lo in range(0, self.n_videos, bsz):
            b = {}
            if self.use_video:
                b["video_feat"], b["video_mask"] = self.pad(self.video_feats[lo:lo + bsz])
            if self.use_sub:
                b["sub_feat"], b["sub_mask"] = self.pad(self.sub_feats[lo:lo + bsz])
            yield b

    def query_batches(self):
        bsz = self.case["q_bsz"]
        for lo in range(0, self.n_queries, bsz):
            yield lo, self.pad(self.query_feats[lo:lo + bsz])


TRAIN_VARIANTS = ("plain", "hard")


class TrainCase:
    """One golden training step (tests/golden/train_step.npz): the batch is rebuilt from the case's data --
    item i = (query i, its ground-truth video) -- and the reference's loss, reported floats and parameter
    gradients are loaded."""

    def __init__(self, name, variant):
        import json as _json
        z = np.load(os.path.join(GOLDEN_DIR, "train_step.npz"))
        g = GoldenCase(name)
        key = "%s/%s/" % (name, variant)
        self.seed = int(z["seed"])
        self.cfg = dict(g.cfg)
        self.cfg.update(_json.loads(str(z["variants_json"]))[variant])
        self.weights = g.weights
        gt = z[key + "gt"].tolist()
        self.inputs = dict(query_feat=g.query_feat, query_mask=g.query_mask, tef_feat=None, tef_mask=None,
                           st_ed_indices=torch.from_numpy(z[key + "st_ed_indices"]))
        for mod, used, feats in (("video", g.use_video, g.video_feats), ("sub", g.use_sub, g.sub_feats)):
            if used:
                self.inputs[mod + "_feat"], self.inputs[mod + "_mask"] = GoldenCase.pad([feats[v] for v in gt])
            else:
                self.inputs[mod + "_feat"] = self.inputs[mod + "_mask"] = None
        self.loss = float(z[key + "loss"])
        self.parts = dict(zip(("loss_st_ed", "loss_neg_ctx", "loss_neg_q", "loss_overall"), z[key + "parts"].tolist()))
        self.grads = {k[len(key) + 2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(key + "g/")}


class AdamCase:
    """tests/golden/bert_adam.npz: the reference BertAdam run for a few steps (make_golden.bert_adam_vectors)."""
    NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")  # reference train.py:152

    def __init__(self):
        import json as _json
        self.z = z = np.load(os.path.join(GOLDEN_DIR, "bert_adam.npz"))
        self.hyper = _json.loads(str(z["hyper_json"]))
        self.names = list(_json.loads(str(z["shapes_json"])))
        self.n_steps = int(z["n_steps"])

    def weight_decay(self, name):
        return 0.0 if any(nd in name for nd in self.NO_DECAY) else 0.01

    def initial(self):
        return {k: torch.from_numpy(self.z["p0/" + k]).clone() for k in self.names}

    def grad(self, step, name):
        if not int(self.z["has_grad/%d/%s" % (step, name)]):
            return None
        return torch.from_numpy(self.z["g/%d/%s" % (step, name)]).clone()

    def t(self, key):
        return torch.from_numpy(self.z[key])


class VisualizationCase:
    """tests/golden/visualization.npz: XML.get_visualization_data of the reference on a training batch."""
    KEYS = ("modular_att_scores", "st_prob", "ed_prob", "similarity_scores", "video_similarity", "sub_similarity")

    def __init__(self):
        z = np.load(os.path.join(GOLDEN_DIR, "visualization.npz"))
        self.train = TrainCase(str(z["case"]), "plain")
        self.items = [{k: z["%d/%s" % (i, k)] for k in self.KEYS + ("st_ed_indices",)} for i in range(int(z["n"]))]

    def check(self, got, rtol, atol):
        """got: list of per-example dicts (the product) or dict of full tensors (the oracle)."""
        i = self.train.inputs
        q_len = i["query_mask"].sum(1).long().tolist()
        c_len = i["video_mask"].sum(1).long().tolist()
        assert len(self.items) == len(q_len)
        for n, want in enumerate(self.items):
            for k in self.KEYS:
                if isinstance(got, dict):
                    g = got[k][n][:(q_len[n] if k == "modular_att_scores" else c_len[n])]
                else:
                    g = got[n][k]
                g = torch.as_tensor(np.asarray(g.detach().cpu() if torch.is_tensor(g) else g))
                assert tuple(g.shape) == want[k].shape, (n, k, g.shape, want[k].shape)
                torch.testing.assert_close(g.double(), torch.from_numpy(want[k]).double(), rtol=rtol, atol=atol)
