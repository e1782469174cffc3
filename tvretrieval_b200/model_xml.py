"""Drop-in `XML` (Cross-modal Moment Localization) whose inference methods run on the xmlb200 sm_100a kernels.

Mirror of the public surface of reference baselines/crossmodal_moment_localization/model_xml.py (class `XML`,
`xml_base_config`, `mask_logits`): same constructor, attribute names, `state_dict` keys
and method signatures (`encode_context`, `encode_query`, `encode_input`, `get_modularized_queries`,
`get_video_level_scores`, `get_merged_st_ed_prob`, `get_st_ed_prob`, `get_pred_from_raw_query`, ...), so the
reference drivers and `baselines/profiling/profile_main.py` can use it unchanged.  Tensors must live on a CUDA
device: there is no CPU or PyTorch-eager fallback.  `forward` (the training step) runs the same kernels and
differentiates them through tvretrieval_b200/autograd.py.  Only `encoder_type="transformer"` with
`span_predictor_type="conv"` (the shipped configuration) is implemented.
"""
import copy
import weakref

import torch
import torch.nn as nn

from . import autograd, ops
from .model_components import BertAttention, BertSelfAttention, LinearLayer, TrainablePositionalEncoding


class AttrDict(dict):
    """Attribute-accessible dict (stands in for easydict.EasyDict, which the reference uses for configs)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


# same keys and defaults as reference model_xml.py:19-49
xml_base_config = AttrDict(
    merge_two_stream=True, cross_att=True, span_predictor_type="conv", encoder_type="transformer",
    add_pe_rnn=False, visual_input_size=2048, query_input_size=768, sub_input_size=768, hidden_size=500,
    conv_kernel_size=5, stack_conv_predictor_conv_kernel_sizes=-1, conv_stride=1, max_ctx_l=100, max_desc_l=30,
    input_drop=0.1, drop=0.1, n_heads=4, ctx_mode="video_sub", margin=0.1, ranking_loss_type="hinge",
    lw_neg_q=1, lw_neg_ctx=1, lw_st_ed=1, use_hard_negative=False, hard_pool_size=20, use_self_attention=True,
    no_modular=False, pe_type="none", initializer_range=0.02)


def mask_logits(target, mask):
    """reference model_xml.py:640-641 (kept as a torch expression: it is part of the module's public API)."""
    return target * mask + (1 - mask) * (-1e10)


def packed_layout(lens, width):
    """Index arithmetic of the packed (ragged) token layout: queries with `lens[i]` valid tokens (clamped to `width`,
    the padded length) -> (rows, pos, cu_seqlens, max_len): for every packed token its row in the padded
    (N * width) layout and its position inside its query; cu_seqlens[i] = first packed row of query i."""
    import numpy as np
    lens = np.minimum(np.asarray(lens, dtype=np.int64), width)
    n = len(lens)
    cu = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=cu[1:])
    pos = np.arange(cu[-1]) - np.repeat(cu[:-1], lens)
    rows = pos + np.repeat(np.arange(n) * width, lens)
    return rows, pos, cu, int(max(1, lens.max(initial=1)))


class XML(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if config.encoder_type != "transformer":
            raise NotImplementedError("xmlb200 implements encoder_type='transformer' only (got %r)"
                                      % (config.encoder_type,))
        if config.span_predictor_type != "conv":
            raise NotImplementedError("xmlb200 implements span_predictor_type='conv' only")
        if config.stack_conv_predictor_conv_kernel_sizes != -1:
            raise NotImplementedError("stacked ConvSE predictors are marked 'do not use' upstream; unsupported")
        if config.conv_stride != 1:
            raise NotImplementedError("ConvSE stride must be 1")
        hsz = config.hidden_size
        att_cfg = AttrDict(hidden_size=hsz, intermediate_size=hsz, hidden_dropout_prob=config.drop,
                           attention_probs_dropout_prob=config.drop, num_attention_heads=config.n_heads)

        def conv_se():
            k = config.conv_kernel_size
            return nn.Conv1d(1, 1, kernel_size=k, stride=1, padding=k // 2, bias=False)

        # registration order follows the reference constructor (model_xml.py:59-165) so that module traversal,
        # and with it seeded initialisation, visits parameters in the same order
        self.query_pos_embed = TrainablePositionalEncoding(config.max_desc_l, hsz, dropout=config.input_drop)
        self.ctx_pos_embed = TrainablePositionalEncoding(config.max_ctx_l, hsz, dropout=config.input_drop)
        self.query_input_proj = LinearLayer(config.query_input_size, hsz, layer_norm=True,
                                            dropout=config.input_drop, relu=True)
        self.query_encoder = BertAttention(att_cfg)
        self.use_video = "video" in config.ctx_mode
        self.use_sub = "sub" in config.ctx_mode
        for name, used, in_dim in (("video", self.use_video, config.visual_input_size),
                                   ("sub", self.use_sub, config.sub_input_size)):
            if not used:
                continue
            setattr(self, name + "_input_proj", LinearLayer(in_dim, hsz, layer_norm=True,
                                                            dropout=config.input_drop, relu=True))
            setattr(self, name + "_encoder1", copy.deepcopy(self.query_encoder))
            setattr(self, name + "_encoder2", copy.deepcopy(self.query_encoder))
            if config.cross_att:
                setattr(self, name + "_cross_att", BertSelfAttention(att_cfg))
                setattr(self, name + "_cross_layernorm", nn.LayerNorm(hsz))
            else:
                setattr(self, name + "_encoder3", copy.deepcopy(self.query_encoder))
            setattr(self, name + "_query_linear", nn.Linear(hsz, hsz))
            if not config.merge_two_stream:
                setattr(self, name + "_st_predictor", conv_se())
                setattr(self, name + "_ed_predictor", conv_se())
        self.modular_vector_mapping = nn.Linear(hsz, self.use_sub + self.use_video, bias=False)
        self.temporal_criterion = nn.CrossEntropyLoss(reduction="mean")
        if config.merge_two_stream:
            self.merged_st_predictor = conv_se()
            self.merged_ed_predictor = conv_se()
        self._norm_cache = {}
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters / config
    def reset_parameters(self):
        """Same scheme as reference model_xml.py:185-201: N(0, initializer_range) for Linear/Embedding weights,
        zero Linear biases, LayerNorm (1, 0), default Conv1d init."""
        std = self.config.initializer_range

        def init(m):
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
            elif isinstance(m, nn.LayerNorm):
                m.bias.data.zero_()
                m.weight.data.fill_(1.0)
            elif isinstance(m, nn.Conv1d):
                m.reset_parameters()
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

        self.apply(init)
        ops.invalidate_weight_caches()  # .data writes do not bump tensor._version

    def _load_from_state_dict(self, *args, **kw):
        super()._load_from_state_dict(*args, **kw)
        ops.invalidate_weight_caches()

    def set_hard_negative(self, use_hard_negative, hard_pool_size):
        self.config.use_hard_negative = use_hard_negative
        self.config.hard_pool_size = hard_pool_size

    def set_train_st_ed(self, lw_st_ed):
        self.config.lw_st_ed = lw_st_ed

    # Linear-layer kernels of the training step (forward and the dX / dW GEMMs of the backward pass): "f16x3" (default)
    # = split-precision tcgen05 GEMMs with K-chunked fp32 accumulation; the backward GEMMs use bf16 halves because
    # output gradients (~1e-6) are out of fp16's range.  Measured at TVR dims against float64
    # (profiles/r02_train_grad_accuracy.txt): median gradient error 6e-6 of each tensor's largest entry; the worst
    # tensor (3.7e-3, an input projection) is where ReLU activations within rounding of zero flip their gradient mask
    # -- torch's own fp32 CPU evaluation of the same batch shows 1.7e-1 on the corresponding tensor.  "f32" = exact
    # SIMT FMA GEMMs (3x slower).
    train_precision = "f16x3"

    def forward(self, query_feat, query_mask, video_feat, video_mask, sub_feat, sub_mask, tef_feat, tef_mask,
                st_ed_indices):
        """The training step, reference model_xml.py:212-251: encode the batch's contexts and queries, in-batch
        video-level scores (N, N) and span logits of each query on its own video (N, L) -> weighted sum of the
        start/end cross-entropy and the two sampled-negative ranking losses.
        -> (loss: 0-dim tensor with grad, {"loss_st_ed", "loss_neg_ctx", "loss_neg_q", "loss_overall": float}).
        Forward values come from the CUDA kernels; gradients flow through tvretrieval_b200/autograd.py.
        tef_feat / tef_mask are accepted and ignored like in the reference."""
        cfg = self.config
        prec = self.train_precision
        video_feat1, video_feat2, sub_feat1, sub_feat2 = self._encode_context(video_feat, video_mask, sub_feat,
                                                                              sub_mask, prec)
        q2c, st_logits, ed_logits = self.get_pred_from_raw_query(
            query_feat, query_mask, video_feat1, video_feat2, video_mask, sub_feat1, sub_feat2, sub_mask,
            cross=False, precision=prec)
        loss_st_ed = 0
        if cfg.lw_st_ed != 0:
            loss_st_ed = self.temporal_criterion(st_logits, st_ed_indices[:, 0]) + \
                self.temporal_criterion(ed_logits, st_ed_indices[:, 1])
        loss_neg_ctx = loss_neg_q = 0
        if cfg.lw_neg_ctx != 0 or cfg.lw_neg_q != 0:
            loss_neg_ctx, loss_neg_q = self.get_video_level_loss(q2c)
        loss_st_ed = cfg.lw_st_ed * loss_st_ed
        loss_neg_ctx = cfg.lw_neg_ctx * loss_neg_ctx
        loss_neg_q = cfg.lw_neg_q * loss_neg_q
        loss = loss_st_ed + loss_neg_ctx + loss_neg_q
        def scalar(x):
            return float(x.detach()) if torch.is_tensor(x) else float(x)
        return loss, {"loss_st_ed": scalar(loss_st_ed), "loss_neg_ctx": scalar(loss_neg_ctx),
                      "loss_neg_q": scalar(loss_neg_q), "loss_overall": scalar(loss)}

    @torch.no_grad()
    def get_visualization_data(self, query_feat, query_mask, video_feat, video_mask, sub_feat, sub_mask, tef_feat,
                               tef_mask, st_ed_indices):
        """reference model_xml.py:253-289: per example the modular token attention, start / end logits and the
        similarity curves of the query on its own video, cut to the valid lengths -> list of N dicts of numpy arrays."""
        assert self.config.merge_two_stream and self.use_video and self.use_sub and not self.config.no_modular
        _, video_feat2, _, sub_feat2 = self.encode_context(video_feat, video_mask, sub_feat, sub_mask)
        encoded_query = self.encode_input(query_feat, query_mask, self.query_input_proj, self.query_encoder,
                                          self.query_pos_embed)
        video_query, sub_query, att = self.get_modularized_queries(encoded_query, query_mask, return_modular_att=True)
        st, ed, sim, v_sim, s_sim = self.get_merged_st_ed_prob(video_query, video_feat2, sub_query, sub_feat2,
                                                               video_mask, cross=False, return_similaity=True)
        data = dict(modular_att_scores=att, st_prob=st, ed_prob=ed, similarity_scores=sim, video_similarity=v_sim,
                    sub_similarity=s_sim, st_ed_indices=st_ed_indices)
        data = {k: v.cpu().numpy() for k, v in data.items()}
        q_len = query_mask.sum(1).to(torch.long).cpu().tolist()
        c_len = video_mask.sum(1).to(torch.long).cpu().tolist()
        return [{k: v[i][:(q_len[i] if k == "modular_att_scores" else c_len[i])] for k, v in data.items()}
                for i in range(len(q_len))]

    # ------------------------------------------------------------------ losses (tiny (N, N) / (N,) tensors)
    def get_video_level_loss(self, query_context_scores):
        """reference model_xml.py:588-605: ranking losses of each positive pair (the diagonal) against one sampled
        negative video per query (rows) and one sampled negative query per video (columns)."""
        n = len(query_context_scores)
        diag = torch.arange(n, device=query_context_scores.device)
        pos = query_context_scores[diag, diag]
        masked = query_context_scores.detach().clone()
        masked[diag, diag] = 999  # the positive sorts first and is skipped by the sampler
        neg_ctx = self.get_neg_scores(query_context_scores, masked)
        neg_q = self.get_neg_scores(query_context_scores.transpose(0, 1), masked.transpose(0, 1))
        return self.get_ranking_loss(pos, neg_ctx), self.get_ranking_loss(pos, neg_q)

    def get_neg_scores(self, scores, scores_masked):
        """reference model_xml.py:607-625: per row one negative drawn uniformly from ranks [1, 1 + hard_pool_size)
        (hard negatives) or [1, N) of the row sorted descending; the draw is torch.randint on the default CPU
        generator, as in the reference, so torch.manual_seed reproduces its sampling."""
        n = len(scores)
        rows = torch.arange(n, device=scores.device)
        order = torch.sort(scores_masked, descending=True, dim=1)[1]
        hi = min(1 + self.config.hard_pool_size, n) if self.config.use_hard_negative else n
        pick = torch.randint(1, hi, size=(n,)).to(scores.device)
        return scores[rows, order[rows, pick]]

    def get_ranking_loss(self, pos_score, neg_score):
        """reference model_xml.py:627-637."""
        kind = self.config.ranking_loss_type
        if kind == "hinge":
            return torch.clamp(self.config.margin + neg_score - pos_score, min=0).sum() / len(pos_score)
        if kind == "lse":
            return torch.log1p(torch.exp(neg_score - pos_score)).sum() / len(pos_score)
        raise NotImplementedError("Only support 'hinge' and 'lse'")

    # ------------------------------------------------------------------ encoders
    def encode_input(self, feat, mask, input_proj_layer, encoder_layer, pos_embed_layer,
                     precision=ops.DEFAULT_PRECISION):
        """reference model_xml.py:377-392: projection -> position + LN -> self-attention block."""
        feat = pos_embed_layer(input_proj_layer(feat, precision=precision))
        return encoder_layer(feat, mask.unsqueeze(1), precision=precision)

    def encode_query(self, query_feat, query_mask, precision=ops.DEFAULT_PRECISION):
        encoded = self.encode_input(query_feat, query_mask, self.query_input_proj, self.query_encoder,
                                    self.query_pos_embed, precision=precision)
        return self.get_modularized_queries(encoded, query_mask)

    PACKED_MAX_LEN = 32  # longest query the packed encoder handles (xmlb_attention_ragged)

    def _upload_ints(self, arr, dev):
        """Small host int32 table -> device through a ring of reusable pinned staging buffers (a pageable source makes
        the copy synchronous with the stream, and pinning a fresh buffer per call costs more than the encoder)."""
        ring = self.__dict__.get("_pin_ring")
        if ring is None:  # all slots at once: pinning a buffer costs about a millisecond and synchronises the device
            ring = self.__dict__["_pin_ring"] = {"next": 0, "slots": [
                [torch.empty(max(arr.size, 1 << 17), dtype=torch.int32).pin_memory(), None] for _ in range(8)]}
        i = ring["next"]
        ring["next"] = (i + 1) % len(ring["slots"])
        slot = ring["slots"][i]
        if slot is None or slot[0].numel() < arr.size:
            slot = [torch.empty(max(arr.size, 1 << 17), dtype=torch.int32).pin_memory(), None]
            ring["slots"][i] = slot
        if slot[1] is not None:
            slot[1].synchronize()  # the copy issued 8 uploads ago has long left this buffer
        slot[0][:arr.size].numpy()[:] = arr
        out = slot[0][:arr.size].to(dev, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(dev))
        return out

    def packed_query_tables(self, lens_cpu, width, dev):
        """Index tables of the packed layout for queries with `lens_cpu` valid tokens out of `width` padded ones:
        (row of every packed token in the padded (N * width, Dq) layout, its position in its query, cu_seqlens,
        longest query), uploaded through pinned staging.  Callers that also stream the query features from the host
        create the tables BEFORE enqueueing those big copies (the H2D engine serves copies in issue order)."""
        import numpy as np
        rows, pos, cu, max_len = packed_layout(lens_cpu, width)
        assert max_len <= self.PACKED_MAX_LEN
        t = self._upload_ints(np.concatenate([rows, pos, cu]).astype(np.int32), dev)
        return t[:len(rows)], t[len(rows):2 * len(rows)], t[2 * len(rows):], max_len

    @torch.no_grad()
    def packed_query_tables_device(self, lens_cpu, width, dev, bounds, lens_dev=None):
        """packed_query_tables for all pieces `bounds` = [(lo, hi), ...] of a block of queries at once, with the
        per-token tables made ON THE DEVICE (cumsum / repeat_interleave over the lengths): the host only needs the few
        sums that size the tensors.  Building them in numpy cost ~2 ms per 8 K queries during which the GPU had
        nothing to do.  lens_dev: the lengths already on the device (else lens_cpu is uploaded, 4 bytes per query).
        -> list of (rows, pos, cu_seqlens, max_len) per piece, equal to packed_layout()'s."""
        import numpy as np
        lens_np = np.minimum(np.asarray(lens_cpu, dtype=np.int64), width)
        n = len(lens_np)
        cu_np = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens_np, out=cu_np[1:])
        total = int(cu_np[-1])
        if lens_dev is None:
            lens_dev = self._upload_ints(lens_np.astype(np.int32), dev)
        lens_dev = lens_dev.to(torch.int64).clamp(max=width)
        cu = torch.zeros(n + 1, device=dev, dtype=torch.int64)
        torch.cumsum(lens_dev, 0, out=cu[1:])
        seq = torch.repeat_interleave(torch.arange(n, device=dev), lens_dev, output_size=total)
        pos = torch.arange(total, device=dev) - cu[seq]
        rows = seq * width + pos
        pos = pos.to(torch.int32)
        tables = []
        for lo, hi in bounds:
            t0, t1 = int(cu_np[lo]), int(cu_np[hi])
            max_len = int(max(1, lens_np[lo:hi].max(initial=1)))
            assert max_len <= self.PACKED_MAX_LEN
            tables.append(((rows[t0:t1] - lo * width).to(torch.int32), pos[t0:t1],
                           (cu[lo:hi + 1] - t0).to(torch.int32), max_len))
        return tables

    @torch.no_grad()
    def encode_query_packed(self, query_feat, lens_cpu=None, tables=None, precision=ops.DEFAULT_PRECISION):
        """encode_query (reference model_xml.py:291-295) on the VALID tokens only: query_feat (N, Lq, Dq) padded
        features on the device, lens_cpu (N,) host ints = valid tokens per query (masks are prefix masks,
        utils/tensor_utils.py:5-53), or `tables` = packed_query_tables(...) made earlier.  The tokens are packed into
        (T, Dq) rows, every row-wise layer runs on T instead of N * Lq rows, attention and pooling are
        sequence-aware kernels.  Same pooled vectors as encode_query (padded tokens carry exactly zero weight
        there); inference only."""
        n, width = query_feat.shape[:2]
        if tables is None:
            tables = self.packed_query_tables(lens_cpu, width, query_feat.device)
        rows_d, pos_d, cu_d, max_len = tables
        assert cu_d.numel() == n + 1
        x = query_feat.reshape(n * width, -1).index_select(0, rows_d)
        pe, proj = self.query_pos_embed, self.query_input_proj
        att, outp = self.query_encoder.self, self.query_encoder.output
        if len(x) >= 256 and precision != "f32" and self.config.hidden_size % 64 == 0:
            # tensor-core pipeline with fused stages: LN -> split | GEMM + ReLU | position LN -> (fp32, split) |
            # ONE GEMM for Q, K and V | fused ragged attention | split | output GEMM + residual | LN | ragged pooling
            bf16 = precision == "bf16x3"
            _, x16 = ops.add_layernorm_split(x, proj.LayerNorm.weight, proj.LayerNorm.bias, eps=proj.LayerNorm.eps,
                                             bf16=bf16, want_f32=False)
            fc = proj.net[1]
            h0, _, _ = ops.linear_tc_ex(x16, ops._weight_split(fc.weight, bf16), fc.bias, relu=True, bf16=bf16)
            h, h16 = ops.add_layernorm_split(h0, pe.LayerNorm.weight, pe.LayerNorm.bias,
                                             add=pe.position_embeddings.weight, add_index=pos_d,
                                             eps=pe.LayerNorm.eps, bf16=bf16)
            w = ops._weight_split_cat([att.query.weight, att.key.weight, att.value.weight], bf16)
            b = ops._bias_cat([att.query.bias, att.key.bias, att.value.bias])
            qkv, _, _ = ops.linear_tc_ex(h16, w, b, bf16=bf16)
            ctx = ops.attention_ragged_qkv(qkv, cu_d, max_len, att.num_attention_heads)
            o, _, _ = ops.linear_tc_ex(ops.split_rows(ctx, bf16=bf16), ops._weight_split(outp.dense.weight, bf16),
                                       outp.dense.bias, residual=h, bf16=bf16)
        else:
            h = proj(x, precision=precision)
            h = ops.add_layernorm_indexed(h, pe.LayerNorm.weight, pe.LayerNorm.bias, pe.position_embeddings.weight,
                                          pos_d, eps=pe.LayerNorm.eps)
            q = ops.linear(h, att.query.weight, att.query.bias, precision=precision)
            k = ops.linear(h, att.key.weight, att.key.bias, precision=precision)
            v = ops.linear(h, att.value.weight, att.value.bias, precision=precision)
            ctx = ops.attention_ragged(q, k, v, cu_d, max_len, att.num_attention_heads)
            o = ops.linear(ctx, outp.dense.weight, outp.dense.bias, residual=h, precision=precision)
        o = ops.add_layernorm(o, outp.LayerNorm.weight, outp.LayerNorm.bias, eps=outp.LayerNorm.eps)
        return ops.modular_pool_ragged(o, cu_d, self.PACKED_MAX_LEN, self.modular_vector_mapping.weight)

    def get_modularized_queries(self, encoded_query, query_mask, return_modular_att=False):
        """reference model_xml.py:399-423."""
        if self.config.no_modular:
            raise NotImplementedError("no_modular=True is not supported")
        pooled = ops.modular_pool(encoded_query, query_mask, self.modular_vector_mapping.weight)
        if not return_modular_att:
            return pooled
        # visualisation path (reference model_xml.py:410-416): the token attention itself, (N, Lq, 2)
        assert self.modular_vector_mapping.weight.shape[0] == 2
        # 2 output features: not a tensor-core shape
        logits = ops.linear(encoded_query, self.modular_vector_mapping.weight, precision="f32")      # (N, Lq, 2)
        logits = mask_logits(logits, query_mask.unsqueeze(2)).transpose(1, 2).contiguous()  # (N, 2, Lq)
        att = ops.softmax_rows(logits).transpose(1, 2).contiguous()
        return pooled[0], pooled[1], att

    def cross_context_encoder(self, main_context_feat, main_context_mask, side_context_feat, side_context_mask,
                              cross_att_layer, norm_layer, self_att_layer, precision=ops.DEFAULT_PRECISION):
        """reference model_xml.py:357-373."""
        cross_mask = main_context_mask.unsqueeze(2) * side_context_mask.unsqueeze(1)  # (N, Lq, Lk) {0,1}
        cross_out = cross_att_layer(main_context_feat, side_context_feat, side_context_feat, cross_mask,
                                    precision=precision)
        residual_out = ops.add_layernorm(cross_out, norm_layer.weight, norm_layer.bias, add=main_context_feat,
                                         eps=norm_layer.eps)
        return self_att_layer(residual_out, main_context_mask.unsqueeze(1), precision=precision)

    def cross_encode_context(self, video_feat, video_mask, sub_feat, sub_mask, precision=ops.DEFAULT_PRECISION):
        """reference model_xml.py:344-355."""
        v1 = self.encode_input(video_feat, video_mask, self.video_input_proj, self.video_encoder1, self.ctx_pos_embed,
                               precision=precision)
        s1 = self.encode_input(sub_feat, sub_mask, self.sub_input_proj, self.sub_encoder1, self.ctx_pos_embed,
                               precision=precision)
        v2 = self.cross_context_encoder(v1, video_mask, s1, sub_mask, self.video_cross_att,
                                        self.video_cross_layernorm, self.video_encoder2, precision=precision)
        s2 = self.cross_context_encoder(s1, sub_mask, v1, video_mask, self.sub_cross_att,
                                        self.sub_cross_layernorm, self.sub_encoder2, precision=precision)
        return v1, v2, s1, s2

    def non_cross_encode_context(self, context_feat, context_mask, module_name="video",
                                 precision=ops.DEFAULT_PRECISION):
        """reference model_xml.py:297-329: feat1 = enc1(...), feat2 = enc3(enc2(feat1))."""
        feat1 = self.encode_input(context_feat, context_mask, getattr(self, module_name + "_input_proj"),
                                  getattr(self, module_name + "_encoder1"), self.ctx_pos_embed, precision=precision)
        m3 = context_mask.unsqueeze(1)
        feat2 = getattr(self, module_name + "_encoder2")(feat1, m3, precision=precision)
        feat2 = getattr(self, module_name + "_encoder3")(feat2, m3, precision=precision)
        return feat1, feat2

    # Kernels used by encode_context: "f16x3" (default) / "bf16x3" = the tensor-core pipeline below (split-precision
    # tcgen05 GEMMs with K-chunked fp32 accumulation + the fused tcgen05 attention kernel), "f32" = exact-fp32 SIMT
    # kernels layer by layer.  Both reproduce the reference's fp32 values to ~1e-5.
    context_precision = "f16x3"

    def encode_context(self, video_feat, video_mask, sub_feat, sub_mask):
        """reference model_xml.py:331-342."""
        return self._encode_context(video_feat, video_mask, sub_feat, sub_mask, self.context_precision)

    def _encode_context(self, video_feat, video_mask, sub_feat, sub_mask, precision):
        if self._tc_context_ok(video_feat if self.use_video else sub_feat, precision):
            return self._encode_context_tc(video_feat, video_mask, sub_feat, sub_mask, precision == "bf16x3")
        if self.config.cross_att:
            assert self.use_video and self.use_sub
            return self.cross_encode_context(video_feat, video_mask, sub_feat, sub_mask, precision=precision)
        v1 = v2 = s1 = s2 = None
        if self.use_video:
            v1, v2 = self.non_cross_encode_context(video_feat, video_mask, module_name="video", precision=precision)
        if self.use_sub:
            s1, s2 = self.non_cross_encode_context(sub_feat, sub_mask, module_name="sub", precision=precision)
        return v1, v2, s1, s2

    # ---- the context encoders on the tensor cores ------------------------------------------------------------
    def _tc_context_ok(self, feat, precision):
        """The fused tensor-core pipeline serves eval-mode batches of at least 256 clip rows whose head size the
        attention kernel supports (64 / 128 / 192 / 256) and whose sequences are at most 256 clips long."""
        if precision == "f32" or feat is None or not feat.is_cuda or self.training or torch.is_grad_enabled():
            return False  # (inference only: the fused pipeline records no autograd graph)
        n, length = feat.shape[:2]
        return (n * length >= 256 and self.config.hidden_size % 64 == 0
                and ops.attention_tc_supported(self.config.hidden_size, self.config.n_heads, length))

    def _tc_input(self, feat, proj, bf16):
        """LinearLayer + TrainablePositionalEncoding (reference model_components.py:156-163, 81-88):
        LN -> split | GEMM + bias + ReLU | + position, LN -> (fp32 rows, split rows)."""
        n, length, din = feat.shape
        _, x16 = ops.add_layernorm_split(feat.reshape(n * length, din), proj.LayerNorm.weight, proj.LayerNorm.bias,
                                         eps=proj.LayerNorm.eps, bf16=bf16, want_f32=False)
        fc = proj.net[1]
        h0, _, _ = ops.linear_tc_ex(x16, ops._weight_split(fc.weight, bf16), fc.bias, relu=True, bf16=bf16)
        pe = self.ctx_pos_embed
        if length > pe.position_embeddings.num_embeddings:
            raise IndexError("sequence length %d exceeds the %d learned positions"
                             % (length, pe.position_embeddings.num_embeddings))
        return ops.add_layernorm_split(h0, pe.LayerNorm.weight, pe.LayerNorm.bias, add=pe.position_embeddings.weight,
                                       add_rows=length, eps=pe.LayerNorm.eps, bf16=bf16)

    def _tc_block(self, x, x16, mask3, enc, n, length, bf16):
        """BertAttention (reference model_components.py:207-216, 266-317) on rows x (fp32) / x16 (split):
        fused QKV GEMM (Q, K split row-major; V split transposed per sequence) | fused attention | output GEMM +
        residual | LN -> (fp32 rows, split rows)."""
        att, outp = enc.self, enc.output
        hid, nh = self.config.hidden_size, att.num_attention_heads
        w = ops._weight_split_cat([att.query.weight, att.key.weight, att.value.weight], bf16)
        b = ops._bias_cat([att.query.bias, att.key.bias, att.value.bias])
        _, qk16, vt16 = ops.linear_tc_ex(x16, w, b, bf16=bf16, want_f32=False, out16_cols=2 * hid, vt_col0=2 * hid,
                                         vt_seq=length)
        _, c16 = ops.attention_tc(qk16, 0, qk16, hid, vt16, mask3, n, length, length, hid, nh, bf16=bf16,
                                  want_f32=False, want_split=True)
        o, _, _ = ops.linear_tc_ex(c16, ops._weight_split(outp.dense.weight, bf16), outp.dense.bias, residual=x,
                                   bf16=bf16)
        return ops.add_layernorm_split(o, outp.LayerNorm.weight, outp.LayerNorm.bias, eps=outp.LayerNorm.eps,
                                       bf16=bf16)

    def _tc_cross(self, main, main16, main_mask, side16, side_mask, att, norm, enc2, n, length, bf16):
        """cross_context_encoder (reference model_xml.py:357-373) on the tensor cores."""
        hid, nh = self.config.hidden_size, att.num_attention_heads
        _, q16, _ = ops.linear_tc_ex(main16, ops._weight_split(att.query.weight, bf16), att.query.bias, bf16=bf16,
                                     want_f32=False, out16_cols=hid)
        wkv = ops._weight_split_cat([att.key.weight, att.value.weight], bf16)
        _, k16, vt16 = ops.linear_tc_ex(side16, wkv, ops._bias_cat([att.key.bias, att.value.bias]), bf16=bf16,
                                        want_f32=False, out16_cols=hid, vt_col0=hid, vt_seq=length)
        cross_mask = (main_mask.unsqueeze(2) * side_mask.unsqueeze(1)).contiguous()  # (N, Lq, Lk) {0,1}
        x, _ = ops.attention_tc(q16, 0, k16, 0, vt16, cross_mask, n, length, length, hid, nh, bf16=bf16)
        r, r16 = ops.add_layernorm_split(x, norm.weight, norm.bias, add=main, eps=norm.eps, bf16=bf16)
        return self._tc_block(r, r16, main_mask.unsqueeze(1).contiguous(), enc2, n, length, bf16)[0]

    @torch.no_grad()
    def _encode_context_tc(self, video_feat, video_mask, sub_feat, sub_mask, bf16):
        hid = self.config.hidden_size
        streams = {}
        for name, used, feat, mask in (("video", self.use_video, video_feat, video_mask),
                                       ("sub", self.use_sub, sub_feat, sub_mask)):
            if not used:
                continue
            n, length = feat.shape[:2]
            h1, h16 = self._tc_input(feat.contiguous(), getattr(self, name + "_input_proj"), bf16)
            m3 = mask.unsqueeze(1).contiguous()
            f1, f16 = self._tc_block(h1, h16, m3, getattr(self, name + "_encoder1"), n, length, bf16)
            streams[name] = (f1, f16, mask, m3, n, length)
        out = {}
        if self.config.cross_att:
            assert self.use_video and self.use_sub
            for main, side in (("video", "sub"), ("sub", "video")):
                f1, f16, mask, m3, n, length = streams[main]
                assert streams[side][4:] == (n, length), "video and subtitle batches must have the same shape"
                f2 = self._tc_cross(f1, f16, mask, streams[side][1], streams[side][2],
                                    getattr(self, main + "_cross_att"), getattr(self, main + "_cross_layernorm"),
                                    getattr(self, main + "_encoder2"), n, length, bf16)
                out[main] = (f1.view(n, length, hid), f2.view(n, length, hid))
        else:
            for name, (f1, f16, mask, m3, n, length) in streams.items():
                f2, f2_16 = self._tc_block(f1, f16, m3, getattr(self, name + "_encoder2"), n, length, bf16)
                f2, _ = self._tc_block(f2, f2_16, m3, getattr(self, name + "_encoder3"), n, length, bf16)
                out[name] = (f1.view(n, length, hid), f2.view(n, length, hid))
        v = out.get("video", (None, None))
        s_ = out.get("sub", (None, None))
        return v[0], v[1], s_[0], s_[1]

    # ------------------------------------------------------------------ scoring
    def _normalized_corpus(self, feat1):
        """The reference re-normalises the whole corpus tensor on every call (model_xml.py:447); the result only
        depends on the tensor, so it is cached per (storage, version)."""
        if autograd.recording(feat1):  # training: differentiable, never cached
            return ops.l2norm_rows(feat1)
        # Entries are valid only for the very tensor object they were made from: kernel-written tensors all have
        # _version 0 and the caching allocator re-uses addresses, so (data_ptr, shape, version) alone would serve
        # the normalisation of a freed corpus to a new one of the same shape.
        key = (feat1.data_ptr(), tuple(feat1.shape), feat1._version)
        hit = self._norm_cache.get(key)
        if hit is not None and hit[0]() is feat1:
            return hit[1]
        if len(self._norm_cache) > 4:
            self._norm_cache.clear()
        out = ops.l2norm_rows(feat1)
        self._norm_cache[key] = (weakref.ref(feat1), out)
        return out

    def get_video_level_scores(self, modularied_query, context_feat1, context_mask):
        """reference model_xml.py:436-453 -> (Nq, Nv)."""
        return ops.vr_scores_f32(ops.l2norm_rows(modularied_query), None, self._normalized_corpus(context_feat1),
                                 None, context_mask, None)

    def get_merged_st_ed_prob(self, video_query, video_feat, sub_query, sub_feat, context_mask, cross=False,
                              return_similaity=False):
        """reference model_xml.py:455-502 -> masked st/ed logits."""
        assert self.use_video and self.use_sub and self.config.span_predictor_type == "conv"
        qv = ops.linear(video_query, self.video_query_linear.weight, self.video_query_linear.bias)
        qs = ops.linear(sub_query, self.sub_query_linear.weight, self.sub_query_linear.bias)
        lists = None if cross else ops.diagonal_pair_lists(len(qv), qv.device)
        st, ed = ops.span_logits(qv, video_feat, context_mask, self.merged_st_predictor.weight,
                                 self.merged_ed_predictor.weight, q_b=qs, feat2_b=sub_feat, mask_b=context_mask,
                                 merged=True, lists=lists)
        if not return_similaity:
            return st, ed
        # visualisation path (reference model_xml.py:484-485,498-500): the similarity curves themselves = the same
        # kernel with an identity "convolution" (one tap of 1.0) and an all-ones mask
        assert not cross
        one = torch.ones(1, device=qv.device)
        ones = torch.ones_like(context_mask)
        sim = ops.span_logits(qv, video_feat, ones, one, one, q_b=qs, feat2_b=sub_feat, mask_b=ones, merged=True,
                              lists=lists)[0]
        v_sim = ops.span_logits(qv, video_feat, ones, one, one, lists=lists)[0]
        s_sim = ops.span_logits(qs, sub_feat, ones, one, one, lists=lists)[0]
        return st, ed, sim, v_sim, s_sim

    def get_st_ed_prob(self, modularied_query, context_feat2, context_mask, module_name="video", cross=False):
        """reference model_xml.py:504-551 (single stream)."""
        fc = getattr(self, module_name + "_query_linear")
        q = ops.linear(modularied_query, fc.weight, fc.bias)
        lists = None if cross else ops.diagonal_pair_lists(len(q), q.device)
        return ops.span_logits(q, context_feat2, context_mask, getattr(self, module_name + "_st_predictor").weight,
                               getattr(self, module_name + "_ed_predictor").weight, lists=lists)

    def span_streams(self, video_query, sub_query, video_feat2, sub_feat2, video_mask, sub_mask,
                     precision=ops.DEFAULT_PRECISION):
        """Arguments of ops.span_logits for this model's stream layout (merged / two streams / one stream)."""
        if self.config.merge_two_stream and self.use_video and self.use_sub:
            qv = ops.linear(video_query, self.video_query_linear.weight, self.video_query_linear.bias,
                            precision=precision)
            qs = ops.linear(sub_query, self.sub_query_linear.weight, self.sub_query_linear.bias, precision=precision)
            return dict(q_a=qv, feat2_a=video_feat2, mask_a=video_mask, w_st_a=self.merged_st_predictor.weight,
                        w_ed_a=self.merged_ed_predictor.weight, q_b=qs, feat2_b=sub_feat2, mask_b=video_mask,
                        merged=True)
        streams = []
        for name, used, q, f2, m in (("video", self.use_video, video_query, video_feat2, video_mask),
                                     ("sub", self.use_sub, sub_query, sub_feat2, sub_mask)):
            if used:
                fc = getattr(self, name + "_query_linear")
                streams.append((ops.linear(q, fc.weight, fc.bias, precision=precision), f2, m,
                                getattr(self, name + "_st_predictor").weight,
                                getattr(self, name + "_ed_predictor").weight))
        args = dict(zip(("q_a", "feat2_a", "mask_a", "w_st_a", "w_ed_a"), streams[0]), merged=False)
        if len(streams) == 2:
            args.update(zip(("q_b", "feat2_b", "mask_b", "w_st_b", "w_ed_b"), streams[1]))
        return args

    def video_scores(self, video_query, sub_query, video_feat1, sub_feat1, video_mask, sub_mask):
        """q2c = mean over modalities of the per-modality masked max cosine (reference model_xml.py:572-574)."""
        return ops.vr_scores_f32(
            ops.l2norm_rows(video_query) if self.use_video else None,
            ops.l2norm_rows(sub_query) if self.use_sub else None,
            self._normalized_corpus(video_feat1) if self.use_video else None,
            self._normalized_corpus(sub_feat1) if self.use_sub else None,
            video_mask if self.use_video else None, sub_mask if self.use_sub else None)

    def get_pred_from_raw_query(self, query_feat, query_mask, video_feat1, video_feat2, video_mask, sub_feat1,
                                sub_feat2, sub_mask, cross=False, precision=ops.DEFAULT_PRECISION):
        """reference model_xml.py:553-586 -> (q2ctx_scores, st_logits, ed_logits); st/ed are masked logits
        (-1e10 at padded clips).  cross=False: (N,N),(N,L),(N,L); cross=True: (Nq,Nv),(Nq,Nv,L),(Nq,Nv,L)."""
        video_query, sub_query = self.encode_query(query_feat, query_mask, precision=precision)
        q2c = self.video_scores(video_query, sub_query, video_feat1, sub_feat1, video_mask, sub_mask)
        args = self.span_streams(video_query, sub_query, video_feat2, sub_feat2, video_mask, sub_mask,
                                 precision=precision)
        lists = None if cross else ops.diagonal_pair_lists(len(video_query), video_query.device)
        st, ed = ops.span_logits(lists=lists, **args)
        return q2c, st, ed
