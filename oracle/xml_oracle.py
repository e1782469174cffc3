"""CPU oracle for the XML inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a functional (stateless) fp32 restatement, in plain torch CPU ops, of the
arithmetic that the reference (jayleicn/TVRetrieval) performs on the path named by
BASELINE.json:north_star.  Every function cites the reference file:line it follows
(paths relative to the reference root).  It exists only so that `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
can check and time the CUDA product against it.  Nothing under `tvretrieval_b200/`
imports it; the product has no CPU fallback.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the real reference from
/root/reference (in the build container), runs it on seeded inputs and commits its
outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
here against those vectors.

Model weights are passed as a flat ``{state_dict key: tensor}`` mapping with the
reference's own key names, so a reference checkpoint can be fed
to the oracle and to the product unchanged.
"""
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]

LN_EPS = 1e-5  # nn.LayerNorm default, model_components.py:73,149,310
ATT_MASK_ADD = -10000.0  # model_components.py:277
LOGIT_MASK_ADD = -1e10  # model_xml.py:640-641


# --------------------------------------------------------------------------------------
# building blocks (model_components.py)
# --------------------------------------------------------------------------------------
def mask_logits(x: Tensor, m: Tensor) -> Tensor:
    """model_xml.py:640-641 -- x*m + (1-m)*(-1e10); identity where m == 1."""
    return x * m + (1 - m) * LOGIT_MASK_ADD


def layer_norm(x: Tensor, w: Weights, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[prefix + ".weight"], w[prefix + ".bias"], LN_EPS)


def input_projection(x: Tensor, w: Weights, prefix: str) -> Tensor:
    """LinearLayer.forward, model_components.py:156-163 (eval: dropout is identity).
    relu(W . LN_Din(x) + b); the Linear sits at ``net.1`` (``net.0`` is the Dropout)."""
    h = layer_norm(x, w, prefix + ".LayerNorm")
    h = F.linear(h, w[prefix + ".net.1.weight"], w[prefix + ".net.1.bias"])
    return torch.relu(h)


def add_position(x: Tensor, w: Weights, prefix: str) -> Tensor:
    """TrainablePositionalEncoding.forward, model_components.py:76-89.
    LN_H(x + E[0:L]) with E the learned table (rows 0..L-1)."""
    seq_len = x.shape[1]
    table = w[prefix + ".position_embeddings.weight"]
    return layer_norm(x + table[:seq_len].unsqueeze(0), w, prefix + ".LayerNorm")


def multi_head_attention(q_in: Tensor, kv_in: Tensor, mask3: Tensor, w: Weights, prefix: str,
                         n_heads: int) -> Tensor:
    """BertSelfAttention.forward, model_components.py:266-303.

    q_in (N, Lq, H); kv_in (N, Lk, H); mask3 (N, Lq or 1, Lk) float {0,1}.
    scores = QK^T / sqrt(dh) + (1 - mask) * -10000 (added in fp32, NOT -inf);
    softmax over keys; heads re-merged to (N, Lq, H).  No output projection here."""
    n, lq, hid = q_in.shape
    lk = kv_in.shape[1]
    dh = hid // n_heads
    add = (1 - mask3.unsqueeze(1)) * ATT_MASK_ADD  # (N, 1, Lq|1, Lk)

    def heads(t, length):
        return t.view(n, length, n_heads, dh).permute(0, 2, 1, 3)

    q = heads(F.linear(q_in, w[prefix + ".query.weight"], w[prefix + ".query.bias"]), lq)
    k = heads(F.linear(kv_in, w[prefix + ".key.weight"], w[prefix + ".key.bias"]), lk)
    v = heads(F.linear(kv_in, w[prefix + ".value.weight"], w[prefix + ".value.bias"]), lk)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh) + add
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v)  # (N, nh, Lq, dh)
    return ctx.permute(0, 2, 1, 3).contiguous().view(n, lq, hid)


def attention_block(x: Tensor, mask3: Tensor, w: Weights, prefix: str, n_heads: int) -> Tensor:
    """BertAttention.forward = BertSelfAttention + BertSelfOutput,
    model_components.py:207-216 and :313-317:  LN(dense(att) + x).  No FFN."""
    att = multi_head_attention(x, x, mask3, w, prefix + ".self", n_heads)
    dense = F.linear(att, w[prefix + ".output.dense.weight"], w[prefix + ".output.dense.bias"])
    return layer_norm(dense + x, w, prefix + ".output.LayerNorm")


# --------------------------------------------------------------------------------------
# encoders (model_xml.py)
# --------------------------------------------------------------------------------------
def encode_input(feat: Tensor, mask: Tensor, w: Weights, proj: str, encoder: str, pos: str,
                 n_heads: int) -> Tensor:
    """XML.encode_input, model_xml.py:377-392 (encoder_type == 'transformer')."""
    h = add_position(input_projection(feat, w, proj), w, pos)
    return attention_block(h, mask.unsqueeze(1), w, encoder, n_heads)


def cross_context(main: Tensor, main_mask: Tensor, side: Tensor, side_mask: Tensor, w: Weights,
                  cross_att: str, cross_ln: str, encoder2: str, n_heads: int) -> Tensor:
    """XML.cross_context_encoder, model_xml.py:357-373.
    mask[b,i,j] = m_main[b,i]*m_side[b,j]; LN(cross(main<-side) + main); then BertAttention."""
    cross_mask = main_mask.unsqueeze(2) * side_mask.unsqueeze(1)
    x = multi_head_attention(main, side, cross_mask, w, cross_att, n_heads)
    r = layer_norm(x + main, w, cross_ln)
    return attention_block(r, main_mask.unsqueeze(1), w, encoder2, n_heads)


def encode_context(cfg, w: Weights, video_feat, video_mask, sub_feat, sub_mask):
    """XML.encode_context, model_xml.py:331-342; cross_encode_context :344-355;
    non_cross_encode_context :297-329.  Returns (video_feat1, video_feat2, sub_feat1, sub_feat2),
    None for an unused modality."""
    nh = cfg["n_heads"]
    use_video = "video" in cfg["ctx_mode"]
    use_sub = "sub" in cfg["ctx_mode"]
    if cfg["cross_att"]:
        assert use_video and use_sub
        v1 = encode_input(video_feat, video_mask, w, "video_input_proj", "video_encoder1", "ctx_pos_embed", nh)
        s1 = encode_input(sub_feat, sub_mask, w, "sub_input_proj", "sub_encoder1", "ctx_pos_embed", nh)
        v2 = cross_context(v1, video_mask, s1, sub_mask, w, "video_cross_att", "video_cross_layernorm",
                           "video_encoder2", nh)
        s2 = cross_context(s1, sub_mask, v1, video_mask, w, "sub_cross_att", "sub_cross_layernorm",
                           "sub_encoder2", nh)
        return v1, v2, s1, s2

    def stacked(feat, mask, name):
        f1 = encode_input(feat, mask, w, name + "_input_proj", name + "_encoder1", "ctx_pos_embed", nh)
        m3 = mask.unsqueeze(1)
        f2 = attention_block(f1, m3, w, name + "_encoder2", nh)
        f2 = attention_block(f2, m3, w, name + "_encoder3", nh)
        return f1, f2

    v1 = v2 = s1 = s2 = None
    if use_video:
        v1, v2 = stacked(video_feat, video_mask, "video")
    if use_sub:
        s1, s2 = stacked(sub_feat, sub_mask, "sub")
    return v1, v2, s1, s2


def modular_queries(encoded: Tensor, query_mask: Tensor, w: Weights) -> Tuple[Tensor, Tensor]:
    """XML.get_modularized_queries, model_xml.py:410-423 (no_modular=False).
    a = softmax_tokens(mask_logits(encoded . W_mod)); pooled[m] = sum_t a[t,m] * encoded[t]."""
    scores = F.linear(encoded, w["modular_vector_mapping.weight"])  # (N, Lq, 1|2)
    att = torch.softmax(mask_logits(scores, query_mask.unsqueeze(2)), dim=1)
    pooled = torch.einsum("blm,bld->bmd", att, encoded)
    if pooled.shape[1] == 2:
        return pooled[:, 0], pooled[:, 1]
    return pooled[:, 0], pooled[:, 0]


def encode_query(cfg, w: Weights, query_feat: Tensor, query_mask: Tensor):
    """XML.encode_query, model_xml.py:291-295."""
    e = encode_input(query_feat, query_mask, w, "query_input_proj", "query_encoder", "query_pos_embed",
                     cfg["n_heads"])
    return modular_queries(e, query_mask, w)


# --------------------------------------------------------------------------------------
# query x corpus scoring (model_xml.py:436-586)
# --------------------------------------------------------------------------------------
def video_level_scores(q: Tensor, feat1: Tensor, mask: Tensor) -> Tensor:
    """XML.get_video_level_scores, model_xml.py:446-452.
    max_l mask_logits(normalize(q) . normalize(feat1[v,l]));  (Nq, Nv)."""
    qn = F.normalize(q, dim=-1)
    cn = F.normalize(feat1, dim=-1)
    s = torch.einsum("md,nld->mln", qn, cn)  # (Nq, L, Nv)
    s = mask_logits(s, mask.transpose(0, 1).unsqueeze(0))
    return s.max(dim=1)[0]


def conv_se(sim: Tensor, weight: Tensor) -> Tensor:
    """nn.Conv1d(1,1,k,stride=1,padding=k//2,bias=False) applied on (..., L): cross-correlation,
    zero padding; model_xml.py:95-100,468-471."""
    shape = sim.shape
    k = weight.shape[-1]
    out = F.conv1d(sim.reshape(-1, 1, shape[-1]), weight.view(1, 1, k), padding=k // 2)
    return out.view(shape)


def merged_st_ed_logits(w: Weights, video_query, video_feat2, sub_query, sub_feat2, ctx_mask, cross: bool):
    """XML.get_merged_st_ed_prob, model_xml.py:455-502 (stack_conv disabled)."""
    qv = F.linear(video_query, w["video_query_linear.weight"], w["video_query_linear.bias"])
    qs = F.linear(sub_query, w["sub_query_linear.weight"], w["sub_query_linear.bias"])
    if cross:
        sim = (torch.einsum("md,nld->mnl", qv, video_feat2) + torch.einsum("md,nld->mnl", qs, sub_feat2)) / 2
    else:
        sim = (torch.einsum("bd,bld->bl", qv, video_feat2) + torch.einsum("bd,bld->bl", qs, sub_feat2)) / 2
    st = conv_se(sim, w["merged_st_predictor.weight"])
    ed = conv_se(sim, w["merged_ed_predictor.weight"])
    return mask_logits(st, ctx_mask), mask_logits(ed, ctx_mask)


def single_st_ed_logits(w: Weights, query, feat2, ctx_mask, name: str, cross: bool):
    """XML._get_st_ed_prob, model_xml.py:512-551 (span_predictor_type == 'conv')."""
    q = F.linear(query, w[name + "_query_linear.weight"], w[name + "_query_linear.bias"])
    sim = torch.einsum("md,nld->mnl", q, feat2) if cross else torch.einsum("bd,bld->bl", q, feat2)
    st = conv_se(sim, w[name + "_st_predictor.weight"])
    ed = conv_se(sim, w[name + "_ed_predictor.weight"])
    return mask_logits(st, ctx_mask), mask_logits(ed, ctx_mask)


def pred_from_raw_query(cfg, w: Weights, query_feat, query_mask, video_feat1, video_feat2, video_mask,
                        sub_feat1, sub_feat2, sub_mask, cross: bool = False):
    """XML.get_pred_from_raw_query, model_xml.py:553-586.  Returns (q2c, st_logits, ed_logits);
    st/ed are masked *logits* (-1e10 at padded clips)."""
    use_video = "video" in cfg["ctx_mode"]
    use_sub = "sub" in cfg["ctx_mode"]
    vq, sq = encode_query(cfg, w, query_feat, query_mask)
    divisor = use_video + use_sub
    s_video = video_level_scores(vq, video_feat1, video_mask) if use_video else 0
    s_sub = video_level_scores(sq, sub_feat1, sub_mask) if use_sub else 0
    q2c = (s_video + s_sub) / divisor
    if cfg["merge_two_stream"] and use_video and use_sub:
        st, ed = merged_st_ed_logits(w, vq, video_feat2, sq, sub_feat2, video_mask, cross)
    else:
        v_st, v_ed = single_st_ed_logits(w, vq, video_feat2, video_mask, "video", cross) if use_video else (0, 0)
        s_st, s_ed = single_st_ed_logits(w, sq, sub_feat2, sub_mask, "sub", cross) if use_sub else (0, 0)
        st = (v_st + s_st) / divisor
        ed = (v_ed + s_ed) / divisor
    return q2c, st, ed


def visualization_data(cfg, w: Weights, query_feat, query_mask, video_feat, video_mask, sub_feat, sub_mask):
    """XML.get_visualization_data, model_xml.py:253-289 (tensors before the per-example cut to valid lengths):
    modular token attention (N, Lq, 2), masked start / end logits, merged / video / subtitle similarity (N, L)."""
    _, v2, _, s2 = encode_context(cfg, w, video_feat, video_mask, sub_feat, sub_mask)
    e = encode_input(query_feat, query_mask, w, "query_input_proj", "query_encoder", "query_pos_embed", cfg["n_heads"])
    att = torch.softmax(mask_logits(F.linear(e, w["modular_vector_mapping.weight"]), query_mask.unsqueeze(2)), dim=1)
    vq, sq = modular_queries(e, query_mask, w)
    st, ed = merged_st_ed_logits(w, vq, v2, sq, s2, video_mask, cross=False)
    qv = F.linear(vq, w["video_query_linear.weight"], w["video_query_linear.bias"])
    qs = F.linear(sq, w["sub_query_linear.weight"], w["sub_query_linear.bias"])
    v_sim = torch.einsum("bd,bld->bl", qv, v2)
    s_sim = torch.einsum("bd,bld->bl", qs, s2)
    return dict(modular_att_scores=att, st_prob=st, ed_prob=ed, similarity_scores=(v_sim + s_sim) / 2,
                video_similarity=v_sim, sub_similarity=s_sim)


# --------------------------------------------------------------------------------------
# training step: XML.forward + losses (model_xml.py:212-251, 588-637)
# --------------------------------------------------------------------------------------
def sampled_negative_scores(scores: Tensor, scores_masked: Tensor, use_hard_negative: bool, hard_pool_size: int):
    """XML.get_neg_scores, model_xml.py:607-625: per row, one negative drawn uniformly from the ranks
    [1, 1 + hard_pool_size) (hard negatives) or [1, N) of the row sorted descending with the positive masked to
    999.  The draw uses torch.randint on the DEFAULT CPU generator exactly like the reference, so that the same
    torch.manual_seed gives the same negatives."""
    bsz = len(scores)
    rows = torch.arange(bsz, device=scores.device)
    order = torch.sort(scores_masked, descending=True, dim=1)[1]
    hi = min(1 + hard_pool_size, bsz) if use_hard_negative else bsz
    pick = torch.randint(1, hi, size=(bsz,)).to(scores.device)
    return scores[rows, order[rows, pick]]


def ranking_loss(cfg, pos: Tensor, neg: Tensor) -> Tensor:
    """XML.get_ranking_loss, model_xml.py:627-637."""
    if cfg["ranking_loss_type"] == "hinge":
        return torch.clamp(cfg["margin"] + neg - pos, min=0).sum() / len(pos)
    if cfg["ranking_loss_type"] == "lse":
        return torch.log1p(torch.exp(neg - pos)).sum() / len(pos)
    raise NotImplementedError("Only support 'hinge' and 'lse'")


def video_level_loss(cfg, q2c: Tensor):
    """XML.get_video_level_loss, model_xml.py:588-605 -> (loss_neg_ctx, loss_neg_q)."""
    bsz = len(q2c)
    diag = torch.arange(bsz, device=q2c.device)
    pos = q2c[diag, diag]
    masked = q2c.detach().clone()
    masked[diag, diag] = 999
    neg_ctx = sampled_negative_scores(q2c, masked, cfg["use_hard_negative"], cfg["hard_pool_size"])
    neg_q = sampled_negative_scores(q2c.transpose(0, 1), masked.transpose(0, 1), cfg["use_hard_negative"],
                                    cfg["hard_pool_size"])
    return ranking_loss(cfg, pos, neg_ctx), ranking_loss(cfg, pos, neg_q)


def train_forward(cfg, w: Weights, query_feat, query_mask, video_feat, video_mask, sub_feat, sub_mask,
                  st_ed_indices):
    """XML.forward, model_xml.py:212-251 (eval-mode arithmetic: dropout = identity).  Differentiable w.r.t. the
    tensors in `w`.  Returns (loss, {loss_st_ed, loss_neg_ctx, loss_neg_q, loss_overall: float})."""
    v1, v2, s1, s2 = encode_context(cfg, w, video_feat, video_mask, sub_feat, sub_mask)
    q2c, st, ed = pred_from_raw_query(cfg, w, query_feat, query_mask, v1, v2, video_mask, s1, s2, sub_mask,
                                      cross=False)
    loss_st_ed = 0
    if cfg["lw_st_ed"] != 0:
        loss_st_ed = F.cross_entropy(st, st_ed_indices[:, 0]) + F.cross_entropy(ed, st_ed_indices[:, 1])
    loss_neg_ctx = loss_neg_q = 0
    if cfg["lw_neg_ctx"] != 0 or cfg["lw_neg_q"] != 0:
        loss_neg_ctx, loss_neg_q = video_level_loss(cfg, q2c)
    loss_st_ed = cfg["lw_st_ed"] * loss_st_ed
    loss_neg_ctx = cfg["lw_neg_ctx"] * loss_neg_ctx
    loss_neg_q = cfg["lw_neg_q"] * loss_neg_q
    loss = loss_st_ed + loss_neg_ctx + loss_neg_q
    val = lambda t: float(t.detach()) if torch.is_tensor(t) else float(t)  # noqa: E731
    return loss, {"loss_st_ed": val(loss_st_ed), "loss_neg_ctx": val(loss_neg_ctx),
                  "loss_neg_q": val(loss_neg_q), "loss_overall": val(loss)}


# --------------------------------------------------------------------------------------
# optimizer: BertAdam.step (optimization.py:273-338) and the warm-up schedules (:31-170)
# --------------------------------------------------------------------------------------
def lr_multiplier(schedule: str, warmup: float, t_total: float, step: int) -> float:
    """_LRSchedule.get_lr, optimization.py:51-66, for ConstantLR / WarmupConstantSchedule / WarmupLinearSchedule /
    WarmupCosineSchedule (cycles = 0.5)."""
    if t_total < 0 or schedule in (None, "none"):
        return 1.0
    warmup = max(warmup, 0.0)
    progress = float(step) / t_total
    if progress < warmup:
        return progress / warmup
    if schedule == "warmup_constant":
        return 1.0
    if schedule == "warmup_linear":
        return max((progress - 1.0) / (warmup - 1.0), 0.0)
    if schedule == "warmup_cosine":
        return 0.5 * (1.0 + math.cos(math.pi * (progress - warmup) / (1 - warmup)))
    raise ValueError(schedule)


def bert_adam_step(param: Tensor, grad: Tensor, state: dict, *, lr: float, weight_decay: float, schedule: str,
                   warmup: float, t_total: float, b1: float, b2: float, e: float, max_grad_norm: float) -> None:
    """One BertAdam update of one parameter, in place (param, grad, state = {step, next_m, next_v}).
    optimization.py:296-330: clip_grad_norm_(p, max_grad_norm) rescales the stored gradient by
    max_norm / (||g||_2 + 1e-6) when that is < 1; moments without bias correction; weight decay added to the
    update (decoupled); lr scaled by the schedule at the parameter's own step count."""
    if not state:
        state.update(step=0, next_m=torch.zeros_like(param), next_v=torch.zeros_like(param))
    if max_grad_norm > 0:
        coef = max_grad_norm / (float(torch.linalg.vector_norm(grad.double())) + 1e-6)
        if coef < 1:
            grad.mul_(coef)
    state["next_m"].mul_(b1).add_(grad, alpha=1 - b1)
    state["next_v"].mul_(b2).addcmul_(grad, grad, value=1 - b2)
    update = state["next_m"] / (state["next_v"].sqrt() + e)
    if weight_decay > 0.0:
        update = update + weight_decay * param
    param.sub_(lr * lr_multiplier(schedule, warmup, t_total, state["step"]) * update)
    state["step"] += 1


# --------------------------------------------------------------------------------------
# driver, tensor section (inference.py)
# --------------------------------------------------------------------------------------
def cat_padded(tensors):
    """cat_tensor inside compute_context_info, inference.py:71-87: zero-pad each batch's output to the
    global max length and stack along dim 0."""
    if len(tensors) == 0:
        return None
    width = max(t.shape[1] for t in tensors)
    total = sum(t.shape[0] for t in tensors)
    out = tensors[0].new_zeros((total, width) + tuple(tensors[0].shape[2:]))
    row = 0
    for t in tensors:
        out[row:row + t.shape[0], :t.shape[1]] = t
        row += t.shape[0]
    return out


def context_info(cfg, w: Weights, batches):
    """compute_context_info, inference.py:32-97.  `batches` yields dicts with video_feat/video_mask/
    sub_feat/sub_mask (already padded per batch like start_end_collate does)."""
    acc = {k: [] for k in ("video_feat1", "video_feat2", "video_mask", "sub_feat1", "sub_feat2", "sub_mask")}
    for b in batches:
        v1, v2, s1, s2 = encode_context(cfg, w, b.get("video_feat"), b.get("video_mask"),
                                        b.get("sub_feat"), b.get("sub_mask"))
        if "video" in cfg["ctx_mode"]:
            acc["video_feat1"].append(v1), acc["video_feat2"].append(v2), acc["video_mask"].append(b["video_mask"])
        if "sub" in cfg["ctx_mode"]:
            acc["sub_feat1"].append(s1), acc["sub_feat2"].append(s2), acc["sub_mask"].append(b["sub_mask"])
    return {k: cat_padded(v) for k, v in acc.items()}


def band_mask(length: int, min_l: int, max_l: int) -> np.ndarray:
    """generate_min_max_length_mask, inference.py:170-192: 1 where min_l <= n - m < max_l."""
    ones = np.ones((length, length), dtype=np.float32)
    return np.triu(ones, k=min_l) * (1 - np.triu(ones, k=max_l))


def stable_desc_order(scores: Tensor) -> Tensor:
    """Canonical ranking used for every parity comparison: score descending, index ascending among
    exact ties (the reference's torch.sort/topk tie order is unspecified)."""
    return torch.sort(scores, dim=-1, descending=True, stable=True)[1]


def query_batch_tensor_section(cfg, w: Weights, ctx: dict, query_feat, query_mask, *, q2c_alpha: float,
                               max_n_videos: int, max_before_nms: int, min_pred_l: int, max_pred_l: int,
                               canonical_ties: bool = True, external_topk=None):
    """Tensor section of compute_query2ctx_info for ONE query batch, inference.py:308-386.

    Returns dict with q2c_scores (Q,Nv) after exp, st/ed probs (Q,Nv,L), the top-`max_n_videos` video
    (meta) indices + scores, and the top-`max_before_nms` flat span indices + scores.
    With canonical_ties=True the two rankings use (score desc, index asc); with False they use
    torch.topk / unstable torch.sort exactly as the reference does (used for timing)."""
    q2c, st, ed = pred_from_raw_query(cfg, w, query_feat, query_mask, ctx["video_feat1"], ctx["video_feat2"],
                                      ctx["video_mask"], ctx["sub_feat1"], ctx["sub_feat2"], ctx["sub_mask"],
                                      cross=True)
    q2c = torch.exp(q2c_alpha * q2c)  # inference.py:317
    st = torch.softmax(st, dim=-1)  # inference.py:321-322
    ed = torch.softmax(ed, dim=-1)
    if external_topk is not None:  # inference.py:349-355 (external VR results)
        top_idx, top_raw = external_topk
        top_sc = torch.exp(q2c_alpha * top_raw)
    elif canonical_ties:
        order = stable_desc_order(q2c)[:, :max_n_videos]
        top_idx, top_sc = order, torch.gather(q2c, 1, order)
    else:
        top_sc, top_idx = torch.topk(q2c, max_n_videos, dim=1, largest=True)  # inference.py:347-348
    rows = torch.arange(len(st), device=st.device).unsqueeze(1)
    st_sel, ed_sel = st[rows, top_idx], ed[rows, top_idx]  # inference.py:365-367
    span = torch.einsum("qvm,qv,qvn->qvmn", st_sel, top_sc, ed_sel)  # inference.py:370
    span = span * torch.from_numpy(band_mask(span.shape[-1], min_pred_l, max_pred_l)).to(span.device)  # :371-374
    flat = span.reshape(len(span), -1)
    if canonical_ties:
        order = stable_desc_order(flat)[:, :max_before_nms]
        flat_idx, flat_sc = order, torch.gather(flat, 1, order)
    else:
        s, i = torch.sort(flat, dim=1, descending=True)  # inference.py:380-381
        flat_idx, flat_sc = i[:, :max_before_nms], s[:, :max_before_nms]
    return dict(q2c=q2c, st_prob=st, ed_prob=ed, top_video_idx=top_idx, top_video_score=top_sc,
                span_flat_idx=flat_idx, span_score=flat_sc)


# --------------------------------------------------------------------------------------
# driver, host section (inference.py:391-445) and SVMR / NMS helpers
# --------------------------------------------------------------------------------------
def decode_vcmr(span_flat_idx: np.ndarray, span_score: np.ndarray, top_video_idx: np.ndarray,
                meta_to_video_idx: np.ndarray, max_n_videos: int, max_ctx_l: int, clip_length: float):
    """inference.py:419-442: flat index -> (rank, st_idx, ed_idx) with shape (max_n_videos, max_ctx_l,
    max_ctx_l); st_sec = st*clip; ed_sec = ed*clip + clip.  Returns (Nq, K, 4) float64 rows
    [video_idx, st_sec, ed_sec, score] exactly as float() boxing in the reference gives them."""
    rank, st_i, ed_i = np.unravel_index(span_flat_idx, (max_n_videos, max_ctx_l, max_ctx_l))
    meta = np.take_along_axis(top_video_idx, rank, axis=1)
    out = np.empty(span_flat_idx.shape + (4,), dtype=np.float64)
    out[..., 0] = meta_to_video_idx[meta]
    out[..., 1] = (st_i.astype(np.float32) * np.float32(clip_length)).astype(np.float64)
    out[..., 2] = (ed_i.astype(np.float32) * np.float32(clip_length) + np.float32(clip_length)).astype(np.float64)
    out[..., 3] = span_score.astype(np.float64)
    return out


def svmr_from_probs(st_prob: np.ndarray, ed_prob: np.ndarray, clip_length: float, min_pred_l: int,
                    max_pred_l: int, top_n: int):
    """get_svmr_res_from_st_ed_probs, inference.py:215-232 + tensor_utils.py:115-141.
    Outer product x band mask, per-query full argsort (ascending) reversed, so exact ties come out with
    the LARGER flat index first (stable ascending argsort reversed).  ed index += 1, then x clip_length.
    Returns (Nq, top_n, 3) float32 rows [st_sec, ed_sec, score]."""
    prod = np.einsum("bm,bn->bmn", st_prob, ed_prob) * band_mask(st_prob.shape[1], min_pred_l, max_pred_l)[None]
    n, length = st_prob.shape
    out = np.zeros((n, top_n, 3), dtype=np.float32)
    for i in range(n):
        order = np.argsort(prod[i], axis=None, kind="stable")[::-1][:top_n]
        rows, cols = np.unravel_index(order, (length, length))
        out[i, :, 0] = rows * clip_length
        out[i, :, 1] = (cols + 1) * clip_length
        out[i, :, 2] = prod[i][rows, cols]
    return out


def temporal_iou(a, b) -> float:
    """compute_temporal_iou, utils/temporal_nms.py:17-22: intersection over convex hull."""
    inter = max(0, min(a[1], b[1]) - max(a[0], b[0]))
    hull = max(a[1], b[1]) - min(a[0], b[0])
    return 0 if hull == 0 else 1.0 * inter / hull


def temporal_nms(preds, thd: float, max_after_nms: int = 100):
    """temporal_non_maximum_suppression, utils/temporal_nms.py:25-74.  preds: [[st, ed, score], ...].
    Greedy: sort by score desc (python stable sort), keep the head, drop every later item with
    IoU > thd, repeat; the last remaining item is appended if there is room."""
    if len(preds) == 1:
        return preds
    todo = sorted(preds, key=lambda p: p[2], reverse=True)
    kept = []
    while len(todo) > 1 and len(kept) < max_after_nms:
        head = todo[0]
        todo = [head] + [p for p in todo[1:] if not temporal_iou(head, p) > thd]
        kept.append(todo.pop(0))
    if len(kept) < max_after_nms and len(todo) >= 1:
        kept.append(todo.pop(0))
    return [[p[0], p[1], p[2]] for p in kept]


def vcmr_nms(preds, thd: float, max_before_nms: int, max_after_nms: int):
    """filter_vcmr_by_nms, baselines/clip_alignment_with_language/inference.py:189-225.
    preds: [[video_idx, st, ed, score], ...] ranked.  Group the first max_before_nms by video (dict
    insertion order), NMS per group, stable re-sort by score desc, keep max_after_nms."""
    groups = {}
    for p in preds[:max_before_nms]:
        groups.setdefault(p[0], []).append(list(p[1:]))
    merged = []
    for vid, g in groups.items():
        for p in temporal_nms(g, thd):
            merged.append([vid] + p)
    return sorted(merged, key=lambda p: p[3], reverse=True)[:max_after_nms]


# --------------------------------------------------------------------------------------
# weights for timing-only runs (bench.py --impl reference): same shapes / init scheme as
# XML.__init__ + reset_parameters (model_xml.py:52-201); values are NOT the reference's RNG stream.
# --------------------------------------------------------------------------------------
def init_weights(cfg, seed: int = 2018) -> Weights:
    g = torch.Generator().manual_seed(seed)
    h, std = cfg["hidden_size"], cfg.get("initializer_range", 0.02)
    w: Weights = {}

    def linear(name, out_dim, in_dim, bias=True):
        w[name + ".weight"] = torch.randn(out_dim, in_dim, generator=g) * std
        if bias:
            w[name + ".bias"] = torch.zeros(out_dim)

    def ln(name, dim):
        w[name + ".weight"], w[name + ".bias"] = torch.ones(dim), torch.zeros(dim)

    def block(name):
        for p in ("query", "key", "value"):
            linear(name + ".self." + p, h, h)
        linear(name + ".output.dense", h, h)
        ln(name + ".output.LayerNorm", h)

    def proj(name, in_dim):
        ln(name + ".LayerNorm", in_dim)
        linear(name + ".net.1", h, in_dim)

    def conv(name):
        k = cfg.get("conv_kernel_size", 5)
        w[name + ".weight"] = (torch.rand(1, 1, k, generator=g) * 2 - 1) / math.sqrt(k)

    for name, n_pos in (("query_pos_embed", cfg["max_desc_l"]), ("ctx_pos_embed", cfg["max_ctx_l"])):
        w[name + ".position_embeddings.weight"] = torch.randn(n_pos, h, generator=g) * std
        ln(name + ".LayerNorm", h)
    proj("query_input_proj", cfg["query_input_size"])
    block("query_encoder")
    streams = [s for s in ("video", "sub") if s in cfg["ctx_mode"]]
    for s in streams:
        proj(s + "_input_proj", cfg["visual_input_size"] if s == "video" else cfg["sub_input_size"])
        block(s + "_encoder1"), block(s + "_encoder2")
        if cfg["cross_att"]:
            for p in ("query", "key", "value"):
                linear(s + "_cross_att." + p, h, h)
            ln(s + "_cross_layernorm", h)
        else:
            block(s + "_encoder3")
        linear(s + "_query_linear", h, h)
        if not cfg["merge_two_stream"]:
            conv(s + "_st_predictor"), conv(s + "_ed_predictor")
    linear("modular_vector_mapping", len(streams), h, bias=False)
    if cfg["merge_two_stream"]:
        conv("merged_st_predictor"), conv("merged_ed_predictor")
    return w
