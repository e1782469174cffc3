"""Thin tensor-level wrappers over the C ABI (include/xmlb200.h).  PyTorch is used only to own device
memory and streams; every computation below is one or more hand-written sm_100a kernels.  All inputs must be
CUDA fp32 (or int32/uint8 where stated) tensors; there is no CPU path."""
import os
import weakref

import torch

from . import _lib, autograd


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.XmlbError("%s must be a CUDA tensor: the xmlb200 kernels have no CPU fallback" % name)
    if t.dtype != torch.float32:
        raise _lib.XmlbError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _i32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.XmlbError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t if t.is_contiguous() else t.contiguous()


def _u8(t, name):
    if t is None:
        return None
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype != torch.uint8 or not t.is_cuda:
        raise _lib.XmlbError("%s must be a CUDA uint8/bool tensor" % name)
    return t if t.is_contiguous() else t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _sched_ws(device):
    """One int of device scratch for the dynamic tile scheduler of a tensor-core kernel (zeroed by the kernel's
    entry point, on the launch stream)."""
    return torch.empty(1, device=device, dtype=torch.int32)


def add_layernorm(x, gamma, beta, add=None, add_rows=None, eps=1e-5):
    """LayerNorm(x + add[row % add_rows]) over the last dim."""
    if autograd.recording(x, gamma, beta, add):
        return autograd.add_layernorm(add_layernorm, x, gamma, beta, add, add_rows, eps)
    x = _f32(x, "x")
    dim = x.shape[-1]
    rows = x.numel() // dim
    add = _f32(add, "add")
    if add is not None and add_rows is None:
        add_rows = add.numel() // dim
    out = torch.empty_like(x)
    rc = _lib.lib().xmlb_add_layernorm(_p(x), _p(add), add_rows or 0, _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")),
                                       _p(out), rows, dim, eps, _stream())
    _lib.check(rc, "xmlb_add_layernorm")
    return out


# Precision of the Linear layers, passed explicitly by every caller (no global state): "f16x3" / "bf16x3" = tcgen05
# tensor cores with hi/lo-split operands (3 MMAs per product) and K-chunked fp32 accumulation (the TMEM accumulator
# truncates; chunks of 128 are summed on the tensor core and added round-to-nearest in registers, csrc/linear_tc.cu),
# "f32" = exact-fp32 SIMT kernel (FMA).  Small problems always use the latter.
PRECISIONS = ("f16x3", "bf16x3", "f32")
DEFAULT_PRECISION = "f16x3"
_TC_MIN_ROWS = 256
_WEIGHT_SPLITS = {}


def invalidate_weight_caches():
    """Drop the cached (hi, lo) weight splits: called after kernels that write parameters in place without going
    through torch (BertAdam.step), which leaves tensor._version unchanged."""
    _WEIGHT_SPLITS.clear()
    _BIAS_CATS.clear()


def _weight_split(weight, bf16):
    """(hi, lo) of a weight matrix, cached per (tensor, version) for module parameters (not for the transposed
    temporaries of the backward pass)."""
    if not isinstance(weight, torch.nn.Parameter):
        return split_rows(weight.detach(), bf16=bf16)
    key = (weight.data_ptr(), tuple(weight.shape), bf16)
    hit = _WEIGHT_SPLITS.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version:
        return hit[2]
    if len(_WEIGHT_SPLITS) > 256:
        _WEIGHT_SPLITS.clear()
    pair = split_rows(weight.detach(), bf16=bf16)
    _WEIGHT_SPLITS[key] = (weakref.ref(weight), weight._version, pair)
    return pair


def _weight_split_cat(weights, bf16):
    """(hi, lo) of several weight matrices stacked along the output dimension (the fused QKV / KV projections),
    cached like _weight_split; -> ((hi, lo), [parameters])."""
    key = tuple(w.data_ptr() for w in weights) + (bf16,)
    hit = _WEIGHT_SPLITS.get(key)
    if hit is not None and all(r() is w for r, w in zip(hit[0], weights)) and hit[1] == [w._version for w in weights]:
        return hit[2]
    if len(_WEIGHT_SPLITS) > 256:
        _WEIGHT_SPLITS.clear()
    pair = split_rows(torch.cat([w.detach() for w in weights], dim=0), bf16=bf16)
    _WEIGHT_SPLITS[key] = ([weakref.ref(w) for w in weights], [w._version for w in weights], pair)
    return pair


_BIAS_CATS = {}


def _bias_cat(biases):
    key = tuple(b.data_ptr() for b in biases)
    hit = _BIAS_CATS.get(key)
    if hit is not None and all(r() is b for r, b in zip(hit[0], biases)) and hit[1] == [b._version for b in biases]:
        return hit[2]
    if len(_BIAS_CATS) > 256:
        _BIAS_CATS.clear()
    cat = torch.cat([b.detach() for b in biases]).contiguous()
    _BIAS_CATS[key] = ([weakref.ref(b) for b in biases], [b._version for b in biases], cat)
    return cat


def linear(x, weight, bias=None, residual=None, relu=False, precision=DEFAULT_PRECISION):
    assert precision in PRECISIONS, precision
    if autograd.recording(x, weight, bias, residual):
        return autograd.linear(linear, x, weight, bias, residual, relu, precision)
    x = _f32(x, "x")
    weight = _f32(weight, "weight")
    out_dim, in_dim = weight.shape
    assert x.shape[-1] == in_dim, (x.shape, weight.shape)
    rows = x.numel() // in_dim
    out = torch.empty(x.shape[:-1] + (out_dim,), device=x.device, dtype=torch.float32)
    residual = _f32(residual, "residual")
    if residual is not None:
        assert residual.shape == out.shape
    if precision != "f32" and rows >= _TC_MIN_ROWS:
        bf16 = precision == "bf16x3"
        w_hi, w_lo = _weight_split(weight, bf16)
        x_hi, x_lo = split_rows(x, kpad=w_hi.shape[1], bf16=bf16)
        rc = _lib.lib().xmlb_linear_tc(_p(x_hi), _p(x_lo), _p(w_hi), _p(w_lo), _p(_f32(bias, "bias")), _p(residual),
                                       _p(out), _p(_sched_ws(x.device)), rows, out_dim, w_hi.shape[1], int(relu),
                                       int(bf16), _stream())
        _lib.check(rc, "xmlb_linear_tc")
        return out
    rc = _lib.lib().xmlb_linear(_p(x), _p(weight), _p(_f32(bias, "bias")), _p(residual), _p(out), rows, out_dim,
                                in_dim, int(relu), _stream())
    _lib.check(rc, "xmlb_linear")
    return out


def pad64(n):
    return (n + 63) // 64 * 64


def add_layernorm_split(x, gamma, beta, add=None, add_rows=None, add_index=None, eps=1e-5, bf16=False, want_f32=True):
    """LayerNorm(x + add[...]) -> (fp32 rows or None, (hi, lo) int16 (rows, pad64(dim))): the normalised rows and
    their 16-bit split in one pass (the split is the operand of the tensor-core Linear that follows)."""
    x = _f32(x, "x")
    dim = x.shape[-1]
    rows = x.numel() // dim
    add = _f32(add, "add")
    if add is not None and add_rows is None:
        add_rows = add.numel() // dim
    kpad = pad64(dim)
    out = torch.empty_like(x) if want_f32 else None
    hi = torch.empty(rows, kpad, device=x.device, dtype=torch.int16)
    lo = torch.empty_like(hi)
    rc = _lib.lib().xmlb_add_layernorm_split(_p(x), _p(add), add_rows or 0, _p(_i32(add_index, "add_index")),
                                             _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")), _p(out), _p(hi), _p(lo),
                                             kpad, int(bf16), rows, dim, eps, _stream())
    _lib.check(rc, "xmlb_add_layernorm_split")
    return out, (hi, lo)


def linear_tc_ex(x16, w16, bias=None, residual=None, relu=False, bf16=False, want_f32=True, out16_cols=0, vt_col0=None,
                 vt_seq=0, k_chunk=0):
    """Tensor-core Linear on pre-split operands with fused output formats (xmlb_linear_tc_ex).  x16 = (hi, lo) of
    (rows, kpad); w16 = (hi, lo) of (out_dim, kpad).  -> (out fp32 (rows, out_dim) or None,
    (hi, lo) (rows, out16_cols) split of the leading output columns or None,
    (hi, lo) (rows / vt_seq * (out_dim - vt_col0), pad64(vt_seq)) transposed split of the trailing columns or None)."""
    rows, kpad = x16[0].shape
    out_dim = w16[0].shape[0]
    assert w16[0].shape[1] == kpad
    dev = x16[0].device
    out = torch.empty(rows, out_dim, device=dev, dtype=torch.float32) if want_f32 else None
    o16 = (None, None)
    if out16_cols:
        o16 = (torch.empty(rows, out16_cols, device=dev, dtype=torch.int16),
               torch.empty(rows, out16_cols, device=dev, dtype=torch.int16))
    vt, vt_ld = (None, None), 0
    if vt_col0 is not None:
        assert rows % vt_seq == 0
        vt_ld = pad64(vt_seq)
        n_rows = rows // vt_seq * (out_dim - vt_col0)
        # positions beyond the sequence are read as zeros by the attention kernel: they must be finite
        alloc = torch.empty if vt_ld == vt_seq else torch.zeros
        vt = (alloc(n_rows, vt_ld, device=dev, dtype=torch.int16), alloc(n_rows, vt_ld, device=dev, dtype=torch.int16))
    residual = _f32(residual, "residual")
    rc = _lib.lib().xmlb_linear_tc_ex(_p(x16[0]), _p(x16[1]), _p(w16[0]), _p(w16[1]), _p(_f32(bias, "bias")),
                                      _p(residual), _p(out), _p(o16[0]), _p(o16[1]), out16_cols, out16_cols, _p(vt[0]),
                                      _p(vt[1]), vt_col0 or 0, vt_seq, vt_ld, _p(_sched_ws(dev)), rows, out_dim, kpad,
                                      int(relu), int(bf16), k_chunk, _stream())
    _lib.check(rc, "xmlb_linear_tc_ex")
    return out, (o16 if out16_cols else None), (vt if vt_col0 is not None else None)


def attention_tc(q16, q_col0, k16, k_col0, vt16, mask3, batch, len_q, len_k, hidden, n_heads, bf16=False,
                 want_f32=True, want_split=False):
    """Fused tensor-core attention (xmlb_attention_tc) on the split outputs of linear_tc_ex: q16 / k16 = (hi, lo) of
    (batch * len, ld) with the heads at columns col0 + h * dh; vt16 = (hi, lo) of (batch * hidden, ld_v);
    mask3 (batch, 1 or len_q, len_k) float {0,1}.  -> (fp32 (batch * len_q, hidden) or None, (hi, lo) or None)."""
    mask3 = _f32(mask3, "mask")
    assert mask3.shape[0] == batch and mask3.shape[2] == len_k and mask3.shape[1] in (1, len_q), mask3.shape
    dev = q16[0].device
    out = torch.empty(batch * len_q, hidden, device=dev, dtype=torch.float32) if want_f32 else None
    o16 = (None, None)
    if want_split:
        assert hidden % 64 == 0
        o16 = (torch.empty(batch * len_q, hidden, device=dev, dtype=torch.int16),
               torch.empty(batch * len_q, hidden, device=dev, dtype=torch.int16))
    mq = 0 if mask3.shape[1] == 1 else len_k
    rc = _lib.lib().xmlb_attention_tc(_p(q16[0]), _p(q16[1]), q16[0].shape[1], q_col0, _p(k16[0]), _p(k16[1]),
                                      k16[0].shape[1], k_col0, _p(vt16[0]), _p(vt16[1]), vt16[0].shape[1], _p(mask3),
                                      mask3.shape[1] * len_k, mq, _p(out), _p(o16[0]), _p(o16[1]), hidden, batch, len_q,
                                      len_k, hidden, n_heads, int(bf16), _stream())
    _lib.check(rc, "xmlb_attention_tc")
    return out, (o16 if want_split else None)


def attention_tc_supported(hidden, n_heads, len_k):
    dh = hidden // max(1, n_heads)
    return hidden % n_heads == 0 and dh % 64 == 0 and dh <= 256 and 1 <= len_k <= 256


def attention(q, k, v, mask3, n_heads, max_batch=8192, dropout_p=0.0, seed=0):
    """q (N, Lq, H), k/v (N, Lk, H), mask3 (N, 1 or Lq, Lk) float {0,1} -> (N, Lq, H).
    dropout_p > 0 (train mode): dropout on the attention probabilities with the counter-based mask of `seed`."""
    if autograd.recording(q, k, v):
        return autograd.attention(attention, q, k, v, mask3, n_heads, max_batch, dropout_p, seed)
    q, k, v, mask3 = _f32(q, "q"), _f32(k, "k"), _f32(v, "v"), _f32(mask3, "mask")
    n, lq, hid = q.shape
    lk = k.shape[1]
    assert mask3.shape[0] == n and mask3.shape[2] == lk and mask3.shape[1] in (1, lq), mask3.shape
    out = torch.empty_like(q)
    step = max(1, min(max_batch, 65535 // n_heads, max(1, (1 << 28) // max(1, n_heads * lq * lk))))
    ws = torch.empty(min(n, step) * n_heads * lq * lk, device=q.device, dtype=torch.float32)
    mq = 0 if mask3.shape[1] == 1 else lk
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        if dropout_p > 0:
            rc = _lib.lib().xmlb_attention_train(_p(q[lo:hi]), _p(k[lo:hi]), _p(v[lo:hi]), _p(mask3[lo:hi]),
                                                 mask3.shape[1] * lk, mq, _p(out[lo:hi]), _p(ws), hi - lo, lq, lk,
                                                 hid, n_heads, dropout_p, seed, lo * n_heads * lq * lk, _stream())
        else:
            rc = _lib.lib().xmlb_attention(_p(q[lo:hi]), _p(k[lo:hi]), _p(v[lo:hi]), _p(mask3[lo:hi]),
                                           mask3.shape[1] * lk, mq, _p(out[lo:hi]), _p(ws), hi - lo, lq, lk, hid,
                                           n_heads, _stream())
        _lib.check(rc, "xmlb_attention")
    return out


def _modular_pool_outputs(encoded, mask, w_mod):
    """-> tuple of n_mod pooled (N, H) tensors."""
    encoded, mask, w_mod = _f32(encoded, "encoded"), _f32(mask, "mask"), _f32(w_mod, "w_mod")
    n, length, hid = encoded.shape
    n_mod = w_mod.shape[0]
    out0 = torch.empty(n, hid, device=encoded.device, dtype=torch.float32)
    out1 = torch.empty_like(out0) if n_mod == 2 else None
    rc = _lib.lib().xmlb_modular_pool(_p(encoded), _p(mask), _p(w_mod), _p(out0), _p(out1), n, length, hid, n_mod,
                                      _stream())
    _lib.check(rc, "xmlb_modular_pool")
    return (out0, out1) if n_mod == 2 else (out0,)


def add_layernorm_indexed(x, gamma, beta, add, add_index, eps=1e-5):
    """LayerNorm(x[r] + add[add_index[r]]): the position encoding of packed (ragged) token rows."""
    x, add = _f32(x, "x"), _f32(add, "add")
    dim = x.shape[-1]
    out = torch.empty_like(x)
    rc = _lib.lib().xmlb_add_layernorm_indexed(_p(x), _p(add), _p(_i32(add_index, "add_index")), add.numel() // dim,
                                               _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")), _p(out),
                                               x.numel() // dim, dim, eps, _stream())
    _lib.check(rc, "xmlb_add_layernorm_indexed")
    return out


def attention_ragged(q, k, v, cu_seqlens, max_len, n_heads):
    """Fused self-attention over packed sequences: q / k / v (T, H), sequence s = rows [cu[s], cu[s+1]), <= 32 tokens."""
    q, k, v = _f32(q, "q"), _f32(k, "k"), _f32(v, "v")
    out = torch.empty_like(q)
    rc = _lib.lib().xmlb_attention_ragged(_p(q), _p(k), _p(v), _p(_i32(cu_seqlens, "cu_seqlens")), _p(out),
                                          cu_seqlens.numel() - 1, max_len, q.shape[-1], n_heads, _stream())
    _lib.check(rc, "xmlb_attention_ragged")
    return out


def attention_ragged_qkv(qkv, cu_seqlens, max_len, n_heads):
    """attention_ragged on the (T, 3H) output of a fused Q|K|V projection -> (T, H)."""
    qkv = _f32(qkv, "qkv")
    hid = qkv.shape[-1] // 3
    out = torch.empty(qkv.shape[0], hid, device=qkv.device, dtype=torch.float32)
    rc = _lib.lib().xmlb_attention_ragged_qkv(_p(qkv), _p(_i32(cu_seqlens, "cu_seqlens")), _p(out),
                                              cu_seqlens.numel() - 1, max_len, hid, n_heads, _stream())
    _lib.check(rc, "xmlb_attention_ragged_qkv")
    return out


def modular_pool_ragged(encoded, cu_seqlens, max_len, w_mod):
    """modular_pool on packed token rows -> (video_query, sub_query), each (n_sequences, H)."""
    encoded, w_mod = _f32(encoded, "encoded"), _f32(w_mod, "w_mod")
    n, hid, n_mod = cu_seqlens.numel() - 1, encoded.shape[-1], w_mod.shape[0]
    out0 = torch.empty(n, hid, device=encoded.device, dtype=torch.float32)
    out1 = torch.empty_like(out0) if n_mod == 2 else None
    rc = _lib.lib().xmlb_modular_pool_ragged(_p(encoded), _p(_i32(cu_seqlens, "cu_seqlens")), _p(w_mod), _p(out0),
                                             _p(out1), n, max_len, hid, n_mod, _stream())
    _lib.check(rc, "xmlb_modular_pool_ragged")
    return (out0, out1) if n_mod == 2 else (out0, out0)


def dropout(x, p, seed, index0=0):
    """x * keep / (1 - p) with the counter-based mask keep(seed, index0 + i) -- nn.Dropout in train mode.  The
    backward pass applies the same mask to the gradient (nothing is stored)."""
    if p <= 0:
        return x
    if autograd.recording(x):
        return autograd.Dropout.apply(x, p, seed, index0)
    x = _f32(x, "x")
    out = torch.empty_like(x)
    rc = _lib.lib().xmlb_dropout(_p(x), _p(out), x.numel(), p, seed, index0, _stream())
    _lib.check(rc, "xmlb_dropout")
    return out


def dropout_mask(shape, p, seed, device, index0=0):
    """The scaled keep mask itself: keep / (1 - p)."""
    out = torch.empty(shape, device=device, dtype=torch.float32)
    rc = _lib.lib().xmlb_dropout(None, _p(out), out.numel(), p, seed, index0, _stream())
    _lib.check(rc, "xmlb_dropout")
    return out


def new_seed():
    """Dropout seed drawn from torch's default CPU generator (follows torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def modular_pool(encoded, mask, w_mod):
    """-> (video_query, sub_query); the same vector twice for a single-modality model (model_xml.py:420-423)."""
    if autograd.recording(encoded, w_mod):
        outs = autograd.modular_pool(_modular_pool_outputs, encoded, mask, w_mod)
    else:
        outs = _modular_pool_outputs(encoded, mask, w_mod)
    return (outs[0], outs[1]) if len(outs) == 2 else (outs[0], outs[0])


def l2norm_rows(x, eps=1e-12):
    if autograd.recording(x):
        return autograd.l2norm_rows(l2norm_rows, x, eps)
    x = _f32(x, "x")
    out = torch.empty_like(x)
    dim = x.shape[-1]
    rc = _lib.lib().xmlb_l2norm_rows(_p(x), _p(out), x.numel() // dim, dim, eps, _stream())
    _lib.check(rc, "xmlb_l2norm_rows")
    return out


def softmax_rows(x):
    x = _f32(x, "x")
    out = torch.empty_like(x)
    dim = x.shape[-1]
    rc = _lib.lib().xmlb_softmax_rows(_p(x), _p(out), x.numel() // dim, dim, _stream())
    _lib.check(rc, "xmlb_softmax_rows")
    return out


def vr_scores_f32(q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask):
    """-> q2c (Nq, Nv).  Inputs already L2-normalised; a modality is skipped when None."""
    if autograd.recording(q_video_n, q_sub_n, feat1_video_n, feat1_sub_n):
        return autograd.vr_scores(vr_scores_f32, q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask)
    ref_q = q_video_n if q_video_n is not None else q_sub_n
    ref_c = feat1_video_n if feat1_video_n is not None else feat1_sub_n
    nq, hid = ref_q.shape
    nv, length, _ = ref_c.shape
    out = torch.empty(nq, nv, device=ref_q.device, dtype=torch.float32)
    ws = torch.empty(2, nq, nv, device=ref_q.device, dtype=torch.float32)
    rc = _lib.lib().xmlb_vr_scores_f32(
        _p(_f32(q_video_n, "q_video")), _p(_f32(q_sub_n, "q_sub")), _p(_f32(feat1_video_n, "feat1_video")),
        _p(_f32(feat1_sub_n, "feat1_sub")), _p(_f32(video_mask, "video_mask")), _p(_f32(sub_mask, "sub_mask")),
        _p(out), _p(ws), nq, nv, length, hid, _stream())
    _lib.check(rc, "xmlb_vr_scores_f32")
    return out


def split_rows(x, group_in=1, group_out=1, kpad=None, normalize=False, bf16=False, row_index=None, out=None,
               out_col0=0, hi_err=False, err_out=None):
    """fp32 rows -> 16-bit (hi, lo) halves with x ~= hi + lo, for the split-precision tensor-core kernels.
    x: (..., k) with rows grouped by `group_in`; output (n_groups * group_out, kpad) int16 tensors (raw bits).
    With row_index (int32, one source row per output row; negative = zero row) the rows are gathered instead.
    hi_err=True additionally returns ||x - hi||_2 per output row (error bound of hi-only products)."""
    x = _f32(x, "x")
    k = x.shape[-1]
    rows = x.numel() // k
    kpad = kpad or (k + 63) // 64 * 64
    if row_index is not None:
        row_index = _i32(row_index, "row_index")
        n_groups, group_in, group_out = row_index.numel(), 1, 1
    else:
        assert rows % group_in == 0
        n_groups = rows // group_in
    if out is None:
        hi = torch.empty(n_groups * group_out, kpad, device=x.device, dtype=torch.int16)
        lo = torch.empty_like(hi)
    else:  # write into columns [out_col0, out_col0 + kpad) of preallocated (rows, out_ld) buffers
        hi, lo = out
        assert hi.shape[0] == n_groups * group_out and hi.shape == lo.shape and hi.is_contiguous()
    err = None
    if hi_err:
        err = err_out if err_out is not None else torch.empty(hi.shape[0], device=x.device, dtype=torch.float32)
        assert err.numel() == hi.shape[0] and err.is_contiguous()
    rc = _lib.lib().xmlb_split_rows(_p(x), _p(row_index), n_groups, group_in, group_out, k, kpad, hi.shape[1],
                                    out_col0, int(normalize), int(bf16), _p(hi), _p(lo), _p(err), _stream())
    _lib.check(rc, "xmlb_split_rows")
    return (hi, lo, err) if hi_err else (hi, lo)


def split_rows_t(x, bf16=False):
    """x (rows, cols) fp32 -> (hi, lo) int16 (cols, pad64(rows)) of the TRANSPOSE (xmlb_split_rows_t)."""
    x = _f32(x, "x")
    rows, cols = x.shape
    rpad = pad64(rows)
    hi = torch.empty(cols, rpad, device=x.device, dtype=torch.int16)
    lo = torch.empty_like(hi)
    rc = _lib.lib().xmlb_split_rows_t(_p(x), rows, cols, rpad, int(bf16), _p(hi), _p(lo), _stream())
    _lib.check(rc, "xmlb_split_rows_t")
    return hi, lo


def gather_rows16(src, row_index):
    """src = (hi, lo) int16 (rows, kpad) -> (hi, lo) gathered at row_index (int32; negative = row left unwritten)."""
    row_index = _i32(row_index, "row_index")
    kpad = src[0].shape[1]
    hi = torch.empty(row_index.numel(), kpad, device=src[0].device, dtype=torch.int16)
    lo = torch.empty_like(hi)
    rc = _lib.lib().xmlb_gather_rows16(_p(src[0]), _p(src[1]), _p(row_index), row_index.numel(), kpad, _p(hi), _p(lo),
                                       _stream())
    _lib.check(rc, "xmlb_gather_rows16")
    return hi, lo


def mask_bits(mask, lp):
    """(Nv, L) float {0,1} -> (Nv, lp // 32) int32 bit masks."""
    mask = _f32(mask, "mask")
    nv, length = mask.shape
    bits = torch.empty(nv, lp // 32, device=mask.device, dtype=torch.int32)
    rc = _lib.lib().xmlb_mask_bits(_p(mask), nv, length, lp, _p(bits), _stream())
    _lib.check(rc, "xmlb_mask_bits")
    return bits


def vr_scores_tc(q_a, c_a, bits_a, n_videos, lp, q_b=None, c_b=None, bits_b=None, bf16=False, max_ctas=0):
    """tcgen05 video-level scores on prepared operands: q_x = (hi, lo) of (Nq, kpad); c_x = (hi, lo) of
    (Nv * lp, kpad); bits_x (Nv, lp/32).  -> q2c (Nq, Nv) fp32."""
    nq, kpad = q_a[0].shape
    assert c_a[0].shape == (n_videos * lp, kpad), (c_a[0].shape, n_videos, lp, kpad)
    out = torch.empty(nq, n_videos, device=q_a[0].device, dtype=torch.float32)
    qb = q_b if q_b is not None else (None, None)
    cb = c_b if c_b is not None else (None, None)
    rc = _lib.lib().xmlb_vr_scores_tc(_p(q_a[0]), _p(q_a[1]), _p(qb[0]), _p(qb[1]), _p(c_a[0]), _p(c_a[1]), _p(cb[0]),
                                      _p(cb[1]), _p(bits_a), _p(bits_b), _p(out), _p(_sched_ws(out.device)), nq,
                                      n_videos, lp, kpad, int(bf16), max_ctas, _stream())
    _lib.check(rc, "xmlb_vr_scores_tc")
    return out


def vr_scores_tc_packed(q_a, c_a, packing, n_videos, q_b=None, c_b=None, bf16=False, max_ctas=0, ordinal=False,
                        hi_only=False, out=None, m_tiles=None):
    """tcgen05 video-level scores on the packed (valid clips only) corpus; `packing` = engine.CorpusPacking.
    The kernel writes the scores in packed-ordinal order (adjacent columns per tile).  ordinal=True returns that
    layout, (Nq, Nv) with column o <-> video packing.order_full[o]; ordinal=False re-orders to video ids."""
    nq, kpad = q_a[0].shape
    restricted = m_tiles is not None  # (list, count) device int32 tensors: recompute only these 128-query tiles
    if out is None:
        out = torch.empty(nq, n_videos, device=q_a[0].device, dtype=torch.float32)
    qb = q_b if q_b is not None else (None, None)
    cb = c_b if c_b is not None else (None, None)
    if hi_only and not restricted and max_ctas == 0 and FILTER_ON_CTA_PAIRS:
        # the filter pass of the two-pass search: CTA pairs (tcgen05 cta_group::2) halve the operand traffic per SM
        rc = _lib.lib().xmlb_vr_filter_pair(_p(q_a[0]), _p(qb[0]), _p(c_a[0]), _p(cb[0]), _p(packing.tile_meta),
                                            _p(packing.tile_starts), _p(out), _p(_sched_ws(out.device)), nq, n_videos,
                                            packing.n_rows, packing.n_tiles, kpad, int(bf16), _stream())
        _lib.check(rc, "xmlb_vr_filter_pair")
        return _vr_packed_finish(out, packing, n_videos, ordinal)
    rc = _lib.lib().xmlb_vr_scores_tc_packed(
        _p(q_a[0]), _p(q_a[1]), _p(qb[0]), _p(qb[1]), _p(c_a[0]), _p(c_a[1]), _p(cb[0]), _p(cb[1]),
        _p(packing.tile_meta), _p(packing.tile_starts), _p(out), _p(_sched_ws(out.device)), nq, n_videos,
        packing.n_rows, packing.n_tiles, int(hi_only), _p(m_tiles[0]) if restricted else None,
        _p(m_tiles[1]) if restricted else None, kpad, int(bf16), max_ctas, _stream())
    _lib.check(rc, "xmlb_vr_scores_tc_packed")
    return _vr_packed_finish(out, packing, n_videos, ordinal, fill=not restricted)


FILTER_ON_CTA_PAIRS = True  # False: run the hi-only pass on the single-CTA kernel (kept for A/B measurements)


def _vr_packed_finish(out, packing, n_videos, ordinal, fill=True):
    if fill and packing.n_packed < n_videos:
        out[:, packing.n_packed:] = -1e10  # videos without a valid clip (reference: every clip masked -> -1e10)
    if ordinal:
        return out
    by_id = torch.empty_like(out)
    by_id[:, packing.order_full.long()] = out
    return by_id


# How the grouped kernels (exact re-scoring, span similarity) get the query rows of a video's inverted list:
#   "warps": four extra warps of the kernel copy them, 16 bytes at a time (cp.async), from the query array into the
#            swizzled shared-memory tile -- no gathered copy in HBM, no work for unlisted table slots (default);
#   "copy" : xmlb_gather_rows16 first materialises the rows in list order, the kernel box-loads them with TMA;
#   "tma"  : the kernel's producer warp issues TMA tile::gather4 (256 B per instruction: 2-3x slower, kept for that
#            measurement).
# All three give the same bits (tests/test_gpu_ops.py).  Measured on B200 in the full search (bench.py, 21.8K videos x
# 10K queries): "warps" vs "copy" -- re-scoring 3.6 vs 5.4 ms, span similarity 6.5 vs 7.9 ms at N=1; 0.44 vs 1.32 ms and
# 1.08 vs 1.39 ms per rank at N=8 (the copy also touches every empty slot of the (Nq, max_candidates) table).
GATHER = os.environ.get("XMLB_GATHER", "warps")


class Candidates:
    """Output of select_candidates: (R, max_cand) column / id / value tables + the overflow flags."""

    def __init__(self, col, ids, val, flag_ws, n_groups, n_rows):
        self.col, self.ids, self.val = col, ids, val
        self.n_flagged = flag_ws[0:1]                          # number of flagged 128-row groups
        self.flagged_groups = flag_ws[1 + n_groups:1 + 2 * n_groups]
        self.row_flags = flag_ws[1 + 2 * n_groups:1 + 2 * n_groups + n_rows]


def select_candidates(approx, k, row_err_a, row_err_b, err_scale, err_const, max_cand, ids=None, rows_per_group=128,
                      row_kth=None):
    """Columns of `approx` (R, C) that can be among the exact top-k of their row given the error bound
    eps[r] = err_scale * (row_err_a[r] + row_err_b[r]) + err_const  (see xmlb_select_candidates).
    row_kth (R,): the k-th largest approximate score of each row over a larger column set (all shards)."""
    approx = _f32(approx, "approx")
    n_rows, n_cols = approx.shape
    dev = approx.device
    n_groups = (n_rows + rows_per_group - 1) // rows_per_group
    col = torch.empty(n_rows, max_cand, device=dev, dtype=torch.int32)
    cid = torch.empty_like(col)
    val = torch.empty(n_rows, max_cand, device=dev, dtype=torch.float32)
    flag_ws = torch.empty(1 + 2 * n_groups + n_rows, device=dev, dtype=torch.int32)
    rc = _lib.lib().xmlb_select_candidates(_p(approx), _p(_i32(ids, "ids")), n_rows, n_cols, k,
                                           _p(_f32(row_err_a, "row_err_a")), _p(_f32(row_err_b, "row_err_b")),
                                           err_scale, err_const, _p(_f32(row_kth, "row_kth")), max_cand,
                                           rows_per_group, _p(col), _p(cid), _p(val), _p(flag_ws), _stream())
    _lib.check(rc, "xmlb_select_candidates")
    return Candidates(col, cid, val, flag_ws, n_groups, n_rows)


def vr_rescore_tc(q_fp32_a, c_a, packing, cand, kpad, q_fp32_b=None, c_b=None, bf16=False, q_split_a=None,
                  q_split_b=None):
    """Overwrites cand.val with the exact split-precision scores of the candidate (query, packed video) pairs.
    q_fp32_* (Nq, H) pooled query vectors (normalised here, gathered per video list); c_* packed corpus (hi, lo),
    row-major (rows, kpad) or k-blocked (kblock_rows: contiguous TMA boxes).
    q_split_* = (hi, lo) of the normalised queries when the caller already has them (then only copied)."""
    lists = build_pair_lists(cand.col, packing.n_packed, chunk=128)
    dev = cand.col.device
    units = torch.empty(lists.max_chunks * 4 + 4, device=dev, dtype=torch.int32)
    rc = _lib.lib().xmlb_build_span_units(_p(lists.vid_ptr), _p(lists.chunk_ptr), packing.n_packed, lists.chunk,
                                          _p(units), _stream())
    _lib.check(rc, "xmlb_build_span_units")
    def halves(q_fp32, q_split):  # (hi, lo) of the normalised queries
        q = q_split if q_split is not None else split_rows(q_fp32, kpad=kpad, normalize=True, bf16=bf16)
        return q if GATHER != "copy" else gather_rows16(q, lists.entry_q)
    qa = halves(q_fp32_a, q_split_a)
    qb = halves(q_fp32_b, q_split_b) if (q_fp32_b is not None or q_split_b is not None) else (None, None)
    cb = c_b if c_b is not None else (None, None)
    kblocked = c_a[0].dim() == 3  # (kpad / 32, packed rows, 32), see kblock_rows
    assert c_b is None or (c_b[0].dim() == 3) == kblocked
    rc = _lib.lib().xmlb_vr_rescore_tc_kb(_p(qa[0]), _p(qa[1]), _p(qb[0]), _p(qb[1]), _p(c_a[0]), _p(c_a[1]), _p(cb[0]),
                                          _p(cb[1]), int(kblocked), _p(packing.row_start), _p(units),
                                          lists.chunk_ptr[packing.n_packed:].data_ptr(), lists.max_chunks,
                                          _p(lists.entry_out), _p(lists.entry_q) if GATHER != "copy" else None,
                                          int(GATHER == "warps"), qa[0].shape[0], _p(cand.val), _p(_sched_ws(dev)),
                                          lists.entry_q.numel(), packing.n_rows, packing.max_len, kpad, int(bf16),
                                          _stream())
    _lib.check(rc, "xmlb_vr_rescore_tc")
    return cand.val


class PairLists:
    """Per-video inverted lists of (query, output row) built on device by xmlb_build_pair_lists."""

    def __init__(self, vid_ptr, chunk_ptr, entry_q, entry_out, max_chunks, n_rows, chunk=32):
        self.vid_ptr, self.chunk_ptr, self.entry_q, self.entry_out = vid_ptr, chunk_ptr, entry_q, entry_out
        self.max_chunks, self.n_rows, self.chunk = max_chunks, n_rows, chunk


def build_pair_lists(top_idx, n_videos, vid_lo=0, slot_valid=None, chunk=32):
    """top_idx (Nq, n_slots) int32 global video ids -> PairLists over the videos [vid_lo, vid_lo + n_videos)."""
    top_idx = _i32(top_idx, "top_idx")
    nq, n_slots = top_idx.shape
    dev = top_idx.device
    n_pairs = nq * n_slots
    # entries of pairs that are not listed (other shards' videos) stay -1: "no row" for the gather kernels
    ints = torch.full((4 * n_videos + 2 + 2 * n_pairs,), -1, device=dev, dtype=torch.int32)
    counts, cursor = ints[:n_videos], ints[n_videos:2 * n_videos]
    vid_ptr = ints[2 * n_videos:3 * n_videos + 1]
    chunk_ptr = ints[3 * n_videos + 1:4 * n_videos + 2]
    entry_q = ints[4 * n_videos + 2:4 * n_videos + 2 + n_pairs]
    entry_out = ints[4 * n_videos + 2 + n_pairs:]
    rc = _lib.lib().xmlb_build_pair_lists(_p(top_idx), _p(_u8(slot_valid, "slot_valid")), nq, n_slots, vid_lo,
                                          n_videos, chunk, _p(counts), _p(cursor), _p(vid_ptr), _p(chunk_ptr),
                                          _p(entry_q), _p(entry_out), _stream())
    _lib.check(rc, "xmlb_build_pair_lists")
    max_chunks = (n_pairs + chunk - 1) // chunk + min(n_videos, n_pairs)
    lists = PairLists(vid_ptr, chunk_ptr, entry_q, entry_out, max_chunks, n_pairs, chunk)
    lists.complete = False  # set by the caller when it knows that EVERY pair is listed (no -1 / foreign ids)
    return lists


def span_clip_rows(mask, ksize):
    """Per video: how many leading clip rows the similarity kernel has to load -- every unmasked clip plus the
    ksize // 2 neighbours its ConvSE taps read (int32, on the mask's device)."""
    length = mask.shape[1]
    last = ((mask != 0) * torch.arange(1, length + 1, device=mask.device)).amax(1)
    return (last + ksize // 2).clamp_(max=length).to(torch.int32).contiguous()


def kblock_rows(t, block=32, swizzle=False):
    """(rows, k) 16-bit operand -> the K-BLOCKED layout (k / block, rows, block) the similarity kernel streams with
    contiguous TMA boxes (one run of box_rows * 64 bytes per k-step).  swizzle=True additionally permutes the four
    16-byte pieces of every 64-byte row the way SWIZZLE_64B places them in shared memory (piece c of row r at
    position c ^ ((r >> 1) & 3)) and returns a 4-D (k / block, rows, 4, 8) tensor: the shared-memory IMAGE of the
    tiles, fetched by plain bulk copies (tiles must start at rows that are multiples of 8)."""
    rows, k = t.shape
    assert k % block == 0
    out = t.view(rows, k // block, block).permute(1, 0, 2).contiguous()
    if not swizzle:
        return out
    assert block == 32
    s = (torch.arange(rows, device=t.device) >> 1) & 3
    src = torch.arange(4, device=t.device)[None, :] ^ s[:, None]               # (rows, 4): piece held by position j
    img = torch.empty(k // block, rows, 4, 8, device=t.device, dtype=t.dtype)
    step = max(1, (1 << 28) // max(1, rows * block))                            # bounded temporaries
    for lo in range(0, k // block, step):
        part = out[lo:lo + step].view(-1, rows, 4, 8)
        img[lo:lo + step] = torch.gather(part, 2, src[None, :, :, None].expand(part.shape[0], rows, 4, 8))
    return img


def span_probs_tc(f2cat, q_cat, lists, mask, w_st, w_ed, ctx_len, softmax=True, bf16=False, out_rows=None,
                  clip_rows=None):
    """tcgen05 similarity curves + ConvSE + mask (+ softmax) for listed (query, video) pairs of the merged model.
    f2cat = (hi, lo) of [feat2_video | feat2_sub] (Nv * L, kcat); q_cat fp32 (Nq, kcat) = [q'_video | q'_sub] with
    each half zero-padded to kcat / 2; lists built with chunk in {32, 64, 128}.  -> st, ed of shape (rows, L).
    clip_rows = span_clip_rows(mask, ksize) (gather-warps mode): only those rows of every video are read."""
    # (kcat / 32, Nv * L, 32) = k-blocked, (kcat / 32, Nv * L, 4, 8) = k-blocked shared-memory image, see kblock_rows
    kblocked = {2: 0, 3: 1, 4: 2}[f2cat[0].dim()]
    if kblocked:
        assert kblocked == 1 or (GATHER == "warps" and ctx_len % 8 == 0)
        n_videos, kcat = f2cat[0].shape[1] // ctx_len, f2cat[0].shape[0] * 32
    else:
        n_videos, kcat = f2cat[0].shape[0] // ctx_len, f2cat[0].shape[1]
    assert q_cat.shape[1] == kcat and lists.chunk in (32, 64, 128)
    dev = q_cat.device
    units = torch.empty(lists.max_chunks * 4 + 4, device=dev, dtype=torch.int32)
    if GATHER != "warps":
        clip_rows = None
    rc = _lib.lib().xmlb_build_span_units_rows(_p(lists.vid_ptr), _p(lists.chunk_ptr), n_videos, lists.chunk,
                                               _p(clip_rows), _p(units), _stream())
    _lib.check(rc, "xmlb_build_span_units")
    # queries: split once, then copied into list order (or gathered by the kernel's producer, see TMA_GATHER)
    qg = split_rows(q_cat, kpad=kcat, bf16=bf16)
    if GATHER == "copy":
        qg = gather_rows16(qg, lists.entry_q)
    rows = lists.n_rows if out_rows is None else out_rows
    # rows that no list entry covers (videos of other shards) must read as zeros
    alloc = torch.empty if (out_rows is None and getattr(lists, "complete", False)) else torch.zeros
    st = alloc(rows, ctx_len, device=dev, dtype=torch.float32)
    ed = alloc(rows, ctx_len, device=dev, dtype=torch.float32)
    w_st, w_ed = _f32(w_st.reshape(-1), "w_st"), _f32(w_ed.reshape(-1), "w_ed")
    rc = _lib.lib().xmlb_span_probs_tc_clipped(
        _p(f2cat[0]), _p(f2cat[1]), _p(qg[0]), _p(qg[1]), _p(_f32(mask, "mask")), _p(w_st), _p(w_ed), w_st.numel(),
        int(softmax), n_videos, ctx_len, kcat, lists.entry_q.numel(), lists.chunk, _p(units),
        lists.chunk_ptr[n_videos:].data_ptr(), lists.max_chunks, _p(lists.entry_out),
        _p(lists.entry_q) if GATHER != "copy" else None, int(GATHER == "warps"), qg[0].shape[0],
        int(clip_rows is not None), int(kblocked), _p(st), _p(ed), _p(_sched_ws(dev)), int(bf16), _stream())
    _lib.check(rc, "xmlb_span_probs_tc")
    return st, ed


def diagonal_pair_lists(n, device):
    """Lists for the in-batch (cross=False) case: query i <-> video i."""
    ar = torch.arange(n + 1, device=device, dtype=torch.int32)
    lists = PairLists(ar, ar, ar[:n], ar[:n], n, n)
    lists.diagonal = True  # autograd.t_span_logits differentiates this case (the training step) only
    return lists


def span_logits(q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b=None, feat2_b=None, mask_b=None, w_st_b=None, w_ed_b=None,
                merged=False, softmax=False, lists=None, out_rows=None):
    """Dense (lists=None): -> st, ed of shape (Nq, Nv, L).  List mode: -> (out_rows, L) each."""
    if autograd.recording(q_a, feat2_a, w_st_a, w_ed_a, q_b, feat2_b, w_st_b, w_ed_b):
        return autograd.span_logits(span_logits, q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b, feat2_b, mask_b, w_st_b,
                                    w_ed_b, merged, softmax, lists, out_rows)
    q_a, feat2_a = _f32(q_a, "q_a"), _f32(feat2_a, "feat2_a")
    nq, hid = q_a.shape
    nv, length, _ = feat2_a.shape
    dev = q_a.device
    w = [_f32(t.reshape(-1), "conv taps") if t is not None else None for t in (w_st_a, w_ed_a, w_st_b, w_ed_b)]
    ksize = w[0].numel()
    if lists is None:
        st = torch.empty(nq, nv, length, device=dev, dtype=torch.float32)
        ed = torch.empty_like(st)
        lp = (None, None, None, None, 0)
    else:
        rows = lists.n_rows if out_rows is None else out_rows
        # rows that no list entry covers (videos outside this shard) must read as zeros
        st = torch.zeros(rows, length, device=dev, dtype=torch.float32)
        ed = torch.zeros_like(st)
        lp = (_p(lists.chunk_ptr), _p(lists.vid_ptr), _p(lists.entry_q), _p(lists.entry_out), lists.max_chunks)
    rc = _lib.lib().xmlb_span_logits(
        _p(q_a), _p(_f32(q_b, "q_b")), _p(feat2_a), _p(_f32(feat2_b, "feat2_b")), _p(_f32(mask_a, "mask_a")),
        _p(_f32(mask_b, "mask_b")), _p(w[0]), _p(w[1]), _p(w[2]), _p(w[3]), ksize, int(merged), int(softmax), nq, nv,
        length, hid, lp[0], lp[1], lp[2], lp[3], lp[4], _p(st), _p(ed), _stream())
    _lib.check(rc, "xmlb_span_logits")
    return st, ed


class PeerSpec:
    """Destination of a ranking kernel's output inside the symmetric workspaces of the ranks of a sharded search
    (see xmlb_topk_rows_ex): per-rank base addresses of the idx / val regions (None = that half is not written),
    mode 1 = "to the owner of each row", 2 = "to all ranks"; `per` = rows owned per rank."""

    def __init__(self, idx_ptrs, val_ptrs, mode, per, self_rank):
        import ctypes
        ref = idx_ptrs if idx_ptrs is not None else val_ptrs
        self.world, self.mode, self.per, self.self_rank = len(ref), mode, per, self_rank
        arr = ctypes.c_longlong * self.world
        self.idx = arr(*idx_ptrs) if idx_ptrs is not None else None
        self.val = arr(*val_ptrs) if val_ptrs is not None else None

    def args(self):
        return (self.idx, self.val, self.world, self.mode, self.per, self.self_rank)


_NO_PEER = (None, None, 0, 0, 0, 0)


def topk_rows(values, k, alpha=1.0, apply_exp=False, ids=None, tie_desc=False, row_flags=None, out=None, n_rows=None,
              segments=None, missing_neg=False, out_slice=None, pad=None, peer=None, want_idx=True):
    """-> (idx int32 (R, k), val fp32 (R, k)) ranked by (value desc, id asc).  ids: None (column index), (R, C)
    per-row ids or a 1-D (C,) table shared by all rows.  row_flags (R,) int32 + out=(idx, val): only the flagged
    rows are recomputed, in place.
    Sharded-search extras (xmlb_topk_rows_ex): segments=(seg_k, seg_stride) reads each row as per-rank lists laid out
    [source rank][row][seg_k] (`values` / `ids` are then the flat workspace regions and n_rows is given explicitly);
    missing_neg: negative ids mark absent entries; out_slice=(first, count): only those ranks of each list are
    written; pad=(to, idx, val): filler entries up to `to` per row; peer=PeerSpec: the lists are stored on other ranks
    instead of (mode 1 / 2) the local output, and None is returned."""
    values = _f32(values, "values")
    if segments is None:
        n_rows, n_cols = values.shape
        seg_k, seg_stride = 0, 0
    else:
        seg_k, seg_stride, n_cols = segments
    first, count = out_slice if out_slice is not None else (0, k)
    pad_to, pad_idx, pad_val = pad if pad is not None else (0, 0, 0.0)
    pitch = max(count, pad_to)
    out_idx = out_val = None
    if peer is None:
        if out is None:
            assert row_flags is None
            out_idx = torch.empty(n_rows, pitch, device=values.device, dtype=torch.int32) if want_idx else None
            out_val = torch.empty(n_rows, pitch, device=values.device, dtype=torch.float32)
        else:
            out_idx, out_val = out
            assert out_idx.shape == (n_rows, pitch) and out_idx.is_contiguous() and out_val.is_contiguous()
    ids = _i32(ids, "ids")
    shared = ids is not None and ids.dim() == 1 and segments is None
    assert ids is None or segments is not None or ids.numel() == (n_cols if shared else n_rows * n_cols)
    rc = _lib.lib().xmlb_topk_rows_ex(_p(values), _p(ids), int(shared), n_rows, n_cols, k, alpha, int(apply_exp),
                                      int(tie_desc), _p(row_flags), seg_k, seg_stride, int(missing_neg), first, count,
                                      pad_to, pad_idx, pad_val, _p(out_idx), _p(out_val),
                                      *(peer.args() if peer is not None else _NO_PEER), _stream())
    _lib.check(rc, "xmlb_topk_rows")
    return None if peer is not None else (out_idx, out_val)


def span_topk(st_prob, ed_prob, video_score, min_l, max_l, k, slot_valid=None, tie_desc=False, zero_fill=True,
              peer=None):
    """st/ed (Nq, n_slots, L) probabilities, video_score (Nq, n_slots) or None -> flat idx int32, score (Nq, k)
    (peer=PeerSpec: stored on the owner ranks instead; returns None)."""
    st_prob, ed_prob = _f32(st_prob, "st_prob"), _f32(ed_prob, "ed_prob")
    nq, n_slots, length = st_prob.shape
    out_idx = out_val = None
    if peer is None:
        out_idx = torch.empty(nq, k, device=st_prob.device, dtype=torch.int32)
        out_val = torch.empty(nq, k, device=st_prob.device, dtype=torch.float32)
    rc = _lib.lib().xmlb_span_topk_ex(_p(st_prob), _p(ed_prob), _p(_f32(video_score, "video_score")),
                                      _p(_u8(slot_valid, "slot_valid")), nq, n_slots, length, min_l, max_l, k,
                                      int(tie_desc), int(zero_fill), _p(out_idx), _p(out_val),
                                      *(peer.args() if peer is not None else _NO_PEER), _stream())
    _lib.check(rc, "xmlb_span_topk")
    return None if peer is not None else (out_idx, out_val)


def span_zero_fill(flat_idx, score, total_cells, tie_desc=False, peer=None):
    assert flat_idx.dtype == torch.int32 and flat_idx.is_contiguous() and score.is_contiguous()
    rc = _lib.lib().xmlb_span_zero_fill_ex(_p(flat_idx), _p(score), flat_idx.shape[0], flat_idx.shape[1], total_cells,
                                           int(tie_desc), *(peer.args() if peer is not None else _NO_PEER), _stream())
    _lib.check(rc, "xmlb_span_zero_fill")
    return flat_idx, score


def peer_copy(src, peer_ptrs):
    """Copies the bytes of a contiguous local tensor to the given address on every rank (xmlb_peer_copy)."""
    import ctypes
    assert src.is_contiguous()
    arr = (ctypes.c_longlong * len(peer_ptrs))(*peer_ptrs)
    rc = _lib.lib().xmlb_peer_copy(_p(src), src.numel() * src.element_size(), arr, len(peer_ptrs), _stream())
    _lib.check(rc, "xmlb_peer_copy")


def temporal_nms(st, ed, score, iou_thd, max_out, video_idx=None, n_valid=None, max_per_group=100):
    """Ranked lists (Nq, n) -> (kept idx int32 (Nq, max_out) into the input lists, count int32 (Nq,))."""
    st, ed, score = _f32(st, "st"), _f32(ed, "ed"), _f32(score, "score")
    nq, n_in = st.shape
    out_idx = torch.empty(nq, max_out, device=st.device, dtype=torch.int32)
    out_cnt = torch.empty(nq, device=st.device, dtype=torch.int32)
    rc = _lib.lib().xmlb_temporal_nms(_p(_i32(video_idx, "video_idx")), _p(st), _p(ed), _p(score),
                                      _p(_i32(n_valid, "n_valid")), nq, n_in, float(iou_thd), max_per_group, max_out,
                                      _p(out_idx), _p(out_cnt), _stream())
    _lib.check(rc, "xmlb_temporal_nms")
    return out_idx, out_cnt
