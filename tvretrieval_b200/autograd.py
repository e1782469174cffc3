"""Autograd support for the kernel ops: what the training step `XML.forward` (reference model_xml.py:212-251,
driven by train.py:77-85) needs on top of the inference kernels.

Forward values always come from the CUDA kernels (include/xmlb200.h).  Backward:
  * Linear layers -- the bulk of the FLOPs -- use the same kernels: dX = dY . W and dW = dY^T . X are two more calls
    of xmlb_linear / xmlb_linear_tc (operands transposed and split by xmlb_split_rows_t);
  * every other op (LayerNorm, attention core, modular pooling, L2 normalisation, masked-max video scores,
    similarity + ConvSE, ReLU masks, bias / LayerNorm / position-table / ConvSE reductions) has a hand-written
    backward kernel (csrc/backward.cu; BACKWARD = "kernels", the default).  BACKWARD = "twins" differentiates them by
    recomputing the equivalent torch expression under autograd instead -- kept only so that the tests can compare
    the two op by op.
Inference never enters this module: the dispatch in ops.py only routes here when autograd is recording.
"""
import torch
import torch.nn.functional as F

BACKWARD = "kernels"   # or "twins" (torch recomputation; test reference)
MASK_FILL = -1e10      # reference model_xml.py:640-641
ATT_MASK_FILL = -10000.0  # reference model_components.py:277


def recording(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


# ------------------------------------------------------------------------------------------------ generic op
class _Slot:
    def __init__(self, i):
        self.i = i


def _split(args, kwargs):
    """Replace tensors by slots so that autograd.Function sees them as positional tensor inputs."""
    tensors = []

    def enc(v):
        if isinstance(v, torch.Tensor):
            tensors.append(v)
            return _Slot(len(tensors) - 1)
        return v
    return ([enc(a) for a in args], {k: enc(v) for k, v in kwargs.items()}), tensors


def _fill(spec, tensors):
    dec = lambda v: tensors[v.i] if isinstance(v, _Slot) else v  # noqa: E731
    return [dec(a) for a in spec[0]], {k: dec(v) for k, v in spec[1].items()}


class _RecomputeOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, torch_fn, spec, *tensors):
        ctx.torch_fn, ctx.spec = torch_fn, spec
        ctx.save_for_backward(*tensors)
        args, kw = _fill(spec, tensors)
        with torch.no_grad():
            out = kernel_fn(*args, **kw)
        return out

    @staticmethod
    def backward(ctx, *grads):
        needs = ctx.needs_input_grad[3:]
        leaves = [t.detach().requires_grad_(bool(n)) for t, n in zip(ctx.saved_tensors, needs)]
        args, kw = _fill(ctx.spec, leaves)
        with torch.enable_grad():
            out = ctx.torch_fn(*args, **kw)
        outs = out if isinstance(out, tuple) else (out,)
        pairs = [(o, g) for o, g in zip(outs, grads) if g is not None and o.requires_grad]
        wanted = [leaf for leaf, n in zip(leaves, needs) if n]
        got = torch.autograd.grad([o for o, _ in pairs], wanted, [g for _, g in pairs], allow_unused=True) \
            if pairs and wanted else [None] * len(wanted)
        it = iter(got)
        return (None, None, None) + tuple(next(it) if n else None for n in needs)


def recompute_op(kernel_fn, torch_fn, args, kwargs):
    spec, tensors = _split(args, kwargs)
    return _RecomputeOp.apply(kernel_fn, torch_fn, spec, *tensors)


# ------------------------------------------------------------------------------------------------ linear
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, x, weight, bias, residual, relu, precision):
        with torch.no_grad():
            out = kernel_fn(x, weight, bias, residual, relu, precision)
        ctx.kernel_fn, ctx.relu = kernel_fn, relu
        ctx.precision = precision  # the backward GEMMs run in the precision of the forward one
        ctx.save_for_backward(x, weight, out if relu else None)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out = ctx.saved_tensors
        lin = ctx.kernel_fn
        g = g.contiguous()
        if ctx.relu:
            g = _relu_backward(g, out) if BACKWARD == "kernels" else g * (out > 0)
        out_dim, in_dim = weight.shape
        g2, x2 = g.reshape(-1, out_dim), x.reshape(-1, in_dim)
        dx = dw = db = dres = None
        prec = ctx.precision
        # output gradients are ~1e-6 and below: out of fp16's range (its halves would underflow), so on the tensor cores
        # the backward GEMMs split both operands into bf16 halves (2 x 8 mantissa bits, fp32's exponent range; a
        # kind::f16 MMA cannot mix fp16 and bf16 operands)
        if prec == "f16x3":
            prec = "bf16x3"
        from . import ops
        with torch.no_grad():
            if ctx.needs_input_grad[1]:
                dx = lin(g2, weight.t().contiguous(), None, None, False, prec).view(x.shape)      # dY . W
            if ctx.needs_input_grad[2]:
                if prec != "f32" and out_dim >= 16 and g2.shape[0] >= 256 and g2.is_cuda:
                    # dY^T . X on the tensor cores: both operands transposed AND split in one pass each
                    bf16 = prec == "bf16x3"
                    dw, _, _ = ops.linear_tc_ex(ops.split_rows_t(g2, bf16=bf16), ops.split_rows_t(x2, bf16=bf16),
                                                bf16=bf16)
                else:
                    dw = lin(g2.t().contiguous(), x2.t().contiguous(), None, None, False, prec)   # dY^T . X
            if ctx.has_bias and ctx.needs_input_grad[3]:
                db = _sum_rows(g2).view(-1) if BACKWARD == "kernels" else g2.sum(0)
            if ctx.has_res and ctx.needs_input_grad[4]:
                dres = g
        return None, dx, dw, db, dres, None, None


def linear(kernel_fn, x, weight, bias, residual, relu, precision):
    return _Linear.apply(kernel_fn, x, weight, bias, residual, relu, precision)


class Dropout(torch.autograd.Function):
    """Forward and backward are the same kernel call: the mask is a function of (seed, element index)."""

    @staticmethod
    def forward(ctx, x, p, seed, index0):
        from . import ops
        ctx.args = (p, seed, index0)
        with torch.no_grad():
            return ops.dropout(x, p, seed, index0)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        with torch.no_grad():
            return ops.dropout(g.contiguous(), *ctx.args), None, None, None


# ------------------------------------------------------------------------------------------------ backward kernels
def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _sum_rows(t, group_rows=1):
    """t (n_groups * group_rows, dim) -> (group_rows, dim): out[p] = sum_g t[g * group_rows + p], fixed order."""
    from . import _lib
    t = t.contiguous()
    dim = t.shape[-1]
    n_groups = t.numel() // dim // group_rows
    out = torch.empty(group_rows, dim, device=t.device, dtype=torch.float32)
    ws = None
    if n_groups > 64:
        ws = torch.empty(2 * ((n_groups + 63) // 64) * group_rows * dim, device=t.device, dtype=torch.float32)
    _lib.check(_lib.lib().xmlb_sum_rows(_p(t), n_groups, group_rows, dim, _p(out), _p(ws), _stream()), "xmlb_sum_rows")
    return out


def _relu_backward(g, out):
    from . import _lib
    dx = torch.empty_like(g)
    _lib.check(_lib.lib().xmlb_relu_backward(_p(g), _p(out.contiguous()), g.numel(), _p(dx), _stream()),
               "xmlb_relu_backward")
    return dx


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, x, gamma, beta, add, add_rows, eps):
        out = kernel_fn(x, gamma, beta, add=add, add_rows=add_rows, eps=eps)
        ctx.save_for_backward(x, gamma, add)
        ctx.add_rows, ctx.eps = add_rows, eps
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        x, gamma, add = ctx.saved_tensors
        x, g = x.contiguous(), g.contiguous()
        dim = x.shape[-1]
        rows = x.numel() // dim
        add_rows = ctx.add_rows
        if add is not None and add_rows is None:
            add_rows = add.numel() // dim
        dx, dyx = torch.empty_like(x), torch.empty_like(x)
        rc = _lib.lib().xmlb_layernorm_backward(_p(x), _p(add.contiguous()) if add is not None else None, add_rows or 0,
                                                _p(gamma.contiguous()), _p(g), rows, dim, ctx.eps, _p(dx), _p(dyx),
                                                _stream())
        _lib.check(rc, "xmlb_layernorm_backward")
        dgamma = _sum_rows(dyx.view(rows, dim)).view_as(gamma) if ctx.needs_input_grad[2] else None
        dbeta = _sum_rows(g.view(rows, dim)).view_as(gamma) if ctx.needs_input_grad[3] else None
        dadd = None
        if add is not None and ctx.needs_input_grad[4]:
            if add_rows == rows and add.numel() == x.numel():
                dadd = dx.view_as(add)
            else:  # a (position) table whose first add_rows rows were broadcast over the batch
                dadd = torch.zeros_like(add)
                dadd.view(-1, dim)[:add_rows] = _sum_rows(dx.view(rows, dim), group_rows=add_rows)
        return None, (dx if ctx.needs_input_grad[1] else None), dgamma, dbeta, dadd, None, None


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, x, eps):
        ctx.save_for_backward(x)
        ctx.eps = eps
        return kernel_fn(x, eps=eps)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        (x,) = ctx.saved_tensors
        x, g = x.contiguous(), g.contiguous()
        dim = x.shape[-1]
        dx = torch.empty_like(x)
        rc = _lib.lib().xmlb_l2norm_backward(_p(x), _p(g), x.numel() // dim, dim, ctx.eps, _p(dx), _stream())
        _lib.check(rc, "xmlb_l2norm_backward")
        return None, dx, None


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, q, k, v, mask3, n_heads, max_batch, dropout_p, seed):
        ctx.save_for_backward(q, k, v, mask3)
        ctx.args = (n_heads, max_batch, dropout_p, seed)
        return kernel_fn(q, k, v, mask3, n_heads, max_batch=max_batch, dropout_p=dropout_p, seed=seed)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        q, k, v, mask3 = (t.contiguous() for t in ctx.saved_tensors)
        n_heads, max_batch, dropout_p, seed = ctx.args
        g = g.contiguous()
        n, lq, hid = q.shape
        lk = k.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        # same batch pieces (and dropout counter offsets) as the forward launch loop in ops.attention
        step = max(1, min(max_batch, 65535 // n_heads, max(1, (1 << 28) // max(1, n_heads * lq * lk))))
        ws = torch.empty(2, min(n, step) * n_heads * lq * lk, device=q.device, dtype=torch.float32)
        mq = 0 if mask3.shape[1] == 1 else lk
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            rc = _lib.lib().xmlb_attention_backward(
                _p(q[lo:hi]), _p(k[lo:hi]), _p(v[lo:hi]), _p(mask3[lo:hi]), mask3.shape[1] * lk, mq, _p(g[lo:hi]),
                _p(dq[lo:hi]), _p(dk[lo:hi]), _p(dv[lo:hi]), _p(ws[0]), _p(ws[1]), hi - lo, lq, lk, hid, n_heads,
                dropout_p, seed, lo * n_heads * lq * lk, _stream())
            _lib.check(rc, "xmlb_attention_backward")
        return None, dq, dk, dv, None, None, None, None, None


class _ModularPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, encoded, mask, w_mod):
        ctx.save_for_backward(encoded, mask, w_mod)
        return kernel_fn(encoded, mask, w_mod)

    @staticmethod
    def backward(ctx, *grads):
        from . import _lib
        encoded, mask, w_mod = (t.contiguous() for t in ctx.saved_tensors)
        n, length, hid = encoded.shape
        n_mod = w_mod.shape[0]
        gs = [(g.contiguous() if g is not None else torch.zeros(n, hid, device=encoded.device)) for g in grads[:n_mod]]
        d_enc = torch.empty_like(encoded)
        dlogit = torch.empty(n, length, 2, device=encoded.device, dtype=torch.float32)
        rc = _lib.lib().xmlb_modular_pool_backward(_p(encoded), _p(mask), _p(w_mod), _p(gs[0]),
                                                   _p(gs[1]) if n_mod == 2 else None, n, length, hid, n_mod, _p(d_enc),
                                                   _p(dlogit), _stream())
        _lib.check(rc, "xmlb_modular_pool_backward")
        dw = None
        if ctx.needs_input_grad[3]:
            dw = torch.empty_like(w_mod)
            rc = _lib.lib().xmlb_modular_mapping_grad(_p(dlogit), _p(encoded), n * length, hid, n_mod, _p(dw), _stream())
            _lib.check(rc, "xmlb_modular_mapping_grad")
        return None, d_enc, None, dw


class _VrScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kernel_fn, q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask):
        ctx.save_for_backward(q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask)
        return kernel_fn(q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        qv, qs, cv, cs, mv, ms = ctx.saved_tensors
        g = g.contiguous()
        n_mod = (qv is not None) + (qs is not None)
        grads = {}
        for name, q, c, m in (("v", qv, cv, mv), ("s", qs, cs, ms)):
            if q is None:
                grads[name] = (None, None)
                continue
            q, c, m = q.contiguous(), c.contiguous(), m.contiguous()
            nq, hid = q.shape
            nv, length, _ = c.shape
            dq, dc = torch.empty_like(q), torch.empty_like(c)
            arg = torch.empty(nq * nv, device=q.device, dtype=torch.int32)
            rc = _lib.lib().xmlb_vr_scores_backward(_p(q), _p(c), _p(m), _p(g), 1.0 / n_mod, nq, nv, length, hid, _p(arg),
                                                    _p(dq), _p(dc), _stream())
            _lib.check(rc, "xmlb_vr_scores_backward")
            grads[name] = (dq, dc)
        return None, grads["v"][0], grads["s"][0], grads["v"][1], grads["s"][1], None, None


class _SpanLogitsDiag(torch.autograd.Function):
    """ops.span_logits in diagonal list mode without softmax (XML.forward, cross=False)."""

    @staticmethod
    def forward(ctx, kernel_fn, lists, merged, q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b, feat2_b, mask_b, w_st_b, w_ed_b):
        ctx.save_for_backward(q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b, feat2_b, mask_b, w_st_b, w_ed_b)
        ctx.merged = merged
        return kernel_fn(q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b=q_b, feat2_b=feat2_b, mask_b=mask_b, w_st_b=w_st_b,
                         w_ed_b=w_ed_b, merged=merged, softmax=False, lists=lists)

    @staticmethod
    def backward(ctx, dst, ded):
        from . import _lib
        q_a, f2_a, mask_a, w_st_a, w_ed_a, q_b, f2_b, mask_b, w_st_b, w_ed_b = ctx.saved_tensors
        c = lambda t: None if t is None else t.contiguous()  # noqa: E731
        q_a, f2_a, mask_a, q_b, f2_b, mask_b = map(c, (q_a, f2_a, mask_a, q_b, f2_b, mask_b))
        n, length, hid = f2_a.shape
        dev = q_a.device
        zeros = lambda: torch.zeros(n, length, device=dev)  # noqa: E731
        dst = dst.contiguous() if dst is not None else zeros()
        ded = ded.contiguous() if ded is not None else zeros()
        ksize = w_st_a.numel()
        flat = lambda t: None if t is None else t.reshape(-1).contiguous()  # noqa: E731
        dq_a, df_a = torch.empty_like(q_a), torch.empty_like(f2_a)
        dq_b = torch.empty_like(q_b) if q_b is not None else None
        df_b = torch.empty_like(f2_b) if f2_b is not None else None
        dw = torch.empty(n, 4 * 31, device=dev, dtype=torch.float32)
        ws = [flat(w_st_a), flat(w_ed_a), flat(w_st_b), flat(w_ed_b)]
        rc = _lib.lib().xmlb_span_logits_diag_backward(
            _p(q_a), _p(q_b), _p(f2_a), _p(f2_b), _p(mask_a), _p(mask_b), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]),
            ksize, int(ctx.merged), n, length, hid, _p(dst), _p(ded), _p(dq_a), _p(dq_b), _p(df_a), _p(df_b), _p(dw),
            _stream())
        _lib.check(rc, "xmlb_span_logits_diag_backward")
        dws = _sum_rows(dw).view(2, 2, 31)[:, :, :ksize]  # [stream][start | end][tap]
        gw = lambda w, x, which: None if w is None else dws[x, which].reshape(w.shape).clone()  # noqa: E731
        return (None, None, None, dq_a, df_a, None, gw(w_st_a, 0, 0), gw(w_ed_a, 0, 1), dq_b, df_b, None,
                gw(w_st_b, 1, 0), gw(w_ed_b, 1, 1))


# ---- dispatch used by ops.py (kernels by default; torch twins for the op-by-op comparison tests) ----------------
def add_layernorm(kernel_fn, x, gamma, beta, add, add_rows, eps):
    if BACKWARD == "kernels":
        return _AddLayerNorm.apply(kernel_fn, x, gamma, beta, add, add_rows, eps)
    return recompute_op(kernel_fn, t_add_layernorm, (x, gamma, beta), dict(add=add, add_rows=add_rows, eps=eps))


def l2norm_rows(kernel_fn, x, eps):
    if BACKWARD == "kernels":
        return _L2Norm.apply(kernel_fn, x, eps)
    return recompute_op(kernel_fn, t_l2norm_rows, (x,), dict(eps=eps))


def attention(kernel_fn, q, k, v, mask3, n_heads, max_batch, dropout_p, seed):
    if BACKWARD == "kernels":
        return _Attention.apply(kernel_fn, q, k, v, mask3, n_heads, max_batch, dropout_p, seed)
    return recompute_op(kernel_fn, t_attention, (q, k, v, mask3, n_heads),
                        dict(max_batch=max_batch, dropout_p=dropout_p, seed=seed))


def modular_pool(kernel_fn, encoded, mask, w_mod):
    if BACKWARD == "kernels":
        return _ModularPool.apply(kernel_fn, encoded, mask, w_mod)
    return recompute_op(kernel_fn, t_modular_pool, (encoded, mask, w_mod), {})


def vr_scores(kernel_fn, q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask):
    if BACKWARD == "kernels":
        return _VrScores.apply(kernel_fn, q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask)
    return recompute_op(kernel_fn, t_vr_scores, (q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask),
                        {})


def span_logits(kernel_fn, q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b, feat2_b, mask_b, w_st_b, w_ed_b, merged, softmax,
                lists, out_rows):
    diagonal = lists is not None and getattr(lists, "diagonal", False)
    if BACKWARD == "kernels" and diagonal and not softmax and out_rows is None:
        return _SpanLogitsDiag.apply(kernel_fn, lists, merged, q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b, feat2_b, mask_b,
                                     w_st_b, w_ed_b)
    return recompute_op(kernel_fn, t_span_logits, (q_a, feat2_a, mask_a, w_st_a, w_ed_a),
                        dict(q_b=q_b, feat2_b=feat2_b, mask_b=mask_b, w_st_b=w_st_b, w_ed_b=w_ed_b, merged=merged,
                             softmax=softmax, lists=lists, out_rows=out_rows))


# ------------------------------------------------------------------------------------------------ torch twins
def t_add_layernorm(x, gamma, beta, add=None, add_rows=None, eps=1e-5):
    dim = x.shape[-1]
    if add is not None:
        if add_rows is None:
            add_rows = add.numel() // dim
        x = (x.reshape(-1, add_rows, dim) + add.reshape(-1, dim)[:add_rows]).reshape(x.shape)
    return F.layer_norm(x, (dim,), gamma, beta, eps)


def t_attention(q, k, v, mask3, n_heads, max_batch=8192, dropout_p=0.0, seed=0):
    """reference model_components.py:277-303 after the projections."""
    n, lq, hid = q.shape
    dh = hid // n_heads
    heads = lambda t: t.view(n, -1, n_heads, dh).permute(0, 2, 1, 3)  # noqa: E731
    scores = torch.matmul(heads(q), heads(k).transpose(-1, -2)) / (dh ** 0.5)
    scores = scores + (1.0 - mask3.unsqueeze(1)) * ATT_MASK_FILL
    probs = torch.softmax(scores, dim=-1)
    if dropout_p > 0:  # the very mask the forward kernel applied
        from . import ops
        probs = probs * ops.dropout_mask(probs.shape, dropout_p, seed, probs.device)
    return torch.matmul(probs, heads(v)).permute(0, 2, 1, 3).reshape(n, lq, hid)


def t_mask_logits(x, m):
    return x * m + (1 - m) * MASK_FILL


def t_modular_pool(encoded, mask, w_mod):
    att = torch.softmax(t_mask_logits(F.linear(encoded, w_mod), mask.unsqueeze(2)), dim=1)
    pooled = torch.einsum("blm,bld->bmd", att, encoded)
    return tuple(pooled[:, m] for m in range(pooled.shape[1]))


def t_l2norm_rows(x, eps=1e-12):
    return F.normalize(x, dim=-1, eps=eps)


def t_vr_scores(q_video_n, q_sub_n, feat1_video_n, feat1_sub_n, video_mask, sub_mask):
    total, n = 0, 0
    for q, c, m in ((q_video_n, feat1_video_n, video_mask), (q_sub_n, feat1_sub_n, sub_mask)):
        if q is not None:
            s = t_mask_logits(torch.einsum("md,nld->mln", q, c), m.transpose(0, 1).unsqueeze(0))
            total, n = total + s.max(dim=1)[0], n + 1
    return total / n


def t_span_logits(q_a, feat2_a, mask_a, w_st_a, w_ed_a, q_b=None, feat2_b=None, mask_b=None, w_st_b=None,
                  w_ed_b=None, merged=False, softmax=False, lists=None, out_rows=None):
    if lists is not None and not getattr(lists, "diagonal", False):
        raise NotImplementedError("gradients through inverted-list span scoring are not needed by XML.forward")
    eq = "md,nld->mnl" if lists is None else "bd,bld->bl"

    def conv(sim, w):
        k = w.numel()
        return F.conv1d(sim.reshape(-1, 1, sim.shape[-1]), w.view(1, 1, k), padding=k // 2).view(sim.shape)
    if merged:
        sim = (torch.einsum(eq, q_a, feat2_a) + torch.einsum(eq, q_b, feat2_b)) / 2
        st, ed = t_mask_logits(conv(sim, w_st_a), mask_a), t_mask_logits(conv(sim, w_ed_a), mask_a)
    else:
        st = ed = 0
        n = 0
        for q, f, m, ws, we in ((q_a, feat2_a, mask_a, w_st_a, w_ed_a), (q_b, feat2_b, mask_b, w_st_b, w_ed_b)):
            if q is not None:
                sim = torch.einsum(eq, q, f)
                st, ed, n = st + t_mask_logits(conv(sim, ws), m), ed + t_mask_logits(conv(sim, we), m), n + 1
        st, ed = st / n, ed / n
    if softmax:
        st, ed = torch.softmax(st, -1), torch.softmax(ed, -1)
    return st, ed
