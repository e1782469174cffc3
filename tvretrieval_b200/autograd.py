"""Autograd support for the kernel ops: what the training step `XML.forward` (reference model_xml.py:212-251,
driven by train.py:77-85) needs on top of the inference kernels.

Forward values always come from the CUDA kernels (include/xmlb200.h).  Backward:
  * Linear layers -- the bulk of the FLOPs -- use the same kernels: dX = dY . W and dW = dY^T . X are two more calls
    of xmlb_linear / xmlb_linear_tc on transposed operands;
  * every other op (LayerNorm, attention core, modular pooling, L2 normalisation, masked-max video scores,
    similarity + ConvSE) is differentiated by RECOMPUTATION: its saved inputs are pushed through the equivalent
    torch expression under autograd (SURVEY.md section 8f rank 3 allows either).  Hand-written backward kernels for
    these memory-bound ops are future work; nothing here runs on the CPU.
Inference never enters this module: the dispatch in ops.py only routes here when autograd is recording.
"""
import torch
import torch.nn.functional as F

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
            g = g * (out > 0)
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
                db = g2.sum(0)
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
