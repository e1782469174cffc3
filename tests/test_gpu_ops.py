"""GPU parity of every C-ABI kernel against the CPU oracle (oracle/xml_oracle.py) on seeded inputs.
Tolerance for floating point: rtol 1e-4 / atol 1e-5 (north_star allows 1e-3 relative); masks, indices and
rankings: exact."""
import numpy as np
import pytest
import torch

from oracle import xml_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from tvretrieval_b200 import ops as _ops
    return _ops


def close(got, want, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(got.detach().cpu().double(), want.detach().cpu().double(), rtol=rtol, atol=atol)


def rand_mask(gen, n, length, min_len=1):
    lens = torch.randint(min_len, length + 1, (n,), generator=gen)
    lens[0] = length
    return (torch.arange(length)[None] < lens[:, None]).float()


@pytest.mark.parametrize("rows,dim", [(7, 64), (300, 768), (33, 3072), (5, 500)])
def test_add_layernorm(ops, rows, dim):
    g = torch.Generator().manual_seed(rows * dim)
    x = torch.randn(rows, dim, generator=g) * 3 + 1
    w, b = torch.randn(dim, generator=g), torch.randn(dim, generator=g)
    want = torch.nn.functional.layer_norm(x, (dim,), w, b, 1e-5)
    close(ops.add_layernorm(x.to(DEV), w.to(DEV), b.to(DEV)), want)
    # + broadcast add (position table): rows = n * period
    period = rows if rows < 8 else rows // 3 if rows % 3 == 0 else rows
    table = torch.randn(period + 4, dim, generator=g)
    want = torch.nn.functional.layer_norm(x + table[:period].repeat(rows // period, 1), (dim,), w, b, 1e-5)
    close(ops.add_layernorm(x.to(DEV), w.to(DEV), b.to(DEV), add=table.to(DEV), add_rows=period), want)


@pytest.mark.parametrize("rows,out_dim,in_dim", [(1, 8, 8), (50, 64, 48), (257, 130, 70), (1000, 768, 3072),
                                                 (129, 768, 768), (64, 500, 500)])
def test_linear(ops, rows, out_dim, in_dim):
    _check_linear(ops, rows, out_dim, in_dim, 1.0, precision="f32")


@pytest.mark.parametrize("rows,out_dim,in_dim", [(256, 64, 48), (257, 130, 70), (1000, 768, 3072), (6000, 768, 768),
                                                 (300, 500, 500), (513, 40, 200), (128 * 9 + 5, 1024, 1024)])
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3"])
def test_linear_tc(ops, rows, out_dim, in_dim, precision):
    """tcgen05 split-precision Linear vs float64; error budget a few fp32 ulps of the accumulated magnitude."""
    _check_linear(ops, rows, out_dim, in_dim, 1.0 if precision == "f16x3" else 4.0, fp64=True, precision=precision)


def _check_linear(ops, rows, out_dim, in_dim, slack, fp64=False, precision="f32"):
    g = torch.Generator().manual_seed(rows + out_dim)
    x = torch.randn(rows, in_dim, generator=g)
    w = torch.randn(out_dim, in_dim, generator=g) * 0.05
    b = torch.randn(out_dim, generator=g)
    r = torch.randn(rows, out_dim, generator=g)
    tol = dict(rtol=1e-4, atol=slack * 1e-5 * max(1.0, in_dim / 256))
    lin = torch.nn.functional.linear
    ref = (lambda *a: lin(*[t.double() for t in a])) if fp64 else lin
    close(ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), precision=precision), ref(x, w, b), **tol)
    close(ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), relu=True, precision=precision), torch.relu(ref(x, w, b)), **tol)
    close(ops.linear(x.to(DEV), w.to(DEV), None, residual=r.to(DEV), precision=precision), ref(x, w) + r, **tol)
    if fp64:
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), precision=precision).cpu().double()
        err = (got - ref(x, w, b)).abs().max().item()
        err32 = (lin(x, w, b).double() - ref(x, w, b)).abs().max().item()
        print("linear_tc %s rows=%d K=%d: max abs err vs fp64 = %.3g (torch CPU fp32: %.3g)"
              % (precision, rows, in_dim, err, err32))


@pytest.mark.parametrize("n,lq,lk,hid,nh,full_mask", [(3, 12, 12, 64, 4, False), (5, 30, 30, 768, 4, False),
                                                      (4, 97, 97, 256, 4, True), (2, 128, 128, 768, 4, True),
                                                      (2, 17, 33, 32, 2, True)])
def test_attention(ops, n, lq, lk, hid, nh, full_mask):
    g = torch.Generator().manual_seed(n * lq + hid)
    w = {}
    for name in ("query", "key", "value"):
        w["a.%s.weight" % name] = torch.randn(hid, hid, generator=g) * 0.05
        w["a.%s.bias" % name] = torch.randn(hid, generator=g) * 0.1
    xq, xk = torch.randn(n, lq, hid, generator=g), torch.randn(n, lk, hid, generator=g)
    mk = rand_mask(g, n, lk)
    if full_mask:  # cross-attention style mask incl. fully masked query rows (reference model_xml.py:369)
        mask3 = rand_mask(g, n, lq).unsqueeze(2) * mk.unsqueeze(1)
    else:
        mask3 = mk.unsqueeze(1)
    want = O.multi_head_attention(xq, xk, mask3, w, "a", nh)
    lin = torch.nn.functional.linear
    q = lin(xq, w["a.query.weight"], w["a.query.bias"]).to(DEV)
    k = lin(xk, w["a.key.weight"], w["a.key.bias"]).to(DEV)
    v = lin(xk, w["a.value.weight"], w["a.value.bias"]).to(DEV)
    close(ops.attention(q, k, v, mask3.to(DEV), nh), want)


@pytest.mark.parametrize("n,length,hid,n_mod", [(4, 12, 64, 2), (9, 30, 768, 2), (3, 8, 32, 1)])
def test_modular_pool(ops, n, length, hid, n_mod):
    g = torch.Generator().manual_seed(n + hid)
    enc = torch.randn(n, length, hid, generator=g)
    mask = rand_mask(g, n, length, 3)
    w = {"modular_vector_mapping.weight": torch.randn(n_mod, hid, generator=g) * 0.1}
    a, b = O.modular_queries(enc, mask, w)
    ga, gb = ops.modular_pool(enc.to(DEV), mask.to(DEV), w["modular_vector_mapping.weight"].to(DEV))
    close(ga, a), close(gb, b)


def test_l2norm_and_softmax(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(37, 768, generator=g)
    x[3] = 0
    close(ops.l2norm_rows(x.to(DEV)), torch.nn.functional.normalize(x, dim=-1), rtol=1e-6, atol=1e-8)
    y = torch.randn(50, 100, generator=g) * 4
    y[:, 60:] = -1e10
    got = ops.softmax_rows(y.to(DEV))
    close(got, torch.softmax(y, -1), rtol=1e-5, atol=1e-8)
    assert (got[:, 60:] == 0).all()


@pytest.mark.parametrize("nq,nv,length,hid,both", [(5, 7, 12, 64, True), (70, 130, 32, 96, True),
                                                  (40, 260, 128, 768, True), (9, 11, 100, 500, False)])
def test_vr_scores_f32(ops, nq, nv, length, hid, both):
    g = torch.Generator().manual_seed(nq * nv)
    qv, qs = torch.randn(nq, hid, generator=g), torch.randn(nq, hid, generator=g)
    fv, fs = torch.randn(nv, length, hid, generator=g), torch.randn(nv, length, hid, generator=g)
    mask = rand_mask(g, nv, length)
    fv = fv * mask.unsqueeze(2)  # rows beyond the batch width are exact zeros in the real corpus
    want = O.video_level_scores(qv, fv, mask)
    if both:
        want = (want + O.video_level_scores(qs, fs, mask)) / 2
    n = lambda t: ops.l2norm_rows(t.to(DEV))  # noqa: E731
    got = ops.vr_scores_f32(n(qv), n(qs) if both else None, n(fv), n(fs) if both else None, mask.to(DEV),
                            mask.to(DEV) if both else None)
    close(got, want, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("nq,nv,length,hid,both", [(5, 7, 12, 64, True), (300, 130, 32, 96, True),
                                                  (200, 300, 128, 768, True), (9, 11, 100, 500, False),
                                                  (130, 40, 256, 1024, True), (64, 1000, 100, 256, True)])
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3"])
def test_vr_scores_tc(ops, nq, nv, length, hid, both, precision):
    """tcgen05 split-precision kernel vs a float64 evaluation of reference get_video_level_scores; the exact-fp32
    SIMT kernel is measured against the same float64 values for comparison."""
    g = torch.Generator().manual_seed(nq * nv + hid)
    qv, qs = torch.randn(nq, hid, generator=g), torch.randn(nq, hid, generator=g)
    fv, fs = torch.randn(nv, length, hid, generator=g), torch.randn(nv, length, hid, generator=g)
    vmask = (torch.rand(nv, length, generator=g) < 0.7).float()  # arbitrary (non-prefix) masks
    vmask[:, 0] = 1
    smask = rand_mask(g, nv, length)
    if nv > 3:
        smask[3] = 0  # a video without any valid clip scores -1e10
    want = O.video_level_scores(qv.double(), fv.double(), vmask.double())
    if both:
        want = (want + O.video_level_scores(qs.double(), fs.double(), smask.double())) / 2
    bf16 = precision == "bf16x3"
    lp, kpad = (length + 31) // 32 * 32, (hid + 63) // 64 * 64
    prep_q = lambda t: ops.split_rows(t.to(DEV), kpad=kpad, normalize=True, bf16=bf16)  # noqa: E731
    prep_c = lambda t: ops.split_rows(t.to(DEV), length, lp, kpad, normalize=True, bf16=bf16)  # noqa: E731
    got = ops.vr_scores_tc(prep_q(qv), prep_c(fv), ops.mask_bits(vmask.to(DEV), lp), nv, lp,
                           q_b=prep_q(qs) if both else None, c_b=prep_c(fs) if both else None,
                           bits_b=ops.mask_bits(smask.to(DEV), lp) if both else None, bf16=bf16)
    torch.cuda.synchronize()
    n = lambda t: ops.l2norm_rows(t.to(DEV))  # noqa: E731
    f32 = ops.vr_scores_f32(n(qv), n(qs) if both else None, n(fv), n(fs) if both else None, vmask.to(DEV),
                            smask.to(DEV) if both else None)
    masked = want < -1e9
    assert torch.equal(got.cpu() < -1e9, masked)
    err_tc = (got.cpu().double() - want)[~masked].abs().max().item()
    err_f32 = (f32.cpu().double() - want)[~masked].abs().max().item()
    print("vr_scores %s: max abs err vs fp64 = %.3g (exact-fp32 SIMT kernel: %.3g)" % (precision, err_tc, err_f32))
    # the TMEM accumulator truncates: the error grows with the number of MMA updates (3 * K / 16)
    assert err_tc <= (2e-6 if bf16 else 6e-7) * max(1.0, hid / 768), (err_tc, err_f32)


@pytest.mark.parametrize("nq,nv,length,hid,both", [(5, 7, 12, 64, True), (300, 130, 32, 96, True),
                                                  (200, 300, 128, 768, True), (9, 11, 100, 500, False),
                                                  (130, 40, 256, 1024, True), (64, 1000, 100, 256, True)])
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3"])
def test_vr_scores_tc_packed(ops, nq, nv, length, hid, both, precision):
    """Packed (valid clips only, whole videos bin-packed into 256-row tiles) tcgen05 kernel vs float64."""
    from tvretrieval_b200.engine import CorpusPacking
    g = torch.Generator().manual_seed(nq * nv + hid + 1)
    qv, qs = torch.randn(nq, hid, generator=g), torch.randn(nq, hid, generator=g)
    fv, fs = torch.randn(nv, length, hid, generator=g), torch.randn(nv, length, hid, generator=g)
    mask = (torch.rand(nv, length, generator=g) < 0.6).float() if nv % 2 else rand_mask(g, nv, length)
    mask[0] = 1
    if nv > 3:
        mask[3] = 0  # a video without any valid clip scores -1e10
    want = O.video_level_scores(qv.double(), fv.double(), mask.double())
    if both:
        want = (want + O.video_level_scores(qs.double(), fs.double(), mask.double())) / 2
    bf16 = precision == "bf16x3"
    kpad = (hid + 63) // 64 * 64
    packing = CorpusPacking(mask.to(DEV))
    assert packing.n_rows == int(mask.sum())
    prep_q = lambda t: ops.split_rows(t.to(DEV), kpad=kpad, normalize=True, bf16=bf16)  # noqa: E731
    prep_c = lambda t: ops.split_rows(t.to(DEV), kpad=kpad, normalize=True, bf16=bf16,  # noqa: E731
                                      row_index=packing.src_rows)
    got = ops.vr_scores_tc_packed(prep_q(qv), prep_c(fv), packing, nv, q_b=prep_q(qs) if both else None,
                                  c_b=prep_c(fs) if both else None, bf16=bf16).cpu()
    masked = want < -1e9
    assert torch.equal(got < -1e9, masked) and (got[masked] == -1e10).all()
    err = (got.double() - want)[~masked].abs().max().item()
    print("vr_scores packed %s: max abs err vs fp64 = %.3g, tile fill %.3f" % (precision, err, packing.fill))
    assert err <= (2e-6 if bf16 else 6e-7) * max(1.0, hid / 768), err


def conv_taps(g, k=5):
    return torch.randn(1, 1, k, generator=g) * 0.4


@pytest.mark.parametrize("nq,nv,length,hid,mode", [(5, 6, 12, 64, "merged"), (37, 9, 32, 96, "merged"),
                                                   (33, 5, 128, 768, "merged"), (7, 4, 100, 500, "two"),
                                                   (6, 5, 24, 32, "one"), (3, 3, 200, 64, "merged")])
@pytest.mark.parametrize("softmax", [False, True])
def test_span_logits_dense(ops, nq, nv, length, hid, mode, softmax):
    g = torch.Generator().manual_seed(nq * nv + length)
    qv, qs = torch.randn(nq, hid, generator=g) * 0.2, torch.randn(nq, hid, generator=g) * 0.2
    fv, fs = torch.randn(nv, length, hid, generator=g), torch.randn(nv, length, hid, generator=g)
    vmask, smask = rand_mask(g, nv, length, 2), rand_mask(g, nv, length, 2)
    taps = [conv_taps(g) for _ in range(4)]
    eye = torch.eye(hid)
    zero = torch.zeros(hid)
    if mode == "merged":
        w = {"video_query_linear.weight": eye, "video_query_linear.bias": zero, "sub_query_linear.weight": eye,
             "sub_query_linear.bias": zero, "merged_st_predictor.weight": taps[0], "merged_ed_predictor.weight": taps[1]}
        st, ed = O.merged_st_ed_logits(w, qv, fv, qs, fs, vmask, cross=True)
        got = ops.span_logits(qv.to(DEV), fv.to(DEV), vmask.to(DEV), taps[0].to(DEV), taps[1].to(DEV), q_b=qs.to(DEV),
                              feat2_b=fs.to(DEV), mask_b=vmask.to(DEV), merged=True, softmax=softmax)
    else:
        w = {"video_query_linear.weight": eye, "video_query_linear.bias": zero, "sub_query_linear.weight": eye,
             "sub_query_linear.bias": zero, "video_st_predictor.weight": taps[0], "video_ed_predictor.weight": taps[1],
             "sub_st_predictor.weight": taps[2], "sub_ed_predictor.weight": taps[3]}
        st, ed = O.single_st_ed_logits(w, qv, fv, vmask, "video", cross=True)
        if mode == "two":
            st2, ed2 = O.single_st_ed_logits(w, qs, fs, smask, "sub", cross=True)
            st, ed = (st + st2) / 2, (ed + ed2) / 2
            got = ops.span_logits(qv.to(DEV), fv.to(DEV), vmask.to(DEV), taps[0].to(DEV), taps[1].to(DEV),
                                  q_b=qs.to(DEV), feat2_b=fs.to(DEV), mask_b=smask.to(DEV), w_st_b=taps[2].to(DEV),
                                  w_ed_b=taps[3].to(DEV), softmax=softmax)
        else:
            got = ops.span_logits(qv.to(DEV), fv.to(DEV), vmask.to(DEV), taps[0].to(DEV), taps[1].to(DEV),
                                  softmax=softmax)
    if softmax:
        st, ed = torch.softmax(st, -1), torch.softmax(ed, -1)
        close(got[0], st, rtol=2e-4, atol=1e-7), close(got[1], ed, rtol=2e-4, atol=1e-7)
        pad = (vmask == 0).unsqueeze(0).expand_as(st) if mode != "two" else None
        if pad is not None:
            assert (got[0].cpu()[pad] == 0).all()
    else:
        close(got[0], st, rtol=1e-4, atol=2e-5), close(got[1], ed, rtol=1e-4, atol=2e-5)
        assert torch.equal(got[0].cpu() == -1e10, st == -1e10)


def test_span_logits_lists_match_dense(ops):
    """Inverted-list mode (selected pairs) must reproduce the dense rows bit for bit."""
    g = torch.Generator().manual_seed(77)
    nq, nv, length, hid, slots = 150, 40, 64, 128, 9
    qv, qs = torch.randn(nq, hid, generator=g).to(DEV), torch.randn(nq, hid, generator=g).to(DEV)
    fv, fs = torch.randn(nv, length, hid, generator=g).to(DEV), torch.randn(nv, length, hid, generator=g).to(DEV)
    mask = rand_mask(g, nv, length, 2).to(DEV)
    t0, t1 = conv_taps(g).to(DEV), conv_taps(g).to(DEV)
    top = torch.stack([torch.randperm(nv, generator=g)[:slots] for _ in range(nq)]).to(torch.int32).to(DEV)
    top[:, 0] = 3  # one very popular video -> several 32-query chunks
    dense_st, dense_ed = ops.span_logits(qv, fv, mask, t0, t1, q_b=qs, feat2_b=fs, mask_b=mask, merged=True,
                                         softmax=True)
    lists = ops.build_pair_lists(top, nv)
    vp = lists.vid_ptr.cpu().numpy()
    assert vp[-1] == nq * slots and (np.diff(vp) == np.bincount(top.cpu().numpy().ravel(), minlength=nv)).all()
    st, ed = ops.span_logits(qv, fv, mask, t0, t1, q_b=qs, feat2_b=fs, mask_b=mask, merged=True, softmax=True,
                             lists=lists)
    rows = torch.arange(nq, device=DEV).unsqueeze(1)
    assert torch.equal(st.view(nq, slots, length), dense_st[rows, top.long()])
    assert torch.equal(ed.view(nq, slots, length), dense_ed[rows, top.long()])
    # diagonal lists (cross=False path)
    n = 30
    lists = ops.diagonal_pair_lists(n, DEV)
    st, _ = ops.span_logits(qv[:n], fv[:n], mask[:n], t0, t1, q_b=qs[:n], feat2_b=fs[:n], mask_b=mask[:n],
                            merged=True, softmax=True, lists=lists)
    idx = torch.arange(n, device=DEV)
    assert torch.equal(st, dense_st[idx, idx])


@pytest.mark.parametrize("nq,nv,length,hid,slots,chunk", [(150, 40, 64, 128, 9, 32), (300, 25, 128, 768, 12, 64),
                                                          (40, 30, 100, 500, 7, 32), (700, 12, 128, 256, 8, 128),
                                                          (64, 64, 32, 64, 1, 32), (260, 20, 256, 1024, 10, 128),
                                                          (90, 16, 200, 256, 6, 64), (35, 9, 129, 128, 4, 32)])
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3"])
def test_span_probs_tc(ops, nq, nv, length, hid, slots, chunk, precision):
    """tcgen05 grouped similarity + ConvSE + softmax kernel vs the oracle (float64) on selected pairs."""
    g = torch.Generator().manual_seed(nq + nv + hid)
    qv, qs = torch.randn(nq, hid, generator=g) * 0.1, torch.randn(nq, hid, generator=g) * 0.1
    fv, fs = torch.randn(nv, length, hid, generator=g), torch.randn(nv, length, hid, generator=g)
    mask = rand_mask(g, nv, length, 2)
    t0, t1 = conv_taps(g), conv_taps(g)
    top = torch.stack([torch.randperm(nv, generator=g)[:slots] for _ in range(nq)]).to(torch.int32)
    if slots > 1:
        top[:, 0] = 3  # one video selected by every query: many chunks
    eye, zero = torch.eye(hid).double(), torch.zeros(hid).double()
    w = {"video_query_linear.weight": eye, "video_query_linear.bias": zero, "sub_query_linear.weight": eye,
         "sub_query_linear.bias": zero, "merged_st_predictor.weight": t0.double(), "merged_ed_predictor.weight": t1.double()}
    st, ed = O.merged_st_ed_logits(w, qv.double(), fv.double(), qs.double(), fs.double(), mask.double(), cross=True)
    st, ed = torch.softmax(st, -1), torch.softmax(ed, -1)
    rows = torch.arange(nq).unsqueeze(1)
    want_st, want_ed = st[rows, top.long()], ed[rows, top.long()]
    bf16 = precision == "bf16x3"
    kpad = (hid + 63) // 64 * 64
    f2cat = (torch.empty(nv * length, 2 * kpad, device=DEV, dtype=torch.int16),
             torch.empty(nv * length, 2 * kpad, device=DEV, dtype=torch.int16))
    ops.split_rows(fv.to(DEV), kpad=kpad, bf16=bf16, out=f2cat, out_col0=0)
    ops.split_rows(fs.to(DEV), kpad=kpad, bf16=bf16, out=f2cat, out_col0=kpad)
    pad = torch.nn.functional.pad
    q_cat = torch.cat([pad(qv, (0, kpad - hid)), pad(qs, (0, kpad - hid))], 1).to(DEV)
    lists = ops.build_pair_lists(top.to(DEV), nv, chunk=chunk)
    got_st, got_ed = ops.span_probs_tc(f2cat, q_cat, lists, mask.to(DEV), t0.to(DEV), t1.to(DEV), length, bf16=bf16)
    saved = ops.GATHER  # the three ways of fetching the listed query rows give the same bits
    try:
        for mode in ("copy", "warps", "tma"):
            ops.GATHER = mode
            g_st, g_ed = ops.span_probs_tc(f2cat, q_cat, lists, mask.to(DEV), t0.to(DEV), t1.to(DEV), length, bf16=bf16)
            assert torch.equal(g_st, got_st) and torch.equal(g_ed, got_ed), mode
        # per-video clip boxes: only the rows an unmasked clip's ConvSE taps can read are loaded; the same bits,
        # even when the rows beyond them hold NaNs (nothing may depend on the padded tail of a video)
        ops.GATHER = "warps"
        clip_rows = ops.span_clip_rows(mask.to(DEV), t0.numel())
        assert int(clip_rows.max()) <= length and int(clip_rows.min()) >= 1
        poisoned = tuple(t.clone().view(nv, length, -1) for t in f2cat)
        beyond = torch.arange(length, device=DEV)[None, :] >= ((clip_rows[:, None] + 15) // 16 * 16)
        for t in poisoned:
            t[beyond] = 0x7e00 if not bf16 else 0x7fc0  # NaN
        poisoned = tuple(t.view(nv * length, -1) for t in poisoned)
        for operands in (f2cat, poisoned):
            g_st, g_ed = ops.span_probs_tc(operands, q_cat, lists, mask.to(DEV), t0.to(DEV), t1.to(DEV), length,
                                           bf16=bf16, clip_rows=clip_rows)
            assert torch.equal(g_st, got_st) and torch.equal(g_ed, got_ed), "clip boxes"
            # ... and from the k-blocked operand layout (contiguous TMA boxes)
            kb = tuple(ops.kblock_rows(t) for t in operands)
            for rows_arg in (None, clip_rows):
                g_st, g_ed = ops.span_probs_tc(kb, q_cat, lists, mask.to(DEV), t0.to(DEV), t1.to(DEV), length,
                                               bf16=bf16, clip_rows=rows_arg)
                assert torch.equal(g_st, got_st) and torch.equal(g_ed, got_ed), "k-blocked layout"
                if length % 8 == 0:  # ... and from the shared-memory image (plain bulk copies)
                    img = tuple(ops.kblock_rows(t, swizzle=True) for t in operands)
                    g_st, g_ed = ops.span_probs_tc(img, q_cat, lists, mask.to(DEV), t0.to(DEV), t1.to(DEV), length,
                                                   bf16=bf16, clip_rows=rows_arg)
                    assert torch.equal(g_st, got_st) and torch.equal(g_ed, got_ed), "shared-memory image layout"
        for mode in ("copy", "tma"):  # the other producers read the k-blocked layout too
            ops.GATHER = mode
            g_st, g_ed = ops.span_probs_tc(tuple(ops.kblock_rows(t) for t in f2cat), q_cat, lists, mask.to(DEV),
                                           t0.to(DEV), t1.to(DEV), length, bf16=bf16)
            assert torch.equal(g_st, got_st) and torch.equal(g_ed, got_ed), "k-blocked layout, " + mode
    finally:
        ops.GATHER = saved
    got_st, got_ed = got_st.view(nq, slots, length).cpu(), got_ed.view(nq, slots, length).cpu()
    err = max((got_st.double() - want_st).abs().max().item(), (got_ed.double() - want_ed).abs().max().item())
    rel = ((got_st.double() - want_st).abs() / want_st.clamp_min(1e-30))[want_st > 1e-6].max().item()
    print("span_probs_tc %s: max abs err %.3g, max rel err %.3g" % (precision, err, rel))
    close(got_st, want_st, rtol=2e-4, atol=1e-7), close(got_ed, want_ed, rtol=2e-4, atol=1e-7)
    pad_pos = (mask == 0)[top.long()]
    assert (got_st[pad_pos] == 0).all()


def test_split_rows_hi_err(ops):
    """hi_err = ||x - hi||_2 (rounded up): the Cauchy-Schwarz error bound of hi-only products holds."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(300, 700, generator=g)
    x[3] *= 1e-4  # fp16 subnormal range after normalisation
    for bf16 in (False, True):
        hi, lo, err = ops.split_rows(x.to(DEV), kpad=704, normalize=True, bf16=bf16, hi_err=True)
        xn = torch.nn.functional.normalize(x.double(), dim=1)
        h = hi.view(torch.bfloat16 if bf16 else torch.float16).double().cpu()[:, :700]
        want = (xn - h).norm(dim=1)
        got = err.double().cpu()
        assert (got >= want * (1 - 1e-6)).all() and (got <= want * 1.001 + 1e-9).all()
        # bound of a hi-only dot product of two unit vectors
        d_exact = xn[:150] @ xn[150:].T
        d_hi = h[:150] @ h[150:].T
        bound = got[:150, None] * 1.0 + 1.001 * got[None, 150:]
        assert ((d_exact - d_hi).abs() <= bound).all()


@pytest.mark.parametrize("rows,cols,k,max_cand", [(5, 300, 10, 64), (130, 2179, 100, 256), (260, 5000, 100, 128)])
def test_select_candidates(ops, rows, cols, k, max_cand):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.rand(rows, cols, generator=g) * 0.12 + 0.03
    x[1, :40] = x[1].max()  # ties at the top
    x[2, 3] = 0.9           # an outlier maximum: fewer than k scores within 1/8 of it -> the exact selection path
    err_a = torch.rand(rows, generator=g) * 2e-4
    err_b = torch.rand(rows, generator=g) * 2e-4
    if rows > 200:
        err_a[200] = 0.05  # a row whose window holds more than max_cand columns: flagged
        err_a[7] = 0.05
    scale, const = 0.5, 1e-4
    ids = torch.randperm(cols, generator=g).to(torch.int32)
    cand = ops.select_candidates(x.to(DEV), k, err_a.to(DEV), err_b.to(DEV), scale, const, max_cand, ids=ids.to(DEV))
    eps = (scale * (err_a + err_b) + const).double()
    kth = torch.sort(x, dim=1, descending=True)[0][:, k - 1].double()
    col, cid, val = cand.col.cpu(), cand.ids.cpu(), cand.val.cpu()
    flags = cand.row_flags.cpu()
    n_flagged_groups = int(cand.n_flagged.cpu())
    # the kernel's k-th largest is a LOWER bound at most one 2^-15 bin (+ rounding slack) below the exact one
    below = 2.0 ** -15 + 1e-6
    flagged_groups, maybe_flagged = set(), set()
    for r in range(rows):
        cut = kth[r] - 2 * eps[r]
        must = set(torch.nonzero(x[r].double() >= cut + 1e-7).flatten().tolist())
        may = set(torch.nonzero(x[r].double() >= cut - below).flatten().tolist())
        got = [c for c in col[r].tolist() if c >= 0]
        assert len(got) == len(set(got))
        if len(may) > max_cand:
            maybe_flagged.add(r // 128)
        if len(must) > max_cand:
            assert flags[r] == 1
            flagged_groups.add(r // 128)
            continue
        if len(may) <= max_cand:
            assert flags[r] == 0
        if flags[r] == 0:
            assert must <= set(got) <= may, r
            n = len(got)
            assert (col[r, n:] == -1).all() and (cid[r, n:] == 2 ** 31 - 1).all() and (val[r, n:] == -1e10).all()
            assert torch.equal(cid[r, :n], ids[col[r, :n].long()]) and torch.equal(val[r, :n], x[r, col[r, :n].long()])
    if rows > 200:
        assert flags[200] == 1 and flags[7] == 1
    got_groups = set(cand.flagged_groups.cpu()[:n_flagged_groups].tolist())
    assert len(got_groups) == n_flagged_groups and flagged_groups <= got_groups <= maybe_flagged
    assert got_groups == {r // 128 for r in range(rows) if flags[r] == 1}


@pytest.mark.parametrize("nq,nv,length,hid,both,k,max_cand", [(300, 900, 64, 256, True, 20, 20),
                                                             (200, 1500, 128, 768, True, 100, 256),
                                                             (150, 700, 100, 500, False, 50, 50),
                                                             (140, 600, 256, 128, True, 30, 40)])
def test_two_pass_top_videos(ops, nq, nv, length, hid, both, k, max_cand):
    """hi-only filter + candidate selection + exact re-scoring (+ overflow fallback) returns exactly what the
    one-pass split-precision kernel + top-k returns."""
    from tvretrieval_b200.engine import CorpusPacking
    g = torch.Generator().manual_seed(nq + nv + hid)
    # clustered scores: a common direction plus noise, so that many videos fall inside the candidate window
    base = torch.randn(hid, generator=g)
    mk = lambda *shape: (0.35 * base + torch.randn(*shape, hid, generator=g)).to(DEV)  # noqa: E731
    qv, qs, fv, fs = mk(nq), mk(nq), mk(nv, length), mk(nv, length)
    mask = rand_mask(g, nv, length).to(DEV)
    mask[5] = 0
    kpad = (hid + 63) // 64 * 64
    pk = CorpusPacking(mask)
    ids = pk.order_full + 1000
    cs = [ops.split_rows(f, kpad=kpad, normalize=True, row_index=pk.src_rows, hi_err=True) for f in (fv, fs)]
    qsp = [ops.split_rows(q, kpad=kpad, normalize=True, hi_err=True) for q in (qv, qs)]
    n_mod = 2 if both else 1
    kw = dict(q_b=qsp[1][:2] if both else None, c_b=cs[1][:2] if both else None, ordinal=True)
    exact = ops.vr_scores_tc_packed(qsp[0][:2], cs[0][:2], pk, nv, **kw)
    want_idx, want_val = ops.topk_rows(exact, k, alpha=20.0, apply_exp=True, ids=ids)
    approx = ops.vr_scores_tc_packed(qsp[0][:2], cs[0][:2], pk, nv, hi_only=True, **kw)
    scale = 1.001 / n_mod
    const = scale * sum(float(c[2].max()) for c in cs[:n_mod]) + 4e-5
    bound = scale * (qsp[0][2] + (qsp[1][2] if both else 0)) + const
    real = exact > -1e9
    assert ((approx - exact).abs()[real] <= bound[:, None].expand_as(exact)[real]).all()
    print("hi-only error: max %.3g, bound min %.3g" % ((approx - exact).abs()[real].max().item(), bound.min().item()))
    cand = ops.select_candidates(approx, k, qsp[0][2], qsp[1][2] if both else None, scale, const, max_cand, ids=ids)
    n_cand = (cand.col >= 0).sum(1)
    print("candidates per query: mean %.1f max %d, flagged rows %d" % (n_cand.float().mean().item(), n_cand.max().item(),
                                                                     int(cand.row_flags.sum())))
    ops.vr_rescore_tc(qv, cs[0][:2], pk, cand, kpad, q_fp32_b=qs if both else None, c_b=cs[1][:2] if both else None)
    val_copy, approx_val = cand.val.clone(), torch.gather(approx, 1, cand.col.clamp(min=0).long())
    cand.val.copy_(approx_val)
    saved = ops.GATHER  # the three ways of fetching the listed query rows give the same bits
    listed = cand.col >= 0
    try:
        for mode in ("copy", "warps", "tma"):
            ops.GATHER = mode
            cand.val.copy_(approx_val)
            ops.vr_rescore_tc(qv, cs[0][:2], pk, cand, kpad, q_fp32_b=qs if both else None,
                              c_b=cs[1][:2] if both else None)
            assert torch.equal(cand.val[listed], val_copy[listed]), mode
            # ... and from the k-blocked corpus layout (contiguous TMA boxes; quarter boxes in gather-warps mode)
            ckb = [tuple(ops.kblock_rows(t) for t in c[:2]) for c in cs]
            cand.val.copy_(approx_val)
            ops.vr_rescore_tc(qv, ckb[0], pk, cand, kpad, q_fp32_b=qs if both else None, c_b=ckb[1] if both else None)
            assert torch.equal(cand.val[listed], val_copy[listed]), mode + ", k-blocked"
    finally:
        ops.GATHER = saved
    cand.val.copy_(val_copy)
    ok = cand.col >= 0
    packed = ok & (cand.col < pk.n_packed)
    assert torch.equal(cand.val[packed], torch.gather(exact, 1, cand.col.clamp(min=0).long())[packed])
    idx, val = ops.topk_rows(cand.val, k, alpha=20.0, apply_exp=True, ids=cand.ids)
    clean = cand.row_flags == 0
    assert clean.any()
    assert torch.equal(idx[clean], want_idx[clean]) and torch.equal(val[clean], want_val[clean])
    # overflowed rows: restricted exact pass + restricted top-k, in place
    ops.vr_scores_tc_packed(qsp[0][:2], cs[0][:2], pk, nv, out=approx, m_tiles=(cand.flagged_groups, cand.n_flagged),
                            **kw)
    ops.topk_rows(approx, k, alpha=20.0, apply_exp=True, ids=ids, row_flags=cand.row_flags, out=(idx, val))
    assert torch.equal(idx, want_idx) and torch.equal(val, want_val)
    if max_cand == k:
        assert int(cand.row_flags.sum()) > 0, "this case is meant to exercise the overflow fallback"


@pytest.mark.parametrize("rows,cols,k", [(3, 100, 100), (17, 2179, 100), (4, 21793, 100), (5, 333, 7), (2, 5000, 1000)])
def test_topk_rows(ops, rows, cols, k):
    g = torch.Generator().manual_seed(rows * cols)
    x = torch.rand(rows, cols, generator=g) * 0.12 + 0.03
    x[0, :50] = x[0, 60]  # exact ties straddling the cut
    if cols > 200:
        x[1, 100:160] = x[1].max() + 0.01
    e = torch.exp(20.0 * x)
    order = O.stable_desc_order(e)[:, :k]
    idx, val = ops.topk_rows(x.to(DEV), k, alpha=20.0, apply_exp=True)
    got_e = val.cpu()
    # the device expf may differ from torch CPU exp in the last ulp: compare values loosely, order exactly
    # wherever the oracle's own gap is not a tie within 2 ulp
    close(got_e, torch.gather(e, 1, order), rtol=5e-7, atol=0)
    want_idx = order.numpy()
    got_idx = idx.cpu().numpy().astype(np.int64)
    same = got_idx == want_idx
    if not same.all():
        ev = e.numpy()
        for r, c in zip(*np.nonzero(~same)):
            a, b = ev[r, got_idx[r, c]], ev[r, want_idx[r, c]]
            assert abs(a - b) <= 4e-7 * abs(b), (r, c, a, b)
    # raw (no exp) selection with explicit ids, ties by id ascending / descending
    ids = torch.randperm(cols, generator=g).to(torch.int32).unsqueeze(0).repeat(rows, 1)
    xi = torch.round(x * 2000) / 2000  # many exact ties
    for tie_desc in (False, True):
        idx, val = ops.topk_rows(xi.to(DEV), min(k, 64), ids=ids.to(DEV), tie_desc=tie_desc)
        key = xi.double() * 1e9 + (ids.double() if tie_desc else -ids.double())
        order = torch.sort(key, dim=1, descending=True)[1][:, :min(k, 64)]
        assert torch.equal(idx.cpu().long(), torch.gather(ids.long(), 1, order))
        assert torch.equal(val.cpu(), torch.gather(xi, 1, order))


def test_topk_rows_rejects_k_larger_than_row(ops):
    from tvretrieval_b200._lib import XmlbError
    with pytest.raises(XmlbError):
        ops.topk_rows(torch.rand(2, 10, device=DEV), 100)


def oracle_span_topk(st, ed, vr, min_l, max_l, k, tie_desc=False):
    span = torch.einsum("qvm,qv,qvn->qvmn", st, vr, ed) if vr is not None else torch.einsum("qvm,qvn->qvmn", st, ed)
    span = span * torch.from_numpy(O.band_mask(st.shape[-1], min_l, max_l))
    flat = span.reshape(len(span), -1)
    if tie_desc:
        order = torch.flip(torch.sort(flat, dim=1, descending=False, stable=True)[1], dims=[1])[:, :k]
    else:
        order = O.stable_desc_order(flat)[:, :k]
    return order, torch.gather(flat, 1, order)


@pytest.mark.parametrize("nq,slots,length,k", [(4, 100, 128, 200), (6, 8, 32, 50), (3, 100, 100, 200), (2, 5, 24, 1000),
                                              (2, 30, 256, 200)])
def test_span_topk_vcmr(ops, nq, slots, length, k):
    g = torch.Generator().manual_seed(nq * slots + length)
    mask = rand_mask(g, nq * slots, length, 3).view(nq, slots, length)
    st = torch.softmax(torch.randn(nq, slots, length, generator=g) * 2 + (mask - 1) * 1e10, -1)
    ed = torch.softmax(torch.randn(nq, slots, length, generator=g) * 2 + (mask - 1) * 1e10, -1)
    vr = torch.exp(20 * (torch.rand(nq, slots, generator=g) * 0.1 + 0.03))
    st[0, 0] = st[0, 1]  # duplicated rows -> exact score ties between slots
    ed[0, 0] = ed[0, 1]
    vr[0, 0] = vr[0, 1]
    order, score = oracle_span_topk(st, ed, vr, 2, 16, k)
    idx, val = ops.span_topk(st.to(DEV), ed.to(DEV), vr.to(DEV), 2, 16, k)
    assert torch.equal(val.cpu(), score)  # (st*vr)*ed is bit-exact
    assert torch.equal(idx.cpu().long(), order)


def test_span_topk_zero_fill_and_svmr_ties(ops):
    g = torch.Generator().manual_seed(3)
    nq, length, k = 5, 20, 60
    lens = [3, 5, 20, 4, 9]
    st = torch.zeros(nq, 1, length)
    ed = torch.zeros(nq, 1, length)
    for i, n in enumerate(lens):
        st[i, 0, :n] = torch.softmax(torch.randn(n, generator=g), 0)
        ed[i, 0, :n] = torch.softmax(torch.randn(n, generator=g), 0)
    for tie_desc in (False, True):
        order, score = oracle_span_topk(st, ed, None, 2, 16, k, tie_desc=tie_desc)
        idx, val = ops.span_topk(st.to(DEV), ed.to(DEV), None, 2, 16, k, tie_desc=tie_desc)
        assert torch.equal(val.cpu(), score)
        assert torch.equal(idx.cpu().long(), order), tie_desc
    # without zero fill the tail is (-1, 0); span_zero_fill completes it identically
    idx, val = ops.span_topk(st.to(DEV), ed.to(DEV), None, 2, 16, k, zero_fill=False)
    assert (idx[0, 1:] == -1).all() and (val[0, 1:] == 0).all()
    order, score = oracle_span_topk(st, ed, None, 2, 16, k)
    idx, val = ops.span_zero_fill(idx, val, length * length)
    assert torch.equal(idx.cpu().long(), order) and torch.equal(val.cpu(), score)


def test_temporal_nms_known_answers(ops):
    from tests.golden_io import GOLDEN_DIR
    z = np.load(GOLDEN_DIR + "/temporal_nms.npz")
    for i in range(int(z["n_cases"])):
        preds = z["in/%d" % i]
        order = np.argsort(-preds[:, 2], kind="stable")  # the kernel takes ranked lists
        p = torch.from_numpy(preds[order]).float().to(DEV)
        kept, cnt = ops.temporal_nms(p[None, :, 0], p[None, :, 1], p[None, :, 2], float(z["thd/%d" % i]), 100)
        got = preds[order][kept[0, :int(cnt[0])].cpu().numpy()]
        assert np.array_equal(got, z["out/%d" % i]), i


@pytest.mark.parametrize("n_in", [900, 3000])
def test_temporal_nms_long_lists_vs_oracle(ops, n_in):
    """Per-video NMS over long ranked lists (up to 4096 predictions per query; the reference's max_before_nms default
    is 1000) against the oracle restatement of filter_vcmr_by_nms."""
    g = torch.Generator().manual_seed(n_in)
    nq, max_out = 3, 100
    vid = torch.randint(0, 25, (nq, n_in), generator=g)
    st = torch.randint(0, 80, (nq, n_in), generator=g).float() * 1.5
    ed = st + torch.randint(2, 16, (nq, n_in), generator=g).float() * 1.5
    score = torch.sort(torch.rand(nq, n_in, generator=g), dim=1, descending=True)[0]
    kept, cnt = ops.temporal_nms(st.to(DEV), ed.to(DEV), score.to(DEV), 0.5, max_out, video_idx=vid.to(DEV))
    for q in range(nq):
        preds = [[int(vid[q, i]), float(st[q, i]), float(ed[q, i]), float(score[q, i])] for i in range(n_in)]
        want = O.vcmr_nms(preds, 0.5, n_in, max_out)
        got = [preds[i] for i in kept[q, :int(cnt[q])].cpu().tolist()]
        assert got == want, q


# ---------------------------------------------------------------------------------------------------------
# packed (ragged) query encoder
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hid,heads", [(768, 4), (64, 2), (256, 1), (64, 4), (40, 2)])
def test_attention_ragged_vs_padded_reference(hid, heads):
    """xmlb_attention_ragged on packed tokens == the reference attention (model_components.py:277-303) of the padded
    batch at every valid token (padded keys have probability exactly 0 there)."""
    from tvretrieval_b200 import autograd, ops
    g = torch.Generator().manual_seed(5)
    n, width = 37, 30
    lens = torch.randint(1, width + 1, (n,), generator=g)
    lens[0], lens[1] = width, 1
    mask = (torch.arange(width)[None] < lens[:, None]).float()
    q, k, v = (torch.randn(n, width, hid, generator=g) for _ in range(3))
    want = autograd.t_attention(q.double(), k.double(), v.double(), mask.double().unsqueeze(1), heads)
    valid = mask.bool()
    cu = torch.zeros(n + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    got = ops.attention_ragged(q[valid].to(DEV), k[valid].to(DEV), v[valid].to(DEV), cu.to(DEV), int(lens.max()), heads)
    torch.testing.assert_close(got.cpu().double(), want[valid], rtol=1e-5, atol=1e-6)
    # and the padded kernel path agrees with it to rounding
    padded = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), mask.unsqueeze(1).to(DEV), heads)
    torch.testing.assert_close(got, padded[valid.to(DEV)], rtol=1e-5, atol=1e-6)


def test_ragged_attention_argument_checks():
    from tvretrieval_b200 import ops
    from tvretrieval_b200._lib import XmlbError
    x = torch.zeros(40, 64, device=DEV)
    cu = torch.tensor([0, 40], dtype=torch.int32, device=DEV)
    with pytest.raises(XmlbError):
        ops.attention_ragged(x, x, x, cu, 40, 2)       # longer than 32 tokens
    y = x[:, :12].contiguous()
    with pytest.raises(XmlbError):
        ops.attention_ragged(y, y, y, cu, 8, 2)        # head size 6: not a multiple of 4


@pytest.mark.parametrize("name", ["video_sub_vcmr", "video_only_svmr"])
def test_packed_query_encoder_equals_padded(name):
    """XML.encode_query_packed (valid tokens only) returns the pooled vectors of XML.encode_query."""
    from tests.golden_io import GoldenCase
    from tvretrieval_b200.model_xml import XML, AttrDict
    g = GoldenCase(name)
    model = XML(AttrDict(g.cfg))
    model.load_state_dict(g.weights)
    model = model.to(DEV).eval()
    qf, qm = g.query_feat.to(DEV), g.query_mask.to(DEV)
    with torch.no_grad():
        want = model.encode_query(qf, qm)
        got = model.encode_query_packed(qf, g.query_mask.sum(1).long().numpy())
    for a, b, key in zip(got, want, ("video_query", "sub_query")):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(a.cpu(), g.t(key), rtol=1e-4, atol=1e-5)  # and the reference's own vectors


def test_packed_tables_made_on_the_device_equal_the_host_layout():
    """XML.packed_query_tables_device (cumsum / repeat_interleave on the GPU, per piece) == packed_layout (numpy)."""
    import numpy as np
    from tvretrieval_b200.model_xml import XML, AttrDict, packed_layout, xml_base_config
    model = XML(AttrDict(xml_base_config, hidden_size=64, visual_input_size=64, query_input_size=64, sub_input_size=64))
    g = torch.Generator().manual_seed(7)
    width = 30
    lens = torch.randint(0, width + 3, (517,), generator=g).numpy()  # some empty, some longer than the padded width
    lens[5] = 0
    lens = np.minimum(lens, width)
    bounds = [(0, 100), (100, 101), (101, 517)]
    for lens_dev in (None, torch.from_numpy(lens).to(DEV)):
        tables = model.packed_query_tables_device(lens, width, torch.device(DEV), bounds, lens_dev=lens_dev)
        assert len(tables) == len(bounds)
        for (lo, hi), (rows, pos, cu, max_len) in zip(bounds, tables):
            w_rows, w_pos, w_cu, w_max = packed_layout(lens[lo:hi], width)
            assert rows.dtype == pos.dtype == cu.dtype == torch.int32
            assert np.array_equal(rows.cpu().numpy(), w_rows) and np.array_equal(pos.cpu().numpy(), w_pos)
            assert np.array_equal(cu.cpu().numpy(), w_cu) and max_len == w_max
