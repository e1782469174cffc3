"""GPU parity of the tensor-core encoder pipeline (csrc/linear_tc.cu with K-chunked accumulation and fused output
formats, csrc/norm.cu LayerNorm + split, csrc/attention_tc.cu) against the oracle evaluated in float64.
Tolerances are absolute errors relative to O(1) activations; the fp32 reference arithmetic itself (torch CPU fp32) is
measured against the same float64 values and printed next to the kernels' error."""
import copy

import pytest
import torch

from oracle import xml_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from tvretrieval_b200 import ops as _ops
    return _ops


def join(pair, bf16=False):
    """(hi, lo) int16 raw bits -> float64 hi + lo."""
    dt = torch.bfloat16 if bf16 else torch.float16
    return pair[0].view(dt).double() + pair[1].view(dt).double()


@pytest.mark.parametrize("rows,dim", [(300, 768), (40, 3072), (9, 500)])
def test_add_layernorm_split(ops, rows, dim):
    g = torch.Generator().manual_seed(rows + dim)
    x = (torch.randn(rows, dim, generator=g) * 2 + 0.5).to(DEV)
    w, b = torch.randn(dim, generator=g).to(DEV), torch.randn(dim, generator=g).to(DEV)
    table = torch.randn(rows // 3 + 1, dim, generator=g).to(DEV)
    period = rows // 3
    want = ops.add_layernorm(x, w, b, add=table, add_rows=period)
    out, pair = ops.add_layernorm_split(x, w, b, add=table, add_rows=period)
    assert torch.equal(out, want)
    kpad = ops.pad64(dim)
    assert pair[0].shape == (rows, kpad)
    got = join(pair)
    assert float((got[:, :dim] - want.double()).abs().max()) <= 2e-6 * float(want.abs().max())
    assert float(got[:, dim:].abs().max()) == 0 if kpad > dim else True
    none, pair2 = ops.add_layernorm_split(x, w, b, add=table, add_rows=period, want_f32=False)
    assert none is None and torch.equal(pair2[0], pair[0]) and torch.equal(pair2[1], pair[1])


@pytest.mark.parametrize("rows,out_dim,in_dim,seq", [(512, 768 * 3, 768, 128), (1000, 768, 3072, 100),
                                                     (384, 256 * 3, 256, 32), (260, 320, 192, 65)])
def test_linear_tc_fused_outputs(ops, rows, out_dim, in_dim, seq):
    """xmlb_linear_tc_ex: fp32 output vs float64, and the split / transposed-split outputs are the split of exactly
    that fp32 output."""
    g = torch.Generator().manual_seed(rows + out_dim)
    x = torch.randn(rows, in_dim, generator=g)
    w = torch.randn(out_dim, in_dim, generator=g) * 0.05
    b = torch.randn(out_dim, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    x16, w16 = ops.split_rows(x.to(DEV)), ops.split_rows(w.to(DEV))
    hid = out_dim // 3 if out_dim % 3 == 0 else 64
    vt_col0 = out_dim - hid
    o16_cols = vt_col0 // 8 * 8
    out, o16, vt = ops.linear_tc_ex(x16, w16, b.to(DEV), out16_cols=o16_cols, vt_col0=vt_col0,
                                    vt_seq=seq if rows % seq == 0 else 1)
    err = float((out.cpu().double() - ref).abs().max())
    err32 = float((torch.nn.functional.linear(x, w, b).double() - ref).abs().max())
    print("linear_tc_ex rows=%d N=%d K=%d: max abs err vs fp64 %.3g (torch CPU fp32 %.3g)" % (rows, out_dim, in_dim, err, err32))
    assert err <= 3e-6 * max(1.0, (in_dim / 768) ** 0.5) * float(ref.abs().max())
    # the 16-bit outputs are the split of the fp32 output (hi + lo reproduces it to 2^-22)
    got16 = join(o16)
    assert float((got16 - out[:, :o16_cols].double()).abs().max()) <= 3e-7 * float(out.abs().max())
    s = seq if rows % seq == 0 else 1
    vt_f = join(vt).view(rows // s, hid, -1)
    want_t = out[:, vt_col0:].double().view(rows // s, s, hid).transpose(1, 2)
    assert float((vt_f[:, :, :s] - want_t).abs().max()) <= 3e-7 * float(out.abs().max())
    assert float(vt_f[:, :, s:].abs().max()) == 0 if vt_f.shape[2] > s else True
    # same values without the fused outputs, and with another K-chunk (accuracy knob, not bit-equal)
    out2, _, _ = ops.linear_tc_ex(x16, w16, b.to(DEV))
    assert torch.equal(out2, out)
    out3, _, _ = ops.linear_tc_ex(x16, w16, b.to(DEV), k_chunk=64)
    torch.testing.assert_close(out3, out, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,lq,lk,hid,nh,full_mask", [(3, 40, 40, 256, 4, False), (5, 128, 128, 768, 4, False),
                                                      (4, 97, 97, 256, 4, True), (2, 128, 128, 768, 4, True),
                                                      (3, 256, 256, 1024, 4, True), (2, 200, 200, 512, 8, False),
                                                      (160, 128, 128, 768, 4, False)])
def test_attention_tc(ops, n, lq, lk, hid, nh, full_mask):
    """Fused tcgen05 attention (projections through linear_tc_ex with the fused split / transposed outputs) vs the
    oracle's BertSelfAttention in float64, incl. fully masked query rows (padded clips of the cross attention)."""
    g = torch.Generator().manual_seed(n * lq + hid)
    w = {}
    for name in ("query", "key", "value"):
        w["a.%s.weight" % name] = torch.randn(hid, hid, generator=g) * 0.05
        w["a.%s.bias" % name] = torch.randn(hid, generator=g) * 0.1
    xq, xk = torch.randn(n, lq, hid, generator=g), torch.randn(n, lk, hid, generator=g)
    lens = torch.randint(1, lk + 1, (n,), generator=g)
    lens[0] = lk
    mk = (torch.arange(lk)[None] < lens[:, None]).float()
    if full_mask:
        lq_ = torch.randint(1, lq + 1, (n,), generator=g)
        mask3 = (torch.arange(lq)[None] < lq_[:, None]).float().unsqueeze(2) * mk.unsqueeze(1)
    else:
        mask3 = mk.unsqueeze(1)
    w64 = {k: v.double() for k, v in w.items()}
    want = O.multi_head_attention(xq.double(), xk.double(), mask3.double(), w64, "a", nh)
    want32 = O.multi_head_attention(xq, xk, mask3, w, "a", nh)
    xq16, xk16 = ops.split_rows(xq.view(-1, hid).to(DEV)), ops.split_rows(xk.view(-1, hid).to(DEV))
    wq = ops.split_rows(w["a.query.weight"].to(DEV))
    wkv = ops.split_rows(torch.cat([w["a.key.weight"], w["a.value.weight"]]).to(DEV))
    _, q16, _ = ops.linear_tc_ex(xq16, wq, w["a.query.bias"].to(DEV), want_f32=False, out16_cols=hid)
    _, k16, vt16 = ops.linear_tc_ex(xk16, wkv, torch.cat([w["a.key.bias"], w["a.value.bias"]]).to(DEV), want_f32=False,
                                    out16_cols=hid, vt_col0=hid, vt_seq=lk)
    out, o16 = ops.attention_tc(q16, 0, k16, 0, vt16, mask3.to(DEV), n, lq, lk, hid, nh, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    got = out.view(n, lq, hid).cpu().double()
    # Query rows whose mask row is all zero (padded clips as cross-attention queries) add -10000 to EVERY logit: in
    # fp32 that quantises the logits to multiples of ulp(1e4) = 9.8e-4, so float64 is not
    # the yardstick for them -- the reference's own fp32 arithmetic is, and two fp32 implementations agree on such
    # rows only to ~1e-3 / sqrt(L) relative (a last-bit difference of a score flips its rounding to the next quantum).
    dead = (mask3.sum(2) == 0).expand(n, lq) if mask3.shape[1] == lq else torch.zeros(n, lq, dtype=torch.bool)
    live = ~dead
    err = float((got - want)[live].abs().max())
    err32 = float((want32.double() - want)[live].abs().max())
    print("attention_tc n=%d L=%d/%d H=%d: max abs err vs fp64 %.3g (torch CPU fp32 %.3g), |out| max %.3g"
          % (n, lq, lk, hid, err, err32, float(want.abs().max())))
    assert err <= 6e-6 * max(1.0, float(want.abs().max()))
    if dead.any():
        e_dead = float((got - want32.double())[dead].abs().max())
        q_dead = float((want32.double() - want)[dead].abs().max())
        print("   fully masked query rows: max abs diff to torch CPU fp32 %.3g (fp32 vs fp64 on those rows: %.3g)"
              % (e_dead, q_dead))
        assert e_dead <= 2 * q_dead + 2e-5  # as close to the fp32 reference as fp32 itself is to float64
    assert float((join(o16).cpu().view(n, lq, hid) - got).abs().max()) <= 3e-7 * float(got.abs().max())


def tvr_model(ctx_mode="video_sub", hidden=768, max_ctx_l=128, video_dim=3072):
    from tvretrieval_b200.model_xml import XML, xml_base_config
    cfg = copy.deepcopy(xml_base_config)
    cfg.update(hidden_size=hidden, max_ctx_l=max_ctx_l, max_desc_l=30, visual_input_size=video_dim, ctx_mode=ctx_mode)
    if ctx_mode != "video_sub":
        cfg.update(merge_two_stream=False, cross_att=False)
    torch.manual_seed(2018)
    model = XML(cfg).eval()
    weights = {k: v.clone() for k, v in model.state_dict().items()}
    return cfg, model.to(DEV), weights


@pytest.mark.parametrize("ctx_mode,hidden,length,video_dim", [("video_sub", 768, 128, 3072), ("video", 768, 32, 2048),
                                                              ("video_sub", 256, 100, 2048),
                                                              ("video_sub", 1024, 256, 3072)])
def test_context_encoder_on_tensor_cores(ctx_mode, hidden, length, video_dim):
    """XML.encode_context on the tensor-core pipeline (the default) vs the oracle in float64 at TVR dims, padded rows
    included (they feed the ConvSE taps); the exact-fp32 SIMT path and torch's own CPU fp32
    are measured by the same yardstick."""
    from tvretrieval_b200.synthetic import corpus_batch, corpus_lengths
    cfg, model, weights = tvr_model(ctx_mode, hidden, length, video_dim)
    n = 24 if length >= 128 else 40
    lens = corpus_lengths(n, length, seed=7)
    video, sub, mask = corpus_batch(lens, 0, n, video_dim, 768, DEV, seed=7, video_split=2048 if video_dim == 3072 else None)
    with torch.no_grad():
        assert model._tc_context_ok(video, model.context_precision)
        got = model.encode_context(video, mask, sub, mask)
        model.context_precision = "f32"
        simt = model.encode_context(video, mask, sub, mask)
        w64 = {k: v.double() for k, v in weights.items()}
        want = O.encode_context(dict(cfg), w64, video.cpu().double(), mask.cpu().double(), sub.cpu().double(),
                                mask.cpu().double())
        w32 = O.encode_context(dict(cfg), weights, video.cpu(), mask.cpu(), sub.cpu(), mask.cpu())
    valid = mask.cpu().bool()
    for name, a, b, c, d in zip(("video_feat1", "video_feat2", "sub_feat1", "sub_feat2"), got, simt, want, w32):
        if c is None:
            assert a is None
            continue
        a, b = a.cpu().double(), b.cpu().double()
        e_tc, e_simt, e_cpu = (float((x - c)[valid].abs().max()) for x in (a, b, d.double()))
        print("%s H=%d L=%d, valid clips: max abs err vs fp64 -- tensor cores %.3g, SIMT fp32 %.3g, torch CPU fp32 %.3g"
              % (name, hidden, length, e_tc, e_simt, e_cpu))
        assert e_tc <= 3e-5, name
        # padded clips: feat2 rows go through fully masked cross-attention rows, whose fp32 -10000 add quantises the
        # logits (see test_attention_tc) -- yardstick = the reference's fp32 arithmetic, agreement ~1e-4
        if (~valid).any():
            p_tc, p_simt = (float((x - d.double())[~valid].abs().max()) for x in (a, b))
            print("   padded clips: max abs diff to torch CPU fp32 -- tensor cores %.3g, SIMT fp32 %.3g (fp32 vs fp64 "
                  "there: %.3g)" % (p_tc, p_simt, float((d.double() - c)[~valid].abs().max())))
            assert p_tc <= 3e-4, name
