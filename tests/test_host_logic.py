"""CPU-only checks: the C-ABI library loads and exports every symbol include/xmlb200.h declares (no compute
calls), the drop-in model keeps the reference's checkpoint layout, and the host-side helpers behave like the
reference's."""
import copy
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import xml_oracle as O
from tests.golden_io import CASE_NAMES, GoldenCase


def test_library_exports_every_declared_symbol():
    from tvretrieval_b200 import _lib, build
    build.build()
    protos = _lib.parse_header()
    assert len(protos) >= 16
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(handle, name), "libxmlb200.so does not export " + name
    lib = _lib.lib()
    assert lib.xmlb_version() >= 100
    assert lib.xmlb_last_error() is not None


def test_invalid_arguments_return_error_codes_not_crashes():
    from tvretrieval_b200 import _lib
    lib = _lib.lib()
    rc = lib.xmlb_topk_rows(None, None, 0, 1, 10, 5, 1.0, 0, 0, None, None, None, None)
    assert rc < 0 and b"null" in lib.xmlb_last_error()
    rc = lib.xmlb_span_topk(None, None, None, None, 1, 1, 8, 2, 16, 5000, 0, 1, None, None, None)
    assert rc < 0
    with pytest.raises(_lib.XmlbError):
        _lib.check(rc, "xmlb_span_topk")


@pytest.mark.parametrize("name", CASE_NAMES)
def test_model_checkpoint_layout_and_seeded_init(name):
    from tvretrieval_b200.model_xml import XML, AttrDict
    g = GoldenCase(name)
    torch.manual_seed(2018)
    model = XML(AttrDict(g.cfg))
    sd = model.state_dict()
    assert list(sd.keys()) == list(g.weights.keys())
    for k, v in sd.items():
        assert v.shape == g.weights[k].shape, k
        assert torch.equal(v, g.weights[k]), "seeded init differs from the reference at " + k
    model.load_state_dict(g.weights)
    for attr in ("query_input_proj", "query_encoder", "query_pos_embed", "modular_vector_mapping", "config",
                 "use_video", "use_sub"):
        assert hasattr(model, attr)
    if g.cfg["merge_two_stream"]:
        out = model.merged_st_predictor(torch.zeros(3, 1, 9))  # callable on (n, 1, L) like profile_main.py uses it
        assert out.shape == (3, 1, 9)


def test_model_has_no_cpu_fallback():
    from tvretrieval_b200._lib import XmlbError
    from tvretrieval_b200.model_xml import XML, AttrDict
    g = GoldenCase("video_only_svmr")
    model = XML(AttrDict(g.cfg)).eval()
    with pytest.raises(XmlbError):
        model.encode_query(g.query_feat, g.query_mask)
    with torch.no_grad(), pytest.raises(XmlbError):
        model.encode_query(g.query_feat, g.query_mask)
    # the training step has no CPU path either (golden batch on CPU tensors)
    from tests.golden_io import TrainCase
    tc = TrainCase("video_only_svmr", "plain")
    with pytest.raises(XmlbError):
        model(**tc.inputs)


def test_bert_adam_surface_and_schedules():
    """Constructor validation, schedules and get_lr() of the fused optimizer mirror reference optimization.py
    (no GPU needed: step() itself is covered by tests/test_gpu_train.py)."""
    from oracle import xml_oracle as O
    from tvretrieval_b200._lib import XmlbError
    from tvretrieval_b200.optimization import SCHEDULES, BertAdam
    p = torch.nn.Parameter(torch.zeros(3))
    for bad in (dict(lr=-1.0), dict(lr=1e-3, schedule="nope"), dict(lr=1e-3, b1=1.0), dict(lr=1e-3, b2=-0.1),
                dict(lr=1e-3, e=-1.0), dict(lr=1e-3, warmup=1.5)):
        with pytest.raises(ValueError):
            BertAdam([p], **bad)
    for name in ("warmup_linear", "warmup_constant", "warmup_cosine", "none"):
        sched = SCHEDULES[name](warmup=0.1, t_total=50)
        for step in (0, 1, 4, 5, 6, 25, 49, 50, 60):
            assert abs(sched.get_lr(step) - O.lr_multiplier(name, 0.1, 50, step)) < 1e-12, (name, step)
    assert SCHEDULES["warmup_linear"](warmup=0.1, t_total=-1).get_lr(7) == 1.0
    opt = BertAdam([p], lr=1e-3, warmup=0.1, t_total=50)
    assert opt.get_lr() == [0]
    opt.step()  # no gradients: nothing to do, no GPU touched
    p.grad = torch.ones(3)
    with pytest.raises(XmlbError):  # CPU parameters: no fallback
        opt.step()


def test_unsupported_variants_fail_loudly():
    from tvretrieval_b200.model_xml import XML, xml_base_config
    for override in (dict(encoder_type="lstm"), dict(span_predictor_type="cat_linear"),
                     dict(stack_conv_predictor_conv_kernel_sizes=[3, 5])):
        cfg = copy.deepcopy(xml_base_config)
        cfg.update(override)
        with pytest.raises(NotImplementedError):
            XML(cfg)


def test_band_mask_and_collate_match_reference_semantics():
    from tvretrieval_b200 import inference as I
    for length, lo, hi in ((100, 2, 16), (128, 2, 16), (7, 0, 3)):
        m = I.generate_min_max_length_mask((5, 8, length, length), lo, hi)
        assert m.shape == (1, 1, length, length)
        assert np.array_equal(m[0, 0], O.band_mask(length, lo, hi))
    assert int(I.generate_min_max_length_mask((100, 100), 2, 16).sum()) == 1281
    assert int(I.generate_min_max_length_mask((128, 128), 2, 16).sum()) == 1673
    seqs = [torch.randn(3, 4), torch.randn(5, 4), torch.randn(1, 4)]
    batch = [dict(meta=dict(i=i), model_inputs=dict(query_feat=s)) for i, s in enumerate(seqs)]
    metas, out = I.start_end_collate(batch)
    feat, mask = out["query_feat"]
    assert feat.shape == (3, 5, 4) and mask.dtype == torch.float32
    assert mask.sum(1).tolist() == [3, 5, 1] and torch.equal(feat[0, :3], seqs[0]) and (feat[0, 3:] == 0).all()
    x = I.prepare_batch_inputs(out, device="cpu")
    assert set(x) == {"query_feat", "query_mask"}
    cat = I.cat_tensor([torch.ones(2, 3, 2), torch.ones(1, 5, 2)])
    assert cat.shape == (3, 5, 2) and (cat[:2, 3:] == 0).all()
    with pytest.raises(ValueError):
        I.cat_tensor([torch.ones(2)])


def _check_packing(pk, lens):
    """Invariants of the ragged corpus layout: every video with clips appears once, its rows are contiguous, tiles
    hold whole videos in <= 256 rows, the start bitmap marks the first column of each video."""
    order, row_start = pk.order.numpy(), pk.row_start.numpy()
    meta, starts = pk.tile_meta.numpy(), pk.tile_starts.numpy().view(np.uint32)
    assert sorted(order.tolist()) == [v for v in range(len(lens)) if lens[v] > 0]
    assert sorted(pk.empty.tolist()) == [v for v in range(len(lens)) if lens[v] == 0]
    assert np.array_equal(np.diff(row_start), lens[order]) and row_start[0] == 0
    assert pk.n_rows == int(lens.sum()) and pk.max_len == int(lens.max())
    o = 0
    for t, (row, first, used, nvid) in enumerate(meta):
        assert first == o and row == row_start[o] and 0 < used <= 256 and 0 < nvid <= 32
        assert used == row_start[o + nvid] - row_start[o]
        cols = set((row_start[o:o + nvid] - row).tolist())
        got = {c for c in range(256) if (starts[t, c // 32] >> (c % 32)) & 1}
        assert got == cols
        o += nvid
    assert o == len(order)
    assert torch.equal(pk.order_full, torch.cat([pk.order, pk.empty]))


def test_corpus_packing_and_incremental_append():
    from tvretrieval_b200.engine import CorpusPacking
    g = torch.Generator().manual_seed(0)

    def masks(n, length):
        lens = torch.randint(0, length + 1, (n,), generator=g)
        lens[0] = length
        return (torch.arange(length)[None] < lens[:, None]).float(), lens.numpy()

    m1, l1 = masks(300, 128)
    m2, l2 = masks(77, 128)
    pk = CorpusPacking(m1)
    _check_packing(pk, l1)
    assert pk.fill > 0.95
    pk.append(CorpusPacking(m2), 300)
    _check_packing(pk, np.concatenate([l1, l2]))
    restored = CorpusPacking.from_state({k: getattr(pk, k) for k in CorpusPacking.STATE}, pk.max_len)
    _check_packing(restored, np.concatenate([l1, l2]))


def test_packed_query_layout():
    """Index tables of the packed query encoder (XML.encode_query_packed): valid tokens in query order."""
    from tvretrieval_b200.model_xml import packed_layout
    lens, width = [3, 0, 5, 1, 9], 5
    rows, pos, cu, max_len = packed_layout(lens, width)
    assert cu.tolist() == [0, 3, 3, 8, 9, 14] and max_len == 5  # 9 is clamped to the padded width
    assert pos.tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 0, 0, 1, 2, 3, 4]
    assert rows.tolist() == [0, 1, 2, 10, 11, 12, 13, 14, 15, 20, 21, 22, 23, 24]
    mask = (np.arange(width)[None] < np.minimum(lens, width)[:, None])
    assert np.array_equal(np.flatnonzero(mask.reshape(-1)), rows)  # == boolean-mask gather of the padded layout
    rows, pos, cu, max_len = packed_layout([], 5)
    assert len(rows) == 0 and cu.tolist() == [0] and max_len == 1


def test_eval_retrieval_matches_reference_evaluator():
    """tvretrieval_b200.eval_metrics.eval_retrieval == standalone_eval.eval.eval_retrieval (reference
    standalone_eval/eval.py:83-276) on the golden submission: same keys in the same order, same rounded values."""
    import json
    from tests.golden_io import GOLDEN_DIR
    from tvretrieval_b200.eval_metrics import eval_retrieval
    z = json.load(open(os.path.join(GOLDEN_DIR, "eval_metrics.json")))
    for key, want in z["expected"].items():
        use_type, thds = key.split("/")
        got = eval_retrieval(copy.deepcopy(z["submission"]), copy.deepcopy(z["ground_truth"]),
                             iou_thds=tuple(float(t) for t in thds.split(",")), verbose=False,
                             use_desc_type=use_type == "True")
        got = [[task, [[k, v] for k, v in m.items()]] for task, m in got.items()]
        assert got == want, key
    # a query without predictions (the reference cannot evaluate it) simply never hits
    sub = copy.deepcopy(z["submission"])
    sub["VCMR"][0]["predictions"] = []
    m = eval_retrieval(sub, z["ground_truth"], verbose=False)
    assert set(m) == {"VCMR", "SVMR", "VR", "VCMR_by_type", "SVMR_by_type", "VR_by_type"}


def test_operand_layout_helpers():
    """ops.kblock_rows (k-blocked and shared-memory-image layouts of a 16-bit operand) and ops.span_clip_rows are
    plain index arithmetic: checked here against their definitions, element by element."""
    from tvretrieval_b200 import ops
    rows, k = 40, 96
    t = torch.arange(rows * k, dtype=torch.int16).view(rows, k)
    kb = ops.kblock_rows(t)
    assert kb.shape == (k // 32, rows, 32) and kb.is_contiguous()
    for b in range(k // 32):
        assert torch.equal(kb[b], t[:, 32 * b:32 * b + 32])
    img = ops.kblock_rows(t, swizzle=True)
    assert img.shape == (k // 32, rows, 4, 8) and img.is_contiguous()
    for r in range(rows):
        s = (r >> 1) & 3  # SWIZZLE_64B: 16-byte piece c of row r lands at position c ^ ((r >> 1) & 3)
        for c in range(4):
            assert torch.equal(img[:, r, c ^ s], kb[:, r, 8 * c:8 * c + 8])
    # rows a ConvSE of `ksize` taps can read from an unmasked clip: last unmasked clip + 1 + ksize // 2, at most L
    mask = torch.zeros(5, 12)
    mask[0, :12] = 1
    mask[1, :3] = 1
    mask[2, :10] = 1
    mask[4, 5] = 1  # (not a prefix mask: still covered)
    assert ops.span_clip_rows(mask, 5).tolist() == [12, 5, 12, 2, 8]
    assert ops.span_clip_rows(mask, 1).tolist() == [12, 3, 10, 0, 6]


def test_host_upload_pieces_cover_the_block_and_ramp_up():
    """VCMRSearcher._piece_bounds: contiguous cover of [0, n); from host buffers the pieces at most double (so that an
    upload hides behind the work on the piece before it) up to encode_chunk, and there is no tiny last piece."""
    from tvretrieval_b200.engine import VCMRSearcher

    class Cfg:
        encode_chunk, min_piece = 2048, 256

    for n in (1, 100, 255, 256, 300, 1250, 2048, 5000, 10000, 16384):
        for host in (True, False):
            cuts = VCMRSearcher._piece_bounds(Cfg, n, host)
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and all(hi > lo for lo, hi in cuts)
            sizes = [hi - lo for lo, hi in cuts]
            if host:
                assert all(b <= 2 * a + a // 2 for a, b in zip(sizes, sizes[1:]))  # (the last piece absorbs a remainder)
                assert max(sizes) <= Cfg.encode_chunk + Cfg.encode_chunk // 2
                assert sizes[0] <= 256
    assert [hi - lo for lo, hi in VCMRSearcher._piece_bounds(Cfg, 10000, True)] == [256, 512, 1024, 2048, 2048, 2048, 2064]
