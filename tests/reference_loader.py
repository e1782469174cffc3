"""Imports the UNMODIFIED reference (jayleicn/TVRetrieval) from baseline/_ref -- TEST / BENCH INFRASTRUCTURE.

baseline/_ref is produced in the build container by tools/vendor_reference.py (a byte-for-byte copy of the
reference's Python sources + the easydict / h5py shims of SURVEY.md Appendix E); it is git-ignored but travels to the
GPU box.  Only tests/ and the reference arms of bench.py use this module; the product never imports the reference.
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(REPO, "baseline", "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "baselines", "crossmodal_moment_localization", "model_xml.py"))


def load():
    """-> namespace with the reference's modules: model_xml, inference (crossmodal_moment_localization), EasyDict."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python tools/vendor_reference.py` in the build container")
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # reference inference.py:289,293 uses the alias numpy 2 removed
    for p in (REF_ROOT, os.path.join(REF_ROOT, "_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    from easydict import EasyDict
    from baselines.crossmodal_moment_localization import inference, model_xml
    assert os.path.abspath(model_xml.__file__).startswith(REF_ROOT), model_xml.__file__
    return types.SimpleNamespace(model_xml=model_xml, inference=inference, EasyDict=EasyDict, XML=model_xml.XML,
                                 xml_base_config=model_xml.xml_base_config)


class SectionTimer:
    """Stands in for `tqdm` inside the reference's inference module (a progress bar, not arithmetic): records the
    wall time (CUDA-synchronised) spent inside each tqdm-wrapped loop, keyed by its `desc` -- "Computing q embedding"
    is the tensor section of compute_query2ctx_info (inference.py:302-389), the "[VR] ..." / "[VCMR] ..." loops are
    its host section (:391-445)."""

    def __init__(self, sync=None):
        self.sync = sync or (lambda: None)
        self.seconds = {}

    def __call__(self, iterable, desc="", total=None, **kw):
        def gen():
            self.sync()
            t0 = time.perf_counter()
            for item in iterable:
                yield item
            self.sync()
            self.seconds[desc] = self.seconds.get(desc, 0.0) + time.perf_counter() - t0
        return gen()


class QueryDataset:
    """Duck-typed stand-in for StartEndEvalDataset in QUERY mode (reference start_end_dataset.py:171-343; protocol in
    SURVEY.md section 8b): items come from padded (Nq, Lq, Dq) features + masks."""

    def __init__(self, query_feat, query_mask, n_videos, max_ctx_l):
        import torch
        self.lens = query_mask.sum(1).to(torch.long).tolist()
        self.feat = query_feat
        self.video2idx = {"vid_%05d" % i: i for i in range(n_videos)}
        self.max_ctx_len = max_ctx_l
        self.query_data = [dict(desc_id=i, desc="q%d" % i) for i in range(len(self.lens))]

    def set_data_mode(self, mode):
        assert mode == "query"

    def load_gt_vid_name_for_query(self, flag):
        assert not flag, "synthetic queries carry no ground-truth video"

    def __len__(self):
        return len(self.lens)

    def __getitem__(self, i):
        return dict(meta=dict(desc_id=i, desc="q%d" % i, vid_name=None),
                    model_inputs=dict(query_feat=self.feat[i, :self.lens[i]]))
