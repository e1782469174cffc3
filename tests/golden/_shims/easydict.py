"""Import shim used ONLY by tests/golden/make_golden.py when it imports the reference in the
build container (the reference does `from easydict import EasyDict`; the package is not installed)."""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
