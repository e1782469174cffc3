"""Import shim used ONLY by tests/golden/make_golden.py: the reference only touches `h5py.File`."""


class File:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py stub: synthetic datasets only")
