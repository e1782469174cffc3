#!/usr/bin/env python
"""Vendors the UNMODIFIED reference (jayleicn/TVRetrieval, Python sources only) into baseline/_ref/ so that it travels
to the GPU box with the gpurun snapshot (baseline/_ref is git-ignored, not gpurun-ignored; /root/reference does not
exist on the box).  Build-container only:

    python tools/vendor_reference.py [--src /root/reference]

Copies every *.py of baselines/, utils/, standalone_eval/ (+ the package __init__.py files) byte for byte, and the
two import shims the reference needs on this image (easydict, h5py.File -- SURVEY.md Appendix E) into
baseline/_ref/_shims/.  `pip install --target baseline/_ref /root/reference` is not possible: the reference ships no
setup.py / pyproject.toml (it is run from its checkout with PYTHONPATH, setup.sh:1-8).  tests/reference_loader.py
imports it from there; nothing under tvretrieval_b200/ does."""
import argparse
import hashlib
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(REPO, "baseline", "_ref")
PACKAGES = ("baselines", "utils", "standalone_eval")


def vendor(src="/root/reference", dest=DEST, quiet=False):
    if not os.path.isdir(src):
        raise SystemExit("reference checkout not found at %s" % src)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    n, digest = 0, hashlib.sha256()
    top_init = os.path.join(src, "__init__.py")
    files = [top_init] if os.path.exists(top_init) else []
    for pkg in PACKAGES:
        for root, _, names in os.walk(os.path.join(src, pkg)):
            files += [os.path.join(root, f) for f in names if f.endswith(".py")]
    for path in sorted(files):
        rel = os.path.relpath(path, src)
        out = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(path, out)
        with open(path, "rb") as fh:
            digest.update(rel.encode() + fh.read())
        n += 1
    shutil.copytree(os.path.join(REPO, "tests", "golden", "_shims"), os.path.join(dest, "_shims"))
    with open(os.path.join(dest, "VENDORED.txt"), "w") as fh:
        fh.write("unmodified copy of %d .py files of %s (sha256 over path+content: %s)\n" % (n, src, digest.hexdigest()))
    if not quiet:
        print("vendored %d files into %s" % (n, dest))
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    vendor(a.src)
