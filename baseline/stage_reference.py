#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for `bench.py --impl reference`.

The reference (slowkow/harmonypy, /root/reference) is a pure-Python package whose build backend (hatchling) is not
in this image's offline wheelhouse, so

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

fails with "No module named 'hatchling'" (recorded in DESIGN.md).  What that install would have produced for a
pure-Python package is a byte-for-byte copy of its module files; this script makes exactly that copy into the
git-ignored `baseline/_ref/` (it travels to the GPU box with the working tree, it never enters the history).
Run by `__graft_entry__.build()` whenever /root/reference is present.  Nothing under harmonypy_b200/ imports it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "harmonypy")


def stage(src="/root/reference/harmonypy", verbose=True):
    if not os.path.isdir(src):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    digest = hashlib.sha1()
    for f in sorted(os.listdir(src)):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
            digest.update(open(os.path.join(src, f), "rb").read())
    with open(os.path.join(HERE, "_ref", "STAGED_FROM"), "w") as fh:
        fh.write(f"{src}\nsha1 of the module files: {digest.hexdigest()}\n")
    if verbose:
        print(f"[stage_reference] {src} -> {DST} ({digest.hexdigest()[:12]})")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
