#!/usr/bin/env python
"""Install the UNMODIFIED reference (its PyTorch model code) into baseline/_ref/ --
the reference arm of bench.py and the callers of the GPU model tests.

The reference is a research tree with no setup.py / pyproject (``pip install
/root/reference`` has nothing to build: recorded in DESIGN.md), so "install" is a verbatim
copy of the files its public API is made of:

    /root/reference/pytorch/core/**.py        -> baseline/_ref/pytorch/core/
    /root/reference/pytorch/GMA/core/**.py    -> baseline/_ref/pytorch/GMA/core/

baseline/_ref/ is git-ignored (reference sources never enter this repository's history)
but NOT gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not
exist.  A manifest with the sha256 of every file proves the copy is unmodified
(tests/test_reference_install.py re-hashes it against /root/reference when that is mounted).

On a machine without /root/reference this is a no-op that reports what is there.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pytorch"
DST = os.path.join(HERE, "_ref", "pytorch")
TREES = ["core", os.path.join("GMA", "core")]
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def installed() -> bool:
    return os.path.exists(os.path.join(DST, "core", "corr.py")) and os.path.exists(MANIFEST)


def install(verbose: bool = False) -> str | None:
    """-> path to add to sys.path (the directory that holds ``core``), or None."""
    if not os.path.isdir(SRC):
        return DST if installed() else None
    files = {}
    for tree in TREES:
        for root, _dirs, names in os.walk(os.path.join(SRC, tree)):
            for n in sorted(names):
                if not n.endswith(".py"):
                    continue
                s = os.path.join(root, n)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not os.path.exists(d) or sha256(d) != sha256(s):
                    shutil.copyfile(s, d)
                    if verbose:
                        print("copied", rel)
                files[rel] = sha256(d)
    commit = None
    try:
        commit = json.load(open("/root/reference/.SUBMODULES.json")).get("commit")
    except Exception:                                   # noqa: BLE001
        pass
    with open(MANIFEST, "w") as f:
        json.dump({"source": SRC, "commit": commit, "files": files}, f, indent=1, sort_keys=True)
    return DST


def path() -> str:
    """sys.path entry of the installed reference; raises when it was never installed."""
    if not installed():
        raise FileNotFoundError(f"{DST} missing: run python baseline/install_ref.py where /root/reference exists")
    return DST


if __name__ == "__main__":
    p = install(verbose="-v" in sys.argv)
    print(p if p else "reference absent and nothing installed")
