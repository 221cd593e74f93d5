"""Swap the reference's correlation classes for the CUDA ones without editing it.

The reference binds the classes by name at import time (``from .corr import
CorrBlock``), at these sites: core/raft.py:8, core/l2l.py:8,
core/gma_network.py:7, core/gma_l2l.py:7, GMA/core/network.py:7.  ``patch_reference``
rebinds the attribute in every already-imported module that holds the reference
class, and installs the ``alt_cuda_corr`` shim.
"""
from __future__ import annotations

import sys

from . import alt_cuda_corr as _alt
from .corr import AlternateCorrBlock, CorrBlock

_NAMES = {"CorrBlock": CorrBlock, "AlternateCorrBlock": AlternateCorrBlock}
_ORIGINALS = []          # (module, attribute, original object) for unpatch_reference()


def patch_reference(verbose: bool = False):
    """Rebind CorrBlock / AlternateCorrBlock in every imported module whose attribute of
    that name is a class defined in a module called ``corr`` / ``gma_corr`` (i.e. the
    reference's).  Returns the list of (module, attribute) pairs patched."""
    _alt.install()
    patched = []
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or mod_name.startswith("flow_supervisor_b200"):
            continue
        for attr, repl in _NAMES.items():
            cur = getattr(mod, attr, None)
            if cur is None or cur is repl or not isinstance(cur, type):
                continue
            origin = getattr(cur, "__module__", "") or ""
            if origin.split(".")[-1] in ("corr", "gma_corr"):
                _ORIGINALS.append((mod, attr, cur))
                setattr(mod, attr, repl)
                patched.append((mod_name, attr))
        if getattr(mod, "alt_cuda_corr", None) is not None and mod_name.split(".")[-1] in ("corr", "gma_corr"):
            mod.alt_cuda_corr = _alt
    if verbose:
        for m, a in patched:
            print(f"[flowcorr] patched {m}.{a}")
    return patched


def unpatch_reference():
    """Undo patch_reference(): restore the reference's own classes (the alt_cuda_corr shim
    stays registered; the reference treats that module as optional, corr.py:5-9)."""
    while _ORIGINALS:
        mod, attr, orig = _ORIGINALS.pop()
        setattr(mod, attr, orig)
