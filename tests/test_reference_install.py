"""CPU: baseline/_ref is a verbatim copy of the reference's model code (what `bench.py --impl
reference` and the GPU model tests run), and the drop-in binds into every import site of it."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from baseline import install_ref  # noqa: E402


def test_install_is_verbatim():
    p = install_ref.install()
    if p is None:
        pytest.skip("no reference and no install")
    man = json.load(open(install_ref.MANIFEST))
    assert "core/corr.py" in man["files"] and "core/l2l.py" in man["files"] and "core/gma_network.py" in man["files"]
    for rel, digest in man["files"].items():
        assert install_ref.sha256(os.path.join(p, rel)) == digest, rel
        src = os.path.join(install_ref.SRC, rel)
        if os.path.exists(src):
            assert install_ref.sha256(src) == digest, f"{rel} differs from /root/reference"


def test_installed_reference_imports_and_patches():
    import refmodels as rm
    import flow_supervisor_b200 as fsb
    core = rm.core()
    ref_cls = core.corr.CorrBlock
    with rm.patched() as names:
        assert {m for m, _ in names} >= {"core.corr", "core.raft", "core.l2l", "core.gma_corr",
                                          "core.gma_network", "core.gma_l2l"}
        assert core.raft.CorrBlock is fsb.CorrBlock and core.gma_l2l.CorrBlock is fsb.CorrBlock
    assert core.raft.CorrBlock is ref_cls or core.raft.CorrBlock.__module__.endswith("corr")
