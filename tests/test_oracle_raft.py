"""CPU: the restated caller (oracle/raft_model.py) reproduces the reference RAFT --
weights under the same seed, and the 12-iteration flow of the committed golden pair."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import corr_torch, raft_model                    # noqa: E402
from oracle.make_golden_raft import synth_pair               # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "raft_seed1234_128x160.npz")
REF = "/root/reference/pytorch"


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(1234)
    return raft_model.Raft().eval()


def test_same_seed_same_weights_as_reference(model):
    g = np.load(GOLD)
    sd = model.state_dict()
    assert sum(v.numel() for v in sd.values()) == int(g["n_params"])
    checksum = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(checksum - float(g["state_checksum"])) <= 1e-9 * float(g["state_checksum"])


def test_flow_matches_reference_golden(model):
    """Oracle RAFT + oracle CorrBlock port == reference RAFT + reference CorrBlock (CPU)."""
    g = np.load(GOLD)
    im1, im2 = synth_pair(128, 160)
    with torch.no_grad():
        low, up = model(im1, im2, iters=12, corr_block=corr_torch.TorchCorrBlock)
    assert np.abs(low.numpy() - g["flow_low"]).max() <= 1e-4
    epe = np.sqrt(((up.numpy() - g["flow_up"].astype(np.float32)) ** 2).sum(1)).mean()
    assert epe <= 5e-3            # golden flow_up is stored as fp16 (ulp 0.016 px at |flow| in [16, 32))
    assert abs(float(up.double().mean()) - float(g["flow_up_mean"])) <= 1e-4


@pytest.mark.skipif(not os.path.isdir(REF + "/core"), reason="live reference only exists in the build container")
def test_state_dict_keys_and_values_equal_live_reference(model):
    sys.path.insert(0, REF)
    from core.raft import RAFT
    torch.manual_seed(1234)
    ref = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    a, b = ref.state_dict(), model.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
