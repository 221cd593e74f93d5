"""GPU: the drop-in CorrBlock inside the full GRU loop.  north_star: "final flow within
0.01 px mean EPE after 12 GRU iterations" against the reference's PyTorch CorrBlock on
identical synthetic inputs and random-init weights (seed 1234, pytorch/train.py:347).

The caller is oracle/raft_model.py (restated RAFT, pinned to the live reference by
tests/test_oracle_raft.py); the reference block is oracle/corr_torch.py (the same torch
library calls as corr.py, bit-exact to it on CPU) running on the same GPU.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import flow_supervisor_b200 as fsb
    from oracle import corr_torch, raft_model
    from oracle.make_golden_raft import synth_pair
    torch.backends.cudnn.allow_tf32 = False          # both arms run the same fp32 convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)
    model = raft_model.Raft().eval().cuda()
    return fsb, corr_torch, model, synth_pair


def epe(a, b):
    return float(torch.sqrt(((a - b) ** 2).sum(1)).mean())


@pytest.mark.parametrize("math", ["3xbf16", "fp32"])
def test_raft_12_iters_epe_small(env, math):
    fsb, corr_torch, model, synth_pair = env
    im1, im2 = (t.cuda() for t in synth_pair(128, 160))
    fsb.CorrBlock.math = math
    try:
        with torch.no_grad():
            low_r, up_r = model(im1, im2, iters=12, corr_block=corr_torch.TorchCorrBlock)
            low_o, up_o = model(im1, im2, iters=12, corr_block=fsb.CorrBlock)
    finally:
        fsb.CorrBlock.math = "auto"
    assert epe(up_o, up_r) <= 0.01, epe(up_o, up_r)
    # and against the flow the UNMODIFIED reference produced on CPU (golden; fp16-stored)
    g = np.load(os.path.join(ROOT, "tests", "golden", "raft_seed1234_128x160.npz"))
    gold = torch.from_numpy(g["flow_up"].astype(np.float32)).cuda()
    assert epe(up_o, gold) <= 0.02, epe(up_o, gold)
    assert float((low_o - torch.from_numpy(g["flow_low"]).cuda()).abs().max()) <= 5e-3


def test_raft_12_iters_epe_sintel_size(env):
    """configs[1] geometry: 436x1024 padded to 440x1024 -> 55x128 tokens, batch 2."""
    fsb, corr_torch, model, synth_pair = env
    a = [synth_pair(440, 1024, seed=s) for s in (1, 2)]
    im1 = torch.cat([p[0] for p in a]).cuda()
    im2 = torch.cat([p[1] for p in a]).cuda()
    with torch.no_grad():
        _, up_r = model(im1, im2, iters=12, corr_block=corr_torch.TorchCorrBlock)
        _, up_o = model(im1, im2, iters=12, corr_block=fsb.CorrBlock)
    assert torch.isfinite(up_o).all()
    assert epe(up_o, up_r) <= 0.01, epe(up_o, up_r)


def test_raft_bf16_mode_epe_stated_separately(env):
    """bf16 arithmetic mode (bf16 operands, fp32 accumulate): tolerance 0.05 px mean EPE."""
    fsb, corr_torch, model, synth_pair = env
    im1, im2 = (t.cuda() for t in synth_pair(128, 160))
    fsb.CorrBlock.math = "bf16"
    try:
        with torch.no_grad():
            _, up_r = model(im1, im2, iters=12, corr_block=corr_torch.TorchCorrBlock)
            _, up_o = model(im1, im2, iters=12, corr_block=fsb.CorrBlock)
    finally:
        fsb.CorrBlock.math = "auto"
    assert epe(up_o, up_r) <= 0.05, epe(up_o, up_r)


def test_training_step_gradients_through_the_gru(env):
    """One supervised step (sequence loss over 4 iterations, train.py:60-96 weighting):
    gradients reaching fnet through our lookup/build backward match torch autograd."""
    fsb, corr_torch, model, synth_pair = env
    im1, im2 = (t.cuda() for t in synth_pair(128, 160))
    target = torch.zeros(1, 2, 128, 160, device="cuda")
    grads = {}
    for name, blk in (("ref", corr_torch.TorchCorrBlock), ("ours", fsb.CorrBlock)):
        model.zero_grad(set_to_none=True)
        preds = model(im1, im2, iters=4, corr_block=blk, return_all=True)
        loss = sum(0.8 ** (len(preds) - 1 - i) * (p - target).abs().mean() for i, p in enumerate(preds))
        loss.backward()
        grads[name] = {k: p.grad.detach().clone() for k, p in model.fnet.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    assert grads["ours"].keys() == grads["ref"].keys() and len(grads["ref"]) > 0
    gmax = max(float(g.abs().max()) for g in grads["ref"].values())
    for k, gr in grads["ref"].items():
        go = grads["ours"][k]
        # biases in front of an instance norm have mathematically zero gradient (1e-8 noise):
        # scale every comparison by at least 1e-3 of the largest gradient in the encoder
        denom = max(float(gr.abs().max()), 1e-3 * gmax)
        assert float((go - gr).abs().max()) / denom <= 2e-3, (k, float((go - gr).abs().max()) / denom)
