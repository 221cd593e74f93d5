"""Test helper: the UNMODIFIED reference models (baseline/_ref, installed by
baseline/install_ref.py from /root/reference/pytorch/core) and the synthetic inputs of
SURVEY.md section 8(d).  Test infrastructure only."""
from __future__ import annotations

import argparse
import contextlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def reference_path() -> str:
    from baseline import install_ref
    p = install_ref.install()
    if p is None:
        pytest.skip("baseline/_ref not installed (run python baseline/install_ref.py where /root/reference exists)")
    return p


def core():
    """Import the reference's ``core`` package (raft.py, l2l.py, gma_network.py, ...)."""
    p = reference_path()
    if p not in sys.path:
        sys.path.insert(0, p)
    import core.corr, core.raft, core.l2l, core.gma_corr, core.gma_network, core.gma_l2l  # noqa: E401,F401
    import core as pkg
    return pkg


def raft_args(**kw):
    a = dict(small=False, mixed_precision=False, alternate_corr=False)
    a.update(kw)
    return argparse.Namespace(**a)


def gma_args(**kw):
    # pytorch/train_gma.py:350-355 defaults
    a = dict(small=False, mixed_precision=False, num_heads=1, position_only=False, position_and_content=False)
    a.update(kw)
    return argparse.Namespace(**a)


def synth_pair(h, w, seed=0, shift=(3, -2), batch=1):
    """Smooth random image and a copy shifted by `shift` px (non-trivial flow), in [0, 255]."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(batch, 3, h // 8 + 2, w // 8 + 2, generator=g)
    big = torch.nn.functional.interpolate(base, size=(h + 16, w + 16), mode="bicubic", align_corners=False)
    big = (255 * (big - big.min()) / (big.max() - big.min())).float()
    sx, sy = shift
    im1 = big[:, :, 8:8 + h, 8:8 + w].contiguous()
    im2 = big[:, :, 8 + sy:8 + sy + h, 8 + sx:8 + sx + w].contiguous()
    return im1, im2


@contextlib.contextmanager
def strict_fp32():
    """Both arms run the same fp32 (non-TF32) convolutions and matmuls."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


@contextlib.contextmanager
def patched():
    """The drop-in classes bound into every reference import site (flow_supervisor_b200.patch)."""
    import flow_supervisor_b200 as fsb
    names = fsb.patch_reference()
    try:
        yield names
    finally:
        fsb.unpatch_reference()


def epe(a, b):
    return float(torch.sqrt(((a - b) ** 2).sum(1)).mean())


def sequence_loss(flow_preds, flow_gt, valid, gamma=0.8, gamma2=1.0, max_flow=400):
    """pytorch/train.py:60-96 (student half weighted gamma^k, teacher half gamma2^k)."""
    n = len(flow_preds) // 2
    m = len(flow_preds) - n
    mag = torch.sum(flow_gt ** 2, dim=1).sqrt()
    mask = (valid >= 0.5) & (mag < max_flow)
    loss = 0.0
    for i in range(n):
        d = flow_preds[i] - flow_gt
        loss = loss + gamma ** (n - i - 1) * (mask[:, None] * (d ** 2 + 0.001 ** 2) ** 0.5).mean()
    for i in range(m):
        d = flow_preds[n + i] - flow_gt
        loss = loss + gamma2 ** (n - i - 1) * (mask[:, None] * (d ** 2 + 0.001 ** 2) ** 0.5).mean()
    return loss


def sequence_loss_unsup(flow_preds, gamma=0.8, unsup_weight=1.0):
    """pytorch/train.py:99-128: the student half regresses the teacher's last prediction."""
    n = len(flow_preds) // 2
    pseudo = flow_preds[-1].detach()
    loss = 0.0
    for i in range(n):
        d = flow_preds[i] - pseudo
        loss = loss + unsup_weight * gamma ** (n - i - 1) * ((d ** 2 + 0.001 ** 2) ** 0.5).mean()
    return loss


def l2l_batch(batch, crop=(368, 768), full=(432, 1024), ox=128, oy=32, seed=0, device="cuda"):
    """Inputs of one L2L call (train.py:270): augmented crops + the full frames they were cut
    from (here: the crop IS the window at (oy, ox) of the full frame, so student and teacher
    see consistent content), shared crop offset (l2l.py:86-87 reads ox[0] / oy[0])."""
    f1, f2 = synth_pair(full[0], full[1], seed=seed, batch=batch)
    c1 = f1[:, :, oy:oy + crop[0], ox:ox + crop[1]].contiguous()
    c2 = f2[:, :, oy:oy + crop[0], ox:ox + crop[1]].contiguous()
    oxs = torch.full((batch,), ox, dtype=torch.int64)
    oys = torch.full((batch,), oy, dtype=torch.int64)
    g = torch.Generator().manual_seed(seed + 1)
    flow_gt = torch.tensor([3.0, -2.0]).view(1, 2, 1, 1) + 0.5 * torch.randn(batch, 2, crop[0], crop[1], generator=g)
    valid = torch.ones(batch, crop[0], crop[1])
    return tuple(t.to(device) for t in (c1, c2, f1, f2, oxs, oys, flow_gt, valid))


def grad_dict(module):
    return {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}


def compare_grads(ours, ref, rel=2e-3):
    """Per-parameter max abs difference, scaled by max(|ref|_max, 1e-3 of the largest gradient)
    (biases in front of an instance norm have mathematically zero gradient)."""
    assert ours.keys() == ref.keys() and len(ref) > 0
    gmax = max(float(g.abs().max()) for g in ref.values())
    worst = 0.0
    for k, gr in ref.items():
        denom = max(float(gr.abs().max()), 1e-3 * gmax)
        e = float((ours[k] - gr).abs().max()) / denom
        worst = max(worst, e)
        assert e <= rel, (k, e)
    return worst


def ddp_l2l_worker(rank, world, port, out_path, n_iters):
    """One DDP rank of the config-3 step (NCCL, one process per GPU): every rank takes its slice
    of a 2*world batch; rank 0 saves the all-reduced gradients."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = core()
        torch.manual_seed(1234)
        model = pkg.l2l.L2L(raft_args()).cuda().train()
        model.freeze_bn()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
        per = 1
        data = l2l_batch(per * world, device="cpu")
        mine = tuple(t[rank * per:(rank + 1) * per].cuda() for t in data)
        c1, c2, f1, f2, ox, oy, gt, valid = mine
        with strict_fp32(), patched():
            preds = ddp(c1, c2, f1, f2, ox, oy, iters=n_iters)
            sequence_loss(preds, gt, valid).backward()
        if rank == 0:
            torch.save({k: v.cpu() for k, v in grad_dict(model).items()}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()
