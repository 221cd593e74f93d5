"""Library-call port of the reference CorrBlock -- TEST INFRASTRUCTURE ONLY.

The reference's arithmetic for this path lives in PyTorch itself (third party,
not vendored; pinned by the reference at torch==1.8.0+cu111,
pytorch/requirements.txt:33; here torch 2.11): ``torch.matmul``,
``F.avg_pool2d`` and ``F.grid_sample``.  This file issues the same library
calls in the same order so that

* on CPU it is what ``bench.py --impl reference`` / ``cpu_baseline`` time
  ("the reference's CPU CorrBlock path on the box's host cores"), and
* on the GPU box (where /root/reference does not exist) it is the
  device-resident stand-in for "the reference run on the same GPU" that the
  ``-m gpu`` parity tests compare against at value level.

Pinned against the live reference by ``tests/test_oracle.py`` (bit-exact on CPU
to ``core.corr.CorrBlock`` through the golden vectors of ``oracle/make_golden.py``).

Follows /root/reference/pytorch/core/corr.py:13-60 and
/root/reference/pytorch/core/utils/utils.py:57-65.  Only importable from
``tests/``, ``bench.py`` (cpu_baseline / --impl reference) and
``__graft_entry__.smoke()``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def build_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4):
    """corr.py:52-60 then :21-27 -> list of (B*N, 1, Hl, Wl)."""
    B, D, H, W = fmap1.shape
    a = fmap1.reshape(B, D, H * W)
    b = fmap2.reshape(B, D, H * W)
    vol = torch.matmul(a.transpose(1, 2), b)
    vol = vol / torch.sqrt(torch.tensor(D).float())          # corr.py:60 (separate kernel)
    vol = vol.reshape(B * H * W, 1, H, W)
    levels = [vol]
    for _ in range(num_levels - 1):
        vol = F.avg_pool2d(vol, 2, stride=2)                  # corr.py:26
        levels.append(vol)
    return levels


def window_offsets(radius: int, device) -> torch.Tensor:
    """corr.py:36-38: delta[i, j] = (dy[i], dx[j]) -- added to (x, y)."""
    d = torch.linspace(-radius, radius, 2 * radius + 1)
    return torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).to(device)


def sample_level(vol: torch.Tensor, pix: torch.Tensor) -> torch.Tensor:
    """utils.py:57-65: pixel coords -> [-1, 1] -> grid_sample(align_corners=True)."""
    Hl, Wl = vol.shape[-2:]
    x, y = pix.split([1, 1], dim=-1)
    x = 2 * x / (Wl - 1) - 1
    y = 2 * y / (Hl - 1) - 1
    return F.grid_sample(vol, torch.cat([x, y], dim=-1), align_corners=True)


def lookup(levels, coords: torch.Tensor, radius: int = 4) -> torch.Tensor:
    """corr.py:29-50 -> (B, L*(2r+1)^2, H, W) fp32 contiguous."""
    B, _, H, W = coords.shape
    c = coords.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
    R = 2 * radius + 1
    feats = []
    for i, vol in enumerate(levels):
        delta = window_offsets(radius, coords.device).view(1, R, R, 2)
        feats.append(sample_level(vol, c / 2 ** i + delta).view(B, H, W, -1))
    return torch.cat(feats, dim=-1).permute(0, 3, 1, 2).contiguous().float()


class TorchCorrBlock:
    """Same constructor / call surface as the reference class, built on the
    functions above; used by tests and the CPU baseline as the oracle object."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        self.corr_pyramid = build_pyramid(fmap1, fmap2, num_levels)

    def __call__(self, coords):
        return lookup(self.corr_pyramid, coords, self.radius)


def coords_grid(B: int, H: int, W: int, device="cpu") -> torch.Tensor:
    """utils.py:74-77."""
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(B, 1, 1, 1).to(device)
