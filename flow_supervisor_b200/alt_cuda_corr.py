"""Drop-in for the reference's ``alt_cuda_corr`` extension module
(/root/reference/pytorch/alt_cuda_corr/correlation.cpp:51-54):

    corr, = alt_cuda_corr.forward(fmap1_nhwc, fmap2_nhwc, coords_b1hw2, r)      # corr.py:86
    d1, d2, dcoords = alt_cuda_corr.backward(fmap1, fmap2, coords, corr_grad, r)

Same argument meaning, same return shapes, same error behaviour: inputs must be CUDA
and contiguous or a RuntimeError is raised (CHECK_INPUT, correlation.cpp:19-21).
``install()`` registers this module as ``sys.modules['alt_cuda_corr']`` so that the
reference's ``import alt_cuda_corr`` (corr.py:5-9) resolves to it.
"""
from __future__ import annotations

import sys

import torch

from . import ops


def _check_input(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def forward(fmap1, fmap2, coords, radius):
    for t, n in ((fmap1, "fmap1"), (fmap2, "fmap2"), (coords, "coords")):
        _check_input(t, n)
    if coords.dim() != 5 or coords.shape[1] != 1 or coords.shape[-1] != 2:
        raise RuntimeError("coords must be (B, 1, H, W, 2)")
    return [ops.altcorr_fwd(fmap1.float(), fmap2.float(), coords.float(), int(radius))]


def backward(fmap1, fmap2, coords, corr_grad, radius):
    for t, n in ((fmap1, "fmap1"), (fmap2, "fmap2"), (coords, "coords"), (corr_grad, "corr_grad")):
        _check_input(t, n)
    d1, d2 = ops.altcorr_bwd(fmap1.float(), fmap2.float(), coords.float(), corr_grad.float(), int(radius))
    # correlation_kernel.cu:307 allocates coords_grad with zeros and never writes it
    return [d1, d2, torch.zeros_like(coords)]


def install() -> None:
    sys.modules["alt_cuda_corr"] = sys.modules[__name__]
