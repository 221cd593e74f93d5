"""Loader of the COMPILED reference extension (oracle/_ref/alt_cuda_corr_ref.so, built by
oracle/build_ref.py from /root/reference/pytorch/alt_cuda_corr where the sources lie) --
TEST INFRASTRUCTURE ONLY.

It is the reference's own CUDA kernels (correlation_kernel.cu:18-119 forward, :122-256
backward) behind the reference's own pybind signatures (correlation.cpp:51-54):

    corr, = mod.forward(fmap1_nhwc, fmap2_nhwc, coords_b1hw2, r)
    d1, d2, dcoords = mod.backward(fmap1, fmap2, coords, corr_grad, r)

Used by tests/test_gpu_ref_ext.py to pin rows a8/a9 (and a7 through AlternateCorrBlock's
own call pattern, corr.py:74-91) and by tools/bench_rows.py as the GPU baseline of the
on-demand path.  Nothing under flow_supervisor_b200/ may import this.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
NAME = "alt_cuda_corr_ref"
PATH = os.path.join(HERE, "_ref", NAME + ".so")

_mod = None


def available() -> bool:
    return os.path.exists(PATH)


def load():
    """-> the extension module (raises FileNotFoundError when it was never built)."""
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{PATH} missing: run python oracle/build_ref.py where /root/reference exists")
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        loader = importlib.machinery.ExtensionFileLoader(NAME, PATH)
        spec = importlib.util.spec_from_file_location(NAME, PATH, loader=loader)
        _mod = importlib.util.module_from_spec(spec)
        loader.exec_module(_mod)
    return _mod


class RefAlternateCorrBlock:
    """Driver of the compiled reference kernel with the call pattern of the reference's
    AlternateCorrBlock (/root/reference/pytorch/core/corr.py:63-91): fmap2 is avg-pooled
    level by level (:68-72), both maps are handed over channels-last (:82-83), level l gets
    coords / 2**l as (B, 1, H, W, 2) (:85), the per-level (B, 81, H, W) results are
    concatenated along channels (:89-90) and divided by sqrt(D) (:91)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        import torch.nn.functional as F
        self.num_levels, self.radius, self.dim = num_levels, radius, fmap1.shape[1]
        self.q_nhwc = fmap1.permute(0, 2, 3, 1).contiguous()
        self.t_nhwc = []
        for _ in range(num_levels):
            self.t_nhwc.append(fmap2.permute(0, 2, 3, 1).contiguous())
            fmap2 = F.avg_pool2d(fmap2, 2, stride=2)

    def level(self, coords, l):
        """(B, 2, H, W) coords -> raw (unscaled) reference output of level l, (B, 81, H, W)."""
        B, _, H, W = coords.shape
        c = (coords.permute(0, 2, 3, 1) / 2 ** l).reshape(B, 1, H, W, 2).contiguous()
        return load().forward(self.q_nhwc, self.t_nhwc[l], c, self.radius)[0][:, 0]

    def __call__(self, coords):
        import torch
        per_level = [self.level(coords, l) for l in range(self.num_levels)]
        return torch.cat(per_level, dim=1) / math.sqrt(self.dim)
