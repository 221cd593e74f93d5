#!/usr/bin/env python
"""Training-shaped timing of the correlation path (config 3 of BASELINE.json: 368x768 crops,
batch 6/GPU -> 46x96 tokens; teacher frame 432x1024 -> 54x128): CorrBlock build + 12 lookups,
then the backward (12 scatter-adds into ONE gradient pyramid, fold, two GEMMs).
Prints one JSON line per geometry; an auxiliary measurement, not the bench.py contract line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb          # noqa: E402
from flow_supervisor_b200 import _lib       # noqa: E402


def run(B, H, W, iters=12, reps=5, D=256):
    g = torch.Generator().manual_seed(0)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda().requires_grad_()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda().requires_grad_()
    coords = [(fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(iters)]
    gout = [torch.randn(B, 324, H, W, generator=g).cuda() for _ in range(2)]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    lib = _lib.load()
    for r in range(reps + 2):
        f1.grad = f2.grad = None
        e0, e1, e2 = ev(), ev(), ev()
        n0 = lib.fc_kernel_launches()
        e0.record()
        blk = fsb.CorrBlock(f1, f2, 4, 4)
        outs = [blk(c) for c in coords]
        e1.record()
        torch.autograd.backward(outs, [gout[i & 1] for i in range(iters)])
        e2.record()
        torch.cuda.synchronize()
        n1 = lib.fc_kernel_launches()
        if r >= 2:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    N = H * W
    print(json.dumps({"geometry": f"B={B} {H}x{W} D={D} iters={iters}", "fwd_ms": sum(tf) / len(tf), "bwd_ms": sum(tb) / len(tb),
                      "launches_per_step": n1 - n0, "bwd_gemm_gflop": 2 * 2.0 * B * N * N * D / 1e9,
                      "grad_pyramid_mb": blk._state.pyramid.numel() * 4 / 1e6}))


if __name__ == "__main__":
    run(6, 46, 96)
    run(6, 54, 128)
