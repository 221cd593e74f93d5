#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, read-only):

    python oracle/make_golden.py

The reference has no tests or fixtures of its own for this path (SURVEY.md
section 4), so these vectors -- produced by importing
/root/reference/pytorch/core/corr.py and calling ``CorrBlock`` exactly as
raft.py:105-107,124 does -- are what pins the oracle (oracle/corr_spec.py,
oracle/corr_torch.py) and, through it, the CUDA kernels.  The reference ran on
CPU tensors here, so coordinate rounding is the 'cpu' flavour (true division in
utils.py:61-62); see oracle/corr_spec.py for the 'cuda' flavour.

Nothing in the GPU tests, smoke() or bench.py reads /root/reference; only the
committed .npz files travel.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference/pytorch"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def ref_modules():
    sys.path.insert(0, REF)
    from core.corr import CorrBlock          # corr.py:12
    from core.utils.utils import coords_grid  # utils.py:74
    return CorrBlock, coords_grid


def h16(t):
    """Round to fp16-representable values so inputs can be stored at half size
    (they are still fed to the reference as fp32)."""
    return t.half().float()


def coords_laws(coords_grid, B, H, W, gen):
    base = coords_grid(B, H, W)
    rnd = base + 3.0 * torch.randn(B, 2, H, W, generator=gen)
    oob = base + 25.0 * torch.randn(B, 2, H, W, generator=gen)
    # fractional lattice (coords that are exact multiples of 1/8): stresses the
    # normalise/un-normalise floor flips at every level
    frac = base + torch.randint(-16, 17, (B, 2, H, W), generator=gen).float() / 8.0
    return {"lattice": base, "random": rnd, "oob": oob, "frac": frac}


def case_forward(name, B, D, H, W, L, r, seed, std=1.57, laws=None, keep_pyramid=True):
    CorrBlock, coords_grid = ref_modules()
    gen = torch.Generator().manual_seed(seed)
    f1 = h16(std * torch.randn(B, D, H, W, generator=gen))
    f2 = h16(std * torch.randn(B, D, H, W, generator=gen))
    blk = CorrBlock(f1, f2, num_levels=L, radius=r)
    rec = {"fmap1": f1.numpy().astype(np.float16), "fmap2": f2.numpy().astype(np.float16),
           "num_levels": np.int32(L), "radius": np.int32(r)}
    for l, lvl in enumerate(blk.corr_pyramid):
        if keep_pyramid:
            rec[f"pyr{l}"] = lvl.reshape(B, H * W, *lvl.shape[-2:]).numpy()
    for law, c in coords_laws(coords_grid, B, H, W, gen).items():
        if laws is not None and law not in laws:
            continue
        rec[f"coords_{law}"] = c.numpy()
        rec[f"out_{law}"] = blk(c).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape") and v.ndim})


def case_backward(name, B, D, H, W, L, r, seed, n_lookups=3):
    """Autograd of the reference through several lookups of ONE CorrBlock
    (what pytorch/train.py:273 exercises): d(sum_t <out_t, g_t>)/d fmap{1,2}."""
    CorrBlock, coords_grid = ref_modules()
    gen = torch.Generator().manual_seed(seed)
    f1 = h16(torch.randn(B, D, H, W, generator=gen)).requires_grad_()
    f2 = h16(torch.randn(B, D, H, W, generator=gen)).requires_grad_()
    blk = CorrBlock(f1, f2, num_levels=L, radius=r)
    base = coords_grid(B, H, W)
    rec = {"fmap1": f1.detach().numpy().astype(np.float16),
           "fmap2": f2.detach().numpy().astype(np.float16),
           "num_levels": np.int32(L), "radius": np.int32(r), "n_lookups": np.int32(n_lookups)}
    loss = 0.0
    for t in range(n_lookups):
        c = base + (2.0 + 3.0 * t) * torch.randn(B, 2, H, W, generator=gen)
        g = h16(torch.randn(B, L * (2 * r + 1) ** 2, H, W, generator=gen))
        loss = loss + (blk(c.detach()) * g).sum()
        rec[f"coords{t}"] = c.numpy()
        rec[f"gout{t}"] = g.numpy().astype(np.float16)
    loss.backward()
    rec["dfmap1"] = f1.grad.numpy()
    rec["dfmap2"] = f2.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "dfmap", rec["dfmap1"].shape, float(np.abs(rec["dfmap1"]).max()))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    # odd dims at every level: 17x19 -> 8x9 -> 4x4 -> 2x2
    case_forward("fwd_odd_d32", B=1, D=32, H=17, W=19, L=4, r=4, seed=11)
    # RAFT-small geometry (raft.py:29-33): D=128, r=3
    case_forward("fwd_small_d128_r3", B=1, D=128, H=16, W=20, L=4, r=3, seed=12,
                 laws=("random", "frac"), keep_pyramid=False)
    # full D=256, batch 2, multiple-of-4 widths
    case_forward("fwd_d256_b2", B=2, D=256, H=16, W=24, L=4, r=4, seed=13,
                 laws=("random",), keep_pyramid=False)
    # 2 samples, odd dims at every level, 3 lookups accumulated into one block
    case_backward("bwd_odd_d64", B=2, D=64, H=17, W=19, L=4, r=4, seed=21)


if __name__ == "__main__":
    main()
