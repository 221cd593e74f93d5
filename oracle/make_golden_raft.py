#!/usr/bin/env python
"""Generate tests/golden/raft_seed1234_128x160.npz by running the UNMODIFIED reference RAFT
(/root/reference/pytorch/core/raft.py with its own CorrBlock) on CPU.  Build container only.

    python oracle/make_golden_raft.py

Pins oracle/raft_model.py (the restated caller of the correlation path): same seed -> same
weights, same synthetic pair -> same flow.  Weights are NOT stored (21 MB); the seed is
(pytorch/train.py:347 uses 1234) together with a checksum of the state_dict.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

REF = "/root/reference/pytorch"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def synth_pair(h, w, seed=0):
    """Smooth random image and a copy shifted by (3, -2) px: gives a non-trivial flow."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, h // 8 + 2, w // 8 + 2, generator=g)
    big = torch.nn.functional.interpolate(base, size=(h + 16, w + 16), mode="bicubic", align_corners=False)
    big = (255 * (big - big.min()) / (big.max() - big.min())).float()
    im1 = big[:, :, 8:8 + h, 8:8 + w].contiguous()
    im2 = big[:, :, 8 - 2:8 - 2 + h, 8 + 3:8 + 3 + w].contiguous()
    return im1, im2


def main():
    sys.path.insert(0, REF)
    from core.raft import RAFT                      # raft.py:24
    torch.manual_seed(1234)
    model = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    sd = model.state_dict()
    checksum = float(sum(v.double().abs().sum() for v in sd.values()))
    im1, im2 = synth_pair(128, 160)
    with torch.no_grad():
        flow_low, flow_up = model(im1, im2, iters=12, test_mode=True)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "raft_seed1234_128x160.npz"),
                        flow_low=flow_low.numpy(), flow_up=flow_up.numpy().astype(np.float16),
                        flow_up_mean=np.float64(flow_up.double().mean()), flow_up_absmax=np.float64(flow_up.abs().max()),
                        state_checksum=np.float64(checksum), n_params=np.int64(sum(v.numel() for v in sd.values())))
    print("flow_low", tuple(flow_low.shape), "mean", float(flow_low.mean()), "absmax", float(flow_low.abs().max()),
          "checksum", checksum)


if __name__ == "__main__":
    main()
