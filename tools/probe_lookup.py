#!/usr/bin/env python
"""Lookup-forward timing under env switches (cfg 2: B=8, 55x128; 12 different coordinate sets back to back like a step).
FLOWCORR_PROBE bit 0 = late stage release, bit 1 = no loads (garbage results).  (r01h also timed a whole-map mode for
small levels -- one swizzled TMA box per tile instead of 32 -- which did not pay and was removed: profiles/README.md.)"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb              # noqa: E402
from flow_supervisor_b200 import _lib, ops      # noqa: E402
from probe_bounds import timed                  # noqa: E402

g = torch.Generator().manual_seed(0)
B, H, W, D, L, R = 8, 55, 128, 256, 4, 4
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
cs = [(fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(4)]
pyr = ops.build(f1, f2, L, _lib.MATH_TC_3XBF16, _lib.VOL_F32)
it = [0]


def fwd():
    it[0] += 1
    return ops.lookup(pyr, cs[it[0] % 4], L, R, _lib.COORD_CUDA)


variants = [{"FLOWCORR_PROBE": p} for p in ("0", "1", "2", "0")]
res = {i: [] for i in range(len(variants))}
for rnd in range(5):
    for i, v in enumerate(variants):
        for k, x in v.items():
            if x == "":
                os.environ.pop(k, None)
            else:
                os.environ[k] = x
        res[i].append(1e3 * timed(fwd, reps=36, warm=12))
for i, v in enumerate(variants):
    print(json.dumps({"kernel": "lookup_fwd", **v, "us_median": statistics.median(res[i]), "us_min": min(res[i])}), flush=True)
