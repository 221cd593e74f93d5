#!/usr/bin/env python
"""Lookup-backward timing, back to back (cfg 3 teacher: B=6, 54x128 and student: 46x96).  FLOWCORR_PROBE 1 = no reduce-add,
2 = TMA store instead of the reduce."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb              # noqa: E402
from flow_supervisor_b200 import _lib, ops      # noqa: E402
from probe_bounds import timed                  # noqa: E402

L, R = 4, 4
g = torch.Generator().manual_seed(0)
for B, H, W in ((6, 54, 128), (6, 46, 96)):
    K = L * (2 * R + 1) ** 2
    cs = [(fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(4)]
    gout = torch.randn(B, K, H, W, generator=g).cuda()
    gp = torch.zeros(ops.pyramid_numel(B, H, W, L), device="cuda")
    it = [0]

    def bwd():
        it[0] += 1
        ops.lookup_bwd(gout, cs[it[0] % 4], gp, L, R, _lib.COORD_CUDA)

    for probe in (0, 1, 2, 0):
        os.environ["FLOWCORR_PROBE"] = str(probe)
        print(json.dumps({"kernel": "lookup_bwd", "geometry": f"B={B} {H}x{W}", "probe": probe,
                          "us": 1e3 * timed(bwd, reps=36, warm=12)}), flush=True)
    os.environ["FLOWCORR_PROBE"] = "0"
