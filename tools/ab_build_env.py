#!/usr/bin/env python
"""Round-robin A/B of one build switch (here FLOWCORR_BUILD_SCHED = 0 | 1; r01j also ran FLOWCORR_BUILD_EPI_WARPS = 4 | 8), cfg 2 geometry."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb              # noqa: E402,F401
from flow_supervisor_b200 import _lib, ops      # noqa: E402
from probe_bounds import timed                  # noqa: E402

g = torch.Generator().manual_seed(0)
B, H, W, D, L = 8, 55, 128, 256, 4
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
for math, mname in ((_lib.MATH_TC_3XBF16, "3xbf16"), (_lib.MATH_TC_BF16, "bf16")):
    res = {"0": [], "1": []}
    for rnd in range(6):
        for k in res:
            os.environ["FLOWCORR_BUILD_SCHED"] = k
            res[k].append(1e3 * timed(lambda: ops.build(f1, f2, L, math, _lib.VOL_F32), reps=12, warm=2))
    for k in res:
        print(json.dumps({"kernel": "build (pack + tc_build)", "math": mname, "sched_strided": int(k),
                          "us_median": statistics.median(res[k]), "us_min": min(res[k])}), flush=True)
os.environ.pop("FLOWCORR_BUILD_SCHED")
