#!/usr/bin/env python
"""Round-robin A/B of build switches (env, read per call), cfg 2 geometry: medians over rounds x reps launches so that
slow drifts of the box hit every variant alike.   python ab_build_env.py VAR=a,b[,c] [VAR2=...]  (cartesian product)"""
import itertools
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

axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[1:]] or [("FLOWCORR_BUILD_SCHED", ["0", "1"])]
variants = [dict(zip([k for k, _ in axes], combo)) for combo in itertools.product(*[v for _, v in axes])]
g = torch.Generator().manual_seed(0)
B, H, W, D, L = 8, 55, 128, 256, 4
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
for math, mname in ((_lib.MATH_TC_3XBF16, "3xbf16"), (_lib.MATH_TC_BF16, "bf16")):
    res = {i: [] for i in range(len(variants))}
    for rnd in range(6):
        for i, v in enumerate(variants):
            os.environ.update(v)
            res[i].append(1e3 * timed(lambda: ops.build(f1, f2, L, math, _lib.VOL_F32), reps=12, warm=2))
    for i, v in enumerate(variants):
        print(json.dumps({"kernel": "build (pack + tc_build)", "math": mname, **v, "us_median": statistics.median(res[i]),
                          "us_min": min(res[i])}), flush=True)
