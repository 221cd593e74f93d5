#!/usr/bin/env python
"""Build-only stage probes (see tools/probe_bounds.py): FLOWCORR_PROBE 0 = real, 1 = no global stores,
2 = no MMAs, 3 = no TMEM reads + no stores, 5 = pooled-level stores off, 6 = level-0 stores off,
7 = no target-operand loads.  argv[1] = comma list of probes, argv[2] = comma list of
FLOWCORR_BUILD_L2HINT masks, argv[3] = comma list of FLOWCORR_L0STORE values.
One JSON line per measurement; results of probed launches are garbage by construction."""
import json
import os
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
arg = lambda i, d: [int(x) for x in sys.argv[i].split(",")] if len(sys.argv) > i else d
probes, hints, directs = arg(1, [0, 1, 2, 3, 5, 6, 7, 0]), arg(2, [0]), arg(3, [0])
for direct in directs:
    os.environ["FLOWCORR_L0STORE"] = str(direct)
    for hint in hints:
        os.environ["FLOWCORR_BUILD_L2HINT"] = str(hint)
        for math, mname in ((_lib.MATH_TC_3XBF16, "3xbf16"), (_lib.MATH_TC_BF16, "bf16")):
            for probe in probes:
                os.environ["FLOWCORR_PROBE"] = str(probe)
                us = 1e3 * timed(lambda: ops.build(f1, f2, L, math, _lib.VOL_F32), reps=10, warm=3)
                print(json.dumps({"kernel": "build (pack + tc_build)", "math": mname, "l0_direct": direct, "l2hint": hint,
                                  "probe": probe, "us": us}), flush=True)
os.environ["FLOWCORR_PROBE"] = "0"
