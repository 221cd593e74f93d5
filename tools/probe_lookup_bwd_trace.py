"""Timeline of one warp of the lookup backward (FC_PROBES library only, see LB_TRACE in fc_lookup.cu):
    make -C flow_supervisor_b200/csrc BUILD=build_probes EXTRA=-DFC_PROBES OUT=../libflowcorr_probes.so
    FLOWCORR_LIB=flow_supervisor_b200/libflowcorr_probes.so python tools/probe_lookup_bwd_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402
from flow_supervisor_b200 import _lib, ops  # noqa: E402

B, H, W, L, R = 6, 46, 96, 4, 4
K = L * (2 * R + 1) ** 2
g = torch.Generator().manual_seed(0)
c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
gout = torch.randn(B, K, H, W, generator=g).cuda()
gp = torch.zeros(ops.pyramid_numel(B, H, W, L), device="cuda")
if len(sys.argv) > 1:                                  # stage probe of the FC_PROBES build (results are then garbage)
    _lib.check(_lib.load().fc_tunable_set(b"probe", int(sys.argv[1])), "fc_tunable_set")
for _ in range(3):
    ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((64, 8), dtype=np.uint64)
lib.fc_debug_lookup_bwd_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_lookup_bwd_trace(buf.ctypes.data) == 0
t = buf.astype(np.int64)
n = int((t[:, 6] > 0).sum())
names = ["taps", "wait reduce read + barrier", "zero + barrier", "window", "fence + barrier", "reduce issue", "next tile's loads + loop"]
print(f"B={B} {H}x{W}: {n} tiles by group 0 of CTA 0; clock64 cycles per phase (mean over tiles 1..{n - 2})")
w = t[1:n - 1]
ph = [w[:, 1] - w[:, 0], w[:, 2] - w[:, 1], w[:, 3] - w[:, 2], w[:, 4] - w[:, 3], w[:, 5] - w[:, 4], w[:, 6] - w[:, 5], t[2:n, 0] - w[:, 6]]
for nm, v in zip(names, ph):
    print(f"  {nm:34s} {v.mean():8.0f}")
print(f"  {'tile period':34s} {np.diff(t[1:n, 0]).mean():8.0f}")
