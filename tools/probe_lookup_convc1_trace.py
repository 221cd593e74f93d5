"""Prologue / main loop split of the fused lookup + convc1 kernel (FC_PROBES library only, LC_TRACE in fc_lookup_conv.cu):
    make -C flow_supervisor_b200/csrc BUILD=build_probes EXTRA=-DFC_PROBES OUT=../libflowcorr_probes.so
    FLOWCORR_LIB=flow_supervisor_b200/libflowcorr_probes.so python tools/probe_lookup_convc1_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402
from flow_supervisor_b200 import _lib, ops  # noqa: E402

B, D, H, W = 8, 256, 55, 128
g = torch.Generator().manual_seed(1)
f1 = torch.randn(B, D, H, W, generator=g).cuda()
f2 = torch.randn(B, D, H, W, generator=g).cuda()
blk = fsb.CorrBlock(f1, f2, 4, 4)
coords = (fsb.coords_grid(B, H, W) + 4.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
wgt = (0.05 * torch.randn(256, 324, 1, 1, generator=g)).cuda()
bias = torch.randn(256, generator=g).cuda()
packed = ops.convc1_prepare(wgt, bias)
for _ in range(3):
    out = blk.lookup_convc1(coords, packed)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((148, 4), dtype=np.uint64)
lib.fc_debug_lookup_convc1_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_lookup_convc1_trace(buf.ctypes.data) == 0
t = buf.astype(np.int64)
print("per CTA, clock64 cycles (mean / max over 148 CTAs): barriers + TMEM alloc %.0f / %d | weights -> tensor memory, B zeroed, cluster sync %.0f / %d | main loop %.0f / %d"
      % ((t[:, 1] - t[:, 0]).mean(), (t[:, 1] - t[:, 0]).max(), (t[:, 2] - t[:, 1]).mean(), (t[:, 2] - t[:, 1]).max(),
         (t[:, 3] - t[:, 2]).mean(), (t[:, 3] - t[:, 2]).max()))
