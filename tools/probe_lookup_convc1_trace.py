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

lb = np.zeros((64, 8), dtype=np.uint64)
lib.fc_debug_lookup_convc1_loop_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_lookup_convc1_loop_trace(lb.ctypes.data) == 0
u = lb.astype(np.int64)
nc = int((u[:, 3] > 0).sum()); ne = int((u[:, 7] > 0).sum())
c = u[1:nc - 1]
print("consumer group 0, warp 0, per lookup tile (%d tiles): interpolate (incl. waiting for the footprints) %.0f | wait for a free B buffer %.0f | "
      "split + swizzled stores + arrive %.0f | tile period %.0f" % (nc, (c[:, 1] - c[:, 0]).mean(), (c[:, 2] - c[:, 1]).mean(), (c[:, 3] - c[:, 2]).mean(),
      np.diff(u[1:nc, 0]).mean()))
e = u[1:ne - 1]
print("epilogue warp 0 (hosts the MMA issue), per pair-tile (%d): issue of the next pair-tile's MMAs (incl. waiting for its B operand) %.0f | "
      "wait for the accumulator %.0f | drain + bias + ReLU + stores %.0f | period %.0f" % (ne, (e[:, 5] - e[:, 4]).mean(), (e[:, 6] - e[:, 5]).mean(),
      (e[:, 7] - e[:, 6]).mean(), np.diff(u[1:ne, 4]).mean()))
