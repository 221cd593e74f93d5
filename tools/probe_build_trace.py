"""Timeline of the forward build's pipeline (FC_PROBES library only):
    make -C flow_supervisor_b200/csrc BUILD=build_probes EXTRA=-DFC_PROBES OUT=../libflowcorr_probes.so
    FLOWCORR_LIB=flow_supervisor_b200/libflowcorr_probes.so python tools/probe_build_trace.py
Per tile (= one accumulator of 256 targets x 256 queries of the pair) of CTA 0: where the MMA thread, the first epilogue
warp and the TMA producer spend their cycles (TB_TRACE in fc_build_tc.cu)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402
from flow_supervisor_b200 import _lib  # noqa: E402

B, D, H, W = 8, 256, 55, 128
g = torch.Generator().manual_seed(3)
f1 = torch.randn(B, D, H, W, generator=g).cuda()
f2 = torch.randn(B, D, H, W, generator=g).cuda()
for _ in range(3):
    blk = fsb.CorrBlock(f1, f2, 4, 4)
torch.cuda.synchronize()
lib = _lib.load()
S = 8
buf = np.zeros((512, S), dtype=np.uint64)
lib.fc_debug_build_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_build_trace(buf.ctypes.data) == 0
t = buf.astype(np.int64)
n = int((t[:, 1] > 0).sum())
print(f"B={B} D={D} {H}x{W}: {n} tiles in CTA 0; clock64 cycles; ideal MMA time per tile = 48 x 131 = 6288")
print("tile  acc_free  issued  ring_wait | epi_ready acc_complete drained stores_issued | producer_wait")
t0 = t[8, 0]
for i in range(8, min(n, 40)):
    r = t[i]
    print(f"{i:4d} {r[0]-t0:9d} {r[1]-t0:8d} {r[2]:9d} | {r[3]-t0:9d} {r[4]-t0:10d} {r[5]-t0:8d} {r[6]-t0:10d} | {r[7]:8d}")
span = t[n - 1, 6] - t[0, 0]
d = np.diff(t[:n, 1])
big = np.argsort(-d)[:8]
print("whole run: %d cycles from tile 0's accumulator to the last tile's stores; median tile period %d; the 8 largest periods: %s" %
      (span, int(np.median(d)), ", ".join(f"tile {int(i) + 1}: {int(d[i])}" for i in sorted(big))))
print("first tile issued %d cycles after the first accumulator wait; last tile: issued -> stores issued %d" %
      (t[0, 1] - t[0, 0], t[n - 1, 6] - t[n - 1, 1]))
w = t[8:n - 2]
per = np.diff(w[:, 1]).mean()
print("mean cycles per tile (issue to issue): %.0f" % per)
print("MMA thread: waiting for a free accumulator %.0f | issuing (incl. ring waits %.0f) %.0f" %
      ((w[1:, 0] - w[:-1, 1]).mean(), w[:, 2].mean(), (w[:, 1] - w[:, 0]).mean()))
print("epilogue warp 0: waiting for the accumulator %.0f | drain (TMEM -> staging -> store issue) %.0f | pooled levels after the hand-back %.0f" %
      ((w[:, 4] - w[:, 3]).mean(), (w[:, 5] - w[:, 4]).mean(), (w[:, 6] - w[:, 5]).mean()))
print("producer: waiting for free ring stages per tile %.0f" % w[:, 7].mean())

cb = np.zeros((8, 8, 8), dtype=np.uint64)
lib.fc_debug_build_chunk_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_build_chunk_trace(cb.ctypes.data) == 0
c = cb.astype(np.int64)
print("epilogue warp 0, tiles 16..23, mean cycles per chunk phase (8 chunks of 32 columns per tile):")
print("  TMEM load + wait %.0f | wait for the staging box %.0f | 8 swizzled shared stores %.0f | proxy fence (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) + syncwarp %.0f | "
      "TMA store issue + commit %.0f | pooled levels (level-1 store, stashes) %.0f | chunk to chunk %.0f"
      % ((c[:, :, 1] - c[:, :, 0]).mean(), (c[:, :, 2] - c[:, :, 1]).mean(), (c[:, :, 5] - c[:, :, 2]).mean(), (c[:, :, 6] - c[:, :, 5]).mean(),
         (c[:, :, 3] - c[:, :, 6]).mean(), (c[:, :, 4] - c[:, :, 3]).mean(), np.diff(c[:, :, 0], axis=1).mean()))
