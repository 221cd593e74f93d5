"""Timeline of the backward GEMM's pipeline (FC_PROBES library only):
    make -C flow_supervisor_b200/csrc BUILD=build_probes EXTRA=-DFC_PROBES OUT=../libflowcorr_probes.so
    FLOWCORR_LIB=flow_supervisor_b200/libflowcorr_probes.so python tools/probe_bwd_trace.py
Prints, for k-blocks 32..63 of CTA 0, the clock64() stamps of the hand-offs (BF_TRACE in fc_bwd_tc.cu) relative to
k-block 32's MMA issue, and the mean distance between the steps of the loop over 160 k-blocks."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flow_supervisor_b200 import _lib, ops  # noqa: E402

B, D, H, W, L = 6, 256, 54, 128, 4
SLOTS = 10
g = torch.Generator().manual_seed(3)
f1 = torch.randn(B, D, H, W, generator=g).cuda()
f2 = torch.randn(B, D, H, W, generator=g).cuda()
gp = torch.randn(ops.pyramid_numel(B, H, W, L), device="cuda")
for _ in range(3):
    ops.build_bwd(gp, f1, f2, L, _lib.MATH_TC_3XBF16)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((2, 256, SLOTS), dtype=np.uint64)
lib.fc_debug_bwd_trace.argtypes = [ctypes.c_void_p]
assert lib.fc_debug_bwd_trace(buf.ctypes.data) == 0
# aligned kernel (default for these maps): slot 0 = coarse boxes landed, slot 4 = all converters past the named barrier;
# generic kernel (FLOWCORR_BWD_FUSED=2): slot 0 = box buffer free, slot 4 = operand stage free
names = ["coarse/box_free", "stage_free", "conv_ready", "box_landed", "barrier/stage_free", "stored", "own_half", "peer_half",
         "features", "issued"]
print(f"B={B} D={D} {H}x{W}, CTA 0 (leader of pair 0); clock64 cycles")
for op in (0, 1):
    t = buf[op].astype(np.int64)
    t0 = t[32, 9]
    print(f"--- dF{op + 1}: columns:", " ".join(names))
    for it in range(32, 64):
        print(f"{it:4d} " + " ".join(f"{int(v - t0):8d}" for v in t[it]))
    w = t[32:192]
    print("mean cycles per k-block (issue to issue):", float(np.diff(w[:, 9]).mean()))
    print("mean: conv_ready -> stored %.0f | stored -> own_half (slowest converter warp) %.0f | own_half -> peer_half %.0f | "
          "peer_half -> issued (12 MMAs + commit) %.0f | issued(i) -> conv_stage_free(i+2) %.0f | stored(i) -> conv_ready(i+1) %.0f"
          % ((w[:, 5] - w[:, 2]).mean(), (w[:, 6] - w[:, 5]).mean(), (w[:, 7] - w[:, 6]).mean(), (w[:, 9] - w[:, 7]).mean(),
             (w[2:, 4] - w[:-2, 9]).mean(), (w[1:, 2] - w[:-1, 5]).mean()))
