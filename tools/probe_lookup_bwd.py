"""Time of one lookup backward (12 back-to-back launches per sample) for the library FLOWCORR_LIB points at; the
A/B builds come from  make -C flow_supervisor_b200/csrc BUILD=build_x EXTRA="-DFC_LB_GROUPS=.. -DFC_LB_STAGES=.. -DFC_LB_PREFETCH=.." OUT=...
Also checks the gradient pyramid against the default library's when a reference file is given."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402
from flow_supervisor_b200 import _lib, ops  # noqa: E402
from tools.bench_rows import timed  # noqa: E402

L, R = 4, 4
K = L * (2 * R + 1) ** 2
out = {"lib": os.path.basename(_lib.LIB_PATH)}
for (B, H, W) in ((6, 46, 96), (6, 54, 128)):
    g = torch.Generator().manual_seed(0)
    c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    gout = torch.randn(B, K, H, W, generator=g).cuda()
    gp = torch.zeros(ops.pyramid_numel(B, H, W, L), device="cuda")
    ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
    torch.cuda.synchronize()
    out[f"{H}x{W}_checksum"] = float(gp.double().abs().sum())
    out[f"{H}x{W}_us"] = 1e3 * timed(lambda: ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA), 10, inner=12)
    # the same 12 launches replayed from a CUDA graph: the eager figure above includes the host's ~45 us per call
    # (torch custom-op dispatch) whenever the kernel is shorter than that
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(12):
            ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
    out[f"{H}x{W}_graph_us"] = 1e3 * timed(gr.replay, 10) / 12
print(json.dumps(out), flush=True)
