"""Host time per lookup call: torch.ops.flowcorr.lookup (dispatcher) vs the direct function, tiny problem (kernel ~5 us)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb
from flow_supervisor_b200 import ops
f = torch.randn(1, 64, 16, 24, device="cuda")
blk = fsb.CorrBlock(f, f)
c = fsb.coords_grid(1, 16, 24, device="cuda") + 0.3
st = blk._state
for name, fn in (("custom_op", ops.lookup), ("direct", ops.lookup_direct), ("CorrBlock.__call__", None)):
    call = (lambda: blk(c)) if fn is None else (lambda: fn(st.pyramid, c, st.L, st.radius, st.coord))
    for _ in range(200): call()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): call()
    torch.cuda.synchronize(); print(name, "us per call", (time.perf_counter() - t0) / 2000 * 1e6)
