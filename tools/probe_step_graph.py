"""Launch gaps of the headline step: build + 12 lookups issued eagerly (what bench.py times) against one CUDA-graph replay
of the same launches, cfg 2 (B = 8, 55x128 tokens)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402

B, D, H, W, T = 8, 256, 55, 128, 12
g = torch.Generator().manual_seed(0)
f1 = torch.randn(B, D, H, W, generator=g).cuda()
f2 = torch.randn(B, D, H, W, generator=g).cuda()
coords = [(fsb.coords_grid(B, H, W) + 4.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(T)]


def step():
    blk = fsb.CorrBlock(f1, f2, 4, 4)
    out = None
    for t in range(T):
        out = blk(coords[t])
    return out


def lookups_only(blk):
    out = None
    for t in range(T):
        out = blk(coords[t])
    return out


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    out = {"eager_step_ms": timed(step)}
    blk = fsb.CorrBlock(f1, f2, 4, 4)
    out["eager_12_lookups_ms"] = timed(lambda: lookups_only(blk))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        res = step()
    out["graph_step_ms"] = timed(gr.replay)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        res2 = lookups_only(blk)
    out["graph_12_lookups_ms"] = timed(g2.replay)
print(json.dumps(out))
