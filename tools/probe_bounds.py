#!/usr/bin/env python
"""Which stage bounds the lookup kernels?  FLOWCORR_PROBE (fc_lookup.cuh) switches one stage
of a kernel off; timing the crippled variants next to the real one says how much of the
launch each stage accounts for.  Results of probed launches are garbage by construction.

  build    (cfg 2: B=8, 55x128):  0 = real, 1 = epilogue without global stores, 2 = no MMAs issued,
                                  3 = epilogue neither reads TMEM nor stores,
                                  5 = pooled-level stores off, 6 = level-0 stores off
  forward  (cfg 2: B=8, 55x128):  timing only (its probes lived until r01f)
  backward (cfg 3 teacher: B=6, 54x128): 0 = real, 1 = no reduce-add, 2 = TMA store instead of reduce
One JSON line per measurement."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _switch import set_switch  # noqa: E402
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb              # noqa: E402
from flow_supervisor_b200 import _lib, ops      # noqa: E402

L, R, D = 4, 4, 256


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def hbm_rates():
    """What the memory system gives a pure write / pure read / copy of the pyramid's size."""
    n = 2_070_000_000 // 4
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    for name, fn, nbytes in (("write-only (fill_)", lambda: a.fill_(1.0), 4 * n),
                             ("read-only (sum)", lambda: a.sum(), 4 * n),
                             ("copy (read + write)", lambda: b.copy_(a), 8 * n)):
        ms = timed(fn, reps=10, warm=3)
        print(json.dumps({"kernel": "hbm " + name, "GB": nbytes / 1e9, "us": 1e3 * ms, "GB/s": nbytes / ms / 1e6}), flush=True)
    del a, b


def main():
    hbm_rates()
    g = torch.Generator().manual_seed(0)
    B, H, W = 8, 55, 128
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    cs = [(fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(4)]
    for math, mname in ((_lib.MATH_TC_3XBF16, "3xbf16"), (_lib.MATH_TC_BF16, "bf16")):
        for probe, what in ((0, "real"), (1, "epilogue without global stores"), (2, "no MMAs issued"),
                            (3, "epilogue neither reads TMEM nor stores"), (5, "pooled-level stores off"),
                            (6, "level-0 stores off"), (0, "real again")):
            set_switch("FLOWCORR_PROBE", str(probe))
            print(json.dumps({"kernel": "build (pack + tc_build)", "math": mname, "geometry": f"B={B} {H}x{W}",
                              "probe": probe, "what": what,
                              "us": 1e3 * timed(lambda: ops.build(f1, f2, L, math, _lib.VOL_F32), reps=10, warm=3)}),
                  flush=True)
    set_switch("FLOWCORR_PROBE", "0")
    pyr = ops.build(f1, f2, L, _lib.MATH_TC_3XBF16, _lib.VOL_F32)
    it = [0]

    def fwd():
        it[0] += 1
        return ops.lookup(pyr, cs[it[0] % 4], L, R, _lib.COORD_CUDA)

    # (forward stage probes and L2 eviction hints were measured in r01f and then removed from the kernel:
    #  profiles/r01f_stage_probes.jsonl)
    print(json.dumps({"kernel": "lookup_fwd", "geometry": f"B={B} {H}x{W}", "us": 1e3 * timed(fwd, reps=36, warm=12)}), flush=True)
    del pyr
    B, H, W = 6, 54, 128
    K = L * (2 * R + 1) ** 2
    c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    gout = torch.randn(B, K, H, W, generator=g).cuda()
    gp = torch.zeros(ops.pyramid_numel(B, H, W, L), device="cuda")
    for probe, what in ((0, "real"), (1, "no reduce-add"), (2, "TMA store instead of reduce"), (0, "real again")):
        set_switch("FLOWCORR_PROBE", str(probe))
        print(json.dumps({"kernel": "lookup_bwd", "geometry": f"B={B} {H}x{W}", "probe": probe, "what": what,
                          "us": 1e3 * timed(lambda: ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA))}), flush=True)
    set_switch("FLOWCORR_PROBE", "0")


if __name__ == "__main__":
    main()
