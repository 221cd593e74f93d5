"""A/B of the build backward: fold + bf16 split inside the GEMMs (default) against the round-1 pipeline (separate
fold + pack pass, FLOWCORR_BWD_FUSED=0).  Prints one JSON line per geometry: both times, the difference between the
two results and against the fp32 CUDA-core mode, and whether the gradient pyramid survived the call."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb  # noqa: E402
from flow_supervisor_b200 import _lib, ops  # noqa: E402
from tools.bench_rows import timed  # noqa: E402

L = 4


def switch(v):
    _lib.check(_lib.load().fc_tunable_set(b"bwd_fused", int(v)), "fc_tunable_set")


def one(B, D, H, W, reps=10, check_fp32=True):
    g = torch.Generator().manual_seed(3)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    numel = ops.pyramid_numel(B, H, W, L)
    src = ops.clear_pads_(torch.randn(numel, device="cuda"), B, H, W, L)      # pads of a gradient pyramid are zeros
    gp = src.clone()
    out = {"geometry": f"B={B} D={D} {H}x{W}"}
    res = {}
    # 1: default (patch-row aligned maps take the TMA-coarse-box kernel), 2: the generic fold-in-GEMM kernel, 0: round-1 pipeline
    for name, v in (("fused", 1), ("generic", 2), ("separate", 0)):
        switch(v)
        gp.copy_(src)
        try:
            res[name] = ops.build_bwd(gp, f1, f2, L, _lib.MATH_TC_3XBF16)
        except RuntimeError as e:
            out[name + "_error"] = str(e)[:120]
            continue
        torch.cuda.synchronize()
        out[name + "_pyramid_preserved"] = bool(torch.equal(gp, src))
        out[name + "_ms"] = timed(lambda: ops.build_bwd(gp, f1, f2, L, _lib.MATH_TC_3XBF16), reps, setup=lambda: gp.copy_(src))
    switch(1)
    if "fused" in res and "separate" in res:
        for i, n in enumerate(("d1", "d2")):
            out[f"{n}_fused_vs_separate"] = float((res["fused"][i] - res["separate"][i]).abs().max() / res["separate"][i].abs().max())
    if check_fp32 and "fused" in res:
        gp.copy_(src)
        r = ops.build_bwd(gp, f1, f2, L, _lib.MATH_FP32)
        for i, n in enumerate(("d1", "d2")):
            out[f"{n}_fused_vs_fp32"] = float((res["fused"][i] - r[i]).abs().max() / r[i].abs().max())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":          # for ncu: one geometry, few launches
        B, Dn, H, W = (int(v) for v in sys.argv[2:6])
        one(B, Dn, H, W, reps=2, check_fp32=False)
        sys.exit(0)
    one(1, 64, 19, 27)
    one(2, 256, 46, 62)
    one(1, 256, 47, 156)
    one(6, 256, 46, 96)
    one(6, 256, 54, 128)
    one(1, 256, 136, 240, reps=3, check_fp32=False)     # cfg-5 map: past the old whole-row-in-shared-memory limit
