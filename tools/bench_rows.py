#!/usr/bin/env python
"""Per-row timings of SURVEY.md section 8 that bench.py's headline line does not carry:

  a6  lookup backward (scatter-add) and build backward (fold + two GEMMs) at config 3
      (368x768 crops, batch 6 -> 46x96 tokens; teacher frame 432x1024 -> 54x128)
  a7  on-demand lookup (AlternateCorrBlock) at config 5 (1088x1920 -> 136x240, batch 2),
      next to the reference's own CUDA kernel (oracle/_ref, when built) driven with the
      reference's call pattern (corr.py:74-91)
  a8/a9  alt_cuda_corr forward / backward, one level, same geometry, vs the compiled reference

Each line states the algorithmic work, the measured time (CUDA events on the
launching stream, warm, mean over reps) and the fraction of the bounding roofline
(MEASURED_PEAKS.json).  Auxiliary measurement: one JSON line per row on stdout.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb              # noqa: E402
from flow_supervisor_b200 import _lib, ops      # noqa: E402

L, R, D = 4, 4, 256
K = L * (2 * R + 1) ** 2


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops_sustained"]
    except Exception:
        return 6650.0, 1400.0


def timed(fn, reps, warm=3, setup=None, inner=1):
    """Mean device time of fn() in ms.  inner > 1: that many back-to-back calls between one pair of events -- for
    kernels of tens of microseconds a single call between two events mostly measures the host's dispatch gap
    (custom-op + ctypes, ~30 us) during which the GPU idles."""
    for _ in range(warm):
        if setup: setup()
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if setup: setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) / inner
    return tot / reps


def inbounds_footprint(c, H, W):
    tot = 0.0
    for l in range(L):
        Hl, Wl = H >> l, W >> l
        x0 = torch.floor(c[:, 0] / 2 ** l) - R
        y0 = torch.floor(c[:, 1] / 2 ** l) - R
        nx = (torch.clamp(x0 + 2 * R + 2, max=Wl) - torch.clamp(x0, min=0)).clamp(min=0)
        ny = (torch.clamp(y0 + 2 * R + 2, max=Hl) - torch.clamp(y0, min=0)).clamp(min=0)
        tot += float((nx * ny).mean())
    return tot


ROWS = []


def emit(**kw):
    """Collect a row (and print it when run as a script)."""
    ROWS.append(kw)
    if __name__ == "__main__":
        print(json.dumps(kw), flush=True)


def collect(reps=5, ref_kernel=True, model=True):
    """All rows as a list of dicts: what bench.py embeds under "rows"."""
    global USE_REF
    USE_REF = ref_kernel
    ROWS.clear()
    row_backward(6, 46, 96, reps)
    row_backward(6, 54, 128, reps)
    row_ondemand(2, 136, 240, reps)
    row_bf16_volume(8, 55, 128, reps)
    row_fnet_tail(8, 55, 128, reps)
    row_lookup_convc1(8, 55, 128, reps)
    if model:
        row_model(8, 436, 1024, max(2, reps // 2))
    return list(ROWS)


def row_bf16_volume(B, H, W, reps, iters=12):
    """bf16 VOLUME mode (FC_VOL_BF16) next to the fp32 volume, same arithmetic (3xbf16 contraction)."""
    g = torch.Generator().manual_seed(0)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    keep = fsb.CorrBlock.math, fsb.CorrBlock.volume
    out = {}
    try:
        fsb.CorrBlock.math = "3xbf16"
        for vol in ("f32", "bf16"):
            fsb.CorrBlock.volume = vol
            ms_b = timed(lambda: fsb.CorrBlock(f1, f2, L, R), reps)
            blk = fsb.CorrBlock(f1, f2, L, R)
            ms_l = timed(lambda: blk(c), reps, inner=12)
            out[vol] = {"build_ms": ms_b, "lookup_ms": ms_l, "step_ms": ms_b + iters * ms_l,
                        "pyramid_gb": blk._state.pyramid.numel() * blk._state.pyramid.element_size() / 1e9}
            if vol == "f32":
                ref = blk(c)
            else:
                out["max_rel_diff_vs_f32_volume"] = float((blk(c) - ref).abs().max() / ref.abs().max())
            del blk
    finally:
        fsb.CorrBlock.math, fsb.CorrBlock.volume = keep
    emit(row="bf16 volume mode (build + lookups), stated tolerance 2^-8 of max / 0.05 px EPE", geometry=f"B={B} {H}x{W}", **out)


def row_lookup_convc1(B, H, W, reps):
    """Row f1: lookup + convc1 (1x1, 324 -> 256) + ReLU.  'separate' = fc_lookup_fwd + torch conv2d + relu (cuDNN, torch
    default math); 'fused' = fc_lookup_convc1_fwd (weights in tensor memory, the 324-channel tensor never written)."""
    hbm, tf = peaks()
    g = torch.Generator().manual_seed(0)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    conv = torch.nn.Conv2d(K, 256, 1).cuda()
    packed = ops.convc1_prepare(conv.weight.detach(), conv.bias.detach())
    blk = fsb.CorrBlock(f1, f2, L, R)
    with torch.no_grad():
        ms_look = timed(lambda: blk(c), reps, inner=12)
        ms_sep = timed(lambda: torch.relu(conv(blk(c))), reps, inner=12)
        ms_fused = timed(lambda: blk.lookup_convc1(c, packed), reps, inner=12)
        ref, out = torch.relu(conv(blk(c))), blk.lookup_convc1(c, packed)
    Q = B * H * W
    foot = inbounds_footprint(c.cpu(), H, W)
    byts = Q * (foot * 4 + 256 * 4 + 8)                       # in-bounds footprint read + 256-channel output + coords
    flop = 2.0 * Q * K * 256
    emit(row="f1 lookup fused into convc1 (1x1, 324 -> 256) + ReLU", geometry=f"B={B} {H}x{W}", lookup_ms=ms_look,
         separate_ms=ms_sep, fused_ms=ms_fused, speedup=ms_sep / ms_fused, bytes_algorithmic=byts, gbs=byts / ms_fused / 1e6,
         bound="hbm", peak=hbm, frac=byts / ms_fused / 1e6 / hbm, gflop=flop / 1e9, tflops_useful=flop / ms_fused / 1e9,
         max_rel_diff_vs_separate_tf32_conv=float((out - ref).abs().max() / ref.abs().max()))


def row_fnet_tail(B, H, W, reps, cin=128):
    """Row f3: fnet's 1x1 output convolution (128 -> 256) fused into the volume build.  'separate' = torch conv2d
    (cuDNN, torch default math) + fc_build (pack + GEMM); 'fused' = fc_build_from_fnet_tail (conv on the tensor cores writing
    the packed operands) -- same pyramid."""
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(2 * B, cin, H, W, generator=g)).cuda()
    conv2 = torch.nn.Conv2d(cin, D, kernel_size=1).cuda()
    packed = ops.fnet_tail_prepare(conv2.weight.detach(), conv2.bias.detach())
    keep = fsb.CorrBlock.math
    fsb.CorrBlock.math = "3xbf16"
    try:
        with torch.no_grad():
            def separate():
                f1, f2 = torch.split(conv2(x), [B, B], dim=0)
                return fsb.CorrBlock(f1.float(), f2.float(), L, R)
            ms_conv = timed(lambda: conv2(x), reps, inner=4)
            ms_sep = timed(separate, reps)
            ms_fused = timed(lambda: fsb.CorrBlock.from_fnet_tail(x, packed, D, L, R), reps)
            a, b = separate()._state.pyramid, fsb.CorrBlock.from_fnet_tail(x, packed, D, L, R)._state.pyramid
            diff = float((a - b).abs().max() / a.abs().max())
    finally:
        fsb.CorrBlock.math = keep
    emit(row="f3 fnet tail fused into the build (conv2 1x1 128->256 + build)", geometry=f"B={B} {H}x{W}", conv2_cudnn_ms=ms_conv,
         separate_ms=ms_sep, fused_ms=ms_fused, saved_ms=ms_sep - ms_fused, max_rel_diff_vs_separate_tf32_conv=diff)


def row_model(B, Himg, Wimg, reps, iters=12):
    """BASELINE.json's metric itself: RAFT pairs/s at Sintel resolution, the UNMODIFIED reference model
    (baseline/_ref, random init, stock torch settings) on this GPU with (1) its own CorrBlock, (2) the drop-in block
    through patch_reference(), (3) RaftRunner eager, (4) RaftRunner with the whole forward in one CUDA graph."""
    try:
        from baseline import install_ref
        p = install_ref.install()
        if p is None:
            raise FileNotFoundError("baseline/_ref not installed")
        if p not in sys.path:
            sys.path.insert(0, p)
        import argparse as _ap
        from core.raft import RAFT
    except Exception as e:                                    # noqa: BLE001
        emit(row="RAFT forward pairs/s (reference model)", unavailable=repr(e))
        return
    Hp, Wp = (Himg + 7) // 8 * 8, (Wimg + 7) // 8 * 8
    torch.manual_seed(1234)
    model = RAFT(_ap.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval().cuda()
    g = torch.Generator().manual_seed(0)
    im1 = (torch.rand(B, 3, Hp, Wp, generator=g) * 255.0).cuda()
    im2 = (torch.rand(B, 3, Hp, Wp, generator=g) * 255.0).cuda()
    res = {}
    with torch.no_grad():
        res["reference_block_ms"] = timed(lambda: model(im1, im2, iters=iters, test_mode=True), reps, warm=2)
        fsb.patch_reference()
        try:
            res["dropin_block_ms"] = timed(lambda: model(im1, im2, iters=iters, test_mode=True), reps, warm=2)
        finally:
            fsb.unpatch_reference()
        eager = fsb.RaftRunner(model, iters=iters, graph=False)
        res["runner_eager_ms"] = timed(lambda: eager(im1, im2), reps, warm=2)
        graphed = fsb.RaftRunner(model, iters=iters, graph=True)
        res["runner_graph_ms"] = timed(lambda: graphed(im1, im2), reps, warm=2)
        tail = fsb.RaftRunner(model, iters=iters, graph=True, fused_fnet_tail=True)
        res["runner_graph_fused_tail_ms"] = timed(lambda: tail(im1, im2), reps, warm=2)
        allf = fsb.RaftRunner(model, iters=iters, graph=True, fused_fnet_tail=True, fused_convc1=True)
        res["runner_graph_fused_tail_convc1_ms"] = timed(lambda: allf(im1, im2), reps, warm=2)
    emit(row="RAFT forward pairs/s @436x1024, 12 iterations: unmodified reference model, same GPU",
         geometry=f"B={B} {Hp}x{Wp} px", conv_math="torch defaults (cuDNN TF32 convolutions)", **res,
         pairs_per_s={k[:-3]: B / v * 1e3 for k, v in res.items()},
         speedup_vs_reference_block={k[:-3]: res["reference_block_ms"] / v for k, v in res.items()})
    del model, eager, graphed, tail, allf
    torch.cuda.empty_cache()


USE_REF = True


def row_backward(B, H, W, reps):
    hbm, tf = peaks()
    g = torch.Generator().manual_seed(0)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    gout = torch.randn(B, K, H, W, generator=g).cuda()
    N, Q = H * W, B * H * W
    numel = ops.pyramid_numel(B, H, W, L)
    gp = torch.zeros(numel, device="cuda")
    foot = inbounds_footprint(c.cpu(), H, W)
    eager_ms = timed(lambda: ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA), reps, inner=12)   # 12 lookups per block
    # the kernel is shorter than the host's ~45 us per custom-op call: 12 back-to-back launches replayed from a CUDA graph
    # give the device time (the eager figure is kept beside it)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(12):
            ops.lookup_bwd(gout, c, gp, L, R, _lib.COORD_CUDA)
    ms = timed(gr.replay, reps) / 12
    byts = Q * (K * 4 + 2 * foot * 4 + 8)
    emit(row="a6 lookup_bwd", geometry=f"B={B} {H}x{W}", ms=ms, eager_ms_host_bound=eager_ms, bytes_algorithmic=byts,
         gbs=byts / ms / 1e6, bound="hbm", peak=hbm, frac=byts / ms / 1e6 / hbm,
         note="per launch, 12 launches replayed from a CUDA graph: grad read + footprint read-modify-write (in-bounds discounted) + coords")
    modes = (("fp32", _lib.MATH_FP32), ("3xbf16", _lib.MATH_TC_3XBF16)) if __name__ == "__main__" else (("3xbf16", _lib.MATH_TC_3XBF16),)
    for math_name, math in modes:
        def setup():
            gp.normal_()
        try:
            ms = timed(lambda: ops.build_bwd(gp, f1, f2, L, math), reps, setup=setup)
        except RuntimeError as e:
            emit(row="a6 build_bwd", math=math_name, error=str(e)); continue
        flop = 2 * 2.0 * B * N * N * D
        byts = numel * 4 + 4 * B * D * N * 4
        emit(row="a6 build_bwd (fold + dF1 + dF2)", math=math_name, geometry=f"B={B} {H}x{W}", ms=ms,
             gflop=flop / 1e9, tflops=flop / ms / 1e9, bytes_algorithmic=byts, gbs=byts / ms / 1e6,
             bound="tensor", peak=tf, frac=flop / ms / 1e9 / tf,
             hbm_floor_ms=byts / hbm / 1e6, tensor_floor_ms=flop / tf / 1e9)


def row_ondemand(B, H, W, reps):
    hbm, tf = peaks()
    g = torch.Generator().manual_seed(1)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    c = (fsb.coords_grid(B, H, W) + 8.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    Q = B * H * W
    keep = fsb.AlternateCorrBlock.route
    fsb.AlternateCorrBlock.route = "ondemand"
    ms_prep = timed(lambda: fsb.AlternateCorrBlock(f1, f2, L, R), reps)
    blk = fsb.AlternateCorrBlock(f1, f2, L, R)
    ms = timed(lambda: blk(c), reps)
    # the same block through the materialised route (tensor-core build once + HBM-bound lookups)
    fsb.AlternateCorrBlock.route = "materialise"
    try:
        ms_build = timed(lambda: fsb.AlternateCorrBlock(f1, f2, L, R), max(2, reps // 2), warm=1)
        mblk = fsb.AlternateCorrBlock(f1, f2, L, R)
        ms_look = timed(lambda: mblk(c), reps, inner=12)
        diff = float((mblk(c) - blk(c)).abs().max() / blk(c).abs().max())
        N = H * W
        emit(row="a7 AlternateCorrBlock route=materialise (build once + FC_COORD_RAW lookups)",
             geometry=f"B={B} {H}x{W}", build_ms=ms_build, lookup_ms=ms_look,
             pyramid_gb=4 * ops.pyramid_numel(B, H, W, L) / 1e9,
             build_tflops_useful=2.0 * B * N * N * D / ms_build / 1e9,
             ms_12_lookups={"materialise": ms_build + 12 * ms_look, "ondemand": ms_prep + 12 * ms},
             ms_32_lookups={"materialise": ms_build + 32 * ms_look, "ondemand": ms_prep + 32 * ms},
             max_rel_diff_vs_ondemand=diff)
        del mblk
    except RuntimeError as e:
        emit(row="a7 AlternateCorrBlock route=materialise", geometry=f"B={B} {H}x{W}", error=str(e))
    finally:
        fsb.AlternateCorrBlock.route = keep
        torch.cuda.empty_cache()
    flop = Q * L * (2 * R + 2) ** 2 * 2.0 * D
    byts = Q * (D * 4 + K * 4 + 8) + sum(B * (H >> l) * (W >> l) * D * 4 for l in range(L))
    line = dict(row="a7 ondemand_fwd (4 levels, one launch)", geometry=f"B={B} {H}x{W}", ms=ms, prepare_ms=ms_prep,
                gflop=flop / 1e9, tflops=flop / ms / 1e9, bytes_algorithmic=byts, hbm_floor_ms=byts / hbm / 1e6,
                query_lookups_per_s=Q / ms * 1e3)
    try:
        from oracle import ref_ext
        if USE_REF and ref_ext.available():
            rblk = ref_ext.RefAlternateCorrBlock(f1, f2, L, R)
            want = rblk(c)
            got = blk(c)
            line["max_rel_diff_vs_reference_kernel"] = float((got - want).abs().max() / want.abs().max())
            line["reference_kernel_ms"] = timed(lambda: rblk(c), max(2, reps // 3), warm=1)
            line["speedup_vs_reference_kernel"] = line["reference_kernel_ms"] / ms
            # single level, raw extension signature (rows a8 / a9)
            from flow_supervisor_b200 import alt_cuda_corr as shim
            mod = ref_ext.load()
            n1, n2 = rblk.q_nhwc, rblk.t_nhwc[0]
            cc = c.permute(0, 2, 3, 1).reshape(B, 1, H, W, 2).contiguous()
            gg = torch.randn(B, 1, 81, H, W, generator=g).cuda()
            a8 = timed(lambda: shim.forward(n1, n2, cc, R), reps)
            a8r = timed(lambda: mod.forward(n1, n2, cc, R), max(2, reps // 3), warm=1)
            a9 = timed(lambda: shim.backward(n1, n2, cc, gg, R), reps)
            a9r = timed(lambda: mod.backward(n1, n2, cc, gg, R), max(2, reps // 3), warm=1)
            emit(row="a8 altcorr_fwd (level 0)", geometry=f"B={B} {H}x{W}", ms=a8, reference_kernel_ms=a8r, speedup=a8r / a8)
            emit(row="a9 altcorr_bwd (level 0)", geometry=f"B={B} {H}x{W}", ms=a9, reference_kernel_ms=a9r, speedup=a9r / a9)
        else:
            line["reference_kernel_ms"] = None
    except Exception as e:                                    # noqa: BLE001
        line["reference_kernel_error"] = repr(e)
    emit(**line)


def row_same_gpu(B, Himg, Wimg, reps, iters=12):
    """Standalone only (imports oracle/): the reference's own PyTorch path (matmul + avg_pool2d + grid_sample,
    oracle/corr_torch.py = corr.py's library calls) ON THIS GPU next to the drop-in block, for the bare correlation
    path and inside the whole RAFT forward (oracle/raft_model.py, random init, BASELINE.json configs[1] geometry)."""
    from oracle import corr_torch, raft_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    H, W = (Himg + 7) // 8, Wimg // 8
    g = torch.Generator().manual_seed(0)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    cs = [(fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda() for _ in range(iters)]

    def path(block):
        blk = block(f1, f2, L, R)
        for c in cs:
            out = blk(c)
        return out

    with torch.no_grad():
        ms_ref = timed(lambda: path(corr_torch.TorchCorrBlock), max(2, reps // 3), warm=1)
        ms_our = timed(lambda: path(fsb.CorrBlock), reps)
    emit(row="corr path on the same GPU (build + 12 lookups): torch ops vs this library", geometry=f"B={B} {H}x{W}",
         torch_ops_ms=ms_ref, ms=ms_our, speedup=ms_ref / ms_our)
    torch.manual_seed(1234)
    model = raft_model.Raft().eval().cuda()
    im1 = torch.rand(B, 3, 8 * H, Wimg, generator=g).cuda() * 255.0
    im2 = torch.rand(B, 3, 8 * H, Wimg, generator=g).cuda() * 255.0
    with torch.no_grad():
        ms_ref = timed(lambda: model(im1, im2, iters=iters, corr_block=corr_torch.TorchCorrBlock), max(2, reps // 3), warm=1)
        ms_our = timed(lambda: model(im1, im2, iters=iters, corr_block=fsb.CorrBlock), max(2, reps // 2), warm=1)
    emit(row="RAFT forward, 12 iterations, fp32 convolutions (oracle/raft_model.py): reference-path block vs drop-in block",
         geometry=f"B={B} {8 * H}x{Wimg} px", torch_block_ms=ms_ref, ms=ms_our, pairs_per_s_ref=B / ms_ref * 1e3,
         pairs_per_s=B / ms_our * 1e3, speedup=ms_ref / ms_our)
    del model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rows", default="bwd,ondemand")
    a = ap.parse_args()
    if "bwd" in a.rows:
        row_backward(6, 46, 96, a.reps)
        row_backward(6, 54, 128, a.reps)
    if "ondemand" in a.rows:
        row_ondemand(2, 136, 240, a.reps)
    if "samegpu" in a.rows:
        row_same_gpu(8, 436, 1024, a.reps)
    if "convc1" in a.rows:
        row_lookup_convc1(8, 55, 128, a.reps)
    if "fnet" in a.rows:
        row_fnet_tail(8, 55, 128, a.reps)
