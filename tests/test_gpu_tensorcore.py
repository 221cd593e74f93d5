"""GPU: the tcgen05 build modes against the fp32 mode and the fp64-exact oracle.

Tolerances: '3xbf16' (parity mode) <= 2e-5 of the volume's max magnitude, i.e. well
inside the 1e-4 contract; 'bf16' (stated separately, BASELINE.json north_star) <= 1e-2
of the max magnitude (bf16 inputs carry 2^-9 relative rounding each)."""
import numpy as np
import pytest
import torch

from oracle import corr_spec

pytestmark = pytest.mark.gpu

SHAPES = [(1, 256, 46, 62), (2, 256, 55, 128), (1, 64, 17, 19), (1, 256, 47, 156), (2, 128, 16, 24),
          (1, 256, 24, 132),                  # row pairs split 192 + 80 accumulator columns
          (1, 64, 19, 240), (1, 64, 10, 250),   # cfg 5 width (256 + 224) and the widest map (256 + 256)
          (1, 64, 18, 78), (1, 64, 12, 44)]     # 5 and 3 chunks per row pair: partially filled level-2/3 patches


@pytest.fixture(scope="module")
def fsb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import flow_supervisor_b200 as m
    return m


def build(fsb, f1, f2, math):
    old = fsb.CorrBlock.math
    fsb.CorrBlock.math = math
    try:
        return fsb.CorrBlock(f1, f2)
    finally:
        fsb.CorrBlock.math = old


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("math,tol", [("3xbf16", 2e-5), ("bf16", 1e-2)])
def test_tc_volume_matches_exact(fsb, shape, math, tol):
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(41)
    f1 = 1.57 * torch.randn(B, D, H, W, generator=gen)
    f2 = 1.57 * torch.randn(B, D, H, W, generator=gen) + 0.3
    exact = corr_spec.all_pairs(f1.numpy(), f2.numpy(), exact=True)
    blk = build(fsb, f1.cuda(), f2.cuda(), math)
    lv = [v.cpu().numpy()[:, 0] for v in blk.corr_pyramid]
    got = lv[0].reshape(exact.shape)
    err = float(np.abs(got - exact).max() / np.abs(exact).max())
    assert err < tol, err
    # the pooled levels are bit-exact 2x2 means of the kernel's own level 0
    ref = corr_spec.pool_pyramid(lv[0], 4)
    for l in range(1, 4):
        assert np.array_equal(lv[l], ref[l]), l
    # pad columns / pad rows of level 0 hold exact zeros (pooled-level pads are never read)
    lvl0 = fsb.ops.level_padded(blk._state.pyramid, B, H, W, 4)[0]
    assert not lvl0[:, H:, :].any() and not lvl0[:, :, W:].any()


def test_tc_lookup_end_to_end(fsb):
    B, D, H, W = 2, 256, 46, 96
    gen = torch.Generator().manual_seed(43)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    c = fsb.coords_grid(B, H, W, device="cuda") + 4.0 * torch.randn(B, 2, H, W, generator=gen).cuda()
    ref = build(fsb, f1, f2, "fp32")(c)
    out = build(fsb, f1, f2, "3xbf16")(c)
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5
    out = build(fsb, f1, f2, "bf16")(c)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-2


@pytest.mark.parametrize("L", [1, 2, 3])
def test_tc_fused_pyramid_fewer_levels(fsb, L):
    B, D, H, W = 1, 128, 21, 50
    gen = torch.Generator().manual_seed(47)
    f1 = torch.randn(B, D, H, W, generator=gen).cuda()
    f2 = torch.randn(B, D, H, W, generator=gen).cuda()
    old = fsb.CorrBlock.math
    fsb.CorrBlock.math = "3xbf16"
    try:
        blk = fsb.CorrBlock(f1, f2, num_levels=L, radius=3)
    finally:
        fsb.CorrBlock.math = old
    lv = [v.cpu().numpy()[:, 0] for v in blk.corr_pyramid]
    assert len(lv) == L
    exact = corr_spec.all_pairs(f1.cpu().numpy(), f2.cpu().numpy(), exact=True)
    assert float(np.abs(lv[0].reshape(exact.shape) - exact).max() / np.abs(exact).max()) < 2e-5
    ref = corr_spec.pool_pyramid(lv[0], L)
    for l in range(1, L):
        assert np.array_equal(lv[l], ref[l]), l


# ------------------------------------------------------------------ backward (fc_bwd_tc.cu)
def _grads(fsb, math, f1, f2, coords, gouts, L=4, r=4):
    a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    old = fsb.CorrBlock.math
    fsb.CorrBlock.math = math
    try:
        blk = fsb.CorrBlock(a, b, num_levels=L, radius=r)
        outs = [blk(c) for c in coords]
        torch.autograd.backward(outs, gouts)
    finally:
        fsb.CorrBlock.math = old
    return a.grad, b.grad


@pytest.mark.parametrize("shape,L,r", [((2, 256, 46, 62), 4, 4),      # cfg 1 geometry, ragged 256-row tiles
                                       ((1, 256, 55, 128), 4, 4),     # Sintel geometry, odd H (pad row)
                                       ((1, 64, 19, 27), 4, 4),       # odd everything, D = 64
                                       ((2, 128, 24, 40), 3, 3),      # RAFT-small
                                       ((1, 192, 16, 24), 1, 4),      # single level: no fold
                                       ((1, 256, 47, 156), 4, 4),     # KITTI geometry (Wp = 160)
                                       ((1, 64, 64, 96), 6, 2),       # six levels: the two coarsest are read at use
                                       ((1, 64, 35, 50), 5, 2)])      # five levels, odd maps
@pytest.mark.parametrize("math,tol", [("3xbf16", 3e-5), ("bf16", 2e-2)])
def test_tc_backward_matches_fp32_mode(fsb, shape, L, r, math, tol):
    """The two tcgen05 GEMMs with the fold + bf16 split of the gradient pyramid inside them (K-major and MN-major
    reads of the same converted tile) against the fp32 CUDA-core mode of the same library; three
    lookups accumulate into one gradient pyramid.  3xbf16 <= 3e-5 of the gradient's max
    magnitude (inside the 1e-4 contract); bf16 stated separately <= 2e-2."""
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(61)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=gen) + 0.2).cuda()
    grid = fsb.coords_grid(B, H, W)
    coords = [(grid + s * torch.randn(B, 2, H, W, generator=gen)).cuda() for s in (0.0, 3.0, 12.0)]
    K = L * (2 * r + 1) ** 2
    gouts = [torch.randn(B, K, H, W, generator=gen).cuda() for _ in coords]
    r1, r2 = _grads(fsb, "fp32", f1, f2, coords, gouts, L, r)
    d1, d2 = _grads(fsb, math, f1, f2, coords, gouts, L, r)
    e1 = float((d1 - r1).abs().max() / r1.abs().max())
    e2 = float((d2 - r2).abs().max() / r2.abs().max())
    assert e1 < tol and e2 < tol, (e1, e2)


def _build_bwd_switch(fsb, value):
    from flow_supervisor_b200 import _lib
    _lib.check(_lib.load().fc_tunable_set(b"bwd_fused", int(value)), "fc_tunable_set")


@pytest.mark.parametrize("shape,L", [((2, 256, 46, 62), 4), ((1, 64, 19, 27), 4), ((1, 256, 47, 156), 4),
                                     ((2, 128, 24, 40), 3), ((1, 64, 64, 96), 6)])
@pytest.mark.parametrize("math", ["3xbf16", "bf16"])
def test_tc_backward_fold_in_gemm_matches_the_separate_fold_pass(fsb, shape, L, math):
    """fc_build_bwd's default (fold + bf16 split inside the GEMMs, gradient pyramid only read; maps with a multiple of 4
    patches per row take the kernel whose coarse cells arrive as TMA boxes -- (46, 62), (19, 27), (47, 156) here --, the
    others the generic one) against the generic kernel (FLOWCORR_BWD_FUSED=2) and the round-1 pipeline (fold + pack pass,
    GEMMs from the in-place planes; FLOWCORR_BWD_FUSED=0): same products of the same rounded operands, only the
    accumulation order of the split-K pieces differs -> <= 2e-6 of the gradient's max; and the default leaves the
    gradient pyramid untouched."""
    from flow_supervisor_b200 import _lib, ops
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(11)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=gen) + 0.1).cuda()
    src = ops.clear_pads_(torch.randn(ops.pyramid_numel(B, H, W, L), generator=gen).cuda(), B, H, W, L)
    m = _lib.MATH_TC_3XBF16 if math == "3xbf16" else _lib.MATH_TC_BF16
    try:
        gp = src.clone()
        a1, a2 = ops.build_bwd(gp, f1, f2, L, m)
        torch.cuda.synchronize()
        assert torch.equal(gp, src), "the default backward must not modify the gradient pyramid"
        others = []
        for sw in (2, 0):            # 2: the generic fold-in-GEMM kernel (any map), 0: the round-1 pipeline
            _build_bwd_switch(fsb, sw)
            others.append(ops.build_bwd(src.clone(), f1, f2, L, m))
    finally:
        _build_bwd_switch(fsb, 1)
    for b1, b2 in others:
        for a, b in ((a1, b1), (a2, b2)):
            assert float((a - b).abs().max() / b.abs().max()) < 2e-6


def test_tc_backward_takes_maps_past_the_old_row_in_shared_memory_limit(fsb):
    """136x240 tokens (cfg 5's 1/8 map, 32 640 padded targets): the round-1 backward fell to the fp32 CUDA-core
    contractions above 16 384 targets; the fold-in-GEMM kernels have no such limit.  Checked against that fp32 mode."""
    from flow_supervisor_b200 import _lib, ops
    B, D, H, W, L = 1, 64, 136, 240, 4
    assert _lib.load().fc_build_bwd_workspace_bytes(B, D, H, W, L, _lib.MATH_TC_3XBF16) > 0     # tensor-core route taken
    gen = torch.Generator().manual_seed(5)
    f1 = torch.randn(B, D, H, W, generator=gen).cuda()
    f2 = torch.randn(B, D, H, W, generator=gen).cuda()
    src = torch.zeros(ops.pyramid_numel(B, H, W, L), device="cuda")
    # a sparse gradient pyramid (a dense one is 5.7 GB of randn): 2 M random cells over all levels
    idx = torch.randint(0, src.numel(), (2_000_000,), generator=gen).cuda()
    src[idx] = torch.randn(idx.numel(), generator=gen).cuda()
    ops.clear_pads_(src, B, H, W, L)
    d1, d2 = ops.build_bwd(src.clone(), f1, f2, L, _lib.MATH_TC_3XBF16)
    r1, r2 = ops.build_bwd(src.clone(), f1, f2, L, _lib.MATH_FP32)
    for a, b in ((d1, r1), (d2, r2)):
        assert float((a - b).abs().max() / b.abs().max()) < 3e-5


@pytest.mark.parametrize("shape", [(1, 256, 46, 96), (1, 256, 54, 128), (2, 128, 23, 50)])
def test_tc_backward_matches_torch_autograd_of_the_reference_ops(fsb, shape):
    """The tcgen05 backward (lookup_bwd + fold + two GEMMs, 3xbf16) against torch autograd through the reference's
    own library calls (matmul / avg_pool2d / grid_sample, oracle/corr_torch.py) on the same GPU, at the configuration-3
    geometries (46x96 student crop, 54x128 teacher frame, D = 256), three accumulated lookups (an independent
    implementation, not this library's fp32 mode).  Tolerance 1e-4 of the gradient's max magnitude."""
    from oracle import corr_torch
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(71)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    cs = [(fsb.coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=gen)).cuda() for _ in range(3)]
    gs = [torch.randn(B, 324, H, W, generator=gen).cuda() for _ in range(3)]
    d1, d2 = _grads(fsb, "3xbf16", f1, f2, cs, gs)
    a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    keep = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        blk = corr_torch.TorchCorrBlock(a, b, 4, 4)
        loss = sum((blk(c) * g).sum() for c, g in zip(cs, gs))
        loss.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = keep
    e1 = float((d1 - a.grad).abs().max() / a.grad.abs().max())
    e2 = float((d2 - b.grad).abs().max() / b.grad.abs().max())
    assert e1 < 1e-4 and e2 < 1e-4, (e1, e2)


def test_tc_backward_matches_oracle(fsb):
    """... and against the numpy oracle's adjoint (oracle/corr_spec.py), fp64-free path."""
    B, D, H, W = 1, 64, 18, 26
    gen = torch.Generator().manual_seed(67)
    f1 = torch.randn(B, D, H, W, generator=gen)
    f2 = torch.randn(B, D, H, W, generator=gen)
    c = fsb.coords_grid(B, H, W) + 2.5 * torch.randn(B, 2, H, W, generator=gen)
    gout = torch.randn(B, 324, H, W, generator=gen)
    d1, d2 = _grads(fsb, "3xbf16", f1.cuda(), f2.cuda(), [c.cuda()], [gout.cuda()])
    G = corr_spec.lookup_backward(gout.numpy(), c.numpy(), corr_spec.level_shapes(H, W, 4), 4, "cuda")
    w1, w2 = corr_spec.build_backward(G, f1.numpy(), f2.numpy())
    assert float(np.abs(d1.cpu().numpy() - w1).max() / np.abs(w1).max()) < 1e-4
    assert float(np.abs(d2.cpu().numpy() - w2).max() / np.abs(w2).max()) < 1e-4


DEFAULTS = {"build_epi_warps": 4, "build_sched": 1, "build_stages": 0, "no_fuse": 0}


@pytest.mark.parametrize("env", [{"build_epi_warps": 8}, {"build_sched": 0}, {"build_stages": 2},
                                 {"build_epi_warps": 8, "build_sched": 0}, {"no_fuse": 1}])
@pytest.mark.parametrize("shape", [(2, 256, 55, 128), (1, 128, 21, 156), (3, 64, 17, 23)])
def test_tc_build_variants_are_bit_identical(fsb, env, shape):
    """The alternative epilogue width, unit schedule, ring depth and the unfused pooling launches (diagnostic switches,
    fc_tunable_set) change who computes a tile and when, never the arithmetic: the pyramid must come out bit for bit
    the same."""
    from flow_supervisor_b200 import _lib
    lib = _lib.load()
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(7)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=gen)).cuda()
    for k, v in DEFAULTS.items():
        _lib.check(lib.fc_tunable_set(k.encode(), v), "fc_tunable_set")
    ref = build(fsb, f1, f2, "3xbf16")._state.pyramid.clone()
    try:
        for k, v in env.items():
            _lib.check(lib.fc_tunable_set(k.encode(), v), "fc_tunable_set")
        got = build(fsb, f1, f2, "3xbf16")._state.pyramid
    finally:
        for k, v in DEFAULTS.items():
            lib.fc_tunable_set(k.encode(), v)
    lv_ref = fsb.ops.level_padded(ref, B, H, W, 4)
    lv_got = fsb.ops.level_padded(got, B, H, W, 4)
    for l in range(4):
        assert torch.equal(lv_ref[l], lv_got[l]), (env, l)
