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
          (1, 256, 24, 132)]


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
