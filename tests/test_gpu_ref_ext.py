"""Rows a7/a8/a9 of SURVEY.md section 8 pinned to the REFERENCE'S OWN CUDA KERNELS.

oracle/_ref/alt_cuda_corr_ref.so is /root/reference/pytorch/alt_cuda_corr compiled for
sm_100a from the sources where they lie (oracle/build_ref.py); these tests run it on the
same GPU, on the same seeded inputs, next to libflowcorr's fc_altcorr_fwd / fc_altcorr_bwd
(through the alt_cuda_corr shim, i.e. through the C ABI) and AlternateCorrBlock.
Tolerance: <= 1e-4 of the tensor's max magnitude (fp32 summation order differs: the
reference accumulates 32-channel chunks through global memory).

Shapes are multiples of the reference's 4 x 8 thread block: its kernels read coordinates
of out-of-image threads without a bounds check (correlation_kernel.cu:62-68 leaves x2s
uninitialised in the forward, :176-177 reads past the tensor in the backward), so ragged
shapes are undefined behaviour THERE; the ragged case below is forward-only and compares
in-image outputs, which that defect cannot reach.
"""
import math

import pytest
import torch

from oracle import ref_ext

pytestmark = pytest.mark.gpu
VAL_TOL = 1e-4


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_ext.available():
        pytest.skip("oracle/_ref/alt_cuda_corr_ref.so not built (python oracle/build_ref.py)")
    import flow_supervisor_b200 as fsb
    from flow_supervisor_b200 import alt_cuda_corr as shim
    return fsb, shim, ref_ext.load()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def case(B, C, H1, W1, H2, W2, seed, flow_std, level=0):
    g = torch.Generator().manual_seed(seed)
    f1 = (1.57 * torch.randn(B, H1, W1, C, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, H2, W2, C, generator=g)).cuda()
    ys, xs = torch.meshgrid(torch.arange(H1), torch.arange(W1), indexing="ij")
    grid = torch.stack([xs, ys], -1).float()[None, None]                       # (1,1,H1,W1,2) [x,y]
    c = ((grid + flow_std * torch.randn(B, 1, H1, W1, 2, generator=g)) / 2 ** level).contiguous().cuda()
    return f1, f2, c


@pytest.mark.parametrize("B,C,H1,W1,H2,W2,r,level,std", [
    (2, 256, 24, 40, 24, 40, 4, 0, 3.0),       # level 0
    (1, 256, 32, 48, 16, 24, 4, 1, 6.0),       # pooled target map (level 1 geometry)
    (1, 128, 16, 32, 4, 8, 3, 2, 10.0),        # RAFT-small radius, level 2, heavy out-of-bounds
    (1, 64, 8, 16, 8, 16, 4, 0, 0.0),          # lattice coordinates (iteration 0): dx = dy = 0
])
def test_forward_matches_compiled_reference(env, B, C, H1, W1, H2, W2, r, level, std):
    fsb, shim, ref = env
    f1, f2, c = case(B, C, H1, W1, H2, W2, seed=41 + level, flow_std=std, level=level)
    want, = ref.forward(f1, f2, c, r)
    got, = shim.forward(f1, f2, c, r)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (B, 1, (2 * r + 1) ** 2, H1, W1)
    assert rel(got, want) < VAL_TOL
    # identical exact-zero pattern = identical out-of-bounds decisions
    assert torch.equal(got == 0, want == 0) or std == 0.0


def test_forward_ragged_shape_in_image_outputs(env):
    fsb, shim, ref = env
    f1, f2, c = case(1, 64, 22, 35, 22, 35, seed=47, flow_std=2.0)
    want, = ref.forward(f1, f2, c, 4)
    got, = shim.forward(f1, f2, c, 4)
    torch.cuda.synchronize()
    assert rel(got, want) < VAL_TOL


@pytest.mark.parametrize("B,C,H1,W1,H2,W2,r,level,std", [
    (2, 256, 16, 24, 16, 24, 4, 0, 2.0),
    (1, 128, 16, 32, 8, 16, 3, 1, 5.0),
    (1, 64, 8, 16, 8, 16, 4, 0, 0.0),
])
def test_backward_matches_compiled_reference(env, B, C, H1, W1, H2, W2, r, level, std):
    """Row a9: the reference's backward kernel has no Python caller (dead code) but is the
    only statement of the on-demand gradient; run it and compare."""
    fsb, shim, ref = env
    f1, f2, c = case(B, C, H1, W1, H2, W2, seed=53 + level, flow_std=std, level=level)
    g = torch.randn(B, 1, (2 * r + 1) ** 2, H1, W1, generator=torch.Generator().manual_seed(3)).cuda()
    w1, w2, wc = ref.backward(f1, f2, c, g, r)
    d1, d2, dc = shim.backward(f1, f2, c, g, r)
    torch.cuda.synchronize()
    assert rel(d1, w1) < VAL_TOL
    assert rel(d2, w2) < VAL_TOL
    assert not wc.any() and not dc.any()            # coords_grad is never written (correlation_kernel.cu:307)


def test_alternate_block_matches_reference_call_pattern(env):
    """Row a7: AlternateCorrBlock (4 levels, one launch here) against the reference's call
    pattern around its own kernel (corr.py:63-91 -> oracle/ref_ext.RefAlternateCorrBlock)."""
    fsb, shim, ref = env
    g = torch.Generator().manual_seed(59)
    B, D, H, W = 2, 256, 48, 64
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    c = (fsb.coords_grid(B, H, W) + 6.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    want = ref_ext.RefAlternateCorrBlock(f1, f2, 4, 4)(c)
    got = fsb.AlternateCorrBlock(f1, f2, 4, 4)(c)
    torch.cuda.synchronize()
    assert got.shape == want.shape == (B, 324, H, W)
    assert rel(got, want) < VAL_TOL
    # and the materialised block agrees with the reference's on-demand kernel at the 1e-4 level
    # away from lattice points
    full = fsb.CorrBlock(f1, f2, 4, 4)(c)
    assert rel(full, want) < VAL_TOL


def test_error_behaviour_matches(env):
    """CHECK_INPUT (correlation.cpp:19-21): CPU or non-contiguous inputs raise RuntimeError in
    both implementations."""
    fsb, shim, ref = env
    f1, f2, c = case(1, 64, 8, 16, 8, 16, seed=61, flow_std=1.0)
    for mod in (ref, shim):
        with pytest.raises(RuntimeError):
            mod.forward(f1.cpu(), f2, c, 4)
        with pytest.raises(RuntimeError):
            mod.forward(f1.permute(0, 2, 1, 3), f2, c, 4)


def test_scale_convention(env):
    """The extension returns UNSCALED dot products; the caller divides (corr.py:91)."""
    fsb, shim, ref = env
    f1, f2, c = case(1, 64, 8, 16, 8, 16, seed=67, flow_std=0.0)
    got, = shim.forward(f1, f2, c, 4)
    centre = got[0, 0, 4 * 9 + 4]                                               # zero offset tap
    want = (f1[0] * f2[0]).sum(-1)
    assert rel(centre, want) < VAL_TOL
    assert math.isfinite(float(centre.sum()))
