"""GPU: the bf16 VOLUME mode (CorrBlock.volume = 'bf16', FC_VOL_BF16; SURVEY.md section 7 step 5,
Appendix A.5): the build epilogue rounds the fp32 accumulators (and the fp32-pooled levels) to bf16
once, the lookup reads 2-byte elements.  Inference only; the headline mode stays fp32.

Stated tolerances (separate from the fp32 contract): volume elements = RN-bf16 of the fp32
volume, bit for bit; lookup values <= 2^-8 (3.9e-3) of the tensor's max magnitude against the
fp32-volume lookup; final flow <= 0.05 px mean EPE after 12 iterations (tests/test_gpu_reference_models.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(2, 256, 55, 128), (1, 256, 46, 62), (1, 64, 17, 19), (1, 256, 47, 156), (1, 64, 19, 240), (1, 64, 18, 78)]


@pytest.fixture(scope="module")
def fsb():
    import flow_supervisor_b200 as m
    return m


def blocks(fsb, f1, f2, L=4, r=4):
    """-> (fp32-volume block, bf16-volume block) built with the same arithmetic."""
    old = fsb.CorrBlock.math, fsb.CorrBlock.volume
    fsb.CorrBlock.math = "3xbf16"
    try:
        fsb.CorrBlock.volume = "f32"
        a = fsb.CorrBlock(f1, f2, num_levels=L, radius=r)
        fsb.CorrBlock.volume = "bf16"
        b = fsb.CorrBlock(f1, f2, num_levels=L, radius=r)
    finally:
        fsb.CorrBlock.math, fsb.CorrBlock.volume = old
    return a, b


@pytest.mark.parametrize("shape", SHAPES)
def test_bf16_volume_is_the_rounded_fp32_volume(fsb, shape):
    B, D, H, W = shape
    g = torch.Generator().manual_seed(5)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g) + 0.3).cuda()
    a, b = blocks(fsb, f1, f2)
    pa, pb = a._state.pyramid, b._state.pyramid
    assert pb.dtype == torch.bfloat16 and pb.numel() == pa.numel()
    assert torch.equal(pb, pa.to(torch.bfloat16))            # every level, pads (zeros) included
    for la, lb in zip(a.corr_pyramid, b.corr_pyramid):
        assert la.shape == lb.shape and lb.dtype == torch.bfloat16


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("law", ["random", "lattice", "border"])
def test_bf16_lookup_matches_fp32_lookup_of_the_same_values(fsb, shape, law):
    """The bf16 lookup kernel and the fp32 lookup kernel run the same arithmetic: fed the same
    (bf16-representable) volume they agree bit for bit; against the unrounded volume the stated
    tolerance holds."""
    B, D, H, W = shape
    g = torch.Generator().manual_seed(6)
    f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
    grid = fsb.coords_grid(B, H, W, device="cuda")
    if law == "random":
        c = grid + 5.0 * torch.randn(B, 2, H, W, generator=g).cuda()
    elif law == "lattice":
        c = grid + torch.randint(-3, 4, (B, 2, H, W), generator=g).float().cuda()
    else:
        c = grid + 40.0 * torch.randn(B, 2, H, W, generator=g).cuda()
    a, b = blocks(fsb, f1, f2)
    out_b = b(c)
    st = a._state
    same_values = fsb.ops.lookup(b._state.pyramid.float(), c, st.L, st.radius, st.coord)
    assert torch.equal(out_b, same_values)
    out_a = a(c)
    scale = float(out_a.abs().max())
    assert float((out_b - out_a).abs().max()) <= 2 ** -8 * scale
    # structural zeros (taps outside the map) are zeros in both; a handful of sums may cancel to zero in one only
    assert float(((out_b == 0) != (out_a == 0)).float().mean()) < 1e-3


def test_bf16_volume_small_radius_and_levels(fsb):
    g = torch.Generator().manual_seed(7)
    f1 = torch.randn(1, 128, 24, 40, generator=g).cuda()
    f2 = torch.randn(1, 128, 24, 40, generator=g).cuda()
    c = fsb.coords_grid(1, 24, 40, device="cuda") + 3.0 * torch.randn(1, 2, 24, 40, generator=g).cuda()
    for L, r in ((4, 3), (2, 4), (1, 2)):
        a, b = blocks(fsb, f1, f2, L, r)
        oa, ob = a(c), b(c)
        assert tuple(ob.shape) == (1, L * (2 * r + 1) ** 2, 24, 40)
        assert float((ob - oa).abs().max()) <= 2 ** -8 * float(oa.abs().max())


def test_bf16_volume_is_inference_only_and_needs_the_tensor_core_build(fsb):
    f = torch.randn(1, 64, 16, 24, device="cuda")
    old = fsb.CorrBlock.volume, fsb.CorrBlock.math
    fsb.CorrBlock.volume = "bf16"
    try:
        with pytest.raises(RuntimeError, match="fp32 volume"):
            fsb.CorrBlock(f.clone().requires_grad_(), f)
        fsb.CorrBlock.math = "fp32"
        with pytest.raises(RuntimeError, match="fp32 volume"):
            fsb.CorrBlock(f, f)
    finally:
        fsb.CorrBlock.volume, fsb.CorrBlock.math = old
