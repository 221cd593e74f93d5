"""GPU: the opt-in pieces next to the correlation path:
``fc_upsample_flow`` against RAFT.upsample_flow's torch ops, and ``RaftRunner`` (whole forward in one
CUDA graph) against its own eager path (bit-identical) and against the unmodified reference forward."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refmodels as rm  # noqa: E402

pytestmark = pytest.mark.gpu


def ref_upsample(flow, mask):
    """raft.py:72-83, verbatim arithmetic."""
    N, _, H, W = flow.shape
    mask = mask.view(N, 1, 9, 8, 8, H, W)
    mask = torch.softmax(mask, dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


@pytest.mark.parametrize("shape", [(1, 46, 62), (2, 55, 128), (1, 7, 5), (3, 16, 33), (1, 1, 1)])
def test_upsample_flow_matches_reference_ops(shape):
    import flow_supervisor_b200 as fsb
    B, H, W = shape
    g = torch.Generator().manual_seed(3)
    flow = (4.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
    mask = (3.0 * torch.randn(B, 576, H, W, generator=g)).cuda()
    out = fsb.ops.upsample_flow(flow, mask)
    ref = ref_upsample(flow, mask)
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    # a peaked mask selects one neighbour exactly; the border reads zeros (unfold padding)
    mask2 = torch.full_like(mask, -1e4)
    mask2[:, 0 * 64:1 * 64] = 0.0                                          # k = 0: neighbour (h-1, w-1)
    out2 = fsb.ops.upsample_flow(flow, mask2)
    assert torch.equal(out2, ref_upsample(flow, mask2))
    assert not out2[:, :, :8].any() and not out2[:, :, :, :8].any()


def test_upsample_flow_rejects_bad_input():
    import flow_supervisor_b200 as fsb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsb.ops.upsample_flow(torch.zeros(1, 2, 4, 4), torch.zeros(1, 576, 4, 4))
    with pytest.raises(ValueError):
        fsb.ops.upsample_flow(torch.zeros(1, 2, 4, 4).cuda(), torch.zeros(1, 64, 4, 4).cuda())


@pytest.fixture(scope="module")
def core():
    return rm.core()


@pytest.mark.parametrize("kind,size", [("raft", (368, 496)), ("raft", (440, 1024)), ("gma", (376, 1248))])
def test_runner_graph_is_bit_identical_and_matches_the_reference_forward(core, kind, size):
    import flow_supervisor_b200 as fsb
    torch.manual_seed(1234)
    model = (core.raft.RAFT(rm.raft_args()) if kind == "raft" else core.gma_network.RAFTGMA(rm.gma_args())).eval().cuda()
    im1, im2 = (t.cuda() for t in rm.synth_pair(size[0], size[1], seed=9, batch=2))
    with rm.strict_fp32(), torch.no_grad():
        low_r, up_r = model(im1, im2, iters=12, test_mode=True)              # reference block, reference forward
        with rm.patched():
            low_p, up_p = model(im1, im2, iters=12, test_mode=True)          # drop-in block, reference forward
        eager = fsb.RaftRunner(model, iters=12, graph=False, fused_upsample=False)
        low_e, up_e = eager(im1, im2)
        graphed = fsb.RaftRunner(model, iters=12, graph=True, fused_upsample=False)
        low_g, up_g = graphed(im1, im2)
        low_g2, up_g2 = graphed(im1.flip(0), im2.flip(0))                     # replay with new inputs
        low_e2, up_e2 = eager(im1.flip(0), im2.flip(0))
        fused = fsb.RaftRunner(model, iters=12, graph=True, fused_upsample=True)
        low_f, up_f = fused(im1, im2)
    # same launches in the same order: the runner reproduces the patched reference forward exactly
    assert torch.equal(low_e, low_p) and torch.equal(up_e, up_p)
    # graph on / off: bit-identical flow
    assert torch.equal(low_g, low_e) and torch.equal(up_g, up_e)
    assert torch.equal(low_g2, low_e2) and torch.equal(up_g2, up_e2)
    assert rm.epe(up_g2, up_e.flip(0)) <= 1e-3                             # (cuDNN results may depend on the batch position)
    # fused upsampling: same low-resolution flow, upsampled flow to rounding
    assert torch.equal(low_f, low_e)
    assert float((up_f - up_e).abs().max()) <= 1e-4
    # and the whole thing against the reference's own block
    assert rm.epe(up_f, up_r) <= 0.01, rm.epe(up_f, up_r)


def test_runner_flow_init_and_mixed_precision(core):
    import flow_supervisor_b200 as fsb
    torch.manual_seed(1234)
    model = core.raft.RAFT(rm.raft_args(mixed_precision=True)).eval().cuda()
    im1, im2 = (t.cuda() for t in rm.synth_pair(368, 496, seed=10))
    init = torch.randn(1, 2, 46, 62, device="cuda")
    with torch.no_grad():
        with rm.patched():
            low_p, up_p = model(im1, im2, iters=6, flow_init=init, test_mode=True)
        low_g, up_g = fsb.RaftRunner(model, iters=6, graph=True, fused_upsample=False)(im1, im2, flow_init=init)
    assert torch.equal(low_g, low_p) and torch.equal(up_g, up_p)


@pytest.mark.parametrize("shape,cin,dout", [((2, 46, 62), 128, 256), ((1, 55, 128), 128, 256), ((1, 24, 40), 64, 128),
                                            ((2, 47, 156), 128, 256), ((1, 17, 20), 128, 64)])
@pytest.mark.parametrize("math", ["3xbf16", "bf16"])
def test_fused_fnet_tail_builds_the_same_pyramid(shape, cin, dout, math):
    """fc_build_from_fnet_tail (the encoder's 1x1 output convolution inside the volume build, row f3) against
    conv2 in fp32 followed by the ordinary build: every level within 1e-4 of the volume's scale in the parity mode
    (the convolution itself carries the 3xbf16 split), pads exactly zero."""
    import flow_supervisor_b200 as fsb
    B, H, W = shape
    g = torch.Generator().manual_seed(21)
    x = torch.relu(torch.randn(2 * B, cin, H, W, generator=g)).cuda()           # post-ReLU activations, like layer3's output
    conv2 = torch.nn.Conv2d(cin, dout, kernel_size=1).cuda()
    with torch.no_grad():
        conv2.weight.copy_(torch.randn(dout, cin, 1, 1, generator=g).cuda() * (2.0 / cin) ** 0.5)
        conv2.bias.copy_(0.1 * torch.randn(dout, generator=g).cuda())
    packed = fsb.ops.fnet_tail_prepare(conv2.weight.detach(), conv2.bias.detach())
    old = fsb.CorrBlock.math
    fsb.CorrBlock.math = math
    try:
        with rm.strict_fp32(), torch.no_grad():
            f1, f2 = torch.split(conv2(x), [B, B], dim=0)
            ref = fsb.CorrBlock(f1.contiguous(), f2.contiguous())
            fused = fsb.CorrBlock.from_fnet_tail(x, packed, dout)
    finally:
        fsb.CorrBlock.math = old
    tol = 1e-4 if math == "3xbf16" else 2e-2
    la = fsb.ops.level_padded(ref._state.pyramid, B, H, W, 4)
    lb = fsb.ops.level_padded(fused._state.pyramid, B, H, W, 4)
    scale = float(la[0].abs().max())
    for l, (a, b) in enumerate(zip(la, lb)):
        assert float((a - b).abs().max()) <= tol * scale, (l, float((a - b).abs().max()) / scale)
        assert not b[:, (H >> l):, :].any() and not b[:, :, (W >> l):].any()
    c = fsb.coords_grid(B, H, W, device="cuda") + 3.0 * torch.randn(B, 2, H, W, generator=g).cuda()
    assert float((fused(c) - ref(c)).abs().max()) <= tol * scale


def test_runner_fused_fnet_tail_matches_reference_forward(core):
    import flow_supervisor_b200 as fsb
    torch.manual_seed(1234)
    model = core.raft.RAFT(rm.raft_args()).eval().cuda()
    im1, im2 = (t.cuda() for t in rm.synth_pair(440, 1024, seed=13, batch=2))
    with rm.strict_fp32(), torch.no_grad():
        _, up_r = model(im1, im2, iters=12, test_mode=True)
        low_e, up_e = fsb.RaftRunner(model, iters=12, graph=False, fused_fnet_tail=True)(im1, im2)
        low_g, up_g = fsb.RaftRunner(model, iters=12, graph=True, fused_fnet_tail=True)(im1, im2)
    assert torch.equal(low_g, low_e) and torch.equal(up_g, up_e)
    assert rm.epe(up_g, up_r) <= 0.01, rm.epe(up_g, up_r)


@pytest.mark.parametrize("shape,law", [((2, 46, 62), "random"), ((1, 55, 128), "random"), ((1, 55, 128), "lattice"),
                                       ((2, 47, 156), "border"), ((1, 24, 40), "random"), ((3, 17, 21), "random")])
@pytest.mark.parametrize("volume", ["f32", "bf16"])
def test_lookup_convc1_matches_relu_conv1x1_of_the_lookup(shape, law, volume):
    """fc_lookup_convc1_fwd (row f1) against relu(conv1x1(lookup)) in fp32: <= 1e-4 of the output's max magnitude
    (three-pass bf16 split of weights and looked-up values, fp32 accumulate), exact zeros where ReLU clips."""
    import flow_supervisor_b200 as fsb
    B, H, W = shape
    g = torch.Generator().manual_seed(31)
    f1 = (1.57 * torch.randn(B, 256, H, W, generator=g)).cuda()
    f2 = (1.57 * torch.randn(B, 256, H, W, generator=g)).cuda()
    grid = fsb.coords_grid(B, H, W)
    if law == "random":
        c = grid + 4.0 * torch.randn(B, 2, H, W, generator=g)
    elif law == "lattice":
        c = grid + torch.randint(-3, 4, (B, 2, H, W), generator=g).float()
    else:
        c = grid + 40.0 * torch.randn(B, 2, H, W, generator=g)
        c[0, :, 0, 0] = float("nan")
    c = c.cuda()
    conv = torch.nn.Conv2d(324, 256, 1).cuda()
    packed = fsb.ops.convc1_prepare(conv.weight.detach(), conv.bias.detach())
    old = fsb.CorrBlock.volume
    fsb.CorrBlock.volume = volume
    try:
        with rm.strict_fp32(), torch.no_grad():
            blk = fsb.CorrBlock(f1, f2)
            ref = torch.relu(conv(blk(c)))
            out = blk.lookup_convc1(c, packed)
    finally:
        fsb.CorrBlock.volume = old
    assert out.shape == ref.shape and out.is_contiguous()
    assert float((out - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), float((out - ref).abs().max()) / float(ref.abs().max())
    assert float(((out == 0) != (ref == 0)).float().mean()) < 1e-3


def test_runner_fused_convc1_matches_reference_forward(core):
    import flow_supervisor_b200 as fsb
    torch.manual_seed(1234)
    model = core.raft.RAFT(rm.raft_args()).eval().cuda()
    im1, im2 = (t.cuda() for t in rm.synth_pair(440, 1024, seed=14, batch=2))
    with rm.strict_fp32(), torch.no_grad():
        _, up_r = model(im1, im2, iters=12, test_mode=True)
        low_e, up_e = fsb.RaftRunner(model, iters=12, graph=False, fused_convc1=True)(im1, im2)
        low_g, up_g = fsb.RaftRunner(model, iters=12, graph=True, fused_convc1=True, fused_fnet_tail=True)(im1, im2)
    assert rm.epe(up_e, up_r) <= 0.01, rm.epe(up_e, up_r)
    assert rm.epe(up_g, up_r) <= 0.01, rm.epe(up_g, up_r)
