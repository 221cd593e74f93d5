"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through
the C ABI (libflowcorr.so via flow_supervisor_b200.ops) and is compared with the oracle
(oracle/corr_spec.py, oracle/corr_torch.py) and the committed golden vectors produced by
the live reference.  Tolerances (BASELINE.json north_star): integer tap indices and
out-of-bounds masks bit-exact; fp32 values <= 1e-4 relative (of the tensor's max
magnitude); gradients <= 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import corr_spec, corr_torch

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-4


@pytest.fixture(scope="module")
def fsb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import flow_supervisor_b200 as m
    return m


def rel_err(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    return float(np.abs(a.astype(np.float64) - b).max()) / max(float(np.abs(b).max()), 1e-30)


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def laws(g):
    return [k[len("coords_"):] for k in g if k.startswith("coords_")]


class mode:
    """Temporarily set CorrBlock class attributes."""

    def __init__(self, fsb, **kw):
        self.cls, self.kw = fsb.CorrBlock, kw

    def __enter__(self):
        self.old = {k: getattr(self.cls, k) for k in self.kw}
        for k, v in self.kw.items():
            setattr(self.cls, k, v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            setattr(self.cls, k, v)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["fwd_odd_d32", "fwd_small_d128_r3", "fwd_d256_b2"])
def test_forward_matches_reference_golden(fsb, golden, name):
    g = golden(name)
    L, r = int(g["num_levels"]), int(g["radius"])
    with mode(fsb, math="fp32", coord_mode="cpu"):
        blk = fsb.CorrBlock(cuda(g["fmap1"]), cuda(g["fmap2"]), num_levels=L, radius=r)
        if "pyr0" in g:
            for l, lvl in enumerate(blk.corr_pyramid):
                ref = g[f"pyr{l}"]
                assert tuple(lvl.shape) == (ref.shape[0] * ref.shape[1], 1) + ref.shape[2:]
                assert rel_err(lvl.reshape(ref.shape), ref) < 1e-5, l
        pyr = corr_spec.build(g["fmap1"], g["fmap2"], L)
        for law in laws(g):
            c = g[f"coords_{law}"]
            out, x0, y0, mask = blk.lookup_debug(cuda(c))
            ref = g[f"out_{law}"]
            assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_contiguous()
            assert rel_err(out, ref) < VAL_TOL, law
            assert np.array_equal(out.cpu().numpy() == 0, ref == 0), law       # OOB pattern
            _, dbg = corr_spec.lookup(pyr, c, r, rounding="cpu", debug=True)
            for l in range(L):
                assert np.array_equal(x0[:, l].cpu().numpy(), dbg[l]["x0"]), (law, l)
                assert np.array_equal(y0[:, l].cpu().numpy(), dbg[l]["y0"]), (law, l)
                assert np.array_equal(mask[:, l].cpu().numpy(), dbg[l]["mask"]), (law, l)
            assert torch.equal(out, blk(cuda(c)))                             # debug path == fast path


@pytest.mark.parametrize("name", ["fwd_small_d128_r3", "fwd_d256_b2"])
def test_forward_matches_reference_golden_in_the_benchmarked_mode(fsb, golden, name):
    """The same golden vectors through the mode bench.py and 'auto' run (tcgen05 build, 3xbf16): values within the
    1e-4 contract, integer part and out-of-bounds pattern bit-exact."""
    g = golden(name)
    L, r = int(g["num_levels"]), int(g["radius"])
    with mode(fsb, math="3xbf16", coord_mode="cpu"):
        blk = fsb.CorrBlock(cuda(g["fmap1"]), cuda(g["fmap2"]), num_levels=L, radius=r)
        pyr = corr_spec.build(g["fmap1"], g["fmap2"], L)
        for law in laws(g):
            c = g[f"coords_{law}"]
            out, x0, y0, mask = blk.lookup_debug(cuda(c))
            ref = g[f"out_{law}"]
            assert rel_err(out, ref) < VAL_TOL, law
            assert np.array_equal(out.cpu().numpy() == 0, ref == 0), law
            _, dbg = corr_spec.lookup(pyr, c, r, rounding="cpu", debug=True)
            for l in range(L):
                assert np.array_equal(x0[:, l].cpu().numpy(), dbg[l]["x0"]), (law, l)
                assert np.array_equal(y0[:, l].cpu().numpy(), dbg[l]["y0"]), (law, l)
                assert np.array_equal(mask[:, l].cpu().numpy(), dbg[l]["mask"]), (law, l)


def test_backward_matches_reference_autograd_golden_in_the_benchmarked_mode(fsb, golden):
    g = golden("bwd_odd_d64")
    L, r, T = int(g["num_levels"]), int(g["radius"]), int(g["n_lookups"])
    f1 = cuda(g["fmap1"]).requires_grad_()
    f2 = cuda(g["fmap2"]).requires_grad_()
    with mode(fsb, math="3xbf16", coord_mode="cpu"):
        blk = fsb.CorrBlock(f1, f2, num_levels=L, radius=r)
        loss = 0.0
        for t in range(T):
            loss = loss + (blk(cuda(g[f"coords{t}"])) * cuda(g[f"gout{t}"])).sum()
        loss.backward()
    assert rel_err(f1.grad, g["dfmap1"]) < VAL_TOL
    assert rel_err(f2.grad, g["dfmap2"]) < VAL_TOL


def test_narrow_maps_take_the_fp32_mode_or_fail_cleanly(fsb):
    """W <= 8 tokens: one tile of two target rows would be a 16-column UMMA with M = 256 (invalid).  'auto' runs the
    CUDA-core mode; asking for the tensor-core mode explicitly is a clean error, not a trap."""
    gen = torch.Generator().manual_seed(1)
    f1 = torch.randn(1, 64, 16, 8, generator=gen).cuda()
    f2 = torch.randn(1, 64, 16, 8, generator=gen).cuda()
    with mode(fsb, math="auto"):
        blk = fsb.CorrBlock(f1, f2, num_levels=2, radius=2)
    ref = corr_torch.TorchCorrBlock(f1, f2, 2, 2)
    c = (corr_torch.coords_grid(1, 16, 8) + 0.7).cuda()
    assert rel_err(blk(c), ref(c)) < VAL_TOL
    with mode(fsb, math="3xbf16"):
        with pytest.raises(RuntimeError, match="tokens"):
            fsb.CorrBlock(f1, f2, num_levels=2, radius=2)
    assert torch.isfinite(blk(c)).all()                                    # the device is still usable


def test_alternate_block_accepts_unit_dimension_levels(fsb):
    """8 x 8 tokens with 4 levels: level 3 is 1 x 1.  The reference's AlternateCorrBlock works there (its indexing never
    divides by size - 1); so do both routes here (CorrBlock itself raises: the reference returns NaN)."""
    gen = torch.Generator().manual_seed(2)
    f1 = torch.randn(1, 64, 8, 8, generator=gen).cuda()
    f2 = torch.randn(1, 64, 8, 8, generator=gen).cuda()
    c = (corr_torch.coords_grid(1, 8, 8) + 0.4 * torch.randn(1, 2, 8, 8, generator=gen)).cuda()
    outs = {}
    for route in ("materialise", "ondemand"):
        fsb.AlternateCorrBlock.route = route
        try:
            outs[route] = fsb.AlternateCorrBlock(f1, f2, num_levels=4, radius=4)(c)
        finally:
            fsb.AlternateCorrBlock.route = "auto"
    assert torch.isfinite(outs["materialise"]).all()
    assert rel_err(outs["materialise"], outs["ondemand"]) < VAL_TOL
    with pytest.raises(RuntimeError, match="unit dimension"):
        fsb.CorrBlock(f1, f2, num_levels=4, radius=4)(c)


def test_pooling_is_bit_exact(fsb, golden):
    g = golden("fwd_odd_d32")
    with mode(fsb, math="fp32"):
        blk = fsb.CorrBlock(cuda(g["fmap1"]), cuda(g["fmap2"]))
    lv = [v.cpu().numpy()[:, 0] for v in blk.corr_pyramid]
    ref = corr_spec.pool_pyramid(lv[0], 4)
    for l in range(4):
        assert np.array_equal(lv[l], ref[l]), l


def test_backward_matches_reference_autograd_golden(fsb, golden):
    g = golden("bwd_odd_d64")
    L, r, T = int(g["num_levels"]), int(g["radius"]), int(g["n_lookups"])
    f1 = cuda(g["fmap1"]).requires_grad_()
    f2 = cuda(g["fmap2"]).requires_grad_()
    with mode(fsb, math="fp32", coord_mode="cpu"):
        blk = fsb.CorrBlock(f1, f2, num_levels=L, radius=r)
        loss = 0.0
        for t in range(T):
            loss = loss + (blk(cuda(g[f"coords{t}"])) * cuda(g[f"gout{t}"])).sum()
        loss.backward()
    assert rel_err(f1.grad, g["dfmap1"]) < VAL_TOL
    assert rel_err(f2.grad, g["dfmap2"]) < VAL_TOL


# ------------------------------------------------------------------ vs the library-call port on the same GPU
def _case(B, D, H, W, seed, flow_std):
    gen = torch.Generator().manual_seed(seed)
    f1 = 1.57 * torch.randn(B, D, H, W, generator=gen)
    f2 = 1.57 * torch.randn(B, D, H, W, generator=gen)
    c = corr_torch.coords_grid(B, H, W) + flow_std * torch.randn(B, 2, H, W, generator=gen)
    return f1.cuda(), f2.cuda(), c.cuda()


@pytest.mark.parametrize("cudnn_on", [True, False])
@pytest.mark.parametrize("shape,flow_std", [((1, 256, 46, 62), 5.0), ((2, 256, 47, 156), 0.0),
                                            ((2, 128, 55, 128), 40.0), ((3, 64, 17, 23), 2.0)])
def test_values_match_torch_ops_on_gpu(fsb, shape, flow_std, cudnn_on):
    """coord_mode='cuda' against matmul/avg_pool2d/grid_sample executed by PyTorch on
    this GPU (what the reference does on a CUDA device), cuDNN sampler on and off."""
    f1, f2, c = _case(*shape, seed=3, flow_std=flow_std)
    old = torch.backends.cudnn.enabled
    torch.backends.cudnn.enabled = cudnn_on
    try:
        ref = corr_torch.TorchCorrBlock(f1, f2)(c)
    finally:
        torch.backends.cudnn.enabled = old
    with mode(fsb, math="fp32", coord_mode="cuda"):
        blk = fsb.CorrBlock(f1, f2)
        out, x0, y0, mask = blk.lookup_debug(c)
    assert rel_err(out, ref) < VAL_TOL
    # integer part against the closed-form spec with CUDA rounding
    B, D, H, W = shape
    shapes = corr_spec.level_shapes(H, W, 4)
    cx = c[:, 0].reshape(-1).cpu().numpy()
    cy = c[:, 1].reshape(-1).cpu().numpy()
    for l, (Hl, Wl) in enumerate(shapes):
        assert np.array_equal(x0[:, l].cpu().numpy(), corr_spec.axis_taps(cx, Wl, l, 4, "cuda")[0]), l
        assert np.array_equal(y0[:, l].cpu().numpy(), corr_spec.axis_taps(cy, Hl, l, 4, "cuda")[0]), l


def test_training_gradients_match_torch_autograd_on_gpu(fsb):
    f1, f2, c = _case(2, 256, 46, 62, seed=5, flow_std=3.0)
    gen = torch.Generator().manual_seed(9)
    gs = [torch.randn(2, 324, 46, 62, generator=gen).cuda() for _ in range(3)]
    cs = [c, c + 1.5, corr_torch.coords_grid(2, 46, 62, "cuda")]

    def run(make):
        a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        blk = make(a, b)
        loss = sum((blk(ci) * gi).sum() for ci, gi in zip(cs, gs))
        loss.backward()
        return a.grad, b.grad

    r1, r2 = run(corr_torch.TorchCorrBlock)
    with mode(fsb, math="fp32", coord_mode="cuda"):
        d1, d2 = run(fsb.CorrBlock)
    assert rel_err(d1, r1) < VAL_TOL
    assert rel_err(d2, r2) < VAL_TOL


def test_two_blocks_two_backwards(fsb):
    """L2L builds two CorrBlocks in one graph and train.py runs two backward passes per
    optimiser step (train.py:273,277): accumulators must not leak between them."""
    f1, f2, c = _case(1, 64, 24, 32, seed=7, flow_std=2.0)
    a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    with mode(fsb, math="fp32"):
        for _ in range(2):
            blk1, blk2 = fsb.CorrBlock(a, b), fsb.CorrBlock(b, a)
            loss = blk1(c).sum() + 2.0 * blk2(c + 0.5).square().sum() + blk1(c - 1.0).mean()
            loss.backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a2, b2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    for _ in range(2):
        blk1, blk2 = corr_torch.TorchCorrBlock(a2, b2), corr_torch.TorchCorrBlock(b2, a2)
        loss = blk1(c).sum() + 2.0 * blk2(c + 0.5).square().sum() + blk1(c - 1.0).mean()
        loss.backward()
    assert rel_err(ga, a2.grad) < VAL_TOL
    assert rel_err(gb, b2.grad) < VAL_TOL


# ------------------------------------------------------------------ properties at larger sizes / edge cases
def test_lattice_lookup_reads_volume_entries(fsb):
    """Iteration-0 coords (integer lattice): tap (a, b') of level 0 is the volume entry at
    (y1+b'-r, x1+a-r) or 0 outside (channel order k = a*9 + b')."""
    B, D, H, W, r = 2, 256, 55, 128, 4
    f1, f2, _ = _case(B, D, H, W, seed=11, flow_std=0.0)
    with mode(fsb, math="fp32"):
        blk = fsb.CorrBlock(f1, f2)
        out = blk(fsb.coords_grid(B, H, W, device="cuda"))
    vol = blk.corr_pyramid[0].reshape(B, H, W, H, W)
    pad = torch.nn.functional.pad(vol, (r, r, r, r))
    ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    for a, b in [(0, 0), (4, 4), (8, 8), (2, 7), (7, 1)]:
        want = pad[:, ys, xs, ys + b, xs + a]
        got = out[:, a * 9 + b]
        assert rel_err(got, want) < 1e-5, (a, b)


def test_far_and_nan_coords_give_zeros(fsb):
    f1, f2, c = _case(1, 32, 16, 24, seed=13, flow_std=1.0)
    with mode(fsb, math="fp32"):
        blk = fsb.CorrBlock(f1, f2)
        c2 = c.clone()
        c2[:, :, :8] = 1.0e9
        c2[:, :, 8:12] = float("nan")
        out, _, _, mask = blk.lookup_debug(c2)
        assert not out[:, :, :12].any() and not mask.view(16, 24, 4, 9, 9)[:12].any()
        assert torch.equal(out[:, :, 12:], blk(c)[:, :, 12:])


def test_linearity_and_scaling(fsb):
    f1, f2, c = _case(1, 128, 46, 96, seed=17, flow_std=4.0)
    with mode(fsb, math="fp32"):
        o1 = fsb.CorrBlock(f1, f2)(c)
        o2 = fsb.CorrBlock(2.0 * f1, f2)(c)          # power-of-two scale is exact in fp32
    assert torch.equal(o2, 2.0 * o1)


@pytest.mark.parametrize("L,r", [(1, 4), (2, 2), (3, 1), (4, 3)])
def test_levels_and_radii(fsb, L, r):
    f1, f2, c = _case(2, 32, 19, 21, seed=19, flow_std=3.0)
    ref = corr_torch.TorchCorrBlock(f1, f2, num_levels=L, radius=r)(c)
    with mode(fsb, math="fp32"):
        out = fsb.CorrBlock(f1, f2, num_levels=L, radius=r)(c)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < VAL_TOL


def test_errors_are_loud(fsb):
    f = torch.zeros(1, 8, 16, 16)
    with pytest.raises(RuntimeError):
        fsb.CorrBlock(f, f)
    with pytest.raises(RuntimeError):            # level 3 would be 1 pixel high
        fsb.CorrBlock(f[:, :, :8].cuda(), f[:, :, :8].cuda())(torch.zeros(1, 2, 8, 16).cuda())
    with pytest.raises(RuntimeError):
        fsb.CorrBlock(f.cuda(), f.cuda(), radius=7)(torch.zeros(1, 2, 16, 16).cuda())
    with pytest.raises(ValueError):
        fsb.CorrBlock(f.cuda(), f.cuda())(torch.zeros(1, 2, 8, 8).cuda())


# ------------------------------------------------------------------ on-demand variant
class altroute:
    """Temporarily pin AlternateCorrBlock.route ('ondemand' = fused dot+sample kernel,
    'materialise' = tensor-core build + FC_COORD_RAW lookup)."""

    def __init__(self, fsb, route):
        self.cls, self.route = fsb.AlternateCorrBlock, route

    def __enter__(self):
        self.old = self.cls.route
        self.cls.route = self.route

    def __exit__(self, *a):
        self.cls.route = self.old


@pytest.mark.parametrize("route", ["ondemand", "materialise"])
@pytest.mark.parametrize("shape,r", [((1, 128, 16, 20), 3), ((2, 256, 24, 32), 4)])
def test_ondemand_matches_spec(fsb, shape, r, route):
    f1, f2, c = _case(*shape, seed=23, flow_std=3.0)
    with altroute(fsb, route):
        blk = fsb.AlternateCorrBlock(f1, f2, num_levels=4, radius=r)
        assert blk.materialised == (route == "materialise")
        out = blk(c)
    ref = corr_spec.ondemand_lookup(f1.cpu().numpy(), f2.cpu().numpy(), c.cpu().numpy(), 4, r)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < VAL_TOL


@pytest.mark.parametrize("route", ["ondemand", "materialise"])
def test_ondemand_matches_corrblock_at_cfg1_size(fsb, route):
    f1, f2, c = _case(1, 256, 46, 62, seed=29, flow_std=6.0)
    with mode(fsb, math="fp32"):
        a = fsb.CorrBlock(f1, f2)(c)
    with altroute(fsb, route):
        b = fsb.AlternateCorrBlock(f1, f2)(c)
    assert rel_err(b, a) < VAL_TOL


def test_materialised_route_keeps_ondemand_indices(fsb):
    """FC_COORD_RAW: tap = floor(c / 2^l) + offset with one fraction per axis
    (correlation_kernel.cu:67-76), bit-exact -- including lattice coordinates, where
    CorrBlock's normalise round trip flips floors."""
    from flow_supervisor_b200 import ops, _lib
    B, D, H, W, r = 1, 64, 24, 40, 4
    f1, f2, c = _case(B, D, H, W, seed=37, flow_std=4.0)
    c[:, :, ::2] = torch.round(c[:, :, ::2])                       # half the rows on the lattice
    pyr = ops.build(f1, f2, 4, _lib.MATH_FP32, _lib.VOL_F32)
    out, x0, y0, mask = ops.lookup_debug(pyr, c, 4, r, _lib.COORD_RAW)
    cn = c.cpu().numpy()
    cx, cy = cn[:, 0].reshape(-1), cn[:, 1].reshape(-1)
    offs = np.arange(-r, r + 1, dtype=np.int32)
    for l in range(4):
        ex = np.floor(cx / np.float32(2 ** l)).astype(np.int32)[:, None] + offs
        ey = np.floor(cy / np.float32(2 ** l)).astype(np.int32)[:, None] + offs
        assert np.array_equal(x0[:, l].cpu().numpy(), ex)
        assert np.array_equal(y0[:, l].cpu().numpy(), ey)
    ref = corr_spec.ondemand_lookup(f1.cpu().numpy(), f2.cpu().numpy(), cn, 4, r)
    assert rel_err(out, ref) < VAL_TOL
    with altroute(fsb, "ondemand"):
        od = fsb.AlternateCorrBlock(f1, f2)(c)
    assert np.array_equal((out == 0).cpu().numpy(), (od == 0).cpu().numpy()) or rel_err(out, od) < 1e-5


def test_alternate_routes_agree_at_cfg5_size(fsb):
    """BASELINE.json config 5 geometry (1088x1920 -> 136x240 tokens; batch 1 here): the fused
    on-demand kernel and the materialised route are independent kernels over the same
    inputs; they must agree to fp32 rounding at full size."""
    f1, f2, c = _case(1, 256, 136, 240, seed=41, flow_std=8.0)
    with altroute(fsb, "ondemand"):
        a = fsb.AlternateCorrBlock(f1, f2)(c)
    with altroute(fsb, "auto"):
        blk = fsb.AlternateCorrBlock(f1, f2)
        assert blk.materialised, "5.7 GB must fit a B200 under the default policy"
        b = blk(c)
    assert rel_err(b, a) < VAL_TOL
    del blk
    torch.cuda.empty_cache()


def test_alt_cuda_corr_shim(fsb):
    from flow_supervisor_b200 import alt_cuda_corr
    f1, f2, c = _case(2, 64, 16, 24, seed=31, flow_std=2.0)
    n1 = f1.permute(0, 2, 3, 1).contiguous()
    n2 = f2.permute(0, 2, 3, 1).contiguous()
    cc = c.permute(0, 2, 3, 1).reshape(2, 1, 16, 24, 2).contiguous()
    corr, = alt_cuda_corr.forward(n1, n2, cc, 4)
    assert corr.shape == (2, 1, 81, 16, 24)
    ref = corr_spec.ondemand_lookup(f1.cpu().numpy(), f2.cpu().numpy(), c.cpu().numpy(), 1, 4, scale=False)
    assert rel_err(corr.reshape(2, 81, 16, 24), ref) < VAL_TOL
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(n1.cpu(), n2, cc, 4)
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(f1.permute(0, 2, 3, 1), n2, cc, 4)
    # backward against the oracle's adjoint of the same kernel semantics
    g = torch.randn(2, 1, 81, 16, 24, generator=torch.Generator().manual_seed(1)).cuda()
    d1, d2, dc = alt_cuda_corr.backward(n1, n2, cc, g, 4)
    r1, r2 = corr_spec.ondemand_backward(f1.cpu().numpy(), f2.cpu().numpy(), c.cpu().numpy(),
                                         g.reshape(2, 81, 16, 24).cpu().numpy(), 4)
    assert rel_err(d1.permute(0, 3, 1, 2), r1) < VAL_TOL
    assert rel_err(d2.permute(0, 3, 1, 2), r2) < VAL_TOL
    assert not dc.any()


# ------------------------------------------------------------------ CUDA graphs
def test_build_and_lookups_capture_into_a_cuda_graph(fsb):
    """include/flowcorr.h: every call is asynchronous, allocates nothing and never synchronises.
    One RAFT-style sequence (CorrBlock + 3 lookups, raft.py:105-124) is captured into a CUDA
    graph, replayed on NEW inputs written into the captured buffers, and must reproduce the
    eager result bit for bit (same kernels, same launch geometry)."""
    B, D, H, W = 2, 256, 24, 40
    f1, f2, c = _case(B, D, H, W, seed=43, flow_std=3.0)
    sf1, sf2 = torch.zeros_like(f1), torch.zeros_like(f2)
    sc = [torch.zeros_like(c) for _ in range(3)]

    def run():
        blk = fsb.CorrBlock(sf1, sf2)
        return [blk(x) for x in sc]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()                                               # warm-up: lazy module load, tensor maps
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = run()
    sf1.copy_(f1); sf2.copy_(f2)
    for t, x in enumerate(sc):
        x.copy_(c + 0.37 * t)
    graph.replay()
    torch.cuda.synchronize()
    blk = fsb.CorrBlock(f1, f2)
    for t in range(3):
        want = blk(c + 0.37 * t)
        assert torch.equal(outs[t], want)
    assert float(outs[0].abs().max()) > 0
