"""CPU: the oracle (oracle/corr_spec.py numpy restatement, oracle/corr_torch.py
library-call port) held to the golden vectors produced by the live reference
(oracle/make_golden.py, /root/reference/pytorch/core/corr.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import corr_spec, corr_torch

FWD = ["fwd_odd_d32", "fwd_small_d128_r3", "fwd_d256_b2"]


def rel_err(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / scale


def laws(g):
    return [k[len("coords_"):] for k in g if k.startswith("coords_")]


@pytest.mark.parametrize("name", FWD)
def test_torch_port_is_bit_exact_to_reference(golden, name):
    g = golden(name)
    blk = corr_torch.TorchCorrBlock(torch.from_numpy(g["fmap1"]), torch.from_numpy(g["fmap2"]),
                                    int(g["num_levels"]), int(g["radius"]))
    for law in laws(g):
        out = blk(torch.from_numpy(g[f"coords_{law}"])).numpy()
        assert np.array_equal(out, g[f"out_{law}"]), law


def test_spec_pyramid_bit_exact(golden):
    g = golden("fwd_odd_d32")
    pyr = corr_spec.pool_pyramid(g["pyr0"], 4)          # pooling restated, same level 0
    for l in range(4):
        assert pyr[l].shape == g[f"pyr{l}"].shape
        assert np.array_equal(pyr[l], g[f"pyr{l}"]), l
    vol = corr_spec.all_pairs(g["fmap1"], g["fmap2"])
    assert rel_err(vol, g["pyr0"]) < 2e-6                # accumulation order differs (MKL)


@pytest.mark.parametrize("name", FWD)
def test_spec_lookup_matches_reference(golden, name):
    g = golden(name)
    L, r = int(g["num_levels"]), int(g["radius"])
    pyr = ([g[f"pyr{l}"] for l in range(L)] if "pyr0" in g
           else corr_spec.build(g["fmap1"], g["fmap2"], L))
    for law in laws(g):
        out = corr_spec.lookup(pyr, g[f"coords_{law}"], r, rounding="cpu")
        ref = g[f"out_{law}"]
        assert out.shape == ref.shape
        assert rel_err(out, ref) < 1e-5, law
        # exact zeros of the reference are out-of-bounds windows (the volume is a.s.
        # non-zero): the oracle's corner masks must reproduce them exactly
        if "pyr0" in g:
            assert np.array_equal(out == 0, ref == 0), law


def test_spec_channel_order_known_answer():
    """Position-encoding volume (value = 1000*y2 + x2) read at lattice coords:
    channel k = l*81 + a*9 + b' with a <-> x offset, b' <-> y offset (corr.py:37-39)."""
    H, W, r = 16, 24, 4
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    vol = np.broadcast_to((1000.0 * yy + xx).astype(np.float32), (1, H * W, H, W)).copy()
    coords = corr_spec.coords_grid(1, H, W)
    out = corr_spec.lookup([vol], coords, r, rounding="cpu")
    y1, x1 = 8, 12
    for a in range(9):
        for b in range(9):
            want = 1000.0 * (y1 + b - r) + (x1 + a - r)
            assert abs(out[0, a * 9 + b, y1, x1] - want) < 1e-2, (a, b)


def test_spec_all_oob_is_exact_zero():
    H, W = 16, 16
    rng = np.random.default_rng(0)
    f = rng.standard_normal((1, 8, H, W)).astype(np.float32)
    pyr = corr_spec.build(f, f, 4)
    coords = corr_spec.coords_grid(1, H, W) + np.float32(1000.0)
    out, dbg = corr_spec.lookup(pyr, coords, 4, debug=True)
    assert not out.any()
    assert all(not d["mask"].any() for d in dbg)


def test_spec_rounding_modes_differ_only_on_lattice(golden):
    """'cuda' (multiply by reciprocal) and 'cpu' (true division) flavours give the
    same taps for generic coordinates and differ on some lattice points."""
    g = golden("fwd_odd_d32")
    cx = g["coords_random"][:, 0].ravel()
    a = corr_spec.axis_taps(cx, 19, 0, 4, "cuda")
    b = corr_spec.axis_taps(cx, 19, 0, 4, "cpu")
    assert np.array_equal(a[0], b[0])
    lat = np.arange(0, 300, dtype=np.float32)
    n_diff = 0
    for size in (62, 96, 128, 156, 55, 47):
        n_diff += int((corr_spec.axis_taps(lat, size, 0, 4, "cuda")[0]
                       != corr_spec.axis_taps(lat, size, 0, 4, "cpu")[0]).sum())
    assert n_diff > 0


def test_spec_backward_matches_reference_autograd(golden):
    g = golden("bwd_odd_d64")
    L, r, T = int(g["num_levels"]), int(g["radius"]), int(g["n_lookups"])
    B, D, H, W = g["fmap1"].shape
    shapes = corr_spec.level_shapes(H, W, L)
    G = None
    for t in range(T):
        G = corr_spec.lookup_backward(g[f"gout{t}"], g[f"coords{t}"], shapes, r, "cpu", G)
    d1, d2 = corr_spec.build_backward(G, g["fmap1"], g["fmap2"])
    assert rel_err(d1, g["dfmap1"]) < 1e-5
    assert rel_err(d2, g["dfmap2"]) < 1e-5


def test_spec_ondemand_agrees_with_corrblock(golden):
    """AlternateCorrBlock semantics (correlation_kernel.cu:59-116) equal CorrBlock up
    to summation order away from floor-flip points."""
    g = golden("fwd_small_d128_r3")
    L, r = int(g["num_levels"]), int(g["radius"])
    out = corr_spec.ondemand_lookup(g["fmap1"], g["fmap2"], g["coords_random"], L, r)
    assert rel_err(out, g["out_random"]) < 1e-4


@pytest.mark.skipif(not os.path.isdir("/root/reference/pytorch/core"),
                    reason="live reference only exists in the build container")
def test_live_reference_still_matches_port():
    import sys
    sys.path.insert(0, "/root/reference/pytorch")
    from core.corr import CorrBlock
    gen = torch.Generator().manual_seed(5)
    f1 = torch.randn(2, 16, 18, 22, generator=gen)
    f2 = torch.randn(2, 16, 18, 22, generator=gen)
    c = corr_torch.coords_grid(2, 18, 22) + 4 * torch.randn(2, 2, 18, 22, generator=gen)
    assert torch.equal(CorrBlock(f1, f2)(c), corr_torch.TorchCorrBlock(f1, f2)(c))
