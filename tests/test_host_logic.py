"""CPU: host-side logic of the drop-in layer -- no-fallback behaviour, reference patching,
batch sharding (single process and world_size-2 gloo)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flow_supervisor_b200 as fsb
from flow_supervisor_b200 import shard

REF = "/root/reference/pytorch"


def test_no_cpu_fallback():
    f = torch.zeros(1, 8, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsb.CorrBlock(f, f)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsb.AlternateCorrBlock(f, f)
    from flow_supervisor_b200 import alt_cuda_corr
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):      # correlation.cpp:19
        alt_cuda_corr.forward(f, f, torch.zeros(1, 1, 16, 16, 2), 4)


def test_custom_ops_registered_with_fake_impls():
    for name in ["build", "lookup", "lookup_bwd", "build_bwd", "ondemand_prepare", "ondemand_lookup",
                 "altcorr_fwd", "altcorr_bwd"]:
        assert hasattr(torch.ops.flowcorr, name)
    # meta tracing needs neither CUDA nor the library
    f = torch.empty(2, 256, 46, 62, device="meta")
    pyr = torch.ops.flowcorr.build(f, f, 4, 0, 0)
    assert pyr.numel() == fsb.ops.pyramid_numel(2, 46, 62, 4)
    out = torch.ops.flowcorr.lookup(pyr, torch.empty(2, 2, 46, 62, device="meta"), 4, 4, 0)
    assert tuple(out.shape) == (2, 324, 46, 62)


def test_coords_grid_axis_order():
    g = fsb.coords_grid(2, 3, 5)
    assert tuple(g.shape) == (2, 2, 3, 5)
    assert torch.equal(g[0, 0, 0], torch.arange(5.0)) and torch.equal(g[1, 1, :, 0], torch.arange(3.0))


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference only exists in the build container")
def test_patch_reference_rebinds_every_import_site():
    sys.path.insert(0, REF)
    import core.corr, core.raft, core.l2l, core.gma_corr, core.gma_network, core.gma_l2l  # noqa: E401
    patched = fsb.patch_reference()
    try:
        _check_patched(core, patched)
    finally:
        fsb.unpatch_reference()
    assert core.raft.CorrBlock is core.corr.CorrBlock and core.corr.CorrBlock is not fsb.CorrBlock


def _check_patched(core, patched):
    for mod in (core.corr, core.raft, core.l2l, core.gma_corr, core.gma_network, core.gma_l2l):
        assert mod.CorrBlock is fsb.CorrBlock, mod.__name__
    for mod in (core.corr, core.raft, core.l2l):
        assert mod.AlternateCorrBlock is fsb.AlternateCorrBlock, mod.__name__
    assert sys.modules["alt_cuda_corr"].forward is not None
    assert ("core.raft", "CorrBlock") in patched
    # the patched model now refuses CPU inputs instead of silently using another path
    import argparse
    model = core.raft.RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 128, 128), torch.zeros(1, 3, 128, 128), iters=1, test_mode=True)


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 8, 9, 64):
        for w in (1, 2, 3, 8):
            parts = shard.partition(n, w)
            assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(0)
        a = torch.randn(n_items, 4, 6, 8, generator=gen)
        b = torch.randn(n_items, 4, 6, 8, generator=gen)
        calls = []

        def fn(x, y):                       # stands in for a per-rank RAFT forward
            calls.append(x.shape[0])
            return (x * y).sum(dim=1, keepdim=True) + x.shape[0] * 0.0
        full = shard.run_sharded(fn, a, b)
        want = (a * b).sum(dim=1, keepdim=True)
        ok = torch.equal(full, want) and calls == [shard.partition(n_items, world)[rank][1]
                                                   - shard.partition(n_items, world)[rank][0]]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5, 1])
def test_run_sharded_world2_gloo(n_items):
    world, port = 2, 29500 + os.getpid() % 2000 + n_items
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_multiply_shift_division_formula_is_exact():
    """fc_lookup.cuh::split_query replaces gq / N by (umulhi(m, gq) + gq) >> s with the Granlund-Montgomery round-up
    multiplier of fill_params.  The formula must be exact for every query index the library accepts (gq < 2^31 - 32)
    and every token count; checked here in integer arithmetic on edge values and a random sample."""
    import random
    rng = random.Random(0)
    Ns = [1, 2, 3, 7, 96, 255, 256, 257, 2852, 4416, 6912, 7040, 7332, 32640, 65535, 65536, 1 << 20, (1 << 24) + 1]
    for N in Ns:
        s = 0
        while (1 << s) < N:
            s += 1
        m = ((1 << 32) * ((1 << s) - N)) // N + 1
        assert m < (1 << 32)
        edge = [0, 1, N - 1, N, N + 1, 2 * N - 1, 2 * N, (1 << 31) - 33, (1 << 31) - 1 - 32]
        edge += [k * N + d for k in (3, 1000, ((1 << 31) - 64) // N) for d in (-1, 0, 1) if 0 <= k * N + d < (1 << 31) - 32]
        sample = edge + [rng.randrange(0, (1 << 31) - 32) for _ in range(2000)]
        for gq in sample:
            t = (m * gq) >> 32
            assert t + gq < (1 << 32)                       # the 32-bit add in the kernel cannot wrap
            assert (t + gq) >> s == gq // N, (N, gq)


def test_build_schedule_covers_every_unit_once():
    """fc_build_tc.cu: a CTA pair takes whole pair-tiles c, c + n, c + 2n, ... and the left-over pair-tiles are dealt
    out unit by unit (unit_of / u_end in the kernel).  Mirror of that arithmetic: every unit exactly once, for any
    number of pair-tiles, groups and resident CTA pairs; the busiest pair gets at most one unit more than the mean
    rounded up plus the groups of an unfinished round."""
    for n_pt in (1, 2, 3, 27, 28, 73, 74, 75, 147, 148, 224, 513):
        for groups in (1, 2, 7, 17):
            for n_clusters in (1, 2, 37, 74):
                units = n_pt * groups
                if n_clusters > units:
                    continue
                full = n_pt // n_clusters
                tail_units = (n_pt - full * n_clusters) * groups
                seen = []
                most = 0
                for c in range(n_clusters):
                    n_local = full * groups + ((tail_units - c + n_clusters - 1) // n_clusters if tail_units > c else 0)
                    most = max(most, n_local)
                    for j in range(n_local):
                        k = j // groups
                        if k < full:
                            u = (c + k * n_clusters) * groups + (j - k * groups)
                        else:
                            u = full * n_clusters * groups + c + (j - full * groups) * n_clusters
                        seen.append(u)
                assert sorted(seen) == list(range(units)), (n_pt, groups, n_clusters)
                assert most <= -(-units // n_clusters) + (groups if tail_units else 0)


def test_clear_pads_zeroes_exactly_the_pad_cells():
    """ops.clear_pads_ restores the pyramid invariant of include/flowcorr.h (pad rows / columns of every level are zeros)
    on a caller-filled buffer: after it, the un-patched levels hold zeros outside H_l x W_l and the data inside."""
    import torch
    from flow_supervisor_b200 import ops
    B, H, W, L = 2, 5, 11, 3
    src = torch.arange(1, ops.pyramid_numel(B, H, W, L) + 1, dtype=torch.float32)
    p = ops.clear_pads_(src.clone(), B, H, W, L)
    for lvl, ref, (h, w, wp) in zip(ops.level_padded(p, B, H, W, L), ops.level_padded(src, B, H, W, L), ops.geometry(H, W, L)):
        assert torch.equal(lvl[:, :h, :w], ref[:, :h, :w])
        assert not lvl[:, h:, :].any() and not lvl[:, :, w:].any()
        assert lvl.shape[1] % 2 == 0 and lvl.shape[2] == wp
