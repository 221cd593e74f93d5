"""CPU oracle for RAFT's correlation hot path -- TEST INFRASTRUCTURE ONLY.

This is a numpy restatement (closed-form, index-explicit) of what the reference
computes with torch library calls.  Nothing under ``flow_supervisor_b200/`` may
import it; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker.

Parity status: the reference (iwbn/flow-supervisor) ships NO tests, golden
vectors or fixtures for this path, so the pin is the reference
itself executed in the build container: ``oracle/make_golden.py`` imports
``/root/reference/pytorch/core/corr.py`` read-only, runs it on seeded inputs and
commits inputs+outputs under ``tests/golden/``; ``tests/test_oracle.py`` holds
this file to those vectors.

What each function follows (paths relative to /root/reference):

* ``all_pairs``        pytorch/core/corr.py:52-60   (CorrBlock.corr)
* ``pool_pyramid``     pytorch/core/corr.py:21-27   (reshape + 3x avg_pool2d(2,2))
* ``axis_taps``        pytorch/core/corr.py:35-43 + pytorch/core/utils/utils.py:57-62
                       + ATen grid_sampler (unnormalise, floor, weights;
                       torch/include/ATen/native/cuda/GridSampler.cuh:23-27)
* ``lookup``           pytorch/core/corr.py:29-50   (CorrBlock.__call__)
* ``lookup_backward``  autograd of the above (no explicit reference code;
                       driven by pytorch/train.py:273,277)
* ``build_backward``   autograd of corr.py:21-27,52-60
* ``ondemand_lookup``  pytorch/core/corr.py:63-91 + pytorch/alt_cuda_corr/
                       correlation_kernel.cu:59-116 (AlternateCorrBlock)

Coordinate rounding.  The reference normalises pixel coordinates to [-1, 1] and
grid_sample un-normalises them again, all in fp32, so the integer tap index is
``floor`` of a round-tripped value, not of the coordinate.  The division
``2x / (W-1)`` (utils.py:61-62) is a true division on CPU tensors but
``2x * fl(1/(W-1))`` on CUDA tensors (ATen's div-by-CPU-scalar fast path), so
the two devices disagree on a few lattice points.  ``rounding='cuda'`` is the
canonical target of the CUDA kernels; ``rounding='cpu'`` reproduces the
reference run on CPU (what the golden vectors were generated with).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- volume
def all_pairs(fmap1: np.ndarray, fmap2: np.ndarray, exact: bool = False) -> np.ndarray:
    """C[b, p, q] = sum_d f1[b,d,p] f2[b,d,q] / sqrt(D)   -> (B, N, H, W) fp32.

    corr.py:52-60.  ``exact=True`` accumulates in fp64 (used to judge which of two
    fp32 results is closer to the truth); otherwise fp32 like the reference.
    """
    B, D, H, W = fmap1.shape
    a = fmap1.reshape(B, D, H * W)
    b = fmap2.reshape(B, D, H * W)
    if exact:
        c = np.matmul(a.astype(np.float64).transpose(0, 2, 1), b.astype(np.float64))
        c = c / np.sqrt(np.float64(D))
        return c.astype(F32).reshape(B, H * W, H, W)
    c = np.matmul(a.astype(F32).transpose(0, 2, 1), b.astype(F32))
    c = c / np.sqrt(F32(D), dtype=F32)
    return c.astype(F32).reshape(B, H * W, H, W)


def pool_pyramid(vol0: np.ndarray, num_levels: int = 4) -> list[np.ndarray]:
    """corr.py:24-27: level l+1 = 2x2 mean of level l, odd trailing row/col dropped.

    The summation order (a+b+c+d)*0.25 reproduces ATen's CPU avg_pool2d bit for bit
    (SURVEY.md A.2).
    """
    pyr = [np.ascontiguousarray(vol0, dtype=F32)]
    for _ in range(num_levels - 1):
        v = pyr[-1]
        h, w = v.shape[-2] // 2, v.shape[-1] // 2
        v = v[..., : 2 * h, : 2 * w]
        s = ((v[..., 0::2, 0::2] + v[..., 0::2, 1::2]) + v[..., 1::2, 0::2]) + v[..., 1::2, 1::2]
        pyr.append((s * F32(0.25)).astype(F32))
    return pyr


def build(fmap1, fmap2, num_levels=4, exact=False):
    """CorrBlock.__init__ (corr.py:13-27): list of (B, N, Hl, Wl) fp32 levels."""
    return pool_pyramid(all_pairs(fmap1, fmap2, exact=exact), num_levels)


# --------------------------------------------------------------------------- taps
def axis_taps(c: np.ndarray, size: int, level: int, radius: int, rounding: str = "cuda"):
    """Integer tap index and the two weights for the 2r+1 window samples on ONE axis.

    c: fp32 centre coordinate (level-0 pixels), any shape.  ``size`` is Wl (x axis)
    or Hl (y axis) of the level being sampled.  Returns (i0 int32, w0, w1 fp32) of
    shape c.shape + (2r+1,):  sample = w0 * v[i0] + w1 * v[i0 + 1].

    Every intermediate is rounded to fp32 exactly where the reference rounds:
      corr.py:41-43      X  = fl(c / 2^l) + (a - r)
      utils.py:61-62     g  = fl(fl(2 X  (/ or *1/)  (size-1)) - 1)
      GridSampler.cuh    ix = fl(fl(fl(g + 1) / 2) * (size-1)); i0 = floor(ix)
                         w1 = ix - i0 ; w0 = (i0 + 1) - ix
    """
    c = np.asarray(c, dtype=F32)
    offs = np.arange(-radius, radius + 1, dtype=F32)
    with np.errstate(all="ignore"):
        cl = (c / F32(2 ** level)).astype(F32)
        X = (cl[..., None] + offs).astype(F32)
        t = (F32(2) * X).astype(F32)
        den = F32(size - 1)
        if rounding == "cuda":
            g = (t * (F32(1) / den)).astype(F32)
        elif rounding == "cpu":
            g = (t / den).astype(F32)
        elif rounding == "none":          # on-demand variant: no round trip
            g = None
        else:
            raise ValueError(rounding)
        if g is None:
            ix = X
        else:
            g = (g - F32(1)).astype(F32)
            ix = (((g + F32(1)).astype(F32) / F32(2)).astype(F32) * den).astype(F32)
        i0f = np.floor(ix)
        w1 = (ix - i0f).astype(F32)
        w0 = ((i0f + F32(1)).astype(F32) - ix).astype(F32)
        i0 = np.where(np.isfinite(i0f), np.clip(i0f, -2.0e9, 2.0e9), -2.0e9).astype(np.int32)
    return i0, w0, w1


def level_shapes(H: int, W: int, num_levels: int):
    out = []
    for _ in range(num_levels):
        out.append((H, W))
        H, W = H // 2, W // 2
    return out


# --------------------------------------------------------------------------- lookup
def lookup(pyramid, coords, radius=4, rounding="cuda", debug=False):
    """CorrBlock.__call__ (corr.py:29-50).

    pyramid: list of (B, N, Hl, Wl) fp32;  coords: (B, 2, H, W) fp32, ch0 = x, ch1 = y.
    Returns (B, L*(2r+1)^2, H, W) fp32 with channel k = l*(2r+1)^2 + a*(2r+1) + b',
    a = x-offset index, b' = y-offset index (corr.py:37-39: delta[i, j] = (dy[i],
    dx[j]) is added to (x, y), so the FIRST window axis moves x).

    debug=True additionally returns per-level dicts with the integer tap indices
    ``x0``/``y0`` (B*N, 2r+1) int32 and the four corner predicates packed as
    bit0 = (y0,x0), bit1 = (y0,x0+1), bit2 = (y0+1,x0), bit3 = (y0+1,x0+1) in an
    uint8 array of shape (B*N, 2r+1 [a], 2r+1 [b']).
    """
    B, two, H, W = coords.shape
    assert two == 2
    N = H * W
    R = 2 * radius + 1
    cx = coords[:, 0].reshape(B * N).astype(F32)
    cy = coords[:, 1].reshape(B * N).astype(F32)
    outs, dbg = [], []
    for l, vol in enumerate(pyramid):
        Hl, Wl = vol.shape[-2:]
        v = vol.reshape(B * N, Hl, Wl)
        x0, wx0, wx1 = axis_taps(cx, Wl, l, radius, rounding)      # (Q, R) over a
        y0, wy0, wy1 = axis_taps(cy, Hl, l, radius, rounding)      # (Q, R) over b'
        q = np.arange(B * N)[:, None, None]
        acc = np.zeros((B * N, R, R), dtype=F32)
        mask = np.zeros((B * N, R, R), dtype=np.uint8)
        bit = 0
        for dy, wy in ((0, wy0), (1, wy1)):
            for dx, wx in ((0, wx0), (1, wx1)):
                xx = (x0 + dx)[:, :, None]                          # (Q, a, 1)
                yy = (y0 + dy)[:, None, :]                          # (Q, 1, b')
                ok = (xx >= 0) & (xx < Wl) & (yy >= 0) & (yy < Hl)
                val = v[q, np.clip(yy, 0, Hl - 1), np.clip(xx, 0, Wl - 1)]
                wgt = (wx[:, :, None] * wy[:, None, :]).astype(F32)
                acc = (acc + np.where(ok, val * wgt, F32(0))).astype(F32)
                mask |= (ok.astype(np.uint8) << bit)
                bit += 1
        outs.append(acc.reshape(B, H, W, R * R))
        dbg.append({"x0": x0, "y0": y0, "mask": mask})
    out = np.concatenate(outs, axis=-1).transpose(0, 3, 1, 2)
    out = np.ascontiguousarray(out, dtype=F32)
    return (out, dbg) if debug else out


# --------------------------------------------------------------------------- backward
def lookup_backward(grad_out, coords, shapes, radius=4, rounding="cuda", grad_pyramid=None):
    """Scatter-add of one lookup's output gradient into the gradient pyramid.

    grad_out: (B, L*R*R, H, W);  shapes: [(Hl, Wl)] per level.  Accumulates into
    ``grad_pyramid`` (list of (B, N, Hl, Wl) fp64 arrays, created when None) so
    that several lookups of one CorrBlock share ONE accumulator.
    fp64 accumulation: the reference's own atomics are order-nondeterministic, the
    oracle is the exact sum.
    """
    B, K, H, W = grad_out.shape
    N = H * W
    R = 2 * radius + 1
    L = len(shapes)
    assert K == L * R * R
    if grad_pyramid is None:
        grad_pyramid = [np.zeros((B, N, hl, wl), dtype=np.float64) for hl, wl in shapes]
    cx = coords[:, 0].reshape(B * N).astype(F32)
    cy = coords[:, 1].reshape(B * N).astype(F32)
    g = grad_out.transpose(0, 2, 3, 1).reshape(B * N, L, R, R).astype(np.float64)
    for l, (Hl, Wl) in enumerate(shapes):
        G = grad_pyramid[l].reshape(B * N, Hl, Wl)
        x0, wx0, wx1 = axis_taps(cx, Wl, l, radius, rounding)
        y0, wy0, wy1 = axis_taps(cy, Hl, l, radius, rounding)
        qidx = np.broadcast_to(np.arange(B * N)[:, None, None], (B * N, R, R))
        for dy, wy in ((0, wy0), (1, wy1)):
            for dx, wx in ((0, wx0), (1, wx1)):
                xx = np.broadcast_to((x0 + dx)[:, :, None], (B * N, R, R))
                yy = np.broadcast_to((y0 + dy)[:, None, :], (B * N, R, R))
                ok = (xx >= 0) & (xx < Wl) & (yy >= 0) & (yy < Hl)
                wgt = (wx[:, :, None] * wy[:, None, :]).astype(F32).astype(np.float64)
                np.add.at(G, (qidx[ok], yy[ok], xx[ok]), (g[:, l] * wgt)[ok])
    return grad_pyramid


def build_backward(grad_pyramid, fmap1, fmap2):
    """Fold the gradient pyramid down to level 0 and contract with the feature maps.

    avg_pool2d backward gives 1/4 of the parent's gradient to each child inside the
    floor-cropped area; then dC = G0 / sqrt(D), dF1[b,:,p] = sum_q dC[b,p,q] F2[b,:,q],
    dF2[b,:,q] = sum_p dC[b,p,q] F1[b,:,p]   (SURVEY.md A.4b).  fp64 throughout.
    """
    B, D, H, W = fmap1.shape
    N = H * W
    G = [np.array(g, dtype=np.float64) for g in grad_pyramid]
    for l in range(len(G) - 1, 0, -1):
        hl, wl = G[l].shape[-2:]
        up = np.repeat(np.repeat(G[l], 2, axis=-2), 2, axis=-1) * 0.25
        G[l - 1][..., : 2 * hl, : 2 * wl] += up
    dC = G[0].reshape(B, N, N) / np.sqrt(np.float64(D))
    f1 = fmap1.reshape(B, D, N).astype(np.float64)
    f2 = fmap2.reshape(B, D, N).astype(np.float64)
    d1 = np.einsum("bpq,bdq->bdp", dC, f2).reshape(B, D, H, W)
    d2 = np.einsum("bpq,bdp->bdq", dC, f1).reshape(B, D, H, W)
    return d1.astype(F32), d2.astype(F32)


# --------------------------------------------------------------------------- on-demand
def pool_features(fmap: np.ndarray, num_levels: int):
    """corr.py:68-72: avg-pooled feature pyramid (B, D, Hl, Wl)."""
    return pool_pyramid(fmap, num_levels)


def ondemand_lookup(fmap1, fmap2, coords, num_levels=4, radius=4, scale=True):
    """AlternateCorrBlock.__call__ (corr.py:74-91) with the kernel semantics of
    correlation_kernel.cu:59-116: raw floor of coords/2^l (no normalise round
    trip), dot products against the POOLED fmap2, bilinear splat of each of the
    (2r+2)^2 integer neighbours into <=4 output taps, division by sqrt(D) after
    sampling.  Output (B, L*R*R, H, W), channel = l*R*R + ix*R + iy (x-major).
    """
    B, D, H, W = fmap1.shape
    N = H * W
    R = 2 * radius + 1
    f1 = fmap1.reshape(B, D, N).astype(F32)
    f2p = pool_features(fmap2, num_levels)
    cx = coords[:, 0].reshape(B, N).astype(F32)
    cy = coords[:, 1].reshape(B, N).astype(F32)
    outs = []
    for l in range(num_levels):
        Hl, Wl = f2p[l].shape[-2:]
        xl = (cx / F32(2 ** l)).astype(F32)
        yl = (cy / F32(2 ** l)).astype(F32)
        xf, yf = np.floor(xl), np.floor(yl)
        dx, dy = (xl - xf).astype(F32), (yl - yf).astype(F32)
        x0, y0 = xf.astype(np.int64), yf.astype(np.int64)
        out = np.zeros((B, N, R, R), dtype=F32)                     # [ix][iy]
        g2 = f2p[l].reshape(B, D, Hl * Wl)
        for iy in range(R + 1):
            for ix in range(R + 1):
                h2 = y0 - radius + iy
                w2 = x0 - radius + ix
                ok = (h2 >= 0) & (h2 < Hl) & (w2 >= 0) & (w2 < Wl)
                lin = np.clip(h2, 0, Hl - 1) * Wl + np.clip(w2, 0, Wl - 1)
                gathered = np.take_along_axis(g2, lin[:, None, :], axis=2)   # (B, D, N)
                s = np.where(ok, np.einsum("bdn,bdn->bn", f1, gathered), F32(0)).astype(F32)
                if iy > 0 and ix > 0:
                    out[:, :, ix - 1, iy - 1] += s * dy * dx
                if iy > 0 and ix < R:
                    out[:, :, ix, iy - 1] += s * dy * (F32(1) - dx)
                if iy < R and ix > 0:
                    out[:, :, ix - 1, iy] += s * (F32(1) - dy) * dx
                if iy < R and ix < R:
                    out[:, :, ix, iy] += s * (F32(1) - dy) * (F32(1) - dx)
        outs.append(out.reshape(B, H, W, R * R))
    out = np.concatenate(outs, axis=-1).transpose(0, 3, 1, 2)
    if scale:
        out = out / np.sqrt(F32(D), dtype=F32)
    return np.ascontiguousarray(out, dtype=F32)


def ondemand_backward(fmap1, fmap2, coords, grad_corr, radius=4):
    """Adjoint of ONE unscaled level of ``ondemand_lookup`` w.r.t. both feature maps:
    the semantics of corr_backward_kernel (correlation_kernel.cu:122-256; exported by
    correlation.cpp:53 but never called by the reference).  grad_corr: (B, R*R, H, W)
    with channel = ix*R + iy.  fp64.  coords get no gradient (the reference leaves
    coords_grad at zero, correlation_kernel.cu:307)."""
    B, D, H, W = fmap1.shape
    N = H * W
    R = 2 * radius + 1
    f1 = fmap1.reshape(B, D, N).astype(np.float64)
    f2 = fmap2.reshape(B, D, N).astype(np.float64)
    g = grad_corr.reshape(B, R, R, N).astype(np.float64)            # [ix][iy]
    cx = coords[:, 0].reshape(B, N).astype(F32)
    cy = coords[:, 1].reshape(B, N).astype(F32)
    xf, yf = np.floor(cx), np.floor(cy)
    dx, dy = (cx - xf).astype(np.float64), (cy - yf).astype(np.float64)
    x0, y0 = xf.astype(np.int64), yf.astype(np.int64)
    d1 = np.zeros_like(f1)
    d2 = np.zeros_like(f2)
    bidx = np.arange(B)[:, None]
    for iy in range(R + 1):
        for ix in range(R + 1):
            ds = np.zeros((B, N))
            if iy > 0 and ix > 0:
                ds += g[:, ix - 1, iy - 1] * dy * dx
            if iy > 0 and ix < R:
                ds += g[:, ix, iy - 1] * dy * (1 - dx)
            if iy < R and ix > 0:
                ds += g[:, ix - 1, iy] * (1 - dy) * dx
            if iy < R and ix < R:
                ds += g[:, ix, iy] * (1 - dy) * (1 - dx)
            h2 = y0 - radius + iy
            w2 = x0 - radius + ix
            ok = (h2 >= 0) & (h2 < H) & (w2 >= 0) & (w2 < W)
            ds = np.where(ok, ds, 0.0)
            lin = np.clip(h2, 0, H - 1) * W + np.clip(w2, 0, W - 1)
            d1 += ds[:, None, :] * np.take_along_axis(f2, lin[:, None, :], axis=2)
            for d in range(D):
                np.add.at(d2[:, d], (bidx, lin), ds * f1[:, d])
    return d1.reshape(B, D, H, W).astype(F32), d2.reshape(B, D, H, W).astype(F32)


def coords_grid(B: int, H: int, W: int) -> np.ndarray:
    """utils.py:74-77: (B, 2, H, W) fp32, channel 0 = x, channel 1 = y."""
    ys, xs = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing="ij")
    g = np.stack([xs, ys], axis=0)
    return np.ascontiguousarray(np.broadcast_to(g[None], (B, 2, H, W)), dtype=F32)
