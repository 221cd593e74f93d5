"""torch custom ops over the C ABI (``torch.ops.flowcorr.*``).

PyTorch is plumbing here: it owns device memory (caching allocator) and the
current stream; every op forwards raw pointers to ``libflowcorr.so`` and launches
on ``torch.cuda.current_stream()``.  No op synchronises or touches the host, so
the lookup is CUDA-graph capturable.  CPU tensors are rejected: there is no
fallback path.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import _lib

_VOL_TORCH = {_lib.VOL_F32: torch.float32, _lib.VOL_BF16: torch.bfloat16}


def geometry(H: int, W: int, L: int):
    """[(Hl, Wl, Wp_l)] -- mirrors fc_level_dims (include/flowcorr.h) without the library,
    so fake/meta tracing needs no CUDA."""
    return [(H >> l, W >> l, ((W >> l) + 7) // 8 * 8) for l in range(L)]


def _hp(h: int) -> int:
    return (h + 1) // 2 * 2


def pyramid_numel(B: int, H: int, W: int, L: int) -> int:
    return sum(B * H * W * _hp(h) * wp for h, _, wp in geometry(H, W, L))


def level_padded(pyramid: Tensor, B: int, H: int, W: int, L: int) -> List[Tensor]:
    """Every level as a dense (B*N, Hp, Wp) tensor INCLUDING pad rows/columns, un-patched
    from the library's 2x8 patch layout (include/flowcorr.h).  Copies."""
    out, off, Q = [], 0, B * H * W
    for h, w, wp in geometry(H, W, L):
        hp = _hp(h)
        n = Q * hp * wp
        t = pyramid[off:off + n].view(Q, hp // 2, wp // 8, 2, 8).permute(0, 1, 3, 2, 4)
        out.append(t.reshape(Q, hp, wp))
        off += n
    return out


def clear_pads_(pyramid: Tensor, B: int, H: int, W: int, L: int) -> Tensor:
    """Zero the pad rows / pad columns of every level IN PLACE (the pyramid invariant of include/flowcorr.h).  The library's
    own kernels keep it; a caller that fills a (gradient) pyramid itself -- tests, probes -- restores it with this."""
    off, Q = 0, B * H * W
    for h, w, wp in geometry(H, W, L):
        hp = _hp(h)
        n = Q * hp * wp
        t = pyramid[off:off + n].view(Q, hp // 2, wp // 8, 2, 8)                    # (q, row pair, patch, sub-row, x)
        y = 2 * torch.arange(hp // 2, device=pyramid.device)[:, None, None, None] + torch.arange(2, device=pyramid.device)[None, None, :, None]
        x = 8 * torch.arange(wp // 8, device=pyramid.device)[None, :, None, None] + torch.arange(8, device=pyramid.device)[None, None, None, :]
        t.mul_(((y < h) & (x < w)).to(pyramid.dtype)[None])
        off += n
    return pyramid


def level_views(pyramid: Tensor, B: int, H: int, W: int, L: int) -> List[Tensor]:
    """The reference's ``corr_pyramid`` list ((B*N, 1, Hl, Wl), corr.py:14-27), gathered out
    of the patch layout (copies; nothing on the hot path reads them)."""
    return [t[:, None, :h, :w] for t, (h, w, _) in zip(level_padded(pyramid, B, H, W, L), geometry(H, W, L))]


def _need_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("flowcorr ops run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


def _f32c(t: Tensor) -> Tensor:
    return t.contiguous().float()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """Context that makes `device` current -- a no-op object when it already is (the common case; torch.cuda.device()
    costs several microseconds of host time per call, as much as a tenth of a lookup kernel)."""
    return _NO_SWITCH if device.index is None or torch.cuda.current_device() == device.index else torch.cuda.device(device)


# ----------------------------------------------------------------------------- build
@custom_op("flowcorr::build", mutates_args=())
def build(fmap1: Tensor, fmap2: Tensor, num_levels: int, math: int, vol_dtype: int) -> Tensor:
    """CorrBlock.__init__ (corr.py:13-27) -> flat pyramid buffer."""
    _need_cuda(fmap1, fmap2)
    f1, f2 = _f32c(fmap1), _f32c(fmap2)
    B, D, H, W = f1.shape
    lib = _lib.load()
    with torch.cuda.device(f1.device):
        pyr = torch.empty(pyramid_numel(B, H, W, num_levels), dtype=_VOL_TORCH[vol_dtype], device=f1.device)
        ws_bytes = lib.fc_build_workspace_bytes(B, D, H, W, num_levels, math)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=f1.device)
        _lib.check(lib.fc_build(f1.data_ptr(), f2.data_ptr(), pyr.data_ptr(), B, D, H, W, num_levels,
                                vol_dtype, math, ws.data_ptr(), ws_bytes, _stream()), "fc_build")
    return pyr


@build.register_fake
def _(fmap1, fmap2, num_levels, math, vol_dtype):
    B, D, H, W = fmap1.shape
    return fmap1.new_empty(pyramid_numel(B, H, W, num_levels), dtype=_VOL_TORCH[vol_dtype])


# ----------------------------------------------------------------------------- lookup
def lookup_direct(pyramid: Tensor, coords: Tensor, num_levels: int, radius: int, coord_mode: int) -> Tensor:
    """CorrBlock.__call__ (corr.py:29-50): (B, 2, H, W) -> (B, L*(2r+1)^2, H, W) fp32.
    The plain function behind ``torch.ops.flowcorr.lookup``: the classes call it directly in eager inference (the
    custom-op dispatcher costs more host time per call than the kernel runs on the GPU)."""
    _need_cuda(pyramid, coords)
    c = _f32c(coords)
    B, _, H, W = c.shape
    vol_dtype = _lib.VOL_F32 if pyramid.dtype == torch.float32 else _lib.VOL_BF16
    with _on(c.device):
        out = torch.empty(B, num_levels * (2 * radius + 1) ** 2, H, W, dtype=torch.float32, device=c.device)
        _lib.check(_lib.load().fc_lookup_fwd(pyramid.data_ptr(), c.data_ptr(), out.data_ptr(), B, H, W,
                                             num_levels, radius, vol_dtype, coord_mode,
                                             None, None, None, _stream()), "fc_lookup_fwd")
    return out


lookup = custom_op("flowcorr::lookup", mutates_args=())(lookup_direct)


@lookup.register_fake
def _(pyramid, coords, num_levels, radius, coord_mode):
    B, _, H, W = coords.shape
    return coords.new_empty(B, num_levels * (2 * radius + 1) ** 2, H, W, dtype=torch.float32)


@custom_op("flowcorr::lookup_debug", mutates_args=())
def lookup_debug(pyramid: Tensor, coords: Tensor, num_levels: int, radius: int,
                 coord_mode: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Lookup plus the integer part: x0/y0 (B*N, L, 2r+1) int32 and the corner-mask
    bytes (B*N, L, 2r+1, 2r+1) (bit layout in include/flowcorr.h)."""
    _need_cuda(pyramid, coords)
    c = _f32c(coords)
    B, _, H, W = c.shape
    R = 2 * radius + 1
    vol_dtype = _lib.VOL_F32 if pyramid.dtype == torch.float32 else _lib.VOL_BF16
    with torch.cuda.device(c.device):
        out = torch.empty(B, num_levels * R * R, H, W, dtype=torch.float32, device=c.device)
        x0 = torch.empty(B * H * W, num_levels, R, dtype=torch.int32, device=c.device)
        y0 = torch.empty_like(x0)
        mask = torch.empty(B * H * W, num_levels, R, R, dtype=torch.uint8, device=c.device)
        _lib.check(_lib.load().fc_lookup_fwd(pyramid.data_ptr(), c.data_ptr(), out.data_ptr(), B, H, W,
                                             num_levels, radius, vol_dtype, coord_mode,
                                             x0.data_ptr(), y0.data_ptr(), mask.data_ptr(), _stream()),
                   "fc_lookup_fwd")
    return out, x0, y0, mask


@lookup_debug.register_fake
def _(pyramid, coords, num_levels, radius, coord_mode):
    B, _, H, W = coords.shape
    R = 2 * radius + 1
    out = coords.new_empty(B, num_levels * R * R, H, W, dtype=torch.float32)
    x0 = coords.new_empty(B * H * W, num_levels, R, dtype=torch.int32)
    return out, x0, torch.empty_like(x0), coords.new_empty(B * H * W, num_levels, R, R, dtype=torch.uint8)


# ----------------------------------------------------------------------------- backward
def lookup_bwd_direct(grad_out: Tensor, coords: Tensor, grad_pyramid: Tensor, num_levels: int, radius: int,
                      coord_mode: int) -> None:
    """Scatter-add one lookup's output gradient into the block's gradient pyramid.  The plain function behind
    ``torch.ops.flowcorr.lookup_bwd``: the autograd functions call it directly in eager mode (the kernel runs 26-58 us,
    the dispatcher alone costs the host about as much per call)."""
    _need_cuda(grad_out, coords, grad_pyramid)
    g, c = _f32c(grad_out), _f32c(coords)
    B, _, H, W = c.shape
    with torch.cuda.device(c.device):
        _lib.check(_lib.load().fc_lookup_bwd(g.data_ptr(), c.data_ptr(), grad_pyramid.data_ptr(), B, H, W,
                                             num_levels, radius, coord_mode, _stream()), "fc_lookup_bwd")


lookup_bwd = custom_op("flowcorr::lookup_bwd", mutates_args=("grad_pyramid",))(lookup_bwd_direct)


@custom_op("flowcorr::build_bwd", mutates_args=("grad_pyramid",))
def build_bwd(grad_pyramid: Tensor, fmap1: Tensor, fmap2: Tensor, num_levels: int,
              math: int) -> Tuple[Tensor, Tensor]:
    """Fold the gradient pyramid (consumed) and contract with both feature maps."""
    _need_cuda(grad_pyramid, fmap1, fmap2)
    f1, f2 = _f32c(fmap1), _f32c(fmap2)
    B, D, H, W = f1.shape
    lib = _lib.load()
    with torch.cuda.device(f1.device):
        d1, d2 = torch.empty_like(f1), torch.empty_like(f2)
        ws_bytes = lib.fc_build_bwd_workspace_bytes(B, D, H, W, num_levels, math)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=f1.device)
        _lib.check(lib.fc_build_bwd(grad_pyramid.data_ptr(), f1.data_ptr(), f2.data_ptr(), d1.data_ptr(),
                                    d2.data_ptr(), B, D, H, W, num_levels, math, ws.data_ptr(), ws_bytes,
                                    _stream()), "fc_build_bwd")
    return d1, d2


@build_bwd.register_fake
def _(grad_pyramid, fmap1, fmap2, num_levels, math):
    return torch.empty_like(fmap1, dtype=torch.float32), torch.empty_like(fmap2, dtype=torch.float32)


# ----------------------------------------------------------------------------- on-demand
@custom_op("flowcorr::ondemand_prepare", mutates_args=())
def ondemand_prepare(fmap1: Tensor, fmap2: Tensor, num_levels: int) -> Tensor:
    """AlternateCorrBlock.__init__ (corr.py:64-72): channels-last fmap1 + pooled fmap2 pyramid."""
    _need_cuda(fmap1, fmap2)
    f1, f2 = _f32c(fmap1), _f32c(fmap2)
    B, D, H, W = f1.shape
    lib = _lib.load()
    with torch.cuda.device(f1.device):
        nbytes = lib.fc_ondemand_workspace_bytes(B, D, H, W, num_levels)
        if nbytes == 0:
            _lib.check(-1, "fc_ondemand_workspace_bytes")
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=f1.device)
        _lib.check(lib.fc_ondemand_prepare(f1.data_ptr(), f2.data_ptr(), B, D, H, W, num_levels,
                                           ws.data_ptr(), nbytes, _stream()), "fc_ondemand_prepare")
    return ws


@ondemand_prepare.register_fake
def _(fmap1, fmap2, num_levels):
    B, D, H, W = fmap1.shape
    n = B * H * W * D + sum(B * (H >> l) * (W >> l) * D for l in range(num_levels))
    return fmap1.new_empty(n, dtype=torch.float32)


@custom_op("flowcorr::ondemand_lookup", mutates_args=())
def ondemand_lookup(workspace: Tensor, coords: Tensor, dim: int, num_levels: int, radius: int) -> Tensor:
    """AlternateCorrBlock.__call__ (corr.py:74-91), all levels, scaled by 1/sqrt(D)."""
    _need_cuda(workspace, coords)
    c = _f32c(coords)
    B, _, H, W = c.shape
    with torch.cuda.device(c.device):
        out = torch.empty(B, num_levels * (2 * radius + 1) ** 2, H, W, dtype=torch.float32, device=c.device)
        _lib.check(_lib.load().fc_ondemand_fwd(workspace.data_ptr(), c.data_ptr(), out.data_ptr(), B, dim, H, W,
                                               num_levels, radius, _stream()), "fc_ondemand_fwd")
    return out


@ondemand_lookup.register_fake
def _(workspace, coords, dim, num_levels, radius):
    B, _, H, W = coords.shape
    return coords.new_empty(B, num_levels * (2 * radius + 1) ** 2, H, W, dtype=torch.float32)


@custom_op("flowcorr::altcorr_fwd", mutates_args=())
def altcorr_fwd(fmap1: Tensor, fmap2: Tensor, coords: Tensor, radius: int) -> Tensor:
    """alt_cuda_corr.forward (correlation.cpp:23-33): one level, channels-last, unscaled."""
    B, H1, W1, Cc = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    R = 2 * radius + 1
    with torch.cuda.device(fmap1.device):
        corr = torch.empty(B, 1, R * R, H1, W1, dtype=torch.float32, device=fmap1.device)
        _lib.check(_lib.load().fc_altcorr_fwd(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                              corr.data_ptr(), B, H1, W1, H2, W2, Cc, radius, _stream()),
                   "fc_altcorr_fwd")
    return corr


@altcorr_fwd.register_fake
def _(fmap1, fmap2, coords, radius):
    B, H1, W1, _ = fmap1.shape
    R = 2 * radius + 1
    return fmap1.new_empty(B, 1, R * R, H1, W1, dtype=torch.float32)


@custom_op("flowcorr::altcorr_bwd", mutates_args=())
def altcorr_bwd(fmap1: Tensor, fmap2: Tensor, coords: Tensor, corr_grad: Tensor,
                radius: int) -> Tuple[Tensor, Tensor]:
    """alt_cuda_corr.backward (correlation.cpp:36-48) -> (fmap1_grad, fmap2_grad)."""
    B, H1, W1, Cc = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    with torch.cuda.device(fmap1.device):
        d1, d2 = torch.empty_like(fmap1), torch.empty_like(fmap2)
        _lib.check(_lib.load().fc_altcorr_bwd(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                              corr_grad.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                                              B, H1, W1, H2, W2, Cc, radius, _stream()), "fc_altcorr_bwd")
    return d1, d2


@altcorr_bwd.register_fake
def _(fmap1, fmap2, coords, corr_grad, radius):
    return torch.empty_like(fmap1), torch.empty_like(fmap2)


# ----------------------------------------------------------------------------- adjacent components (section 8 f)
@custom_op("flowcorr::upsample_flow", mutates_args=())
def upsample_flow(flow: Tensor, mask: Tensor) -> Tensor:
    """RAFT.upsample_flow (raft.py:72-83): (B,2,H,W) flow + (B,576,H,W) mask -> (B,2,8H,8W).  Forward only."""
    _need_cuda(flow, mask)
    f, m = _f32c(flow), _f32c(mask)
    B, _, H, W = f.shape
    if tuple(m.shape) != (B, 576, H, W):
        raise ValueError(f"mask must be ({B}, 576, {H}, {W}); got {tuple(m.shape)}")
    with torch.cuda.device(f.device):
        out = torch.empty(B, 2, 8 * H, 8 * W, dtype=torch.float32, device=f.device)
        _lib.check(_lib.load().fc_upsample_flow(f.data_ptr(), m.data_ptr(), out.data_ptr(), B, H, W, _stream()),
                   "fc_upsample_flow")
    return out


@upsample_flow.register_fake
def _(flow, mask):
    B, _, H, W = flow.shape
    return flow.new_empty(B, 2, 8 * H, 8 * W, dtype=torch.float32)


@custom_op("flowcorr::fnet_tail_prepare", mutates_args=())
def fnet_tail_prepare(weight: Tensor, bias: Tensor) -> Tensor:
    """Pre-pack the weights of the encoder's 1x1 output convolution (extractor.py:145 ``conv2``: (D, C, 1, 1) + (D,))
    as K-major bf16 hi/lo + fp32 bias.  Once per model."""
    _need_cuda(weight, bias)
    w = _f32c(weight).reshape(weight.shape[0], -1)
    b = _f32c(bias)
    D, Cin = w.shape
    lib = _lib.load()
    with _on(w.device):
        packed = torch.empty(lib.fc_fnet_tail_weights_bytes(Cin, D), dtype=torch.uint8, device=w.device)
        _lib.check(lib.fc_fnet_tail_prepare(w.data_ptr(), b.data_ptr(), Cin, D, packed.data_ptr(), packed.numel(), _stream()),
                   "fc_fnet_tail_prepare")
    return packed


@fnet_tail_prepare.register_fake
def _(weight, bias):
    D, Cin = weight.shape[0], weight.shape[1]
    return weight.new_empty(D * Cin * 4 + D * 4, dtype=torch.uint8)


def fnet_tail_supported(Cin: int, D: int, H: int, W: int) -> bool:
    return bool(_lib.load().fc_fnet_tail_supported(Cin, D, H, W))


@custom_op("flowcorr::build_from_fnet_tail", mutates_args=())
def build_from_fnet_tail(x: Tensor, packed_weights: Tensor, out_dim: int, num_levels: int, math: int, vol_dtype: int) -> Tensor:
    """CorrBlock.__init__ fused with the encoder's output convolution: x = (2B, C, H, W) activations in front of ``conv2``
    (frames of image 1, then image 2) -> the pyramid of CorrBlock(conv2(x)[:B], conv2(x)[B:])."""
    _need_cuda(x, packed_weights)
    xc = _f32c(x)
    F2, Cin, H, W = xc.shape
    B = F2 // 2
    lib = _lib.load()
    with _on(xc.device):
        pyr = torch.empty(pyramid_numel(B, H, W, num_levels), dtype=_VOL_TORCH[vol_dtype], device=xc.device)
        ws_bytes = lib.fc_build_workspace_bytes(B, out_dim, H, W, num_levels, math)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xc.device)
        _lib.check(lib.fc_build_from_fnet_tail(xc.data_ptr(), packed_weights.data_ptr(), pyr.data_ptr(), B, Cin, out_dim, H, W,
                                               num_levels, vol_dtype, math, ws.data_ptr(), ws_bytes, _stream()),
                   "fc_build_from_fnet_tail")
    return pyr


@build_from_fnet_tail.register_fake
def _(x, packed_weights, out_dim, num_levels, math, vol_dtype):
    F2, _, H, W = x.shape
    return x.new_empty(pyramid_numel(F2 // 2, H, W, num_levels), dtype=_VOL_TORCH[vol_dtype])


@custom_op("flowcorr::convc1_prepare", mutates_args=())
def convc1_prepare(weight: Tensor, bias: Tensor) -> Tensor:
    """Pre-pack the motion encoder's first convolution (update.py:83 ``convc1``: (256, 324, 1, 1) + (256,)) for
    lookup_convc1.  Once per model."""
    _need_cuda(weight, bias)
    w = _f32c(weight).reshape(weight.shape[0], -1)
    b = _f32c(bias)
    if tuple(w.shape) != (256, 324):
        raise ValueError(f"convc1 weight must be (256, 324[, 1, 1]); got {tuple(weight.shape)}")
    lib = _lib.load()
    with _on(w.device):
        packed = torch.empty(lib.fc_convc1_weights_bytes(), dtype=torch.uint8, device=w.device)
        _lib.check(lib.fc_convc1_prepare(w.data_ptr(), b.data_ptr(), packed.data_ptr(), packed.numel(), _stream()),
                   "fc_convc1_prepare")
    return packed


@convc1_prepare.register_fake
def _(weight, bias):
    return weight.new_empty(256 * 384 * 4 + 256 * 4, dtype=torch.uint8)


def lookup_convc1_direct(pyramid: Tensor, coords: Tensor, packed_weights: Tensor, num_levels: int, radius: int,
                         coord_mode: int) -> Tensor:
    """relu(convc1(CorrBlock.__call__(coords))) in one kernel (update.py:90 after raft.py:124): -> (B, 256, H, W)."""
    _need_cuda(pyramid, coords, packed_weights)
    c = _f32c(coords)
    B, _, H, W = c.shape
    vol_dtype = _lib.VOL_F32 if pyramid.dtype == torch.float32 else _lib.VOL_BF16
    with _on(c.device):
        out = torch.empty(B, 256, H, W, dtype=torch.float32, device=c.device)
        _lib.check(_lib.load().fc_lookup_convc1_fwd(pyramid.data_ptr(), c.data_ptr(), packed_weights.data_ptr(), out.data_ptr(),
                                                    B, H, W, num_levels, radius, vol_dtype, coord_mode, _stream()),
                   "fc_lookup_convc1_fwd")
    return out


lookup_convc1 = custom_op("flowcorr::lookup_convc1", mutates_args=())(lookup_convc1_direct)


@lookup_convc1.register_fake
def _(pyramid, coords, packed_weights, num_levels, radius, coord_mode):
    B, _, H, W = coords.shape
    return coords.new_empty(B, 256, H, W, dtype=torch.float32)
