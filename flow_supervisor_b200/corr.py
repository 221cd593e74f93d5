"""Host-side mirror of the reference's correlation interface.

``CorrBlock`` / ``AlternateCorrBlock`` keep the constructor and call signature of
/root/reference/pytorch/core/corr.py:12-91 (and its GMA twin gma_corr.py:15-63) so
RAFT / L2L / RAFTGMA / GMAL2L and the train / evaluate scripts run unchanged:

    corr_fn = CorrBlock(fmap1, fmap2, num_levels=4, radius=4)   # raft.py:105-107
    corr = corr_fn(coords1)                                    # raft.py:124

Arithmetic lives in libflowcorr.so (CUDA, sm_100a).  Differences to the reference
that a caller can observe are listed in DESIGN.md ("Behavioural notes").

Modes (class attributes, or environment variables read at import):
    CorrBlock.math       'auto' | 'fp32' | '3xbf16' | 'bf16'   (FLOWCORR_MATH; default 'auto' =
                          '3xbf16' tensor-core parity mode when the shape allows it
                          (D % 64 == 0, D <= 256, W <= 256 tokens), else 'fp32' CUDA-core mode;
                          both are this library's own kernels and both meet the 1e-4 contract)
    CorrBlock.volume     'f32'  | 'bf16'              (FLOWCORR_VOLUME; 'bf16' = inference-only bf16 VOLUME written by the
                          tensor-core build, read by the same lookup kernel at half the bytes; values within 2^-8
                          of the fp32-volume lookup, final flow within 0.05 px -- stated separately from the fp32 contract)
    CorrBlock.coord_mode 'cuda' | 'cpu'               (FLOWCORR_COORD; which device's
                          rounding of utils.py:61-62 to reproduce; default 'cuda')
"""
from __future__ import annotations

import os
import sys

import torch

from . import _lib, ops

_MATH = {"fp32": _lib.MATH_FP32, "3xbf16": _lib.MATH_TC_3XBF16, "bf16": _lib.MATH_TC_BF16}
_VOL = {"f32": _lib.VOL_F32, "bf16": _lib.VOL_BF16}
_COORD = {"cuda": _lib.COORD_CUDA, "cpu": _lib.COORD_CPU}


_noted = set()


def resolve_math(name: str, D: int, W: int) -> int:
    if name == "auto":
        # tensor-core kernel: D a multiple of 64 up to 256, 16 <= padded width <= 256 tokens (a tile of two
        # target rows is one UMMA N, 32..256); everything else runs the CUDA-core fp32 mode
        name = "3xbf16" if (D % 64 == 0 and D <= 256 and 16 <= (W + 7) // 8 * 8 <= 256) else "fp32"
        if name == "fp32" and (D, W) not in _noted and os.environ.get("FLOWCORR_VERBOSE", "1") != "0":
            _noted.add((D, W))
            print(f"[flowcorr] CorrBlock: D={D}, {W} tokens wide is outside the tensor-core build's range "
                  "(D % 64 == 0, D <= 256, 9..256 tokens): using the fp32 CUDA-core mode", file=sys.stderr)
    return _MATH[name]


def coords_grid(batch, ht, wd, device=None):
    """utils.py:74-77: (B, 2, H, W) fp32, channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def _lookup_fn():
    """torch.ops.flowcorr.lookup while a compiler traces, else the plain function behind it (no dispatcher hop)."""
    return ops.lookup if torch.compiler.is_compiling() else ops.lookup_direct


class _BlockState:
    """Per-CorrBlock state shared by the autograd nodes: the volume itself (never an
    autograd leaf) and ONE lazily zeroed gradient pyramid that every lookup's backward
    accumulates into (the reference materialises a volume-sized gradient per lookup
    per level)."""

    __slots__ = ("pyramid", "grad_pyramid", "B", "H", "W", "L", "radius", "math", "coord")

    def __init__(self):
        self.pyramid = None
        self.grad_pyramid = None


class _Build(torch.autograd.Function):
    """fmap1, fmap2 -> 1-element token.  The token carries the autograd dependency of
    every lookup on the feature maps; the volume travels in ``state``."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, state):
        state.pyramid = ops.build(fmap1, fmap2, state.L, state.math, _lib.VOL_F32)
        ctx.save_for_backward(fmap1, fmap2)
        ctx.state = state
        return fmap1.new_zeros(1)

    @staticmethod
    def backward(ctx, _grad_token):
        state = ctx.state
        fmap1, fmap2 = ctx.saved_tensors
        if state.grad_pyramid is None:            # no lookup contributed a gradient
            return torch.zeros_like(fmap1), torch.zeros_like(fmap2), None
        gp, state.grad_pyramid = state.grad_pyramid, None      # consumed; a second backward re-accumulates
        d1, d2 = ops.build_bwd(gp, fmap1, fmap2, state.L, state.math)
        return d1, d2, None


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, token, coords, state):
        ctx.save_for_backward(coords)
        ctx.state = state
        return _lookup_fn()(state.pyramid, coords, state.L, state.radius, state.coord)

    @staticmethod
    def backward(ctx, grad_out):
        state = ctx.state
        (coords,) = ctx.saved_tensors
        if state.grad_pyramid is None:
            state.grad_pyramid = torch.zeros(state.pyramid.numel(), dtype=torch.float32,
                                             device=state.pyramid.device)
        (ops.lookup_bwd if torch.compiler.is_compiling() else ops.lookup_bwd_direct)(
            grad_out, coords, state.grad_pyramid, state.L, state.radius, state.coord)
        # coords get no gradient: the reference always passes them detached (raft.py:123)
        return grad_out.new_zeros(1), None, None


class CorrBlock:
    math = os.environ.get("FLOWCORR_MATH", "auto")
    volume = os.environ.get("FLOWCORR_VOLUME", "f32")
    coord_mode = os.environ.get("FLOWCORR_COORD", "cuda")

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        if not (fmap1.is_cuda and fmap2.is_cuda):
            raise RuntimeError("flow_supervisor_b200.CorrBlock needs CUDA feature maps: this package "
                               "has no CPU fallback (use the reference on CPU)")
        if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
            raise ValueError(f"fmap1/fmap2 must both be (B, D, H, W); got {tuple(fmap1.shape)} "
                             f"and {tuple(fmap2.shape)}")
        self.num_levels = num_levels
        self.radius = radius
        st = self._state = _BlockState()
        st.B, _, st.H, st.W = fmap1.shape
        st.L, st.radius = num_levels, radius
        st.math, st.coord = resolve_math(self.math, fmap1.shape[1], st.W), _COORD[self.coord_mode]
        self._vol_dtype = _VOL[self.volume]
        fmap1, fmap2 = fmap1.float(), fmap2.float()
        self._token = None
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            if self._vol_dtype != _lib.VOL_F32:
                raise RuntimeError("training needs the fp32 volume (CorrBlock.volume = 'f32')")
            self._token = _Build.apply(fmap1, fmap2, st)
        else:
            st.pyramid = ops.build(fmap1.detach(), fmap2.detach(), num_levels, st.math, self._vol_dtype)

    @classmethod
    def from_fnet_tail(cls, x, packed_weights, out_dim, num_levels=4, radius=4):
        """Inference-only constructor fused with the feature encoder's 1x1 output convolution (SURVEY.md section 8 row
        f3; extractor.py:184, raft.py:99-107): ``x`` = (2B, C, H, W) activations in front of ``fnet.conv2`` (frames of
        image 1, then image 2), ``packed_weights`` = ops.fnet_tail_prepare(conv2.weight, conv2.bias).  Equivalent to
        ``CorrBlock(*torch.split(conv2(x), B), num_levels, radius)`` without the fp32 feature maps and the pack pre-pass."""
        if not x.is_cuda:
            raise RuntimeError("flow_supervisor_b200.CorrBlock needs CUDA tensors: this package has no CPU fallback")
        self = cls.__new__(cls)
        self.num_levels, self.radius = num_levels, radius
        st = self._state = _BlockState()
        st.B, st.H, st.W = x.shape[0] // 2, x.shape[2], x.shape[3]
        st.L, st.radius = num_levels, radius
        st.math, st.coord = resolve_math(cls.math, out_dim, st.W), _COORD[cls.coord_mode]
        if st.math == _lib.MATH_FP32 or not ops.fnet_tail_supported(x.shape[1], out_dim, st.H, st.W):
            raise RuntimeError(f"from_fnet_tail: C={x.shape[1]}, D={out_dim}, {st.H}x{st.W} tokens with math "
                               f"'{cls.math}' is outside the fused tail's range (tensor-core math, C in {{64, 128}}, "
                               "D % 64 == 0, H*W % 4 == 0): run conv2 and CorrBlock(fmap1, fmap2)")
        self._vol_dtype = _VOL[cls.volume]
        self._token = None
        st.pyramid = ops.build_from_fnet_tail(x.detach(), packed_weights, out_dim, num_levels, st.math, self._vol_dtype)
        return self

    @property
    def corr_pyramid(self):
        """corr.py:16,24,27: list of (B*N, 1, Hl, Wl) levels (gathered copies; the hot path
        never materialises them)."""
        st = self._state
        return ops.level_views(st.pyramid, st.B, st.H, st.W, st.L)

    def __call__(self, coords):
        st = self._state
        if coords.dim() != 4 or coords.shape[1] != 2 or coords.shape[0] != st.B \
                or coords.shape[2] != st.H or coords.shape[3] != st.W:
            raise ValueError(f"coords must be ({st.B}, 2, {st.H}, {st.W}); got {tuple(coords.shape)}")
        if self._token is not None and torch.is_grad_enabled():
            return _Lookup.apply(self._token, coords.detach(), st)
        return _lookup_fn()(st.pyramid, coords.detach(), st.L, st.radius, st.coord)

    def lookup_convc1(self, coords, packed_weights):
        """Inference-only: ``relu(convc1(self(coords)))`` (update.py:90) in one kernel, the 324-channel tensor never written
        (SURVEY.md section 8 row f1).  ``packed_weights`` = ops.convc1_prepare(convc1.weight, convc1.bias)."""
        st = self._state
        if not _lib.load().fc_lookup_convc1_supported(st.L, st.radius, 256):
            raise RuntimeError(f"lookup_convc1 is built for num_levels = 4, radius = 4 (got {st.L}, {st.radius})")
        fn = ops.lookup_convc1 if torch.compiler.is_compiling() else ops.lookup_convc1_direct
        return fn(st.pyramid, coords.detach(), packed_weights, st.L, st.radius, st.coord)

    def lookup_debug(self, coords):
        """(out, x0, y0, corner_mask): the lookup plus its integer part, for parity tests."""
        st = self._state
        return ops.lookup_debug(st.pyramid, coords.detach(), st.L, st.radius, st.coord)

    @staticmethod
    def corr(fmap1, fmap2):
        """corr.py:52-60: (B, H, W, 1, H, W) all-pairs volume (level 0 only)."""
        B, D, H, W = fmap1.shape
        pyr = ops.build(fmap1.detach().float(), fmap2.detach().float(), 1, resolve_math(CorrBlock.math, D, W),
                        _lib.VOL_F32)
        return ops.level_views(pyr, B, H, W, 1)[0].reshape(B, H, W, 1, H, W)


class AlternateCorrBlock:
    """On-demand correlation (corr.py:63-91).  Like the reference's, this block is
    inference-only (alt_cuda_corr.forward is a raw call with no autograd, corr.py:86).

    Two routes with the same indices and weights (``floor(coords / 2^l)``, no normalise
    round trip -- FC_COORD_RAW) and values equal up to fp32 rounding:

    * ``'ondemand'``: nothing volume-sized is stored; every call runs the fused
      dot-product + sampling kernel (fc_ondemand_fwd).
    * ``'materialise'``: the reference switches to this class when the volume does not fit
      its GPU; a B200 holds 180 GB, where config 5 of BASELINE.json (1088x1920, batch 2)
      needs 11.3 GB.  Building the pyramid once on the tensor cores and answering every
      call with the HBM-bound lookup costs less than three on-demand calls.

    ``AlternateCorrBlock.route`` (FLOWCORR_ALT_ROUTE) = 'auto' picks 'materialise' when the
    pyramid plus the build workspace fit in ``materialise_fraction`` (default 1/4) of the
    device memory that is free or cached-but-unused right now, else 'ondemand'.
    """

    route = os.environ.get("FLOWCORR_ALT_ROUTE", "auto")
    materialise_fraction = float(os.environ.get("FLOWCORR_ALT_FRACTION", "0.25"))

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        if not (fmap1.is_cuda and fmap2.is_cuda):
            raise RuntimeError("flow_supervisor_b200.AlternateCorrBlock needs CUDA feature maps "
                               "(no CPU fallback)")
        if self.route not in ("auto", "ondemand", "materialise"):
            raise ValueError(f"AlternateCorrBlock.route must be auto|ondemand|materialise, got {self.route!r}")
        self.num_levels = num_levels
        self.radius = radius
        self._shape = tuple(fmap1.shape)
        B, D, H, W = self._shape
        f1, f2 = fmap1.detach().float(), fmap2.detach().float()
        self.materialised = self.route == "materialise" or (
            self.route == "auto" and self._fits(f1.device, B, D, H, W, num_levels))
        if self.materialised:
            self._math = resolve_math(CorrBlock.math, D, W)
            self._pyramid = ops.build(f1, f2, num_levels, self._math, _lib.VOL_F32)
        else:
            self._ws = ops.ondemand_prepare(f1, f2, num_levels)

    @classmethod
    def _fits(cls, device, B, D, H, W, L) -> bool:
        math = resolve_math(CorrBlock.math, D, W)
        need = 4 * ops.pyramid_numel(B, H, W, L) + _lib.load().fc_build_workspace_bytes(B, D, H, W, L, math)
        free, _total = torch.cuda.mem_get_info(device)
        cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
        return need <= cls.materialise_fraction * (free + cached)

    def __call__(self, coords):
        B, D, H, W = self._shape
        if self.materialised:
            return _lookup_fn()(self._pyramid, coords.detach(), self.num_levels, self.radius, _lib.COORD_RAW)
        return ops.ondemand_lookup(self._ws, coords.detach(), D, self.num_levels, self.radius)
