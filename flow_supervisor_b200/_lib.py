"""ctypes binding of libflowcorr.so (the C ABI declared in include/flowcorr.h).

There is deliberately NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` (or ``make -C flow_supervisor_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOWCORR_LIB") or os.path.join(_HERE, "libflowcorr.so")   # FLOWCORR_LIB: A/B builds (tools/)

# enums of include/flowcorr.h
VOL_F32, VOL_BF16 = 0, 1
MATH_FP32, MATH_TC_3XBF16, MATH_TC_BF16 = 0, 1, 2
COORD_CUDA, COORD_CPU, COORD_RAW = 0, 1, 2
MAX_LEVELS, MAX_RADIUS = 6, 4
ABI_VERSION = 1

_p, _i, _z = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); the single source the ABI test checks against the header
SIGNATURES = {
    "fc_abi_version": (_i, []),
    "fc_last_error": (C.c_char_p, []),
    "fc_tunable_set": (_i, [C.c_char_p, _i]),
    "fc_kernel_launches": (C.c_ulonglong, []),
    "fc_level_dims": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "fc_pyramid_bytes": (_z, [_i, _i, _i, _i, _i, C.POINTER(_z)]),
    "fc_build_workspace_bytes": (_z, [_i, _i, _i, _i, _i, _i]),
    "fc_build": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "fc_lookup_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fc_lookup_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "fc_build_bwd_workspace_bytes": (_z, [_i, _i, _i, _i, _i, _i]),
    "fc_build_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "fc_ondemand_workspace_bytes": (_z, [_i, _i, _i, _i, _i]),
    "fc_ondemand_prepare": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _z, _p]),
    "fc_ondemand_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "fc_altcorr_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fc_altcorr_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fc_upsample_flow": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "fc_convc1_weights_bytes": (_z, []),
    "fc_lookup_convc1_supported": (_i, [_i, _i, _i]),
    "fc_convc1_prepare": (_i, [_p, _p, _p, _z, _p]),
    "fc_lookup_convc1_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fc_fnet_tail_weights_bytes": (_z, [_i, _i]),
    "fc_fnet_tail_supported": (_i, [_i, _i, _i, _i]),
    "fc_fnet_tail_prepare": (_i, [_p, _p, _i, _i, _p, _z, _p]),
    "fc_build_from_fnet_tail": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the library once and attach prototypes; raise loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"flow_supervisor_b200: {LIB_PATH} is missing. This package has no CPU or "
                    "PyTorch fallback; build the CUDA library first "
                    "(python -c 'import __graft_entry__ as g; g.build()').")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)       # AttributeError = ABI mismatch, also loud
                fn.restype, fn.argtypes = res, args
            if lib.fc_abi_version() != ABI_VERSION:
                raise RuntimeError(f"libflowcorr ABI {lib.fc_abi_version()} != binding {ABI_VERSION}")
            _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({status}): {msg}")


def level_dims(H: int, W: int, level: int):
    h, w, wp = _i(), _i(), _i()
    check(load().fc_level_dims(H, W, level, C.byref(h), C.byref(w), C.byref(wp)), "fc_level_dims")
    return h.value, w.value, wp.value


def pyramid_layout(B: int, H: int, W: int, L: int, vol_dtype: int):
    """-> (total_bytes, [byte offset per level])"""
    offs = (_z * L)()
    total = load().fc_pyramid_bytes(B, H, W, L, vol_dtype, offs)
    if total == 0:
        check(-1, "fc_pyramid_bytes")
    return total, list(offs)
