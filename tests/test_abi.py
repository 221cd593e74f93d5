"""CPU: the C-ABI library loads and exports exactly what include/flowcorr.h declares, the
ctypes prototypes agree with the header, and the geometry helpers (pure host code) work
without a GPU.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flowcorr.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|size_t|unsigned long long|const char\s*\*)\s+(fc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    out = {}
    for name, args in decls:
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


@pytest.fixture(scope="module")
def lib():
    from flow_supervisor_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_declares_the_path():
    fns = header_functions()
    for name in ["fc_build", "fc_lookup_fwd", "fc_lookup_bwd", "fc_build_bwd", "fc_ondemand_prepare",
                 "fc_ondemand_fwd", "fc_altcorr_fwd", "fc_altcorr_bwd", "fc_pyramid_bytes", "fc_last_error"]:
        assert name in fns


def test_library_exports_every_declared_symbol(lib):
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(cdll, name), f"{name} declared in flowcorr.h but not exported"


def test_binding_matches_header(lib):
    fns = header_functions()
    assert set(fns) == set(lib.SIGNATURES), set(fns) ^ set(lib.SIGNATURES)
    for name, nargs in fns.items():
        assert len(lib.SIGNATURES[name][1]) == nargs, name
    assert lib.load().fc_abi_version() == lib.ABI_VERSION
    src = open(HEADER).read()
    assert f"#define FC_ABI_VERSION {lib.ABI_VERSION}" in src
    for macro, val in [("FC_VOL_F32", lib.VOL_F32), ("FC_VOL_BF16", lib.VOL_BF16), ("FC_MATH_FP32", lib.MATH_FP32),
                       ("FC_MATH_TC_3XBF16", lib.MATH_TC_3XBF16), ("FC_MATH_TC_BF16", lib.MATH_TC_BF16),
                       ("FC_COORD_CUDA", lib.COORD_CUDA), ("FC_COORD_CPU", lib.COORD_CPU),
                       ("FC_MAX_LEVELS", lib.MAX_LEVELS), ("FC_MAX_RADIUS", lib.MAX_RADIUS)]:
        assert re.search(rf"#define {macro} {val}\b", src), macro


def test_geometry_host_functions(lib):
    from flow_supervisor_b200 import ops
    # Sintel 55x128: 55x128, 27x64, 13x32, 6x16
    dims = [lib.level_dims(55, 128, l) for l in range(4)]
    assert [(h, w) for h, w, _ in dims] == [(55, 128), (27, 64), (13, 32), (6, 16)]
    assert dims == ops.geometry(55, 128, 4)
    # odd widths are padded to a multiple of 8 columns
    assert lib.level_dims(46, 62, 0) == (46, 62, 64) and lib.level_dims(46, 62, 3) == (5, 7, 8)
    total, offs = lib.pyramid_layout(8, 55, 128, 4, lib.VOL_F32)
    assert total == 4 * ops.pyramid_numel(8, 55, 128, 4)
    assert total == 4 * 8 * 7040 * (56 * 128 + 28 * 64 + 14 * 32 + 6 * 16)      # rows padded to even
    assert offs[0] == 0 and offs == sorted(offs)
    assert lib.pyramid_layout(8, 55, 128, 4, lib.VOL_BF16)[0] * 2 == total


def test_errors_are_reported_not_thrown(lib):
    L = lib.load()
    assert L.fc_pyramid_bytes(0, 55, 128, 4, 0, None) == 0
    assert b"bad geometry" in L.fc_last_error()
    # null pointers are rejected before anything touches the GPU
    assert L.fc_lookup_fwd(None, None, None, 1, 16, 16, 4, 4, 0, 0, None, None, None, None) == -1
    assert b"null" in L.fc_last_error()
    with pytest.raises(RuntimeError):
        lib.check(-1, "probe")
