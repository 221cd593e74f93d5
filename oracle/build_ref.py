#!/usr/bin/env python
"""Build the reference's ONLY native module (pytorch/alt_cuda_corr) for sm_100a --
TEST INFRASTRUCTURE ONLY (the checker for rows a8/a9 of SURVEY.md section 8 and the GPU
baseline of the on-demand path).

The two sources are compiled WHERE THEY LIE under /root/reference
(pytorch/alt_cuda_corr/correlation.cpp, correlation_kernel.cu); nothing is copied and the
reference's own build system (its setup.py) is not run.  Outputs go to oracle/_ref/ only
(git-ignored, not gpurun-ignored: the .so travels to the GPU box, the reference tree does
not).  On a machine without /root/reference (the GPU box) this is a no-op.

    python oracle/build_ref.py            # ~4 min (torch/extension.h translation units)

Load it with ``oracle.ref_ext.load()``.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pytorch/alt_cuda_corr"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "alt_cuda_corr_ref"          # module name inside the .so (PYBIND11_MODULE(TORCH_EXTENSION_NAME, ..))


def built_path() -> str:
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose: bool = False) -> str | None:
    srcs = [os.path.join(REF_SRC, "correlation.cpp"), os.path.join(REF_SRC, "correlation_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return built_path() if os.path.exists(built_path()) else None
    if os.path.exists(built_path()) and all(
            os.path.getmtime(built_path()) >= os.path.getmtime(s) for s in srcs):
        return built_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")      # no GPU here: do not probe one
    os.environ.setdefault("MAX_JOBS", "2")
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=srcs, build_directory=OUT_DIR, with_cuda=True, is_python_module=False,
         extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose)
    return built_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(p if p else "reference sources absent and no prebuilt oracle/_ref: nothing built")
