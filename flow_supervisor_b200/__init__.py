"""flow_supervisor_b200 -- B200-native (sm_100a) correlation block for RAFT / flow-supervisor.

Host-side mirror of the reference interface for ONE path: ``pytorch/core/corr.py``
(``CorrBlock`` / ``AlternateCorrBlock``) and the ``alt_cuda_corr`` extension module.
All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/flowcorr.h`` (``libflowcorr.so``); there is no CPU or PyTorch fallback.
"""
from . import ops  # noqa: F401
from .corr import AlternateCorrBlock, CorrBlock, coords_grid  # noqa: F401
from .patch import patch_reference, unpatch_reference  # noqa: F401
from .runner import RaftRunner  # noqa: F401

__all__ = ["CorrBlock", "AlternateCorrBlock", "coords_grid", "patch_reference", "unpatch_reference", "RaftRunner"]
