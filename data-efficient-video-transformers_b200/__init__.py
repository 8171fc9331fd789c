"""B200-native hot path of ed-fish/data-efficient-video-transformers: hand-written sm_100a kernels behind
the reference's model API.  Import as ``tvt_b200`` (see tvt_b200.py at the repo root) — the directory name
mirrors the reference repository and is not a valid Python identifier.

Layout:  csrc/ (CUDA kernels + C-ABI, built into libtvt_b200.so by build.py), capi.py (ctypes binding),
ops.py (tensor-level wrappers), functions.py (autograd Functions), hostapi/ (reference-facing modules),
ddp.py (clip-batch data parallelism over NCCL).
"""
from . import capi, ops  # noqa: F401
from .capi import TvtError  # noqa: F401
