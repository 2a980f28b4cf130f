"""blasr_b200 -- B200-native refinement DP for BLASR (guided / k-band / SW alignment + traceback).

Only the hot path lives here: csrc/ (sm_100a kernels + the C ABI), capi.py (ctypes binding),
align.py (host-side mirror of the reference's aligner interface) and synth.py (seeded workloads).
"""
from . import capi  # noqa: F401
from .capi import BgpuError  # noqa: F401
from .align import (Aligner, Alignment, BatchResult, DistanceMatrixScoreFunction, IDSScoreFunction,  # noqa: F401
                    JobBatch, QualityValueScoreFunction, SMRTDistanceMatrix)

__version__ = "0.1.0"
