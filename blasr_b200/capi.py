"""ctypes binding of the C ABI in include/blasr_gpu.h (libblasr_gpu.so, built in-tree).

Nothing here computes an alignment: every call goes to the CUDA library, and loading fails
loudly when the library (or a GPU, at bgpu_create time) is missing -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BGPU_LIB_PATH points experiments at a variant build of the same library (kernel A/B timing); default = the in-tree build
LIB_PATH = os.environ.get("BGPU_LIB_PATH") or os.path.join(_HERE, "libblasr_gpu.so")
CSRC = os.path.join(_HERE, "csrc")

# --- enums (mirror include/blasr_gpu.h) ---
GUIDED, AFFINE_GUIDED, KBAND, SW, AFFINE_KBAND = 0, 1, 2, 3, 4
LOCAL, GLOBAL, QUERYFIT, TARGETFIT, OVERLAP, FRONTANCHORED, ENDANCHORED, FIT, TSUFFIXQPREFIX, TPREFIXQSUFFIX = range(10)
FN_DISTANCE, FN_QUALITY, FN_IDS = 0, 1, 2
JOB_OK, JOB_EMPTY_GUIDE, JOB_PATH_AWRY, JOB_BAD_INPUT, JOB_REF_UNDEFINED, JOB_TOO_WIDE, JOB_RANGE = range(7)
E_NO_DEVICE, E_CUDA, E_INVALID, E_OOM, E_BUSY = -1, -2, -3, -4, -5

EXPORTS = [
    "bgpu_create", "bgpu_destroy", "bgpu_last_error", "bgpu_version", "bgpu_submit", "bgpu_submit_jobs",
    "bgpu_collect", "bgpu_release", "bgpu_rerun", "bgpu_timing_of", "bgpu_align", "bgpu_device_count",
    "bgpu_measure_int_peak", "bgpu_int_peak_modes", "bgpu_cigar", "bgpu_base_code", "bgpu_query", "bgpu_trim", "bgpu_cigar_clipped", "bgpu_strings",
    "bgpu_sdp_align", "bgpu_set_reference", "bgpu_set_suffix_array", "bgpu_map_reads", "bgpu_map_timing", "bgpu_map_rerun", "bgpu_build_lookup_table", "bgpu_rescore",
]


class ScoreFn(C.Structure):
    _fields_ = [("M", C.c_int32 * 25), ("ins", C.c_int32), ("del_", C.c_int32), ("affineOpen", C.c_int32),
                ("affineExtend", C.c_int32), ("kind", C.c_int32), ("substitutionPrior", C.c_int32),
                ("globalDeletionPrior", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("algo", C.c_int32), ("alignType", C.c_int32), ("band", C.c_int32), ("bndIns", C.c_int32),
                ("bndDel", C.c_int32), ("doStats", C.c_int32), ("statsAffine", C.c_int32), ("hpInsOpen", C.c_int32),
                ("hpInsExtend", C.c_int32), ("insOpen", C.c_int32), ("insExtend", C.c_int32), ("compactResults", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("nJobs", C.c_uint32), ("qBases", C.c_void_p), ("qOff", C.c_void_p), ("tBases", C.c_void_p),
                ("tOff", C.c_void_p), ("qual", C.c_void_p), ("guide", C.c_void_p), ("guideOff", C.c_void_p),
                ("band", C.c_void_p), ("insQV", C.c_void_p), ("delQV", C.c_void_p), ("subQV", C.c_void_p),
                ("delTag", C.c_void_p), ("subTag", C.c_void_p), ("guidePacked", C.c_void_p), ("guideWide", C.c_void_p), ("nGuideWide", C.c_uint64),
                ("tRefOff", C.c_void_p), ("tRefRc", C.c_void_p)]


class Job(C.Structure):
    _fields_ = [("q", C.c_void_p), ("qLen", C.c_uint32), ("t", C.c_void_p), ("tLen", C.c_uint32),
                ("qual", C.c_void_p), ("guide", C.c_void_p), ("nGuide", C.c_uint32), ("band", C.c_int32),
                ("insQV", C.c_void_p), ("delQV", C.c_void_p), ("subQV", C.c_void_p), ("delTag", C.c_void_p),
                ("subTag", C.c_void_p)]


class Arena(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("nBlocks", C.c_uint64), ("gapCounts", C.c_void_p),
                ("nGapLists", C.c_uint64), ("gaps", C.c_void_p), ("nGaps", C.c_uint64), ("runs", C.c_void_p), ("nRuns", C.c_uint64)]


class SdpParams(C.Structure):   # bgpu_sdp_params: SDPAlign's parameter list (SDPAlign.h:95-107)
    _fields_ = [("wordSize", C.c_int32), ("sdpIns", C.c_int32), ("sdpDel", C.c_int32), ("indelRate", C.c_float),
                ("alignType", C.c_int32), ("detailed", C.c_int32), ("extendFront", C.c_int32), ("sdpPrefix", C.c_int32),
                ("recurse", C.c_int32), ("noRecurseUnder", C.c_int32), ("maxMatches", C.c_int32)]


class AnchorParams(C.Structure):   # bgpu_anchor_params: MapReadToGenome's scalar arguments (AnchorParameters.h:10-27)
    _fields_ = [("minPrefixMatchLength", C.c_uint32), ("minMatchLength", C.c_uint32), ("expand", C.c_int32), ("useLookupTable", C.c_int32),
                ("maxAnchorsPerPosition", C.c_int32), ("advanceExactMatches", C.c_int32), ("maxLCPLength", C.c_int32),
                ("stopMappingOnceUnique", C.c_int32), ("removeEncompassedMatches", C.c_int32)]


MATCH_DTYPE = np.dtype([("t", "<u4"), ("q", "<u4"), ("l", "<u4")])


class Timing(C.Structure):
    _fields_ = [("msPrep", C.c_double), ("msFill", C.c_double), ("msTrace", C.c_double), ("msEmit", C.c_double),
                ("msTotal", C.c_double), ("cells", C.c_uint64), ("fillCells", C.c_uint64),
                ("kernelLaunches", C.c_uint32), ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64),
                ("msHostSubmit", C.c_double), ("msHostCollect", C.c_double), ("devAllocs", C.c_uint32),
                ("pinAllocs", C.c_uint32)]


# numpy view of bgpu_result (natural C alignment)
RESULT_DTYPE = np.dtype({
    "names": ["status", "score", "qPos", "tPos", "nCells", "nMatch", "nMismatch", "nIns", "nDel", "pctSimilarity",
              "statsScore", "nBlocks", "blockOff", "nGapLists", "gapListOff", "nGaps", "gapOff"],
    "formats": ["<i4", "<i4", "<u4", "<u4", "<i4", "<i4", "<i4", "<i4", "<i4", "<f4", "<i4", "<u4", "<u8", "<u4",
                "<u8", "<u4", "<u8"],
    "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 56, 64, 72, 80],
    "itemsize": 88,
})
BLOCK_DTYPE = np.dtype([("qPos", "<u4"), ("tPos", "<u4"), ("length", "<u4")])
GAP_DTYPE = np.dtype([("seq", "<i4"), ("length", "<i4")])


def build(force: bool = False) -> str:
    """Compile blasr_b200/csrc/*.cu for sm_100a into blasr_b200/libblasr_gpu.so (nvcc cross-compiles on CPU boxes)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "blasr_gpu.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the product path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.bgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.bgpu_destroy.argtypes = [C.c_void_p]
    L.bgpu_destroy.restype = None
    L.bgpu_last_error.argtypes = [C.c_void_p]
    L.bgpu_last_error.restype = C.c_char_p
    L.bgpu_submit.argtypes = [C.c_void_p, C.POINTER(ScoreFn), C.POINTER(Params), C.POINTER(Batch), C.POINTER(C.c_void_p)]
    L.bgpu_submit_jobs.argtypes = [C.c_void_p, C.POINTER(ScoreFn), C.POINTER(Params), C.POINTER(Job), C.c_uint32,
                                   C.POINTER(C.c_void_p)]
    L.bgpu_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Arena)]
    L.bgpu_release.argtypes = [C.c_void_p, C.c_void_p]
    L.bgpu_query.argtypes = [C.c_void_p, C.c_void_p]
    L.bgpu_trim.argtypes = [C.c_void_p]
    L.bgpu_cigar_clipped.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.bgpu_strings.argtypes = [C.c_void_p, C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
    L.bgpu_rerun.argtypes = [C.c_void_p, C.c_void_p]
    L.bgpu_cigar.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.bgpu_timing_of.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Timing)]
    L.bgpu_align.argtypes = [C.c_void_p, C.POINTER(ScoreFn), C.POINTER(Params), C.POINTER(Batch), C.c_void_p,
                             C.POINTER(Arena)]
    L.bgpu_sdp_align.argtypes = [C.c_void_p, C.POINTER(ScoreFn), C.POINTER(SdpParams), C.POINTER(Batch), C.c_void_p, C.POINTER(Arena)]
    L.bgpu_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.bgpu_set_suffix_array.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
    L.bgpu_map_reads.argtypes = [C.c_void_p, C.POINTER(AnchorParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_void_p)]
    L.bgpu_map_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double * 2), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.bgpu_map_rerun.argtypes = [C.c_void_p]
    L.bgpu_rescore.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ScoreFn), C.c_int, C.c_void_p]
    L.bgpu_build_lookup_table.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.bgpu_measure_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bgpu_int_peak_modes.argtypes = [C.POINTER(C.c_double * 4)]
    _lib = L
    return L


class BgpuError(RuntimeError):
    pass
