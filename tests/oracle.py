"""ctypes bindings of the two CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

  orc  -> oracle/liborc.so              plain-C restatement (oracle/orc_align.c)
  ref  -> oracle/_ref/libblasr_ref.so   the unmodified reference templates (oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_PATH = os.path.join(ROOT, "oracle", "liborc.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libblasr_ref.so")


class OrcScoreFn(C.Structure):
    _fields_ = [("M", C.c_int32 * 25), ("ins", C.c_int32), ("del_", C.c_int32), ("affineOpen", C.c_int32),
                ("affineExtend", C.c_int32), ("kind", C.c_int32), ("substitutionPrior", C.c_int32),
                ("globalDeletionPrior", C.c_int32)]


class OrcJob(C.Structure):
    _fields_ = [("algo", C.c_int32), ("alignType", C.c_int32), ("band", C.c_int32), ("bndIns", C.c_int32),
                ("bndDel", C.c_int32), ("doStats", C.c_int32), ("statsAffine", C.c_int32),
                ("q", C.c_void_p), ("qLen", C.c_uint32), ("t", C.c_void_p), ("tLen", C.c_uint32),
                ("qual", C.c_void_p), ("guide", C.c_void_p), ("nGuide", C.c_uint32),
                ("insQV", C.c_void_p), ("delQV", C.c_void_p), ("subQV", C.c_void_p), ("delTag", C.c_void_p),
                ("subTag", C.c_void_p), ("hpInsOpen", C.c_int32), ("hpInsExtend", C.c_int32), ("insOpen", C.c_int32),
                ("insExtend", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("score", C.c_int32), ("alnScore", C.c_int32), ("qPos", C.c_uint32),
                ("tPos", C.c_uint32), ("nCells", C.c_int32), ("nMatch", C.c_int32), ("nMismatch", C.c_int32),
                ("nIns", C.c_int32), ("nDel", C.c_int32), ("pctSimilarity", C.c_float), ("statsScore", C.c_int32),
                ("nBlocks", C.c_uint32), ("nGapLists", C.c_uint32), ("nGaps", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)


_libs = {}


def _load(which: str):
    if which in _libs:
        return _libs[which]
    path = ORC_PATH if which == "orc" else REF_PATH
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    fn = getattr(L, f"{which}_align")
    fn.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.POINTER(OrcResult), C.c_void_p, C.c_uint32, C.c_void_p,
                   C.c_uint32, C.c_void_p, C.c_uint32]
    gr = getattr(L, f"{which}_guide_rows")
    gr.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_int64)]
    rp = getattr(L, f"{which}_replay")
    rp.argtypes = [C.POINTER(OrcScoreFn), C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_int64)]
    rp.restype = C.c_int64
    sc = getattr(L, f"{which}_sdp_chain")
    sc.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    if which == "ref":
        L.ref_block_strings.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint32]
        L.ref_alignment_strings.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_sdp_fragments.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(OrcScoreFn), C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]
        L.ref_rescore.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.POINTER(OrcScoreFn), C.c_int, C.POINTER(C.c_int32)]
        L.ref_cigar.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.c_void_p, C.c_uint32]
        L.ref_sdp_guide.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(OrcScoreFn), C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_void_p, C.c_uint32]
    else:
        L.orc_guided_s16_model.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.c_void_p]
        L.orc_alignment_strings.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_sdp_align.argtypes = [C.POINTER(OrcScoreFn), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_sdp_fragments.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        L.orc_cigar_from.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    _libs[which] = L
    return L


def have_ref() -> bool:
    return _load("ref") is not None


def score_fn(M, ins, del_, affineOpen=0, affineExtend=0, kind=0, substitutionPrior=20, globalDeletionPrior=13) -> OrcScoreFn:
    f = OrcScoreFn()
    m = np.asarray(M, dtype=np.int32).reshape(25)
    for i in range(25):
        f.M[i] = int(m[i])
    f.ins, f.del_, f.affineOpen, f.affineExtend, f.kind = ins, del_, affineOpen, affineExtend, kind
    f.substitutionPrior, f.globalDeletionPrior = substitutionPrior, globalDeletionPrior
    return f


def make_job(algo, alignType, band, q: np.ndarray, t: np.ndarray, guide=None, qual=None, bndIns=0, bndDel=0, doStats=1,
             statsAffine=0, tracks=None, affineKBand=None):
    """Returns (OrcJob, keepalive).  tracks: dict of the rich QV tracks (insQV, delQV, subQV, delTag, subTag)."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    keep = [q, t]
    j = OrcJob()
    j.algo, j.alignType, j.band, j.bndIns, j.bndDel, j.doStats, j.statsAffine = algo, alignType, band, bndIns, bndDel, doStats, statsAffine
    j.q, j.qLen, j.t, j.tLen = q.ctypes.data, len(q), t.ctypes.data, len(t)
    if qual is not None:
        qual = np.ascontiguousarray(qual, np.uint8); keep.append(qual); j.qual = qual.ctypes.data
    if guide is not None and len(guide):
        guide = np.ascontiguousarray(guide, np.uint32).reshape(-1, 3); keep.append(guide)
        j.guide, j.nGuide = guide.ctypes.data, len(guide)
    if affineKBand is not None:   # (hpInsOpen, hpInsExtend, insOpen, insExtend); del travels as bndDel
        j.hpInsOpen, j.hpInsExtend, j.insOpen, j.insExtend = [int(x) for x in affineKBand]
    for name, arr in (tracks or {}).items():
        if arr is not None:
            arr = np.ascontiguousarray(arr, np.uint8); keep.append(arr); setattr(j, name, arr.ctypes.data)
    return j, keep


def align(which: str, fn: OrcScoreFn, job: OrcJob):
    """Run one job; returns dict with the result fields, blocks (n,3), gaps (list of lists of (seq,len))."""
    L = _load(which)
    if L is None:
        raise RuntimeError(f"oracle library '{which}' unavailable")
    cap = int(job.qLen) + int(job.tLen) + 8
    blocks = np.zeros((cap, 3), np.uint32); cnt = np.zeros(cap + 1, np.uint32); gaps = np.zeros((2 * cap, 2), np.int32)
    r = OrcResult()
    rc = getattr(L, f"{which}_align")(C.byref(fn), C.byref(job), C.byref(r), blocks.ctypes.data, cap, cnt.ctypes.data,
                                      cap + 1, gaps.ctypes.data, 2 * cap)
    if rc != 0:
        raise RuntimeError(f"{which}_align overflow rc={rc}")
    out = {k: getattr(r, k) for k, _ in OrcResult._fields_}
    out["blocks"] = blocks[:r.nBlocks].copy()
    gl, p = [], 0
    for i in range(r.nGapLists):
        gl.append([(int(a), int(b)) for a, b in gaps[p:p + int(cnt[i])]]); p += int(cnt[i])
    out["gaps"] = gl
    return out


def guided_s16_model(fn: OrcScoreFn, job: OrcJob):
    """oracle/orc_s16.c: dict of the model's counters, or None on unsupported input."""
    L = _load("orc")
    out = np.zeros(9, np.int64)
    if L.orc_guided_s16_model(C.byref(fn), C.byref(job), out.ctypes.data) != 0:
        return None
    keys = ("cells", "arrow_mismatches", "score_mismatches", "end32", "end16", "min_rel", "max_rel", "max_big", "rebases")
    return dict(zip(keys, (int(x) for x in out)))


def ref_alignment_strings(fn: OrcScoreFn, job: OrcJob):
    """(text, pattern, query) strings of CreateAlignmentStrings on the result of the job's reference aligner."""
    L = _load("ref")
    cap = int(job.qLen) + int(job.tLen) + 8
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    n = L.ref_alignment_strings(C.byref(fn), C.byref(job), bufs[0], bufs[1], bufs[2], cap)
    if n < 0:
        raise RuntimeError("ref_alignment_strings overflow")
    return tuple(b.raw[:n] for b in bufs)


def ref_block_strings(q: np.ndarray, t: np.ndarray, blocks: np.ndarray):
    L = _load("ref")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    blocks = np.ascontiguousarray(blocks, np.uint32).reshape(-1, 3)
    cap = len(q) + len(t) + 8
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    n = L.ref_block_strings(q.ctypes.data, len(q), t.ctypes.data, len(t), blocks.ctypes.data, len(blocks), bufs[0], bufs[1], bufs[2], cap)
    if n < 0:
        raise RuntimeError("ref_block_strings overflow")
    return tuple(b.raw[:n] for b in bufs)


def orc_alignment_strings(q: np.ndarray, t: np.ndarray, aln: dict, with_gaps=True):
    L = _load("orc")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    blocks = np.ascontiguousarray(aln["blocks"], np.uint32).reshape(-1, 3)
    gl = aln["gaps"] if with_gaps else []
    cnt = np.asarray([len(g) for g in gl], np.uint32)
    flat = np.asarray([x for g in gl for x in g], np.int32).reshape(-1, 2)
    cap = len(q) + len(t) + 8
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    n = L.orc_alignment_strings(q.ctypes.data, t.ctypes.data, int(aln["qPos"]), int(aln["tPos"]), blocks.ctypes.data, len(blocks),
                                cnt.ctypes.data, len(cnt), flat.ctypes.data, bufs[0], bufs[1], bufs[2], cap)
    if n < 0:
        raise RuntimeError("orc_alignment_strings overflow")
    return tuple(b.raw[:n] for b in bufs)


def ref_cigar(fn: OrcScoreFn, job: OrcJob) -> np.ndarray:
    """The job's aligner, then the reference's own CreateNoClippingCigarOps on its result: BAM-packed ops."""
    L = _load("ref")
    cap = int(job.qLen) + int(job.tLen) + 8
    ops = np.zeros(cap, np.uint32)
    n = L.ref_cigar(C.byref(fn), C.byref(job), ops.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("ref_cigar overflow")
    return ops[:n].copy()


def ref_cigar_string(fn: OrcScoreFn, job: OrcJob, clipping: int, tStrand: int, qSeqPos: int, readLength: int, lowQPrefix: int,
                     lowQSuffix: int):
    """The reference's whole CreateCIGARString on the job's alignment placed inside a longer read: (text, clips[4])."""
    L = _load("ref")
    cap = 16 * (int(job.qLen) + int(job.tLen)) + 64
    buf = C.create_string_buffer(cap)
    clips = np.zeros(4, np.uint32)
    L.ref_cigar_string.argtypes = [C.POINTER(OrcScoreFn), C.POINTER(OrcJob), C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    n = L.ref_cigar_string(C.byref(fn), C.byref(job), clipping, tStrand, qSeqPos, readLength, lowQPrefix, lowQSuffix, buf, cap,
                           clips.ctypes.data)
    if n < 0:
        raise RuntimeError("ref_cigar_string overflow")
    return buf.raw[:n].decode(), clips


def orc_cigar_from(q: np.ndarray, t: np.ndarray, aln: dict) -> np.ndarray:
    """C restatement of the printer, from an alignment dict as align() returns it."""
    L = _load("orc")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    blocks = np.ascontiguousarray(aln["blocks"], np.uint32).reshape(-1, 3)
    cnt = np.asarray([len(g) for g in aln["gaps"]], np.uint32)
    flat = np.asarray([x for g in aln["gaps"] for x in g], np.int32).reshape(-1, 2)
    cap = len(q) + len(t) + 8
    ops = np.zeros(cap, np.uint32)
    n = L.orc_cigar_from(q.ctypes.data, t.ctypes.data, int(aln["qPos"]), int(aln["tPos"]), blocks.ctypes.data, len(blocks),
                         cnt.ctypes.data, len(cnt), flat.ctypes.data, ops.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("orc_cigar_from overflow")
    return ops[:n].copy()


def sdp_chain(which: str, frags: np.ndarray, queryLength: int, fragmentLength: int, ins: int, del_: int, match: int,
              alignType: int) -> np.ndarray:
    """SDPLongestCommonSubsequence over frags (n x {x, y, length, weight}, unique (x, y)): chain of indices into the
    (x, y)-sorted set."""
    L = _load(which)
    frags = np.ascontiguousarray(frags, np.uint32).reshape(-1, 4)
    chain = np.zeros(len(frags) + 1, np.int32)
    n = getattr(L, f"{which}_sdp_chain")(frags.ctypes.data, len(frags), queryLength, fragmentLength, ins, del_, match, alignType,
                                         chain.ctypes.data, len(chain))
    if n < 0:
        raise RuntimeError("sdp_chain overflow")
    return chain[:n].copy()


def ref_sdp_fragments(q: np.ndarray, t: np.ndarray, fn: OrcScoreFn, wordSize=11, sdpIns=5, sdpDel=10, alignType=0):
    """The fragment set the reference's SDPAlign leaves in its buffers (sorted, de-duplicated) and its chain."""
    L = _load("ref")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    cap = 64 * (len(q) + len(t)) + 4096
    frags = np.zeros((cap, 4), np.uint32); chain = np.zeros(len(q) + len(t) + 8, np.int32); nc = C.c_int32(0)
    n = L.ref_sdp_fragments(q.ctypes.data, len(q), t.ctypes.data, len(t), C.byref(fn), wordSize, sdpIns, sdpDel, alignType,
                            frags.ctypes.data, cap, chain.ctypes.data, len(chain), C.byref(nc))
    if n < 0:
        raise RuntimeError("ref_sdp_fragments overflow")
    return frags[:n].copy(), chain[:nc.value].copy()


def orc_sdp_fragments(q: np.ndarray, t: np.ndarray, wordSize=11, sdpPrefixLength=50):
    """C restatement of the fragment-set construction; None when the restated std::sort is unpinned for this input."""
    L = _load("orc")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    cap = 64 * (len(q) + len(t)) + 4096
    frags = np.zeros((cap, 4), np.uint32)
    n = L.orc_sdp_fragments(q.ctypes.data, len(q), t.ctypes.data, len(t), wordSize, sdpPrefixLength, frags.ctypes.data, cap)
    if n == -2:
        return None
    if n < 0:
        raise RuntimeError("orc_sdp_fragments overflow")
    return frags[:n].copy()


def orc_sdp_guide(q: np.ndarray, t: np.ndarray, fn: OrcScoreFn, tupleSize=11, sdpIns=5, sdpDel=10, indelRate=0.9):
    """C restatement of SDPAlign with blasr's argument pattern (Blasr.cpp:1716-1722): absolute blocks, like sdp_guide()."""
    L = _load("orc")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    cap = len(q) + len(t) + 8
    blocks = np.zeros((cap, 3), np.uint32); qp = C.c_uint32(0); tp = C.c_uint32(0)
    n = L.orc_sdp_align(C.byref(fn), q.ctypes.data, len(q), t.ctypes.data, len(t), tupleSize, sdpIns, sdpDel, C.c_float(indelRate),
                        0, 1, 0, 50, 2, 1000, blocks.ctypes.data, cap, C.byref(qp), C.byref(tp))
    if n < 0:
        raise RuntimeError(f"orc_sdp_align rc={n}")
    out = blocks[:n].copy()
    out[:, 0] += qp.value; out[:, 1] += tp.value
    return out


def guide_rows(which: str, guide: np.ndarray, band: int):
    L = _load(which)
    guide = np.ascontiguousarray(guide, np.uint32).reshape(-1, 3)
    cap = int(guide[-1, 0] + guide[-1, 2]) + 8 if len(guide) else 8
    rows = np.zeros((cap, 4), np.int32); nc = C.c_int64(0)
    n = getattr(L, f"{which}_guide_rows")(guide.ctypes.data, len(guide), band, rows.ctypes.data, cap, C.byref(nc))
    return rows[:max(n, 0)].copy(), nc.value


def sdp_guide(q: np.ndarray, t: np.ndarray, fn: OrcScoreFn, tupleSize=11, sdpIns=5, sdpDel=10, indelRate=0.9):
    """Reference SDPAlign blocks (absolute), the guide blasr hands to RefineAlignment."""
    L = _load("ref")
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    cap = len(q) + len(t) + 8
    blocks = np.zeros((cap, 3), np.uint32)
    n = L.ref_sdp_guide(q.ctypes.data, len(q), t.ctypes.data, len(t), C.byref(fn), tupleSize, sdpIns, sdpDel,
                        C.c_float(indelRate), blocks.ctypes.data, cap)
    return blocks[:max(n, 0)].copy()


def replay(which: str, fn: OrcScoreFn, jobs, nThreads: int):
    """jobs: list of OrcJob. Returns (total nCells, checksum)."""
    L = _load(which)
    arr = (OrcJob * len(jobs))(*jobs)
    s = C.c_int64(0)
    cells = getattr(L, f"{which}_replay")(C.byref(fn), arr, len(jobs), nThreads, C.byref(s))
    return int(cells), int(s.value)


def rescore(fn: OrcScoreFn, job: OrcJob, fn2: OrcScoreFn, useAffine: bool) -> int:
    """The reference's ComputeAlignmentScore(alignment, q, t, fn2, useAffine) (AlignmentUtils.h:127-169) of the alignment the job
    gets under fn: the rescoring step of StoreMapQVs."""
    out = C.c_int32(0)
    _load("ref").ref_rescore(C.byref(fn), C.byref(job), C.byref(fn2), int(useAffine), C.byref(out))
    return int(out.value)
