"""Host-side mirror of the reference's aligner interface for the refinement path.

The reference calls one template function per candidate:

    AffineGuidedAlign(q, t, guide, scoreFn, bandSize, buffers, out, Global, false)   Blasr.cpp:863
    GuidedAlign(q, t, guide, scoreFn, bandSize, buffers, out, Global, false)         Blasr.cpp:869
    KBandAlign(q, t, matchMat, ins, del, k, scoreMat, pathMat, out, scoreFn, type)   Blasr.cpp:717,820
    SWAlign(q, t, scoreMat, pathMat, out, scoreFn, type)                             SDPAlign.h:440,503,563
    ComputeAlignmentStats(out, q, t, scoreFn, useAffine)                             Blasr.cpp:875

Here the same names take a *batch* of candidates and run them on the GPU through the C ABI
(include/blasr_gpu.h); argument meaning, defaults and returned fields follow the reference.
This module holds no alignment arithmetic of its own.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import capi
from .capi import (AFFINE_GUIDED, AFFINE_KBAND, FN_DISTANCE, FN_IDS, FN_QUALITY, GLOBAL, GUIDED, KBAND, SW, BgpuError)

# ScoreMatrices.h:20-26
SMRTDistanceMatrix = np.array([[-5, 6, 6, 6, 0], [6, -5, 6, 6, 0], [6, 6, -5, 6, 0], [6, 6, 6, -5, 0],
                               [0, 0, 0, 0, 0]], dtype=np.int32)


@dataclass
class DistanceMatrixScoreFunction:
    """DistanceMatrixScoreFunction<DNASequence,FASTQSequence> (DistanceMatrixScoreFunction.h:11-105)."""
    scoreMatrix: np.ndarray = field(default_factory=lambda: SMRTDistanceMatrix.copy())
    ins: int = 5
    del_: int = 5
    affineOpen: int = 0
    affineExtend: int = 0
    kind: int = FN_DISTANCE
    substitutionPrior: int = 0      # BaseScoreFunction.h:8-9 (only IDSScoreFunction reads them)
    globalDeletionPrior: int = 0

    def c_struct(self) -> capi.ScoreFn:
        s = capi.ScoreFn()
        m = np.ascontiguousarray(self.scoreMatrix, dtype=np.int32).reshape(25)
        for i in range(25):
            s.M[i] = int(m[i])
        s.ins, s.del_, s.affineOpen, s.affineExtend, s.kind = self.ins, self.del_, self.affineOpen, self.affineExtend, self.kind
        s.substitutionPrior, s.globalDeletionPrior = self.substitutionPrior, self.globalDeletionPrior
        return s


@dataclass
class QualityValueScoreFunction(DistanceMatrixScoreFunction):
    """QualityValueScoreFunction<DNASequence,FASTQSequence> (QualityValueScoreFunction.h:9-85): match cost is
    QVDistanceMatrix[q][t] * qual[q]; gap costs are the constants ins/del.  scoreMatrix is only used by the
    stats pass, exactly as blasr rescoring always uses the distance function (Blasr.cpp:875)."""
    kind: int = FN_QUALITY


@dataclass
class IDSScoreFunction(DistanceMatrixScoreFunction):
    """IDSScoreFunction<DNASequence,FASTQSequence> (IDSScoreFunction.h:21-139): Match is 0 on equal raw bytes, else
    substitutionQV[q] when substitutionTag[q] is the target base, else substitutionPrior; Insertion is insertionQV[q];
    Deletion is deletionQV[q] when deletionTag[q] is the target base, else globalDeletionPrior (the constant del
    without deletion tracks).  Needs batch.insQV / subQV / subTag (delQV + delTag optional).  scoreMatrix, ins, del
    are used by the boundary rows and by the stats pass, which always rescores with the distance function
    (Blasr.cpp:875)."""
    kind: int = FN_IDS
    substitutionPrior: int = 20     # IDSScoreFunction.h:30-31
    globalDeletionPrior: int = 13


@dataclass
class JobBatch:
    """Structure-of-arrays batch (bgpu_batch)."""
    q: np.ndarray
    qOff: np.ndarray
    t: np.ndarray
    tOff: np.ndarray
    guide: Optional[np.ndarray] = None      # (nBlocksTotal, 3) uint32 {qPos,tPos,length}
    guideOff: Optional[np.ndarray] = None
    qual: Optional[np.ndarray] = None
    band: Optional[np.ndarray] = None
    # rich QV tracks of FASTQSequence (FASTQSequence.h:19-26), parallel to q; read by IDSScoreFunction only
    insQV: Optional[np.ndarray] = None
    delQV: Optional[np.ndarray] = None
    subQV: Optional[np.ndarray] = None
    delTag: Optional[np.ndarray] = None
    subTag: Optional[np.ndarray] = None

    TRACKS = ("insQV", "delQV", "subQV", "delTag", "subTag")

    @property
    def n(self) -> int:
        return len(self.qOff) - 1

    @staticmethod
    def from_lists(qs: Sequence[bytes], ts: Sequence[bytes], guides: Optional[Sequence[np.ndarray]] = None,
                   quals: Optional[Sequence[np.ndarray]] = None, bands: Optional[Sequence[int]] = None) -> "JobBatch":
        def cat(seqs):
            off = np.zeros(len(seqs) + 1, dtype=np.uint64)
            if len(seqs):
                off[1:] = np.cumsum([len(s) for s in seqs])
            data = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
            return data, off
        q, qOff = cat(qs)
        t, tOff = cat(ts)
        g = gOff = None
        if guides is not None:
            gOff = np.zeros(len(guides) + 1, dtype=np.uint64)
            gOff[1:] = np.cumsum([len(x) for x in guides])
            nz = [np.asarray(x, dtype=np.uint32).reshape(-1, 3) for x in guides]
            g = np.concatenate(nz, axis=0) if nz else np.zeros((0, 3), np.uint32)
        ql = None
        if quals is not None:
            ql = np.concatenate([np.asarray(x, dtype=np.uint8) for x in quals]) if len(quals) else np.zeros(0, np.uint8)
        bd = np.asarray(bands, dtype=np.int32) if bands is not None else None
        return JobBatch(q, qOff, t, tOff, g, gOff, ql, bd)

    def slice(self, idx: Sequence[int]) -> "JobBatch":
        qs = [self.q[int(self.qOff[i]):int(self.qOff[i + 1])].tobytes() for i in idx]
        ts = [self.t[int(self.tOff[i]):int(self.tOff[i + 1])].tobytes() for i in idx]
        gs = [self.guide[int(self.guideOff[i]):int(self.guideOff[i + 1])] for i in idx] if self.guide is not None else None
        qv = [self.qual[int(self.qOff[i]):int(self.qOff[i + 1])] for i in idx] if self.qual is not None else None
        bd = [int(self.band[i]) for i in idx] if self.band is not None else None
        out = JobBatch.from_lists(qs, ts, gs, qv, bd)
        for name in JobBatch.TRACKS:
            a = getattr(self, name)
            if a is not None:
                parts = [a[int(self.qOff[i]):int(self.qOff[i + 1])] for i in idx]
                setattr(out, name, np.concatenate(parts) if parts else np.zeros(0, np.uint8))
        return out


CIGAR_CHARS = "MIDNSHP=X"


def cigar_string(ops) -> str:
    """BAM-packed ops -> the text CigarOpsToString prints (SAMPrinter.h:318-327)."""
    return "".join(f"{int(o) >> 4}{CIGAR_CHARS[int(o) & 15]}" for o in ops)


@dataclass
class Alignment:
    """The fields of the reference's Alignment the DP fills (datastructures/alignment/Alignment.h:17-41)."""
    status: int
    score: int
    qPos: int
    tPos: int
    nCells: int
    blocks: np.ndarray                  # (n,3) uint32
    gaps: List[List[tuple]]             # gaps[i] = [(seq, length), ...]; seq 0 = Gap::Query, 1 = Gap::Target
    nMatch: int = 0
    nMismatch: int = 0
    nIns: int = 0
    nDel: int = 0
    pctSimilarity: float = 0.0
    statsScore: int = 0


def pack_guide(guide: np.ndarray, guideOff: np.ndarray):
    """bgpu_batch::guidePacked: per block (gap to the previous block's end in q, in t, length) as three bytes; blocks that do
    not fit go to the side list {block, dq, dt, length}.  Returns (packed uint8 (3n,), wide uint32 (m, 4))."""
    g = np.ascontiguousarray(guide, np.uint32).reshape(-1, 3).astype(np.int64)
    n = len(g)
    if n == 0:
        return np.zeros(0, np.uint8), np.zeros((0, 4), np.uint32)
    first = np.zeros(n, bool)
    off = np.asarray(guideOff, np.int64)
    first[off[:-1][off[:-1] < n]] = True
    qe = np.zeros(n, np.int64); te = np.zeros(n, np.int64)
    qe[1:] = g[:-1, 0] + g[:-1, 2]; te[1:] = g[:-1, 1] + g[:-1, 2]
    qe[first] = 0; te[first] = 0
    dq, dt, ln = g[:, 0] - qe, g[:, 1] - te, g[:, 2]
    wide = (dq < 0) | (dq >= 255) | (dt < 0) | (dt >= 255) | (ln >= 255)
    packed = np.stack([dq, dt, ln], axis=1)
    packed[wide] = 255
    w = np.stack([np.flatnonzero(wide), dq[wide], dt[wide], ln[wide]], axis=1).astype(np.int64)
    return packed.astype(np.uint8).reshape(-1), (w & 0xffffffff).astype(np.uint32)


class BatchResult:
    def __init__(self, results: np.ndarray, blocks: np.ndarray, gapCounts: np.ndarray, gaps: np.ndarray, timing=None, runs=None):
        self.results, self.blocks, self.gapCounts, self.gaps, self.timing, self.runs = results, blocks, gapCounts, gaps, timing, runs

    def __len__(self):
        return len(self.results)

    def alignment(self, i: int) -> Alignment:
        r = self.results[i]
        if self.runs is not None:
            # compact results: the path as runs (type << 30 | length); expanded the way blasr_gpu::RefineBatch::Store does
            lo = int(r["blockOff"]) + int(r["gapOff"])
            run = self.runs[lo:lo + int(r["nBlocks"]) + int(r["nGaps"])].astype(np.int64)
            ty, ln = run >> 30, run & 0x3fffffff
            qa = np.cumsum(np.where(ty != 2, ln, 0)) - np.where(ty != 2, ln, 0)
            ta = np.cumsum(np.where(ty != 1, ln, 0)) - np.where(ty != 1, ln, 0)
            d = ty == 0
            blocks = np.stack([qa[d], ta[d], ln[d]], axis=1).astype(np.uint32) if d.any() else np.zeros((0, 3), np.uint32)
            gaps = [[] for _ in range(int(r["nGapLists"]))]
            bidx = np.cumsum(d)
            for k in np.flatnonzero(~d):
                gaps[int(bidx[k])].append((1 if ty[k] == 1 else 0, int(ln[k])))
            return Alignment(int(r["status"]), int(r["score"]), int(r["qPos"]), int(r["tPos"]), int(r["nCells"]), blocks, gaps,
                             int(r["nMatch"]), int(r["nMismatch"]), int(r["nIns"]), int(r["nDel"]), float(r["pctSimilarity"]),
                             int(r["statsScore"]))
        b = self.blocks[int(r["blockOff"]):int(r["blockOff"]) + int(r["nBlocks"])]
        blocks = np.stack([b["qPos"], b["tPos"], b["length"]], axis=1) if len(b) else np.zeros((0, 3), np.uint32)
        cnt = self.gapCounts[int(r["gapListOff"]):int(r["gapListOff"]) + int(r["nGapLists"])]
        g = self.gaps[int(r["gapOff"]):int(r["gapOff"]) + int(r["nGaps"])]
        gaps, p = [], 0
        for c in cnt:
            gaps.append([(int(x["seq"]), int(x["length"])) for x in g[p:p + int(c)]])
            p += int(c)
        return Alignment(int(r["status"]), int(r["score"]), int(r["qPos"]), int(r["tPos"]), int(r["nCells"]), blocks, gaps,
                         int(r["nMatch"]), int(r["nMismatch"]), int(r["nIns"]), int(r["nDel"]), float(r["pctSimilarity"]),
                         int(r["statsScore"]))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Aligner:
    """One GPU context (bgpu_ctx).  Raises BgpuError when no CUDA device is present: there is no CPU fallback."""

    def __init__(self, device: int = 0):
        self._lib = capi.lib()
        self._ctx = C.c_void_p()
        rc = self._lib.bgpu_create(C.byref(self._ctx), device)
        if rc != 0:
            self._ctx = None
            raise BgpuError(f"bgpu_create(device={device}) failed with {rc}"
                            + (" (no CUDA device: the refinement kernels cannot run)" if rc == capi.E_NO_DEVICE else ""))

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.bgpu_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def _err(self, rc: int, what: str):
        raise BgpuError(f"{what} failed with {rc}: {self._lib.bgpu_last_error(self._ctx).decode()}")

    # ---- low level: ticket API ----
    def submit(self, batch: JobBatch, fn: DistanceMatrixScoreFunction, algo: int, alignType: int = GLOBAL, band: int = 16,
               bndIns: int = 0, bndDel: int = 0, doStats: bool = True, statsAffine: Optional[bool] = None,
               affineKBand: Sequence[int] = (0, 0, 0, 0), compact: bool = False, packed: bool = False):
        """compact: results come back as run-length paths (bgpu_params.compactResults); packed: the guide goes over as three
        bytes per block (bgpu_batch.guidePacked; batch.guidePacked / guideWide are used when present, else packed here)."""
        keep = dict(q=np.ascontiguousarray(batch.q, np.uint8), qOff=np.ascontiguousarray(batch.qOff, np.uint64),
                    t=np.ascontiguousarray(batch.t, np.uint8), tOff=np.ascontiguousarray(batch.tOff, np.uint64))
        n = batch.n
        if algo in (GUIDED, AFFINE_GUIDED):
            if batch.guide is None:
                raise ValueError("guided aligners need batch.guide")
            keep["guide"] = np.ascontiguousarray(batch.guide, np.uint32)
            keep["guideOff"] = np.ascontiguousarray(batch.guideOff, np.uint64)
            if packed:
                gp = getattr(batch, "guidePacked", None)
                gw = getattr(batch, "guideWide", None)
                if gp is None:
                    gp, gw = pack_guide(batch.guide, batch.guideOff)
                keep["guidePacked"] = np.ascontiguousarray(gp, np.uint8)
                keep["guideWide"] = np.ascontiguousarray(gw, np.uint32).reshape(-1, 4)
        if batch.qual is not None and fn.kind == FN_QUALITY:      # only QualityValueScoreFunction reads the QV track
            keep["qual"] = np.ascontiguousarray(batch.qual, np.uint8)
        if batch.band is not None:
            keep["band"] = np.ascontiguousarray(batch.band, np.int32)
        for name in JobBatch.TRACKS:
            if getattr(batch, name, None) is not None:
                keep[name] = np.ascontiguousarray(getattr(batch, name), np.uint8)
        t_ref = getattr(batch, "tRefOff", None)                    # targets = windows of the reference set with set_reference()
        if t_ref is not None:
            keep["tRefOff"] = np.ascontiguousarray(t_ref, np.uint64)
            if getattr(batch, "tRefRc", None) is not None:
                keep["tRefRc"] = np.ascontiguousarray(batch.tRefRc, np.uint8)
        b = capi.Batch(n, _ptr(keep["q"]), _ptr(keep["qOff"]), None if t_ref is not None else _ptr(keep["t"]), _ptr(keep["tOff"]), _ptr(keep.get("qual")),
                       _ptr(keep.get("guide")) if not packed else None, _ptr(keep.get("guideOff")), _ptr(keep.get("band")),
                       *[_ptr(keep.get(name)) for name in JobBatch.TRACKS],
                       _ptr(keep.get("guidePacked")), _ptr(keep.get("guideWide")), len(keep["guideWide"]) if "guideWide" in keep else 0,
                       _ptr(keep.get("tRefOff")), _ptr(keep.get("tRefRc")))
        if statsAffine is None:
            statsAffine = algo == AFFINE_GUIDED
        p = capi.Params(algo, alignType, band, bndIns, bndDel, int(doStats), int(statsAffine), *[int(x) for x in affineKBand], int(compact))
        f = fn.c_struct()
        tk = C.c_void_p()
        rc = self._lib.bgpu_submit(self._ctx, C.byref(f), C.byref(p), C.byref(b), C.byref(tk))
        if rc != 0:
            self._err(rc, "bgpu_submit")
        return tk, n

    def collect(self, ticket, copy: bool = False) -> BatchResult:
        """Blocks until the ticket is done.  With copy=False the block / gap arrays are zero-copy views of the
        library's pinned result arena and stay valid until release(ticket)."""
        tk, n = ticket
        res = np.zeros(n, dtype=capi.RESULT_DTYPE)
        arena = capi.Arena()
        rc = self._lib.bgpu_collect(self._ctx, tk, res.ctypes.data_as(C.c_void_p), C.byref(arena))
        if rc != 0:
            self._err(rc, "bgpu_collect")

        def view(ptr, count, dt):
            if not count:
                return np.zeros(0, dtype=dt)
            buf = (C.c_ubyte * (int(count) * dt.itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dt)
            return a.copy() if copy else a
        if arena.runs:
            return BatchResult(res, None, None, None, self.timing(ticket), runs=view(arena.runs, arena.nRuns, np.dtype("<u4")))
        return BatchResult(res, view(arena.blocks, arena.nBlocks, capi.BLOCK_DTYPE),
                           view(arena.gapCounts, arena.nGapLists, np.dtype("<u4")),
                           view(arena.gaps, arena.nGaps, capi.GAP_DTYPE), self.timing(ticket))

    def cigar(self, ticket):
        """SAM CIGAR core of every alignment of a collected guided ticket (bgpu_cigar; SAMPrinter.h:203-293): returns
        (ops, off) -- BAM-packed uint32 ops (length << 4 | code) and nJobs+1 offsets, copied out of the library's arena."""
        tk, n = ticket
        ops, off = C.c_void_p(), C.c_void_p()
        rc = self._lib.bgpu_cigar(self._ctx, tk, C.byref(ops), C.byref(off))
        if rc != 0:
            self._err(rc, "bgpu_cigar")
        o = np.frombuffer((C.c_ubyte * (8 * (n + 1))).from_address(off.value), dtype="<u8").copy()
        tot = int(o[-1])
        c = np.frombuffer((C.c_ubyte * (4 * tot)).from_address(ops.value), dtype="<u4").copy() if tot else np.zeros(0, "<u4")
        return c, o

    def cigar_clipped(self, ticket, clips=None, tStrand=None):
        """CreateCIGARString (SAMPrinter.h:345-400): clips (n, 4) uint32 = hard prefix, soft prefix, soft suffix, hard suffix per job
        (None = -clipping none), tStrand (n,) uint8.  Returns (ops, off) like cigar()."""
        tk, n = ticket
        ops, off = C.c_void_p(), C.c_void_p()
        cl = np.ascontiguousarray(clips, np.uint32).reshape(n, 4) if clips is not None else None
        st = np.ascontiguousarray(tStrand, np.uint8) if tStrand is not None else None
        rc = self._lib.bgpu_cigar_clipped(self._ctx, tk, _ptr(cl), _ptr(st), C.byref(ops), C.byref(off))
        if rc != 0:
            self._err(rc, "bgpu_cigar_clipped")
        o = np.frombuffer((C.c_ubyte * (8 * (n + 1))).from_address(off.value), dtype="<u8").copy()
        tot = int(o[-1])
        c = np.frombuffer((C.c_ubyte * (4 * tot)).from_address(ops.value), dtype="<u4").copy() if tot else np.zeros(0, "<u4")
        return c, o

    def strings(self, ticket):
        """CreateAlignmentStrings (AlignmentUtils.h:390-533) of every alignment of a collected guided ticket (bgpu_strings):
        returns (text, align, query, off): three uint8 arrays and nJobs+1 offsets."""
        tk, n = ticket
        p = [C.c_void_p() for _ in range(4)]
        rc = self._lib.bgpu_strings(self._ctx, tk, *[C.byref(x) for x in p])
        if rc != 0:
            self._err(rc, "bgpu_strings")
        o = np.frombuffer((C.c_ubyte * (8 * (n + 1))).from_address(p[3].value), dtype="<u8").copy()
        tot = int(o[-1])
        arrs = [np.frombuffer((C.c_ubyte * tot).from_address(x.value), dtype=np.uint8).copy() if tot else np.zeros(0, np.uint8) for x in p[:3]]
        return arrs[0], arrs[1], arrs[2], o

    def rescore(self, ticket, fn, useAffinePenalty: bool = False) -> np.ndarray:
        """bgpu_rescore: ComputeAlignmentScore(alignment, q, t, fn, useAffinePenalty) (AlignmentUtils.h:127-169) of every alignment of
        a collected guided ticket under another score function -- the rescoring StoreMapQVs does with SMRTLogProbMatrix."""
        tk, n = ticket
        out = np.zeros(n, np.int32)
        f = fn.c_struct()
        rc = self._lib.bgpu_rescore(self._ctx, tk, C.byref(f), int(useAffinePenalty), _ptr(out))
        if rc != 0:
            self._err(rc, "bgpu_rescore")
        return out

    def rerun(self, ticket):
        rc = self._lib.bgpu_rerun(self._ctx, ticket[0])
        if rc != 0:
            self._err(rc, "bgpu_rerun")
        return self.timing(ticket)

    def timing(self, ticket) -> capi.Timing:
        tm = capi.Timing()
        self._lib.bgpu_timing_of(self._ctx, ticket[0], C.byref(tm))
        return tm

    def release(self, ticket):
        self._lib.bgpu_release(self._ctx, ticket[0])

    def trim(self):
        """Give the context's cached (idle) device / pinned slabs back to the driver."""
        self._lib.bgpu_trim(self._ctx)

    def int_peak(self):
        ops, mhz = C.c_double(), C.c_double()
        rc = self._lib.bgpu_measure_int_peak(self._ctx, C.byref(ops), C.byref(mhz))
        if rc != 0:
            self._err(rc, "bgpu_measure_int_peak")
        modes = (C.c_double * 4)()
        self._lib.bgpu_int_peak_modes(C.byref(modes))
        self.int_peak_modes = dict(zip(("add", "min", "mad", "add+mad"), [float(x) for x in modes]))
        return ops.value, mhz.value

    def _run(self, batch, fn, algo, **kw) -> BatchResult:
        tk = self.submit(batch, fn, algo, **kw)
        try:
            return self.collect(tk, copy=True)
        finally:
            self.release(tk)

    # ---- the reference's names ----
    def AffineGuidedAlign(self, batch: JobBatch, scoreFn, bandSize: int = 16, alignType: int = GLOBAL,
                          computeStats: bool = True) -> BatchResult:
        """AffineGuidedAlign.h:31; blasr defaults bandSize=16, ins=del=5, affineOpen=50, affineExtend=0, Global."""
        return self._run(batch, scoreFn, AFFINE_GUIDED, alignType=alignType, band=bandSize, doStats=computeStats,
                         statsAffine=True)

    def GuidedAlign(self, batch: JobBatch, scoreFn, bandSize: int = 10, alignType: int = GLOBAL,
                    computeStats: bool = True, statsAffine: bool = False) -> BatchResult:
        """GuidedAlign.h:278 (computeProb=false)."""
        return self._run(batch, scoreFn, GUIDED, alignType=alignType, band=bandSize, doStats=computeStats,
                         statsAffine=statsAffine)

    def KBandAlign(self, batch: JobBatch, scoreFn, ins: int, del_: int, k: int, alignType: int = GLOBAL,
                   computeStats: bool = False) -> BatchResult:
        """KBandAlign.h:75; ins/del are the boundary-cost *parameters*, the fill uses scoreFn.ins/del."""
        return self._run(batch, scoreFn, KBAND, alignType=alignType, band=k, bndIns=ins, bndDel=del_, doStats=computeStats,
                         statsAffine=False)

    def AffineKBandAlign(self, batch: JobBatch, matchMat, hpInsOpen: int, hpInsExtend: int, insOpen: int, insExtend: int,
                         del_: int, k: int, alignType: int = GLOBAL, computeStats: bool = False, scoreFn=None) -> BatchResult:
        """AffineKBandAlign.h:12 (argument order of Blasr.cpp:1067-1076): matchMat[5][5] (row = query), the five int gap
        parameters, k.  Global and QueryFit.  scoreFn is only used by the optional stats pass."""
        fn = scoreFn if scoreFn is not None else DistanceMatrixScoreFunction()
        fn = DistanceMatrixScoreFunction(np.asarray(matchMat, np.int32).copy(), fn.ins, fn.del_, fn.affineOpen, fn.affineExtend)
        return self._run(batch, fn, AFFINE_KBAND, alignType=alignType, band=k, bndDel=del_, doStats=computeStats,
                         statsAffine=False, affineKBand=(hpInsOpen, hpInsExtend, insOpen, insExtend))

    def set_reference(self, bases) -> None:
        """bgpu_set_reference: the genome later batches take their targets from (JobBatch.tRefOff / tRefRc), resident on
        this aligner's device and shared by every context on it.  None / empty frees it."""
        a = np.ascontiguousarray(bases if bases is not None else np.zeros(0, np.uint8), np.uint8)
        rc = self._lib.bgpu_set_reference(self._ctx, a.ctypes.data_as(C.c_void_p) if len(a) else None, len(a))
        if rc != 0:
            self._err(rc, "bgpu_set_reference")

    def set_suffix_array(self, index, startPosTable=None, endPosTable=None, lookupPrefixLength: int = 0) -> None:
        """bgpu_set_suffix_array: the members of the reference's SuffixArray the anchoring reads (index, startPosTable,
        endPosTable, lookupPrefixLength), resident on this aligner's device next to the genome of set_reference()."""
        ix = np.ascontiguousarray(index if index is not None else np.zeros(0, np.uint32), np.uint32)
        st = None if startPosTable is None else np.ascontiguousarray(startPosTable, np.uint32)
        en = None if endPosTable is None else np.ascontiguousarray(endPosTable, np.uint32)
        rc = self._lib.bgpu_set_suffix_array(self._ctx, _ptr(ix) if len(ix) else None, len(ix), _ptr(st), _ptr(en), lookupPrefixLength)
        if rc != 0:
            self._err(rc, "bgpu_set_suffix_array")

    def MapReadToGenome(self, reads, readOff, minPrefixMatchLength: int = 8, minMatchLength: int = 12, expand: int = 0,
                        useLookupTable: bool = True, maxAnchorsPerPosition: int = 1000, advanceExactMatches: int = 0,
                        maxLCPLength: int = 0, stopMappingOnceUnique: bool = True, removeEncompassedMatches: bool = False,
                        subreadStart=None, subreadEnd=None, copy: bool = True):
        """MapReadToGenome (MapBySuffixArray.h:209-309) for every read of the batch, blasr's defaults (MappingParameters.h):
        returns (matchOff[n + 1], matches) with matches a MATCH_DTYPE array (t, q, l), read i owning
        matches[matchOff[i]:matchOff[i + 1]] in the reference's matchPosList order.  Pass reverse complements as reads.
        copy=False returns a view of the library's pinned result buffer (valid until the next call on this aligner)."""
        reads = np.ascontiguousarray(reads, np.uint8)
        readOff = np.ascontiguousarray(readOff, np.uint64)
        n = len(readOff) - 1
        ss = None if subreadStart is None else np.ascontiguousarray(subreadStart, np.uint32)
        se = None if subreadEnd is None else np.ascontiguousarray(subreadEnd, np.uint32)
        p = capi.AnchorParams(minPrefixMatchLength, minMatchLength, expand, int(useLookupTable), maxAnchorsPerPosition, advanceExactMatches,
                              maxLCPLength, int(stopMappingOnceUnique), int(removeEncompassedMatches))
        off = np.zeros(n + 1, np.uint64)
        out = C.c_void_p()
        rc = self._lib.bgpu_map_reads(self._ctx, C.byref(p), _ptr(reads), _ptr(readOff), n, _ptr(ss), _ptr(se), _ptr(off), C.byref(out))
        if rc != 0:
            self._err(rc, "bgpu_map_reads")
        total = int(off[-1])
        if total == 0:
            return off, np.zeros(0, capi.MATCH_DTYPE)
        buf = (C.c_uint8 * (12 * total)).from_address(out.value)
        m = np.frombuffer(buf, dtype=capi.MATCH_DTYPE)          # the library's pinned buffer: valid until the next call on this aligner
        return off, (m.copy() if copy else m)

    def map_timing(self):
        """(ms of the search kernel, ms of count + scan + emit, positions searched, H2D bytes, D2H bytes) of the last MapReadToGenome."""
        ms = (C.c_double * 2)()
        pos, h, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = self._lib.bgpu_map_timing(self._ctx, C.byref(ms), C.byref(pos), C.byref(h), C.byref(d))
        if rc != 0:
            self._err(rc, "bgpu_map_timing")
        return ms[0], ms[1], pos.value, h.value, d.value

    def map_rerun(self):
        rc = self._lib.bgpu_map_rerun(self._ctx)
        if rc != 0:
            self._err(rc, "bgpu_map_rerun")

    def SDPAlign(self, batch: JobBatch, scoreFn, wordSize: int = 11, sdpIns: int = 5, sdpDel: int = 10, indelRate: float = 0.30,
                 alignType: int = capi.LOCAL, detailedAlignment: bool = True, extendFrontByLocalAlignment: bool = False,
                 sdpPrefixLength: int = 50, recurse: int = 2, noRecurseUnder: int = 1000, maxMatchesPerPosition: int = 0):
        """SDPAlign.h:95-107 with MappingParameters' defaults (the call of Blasr.cpp:1716-1722): the guide the refinement
        receives.  Returns (results, blocks): results[i] holds status / qPos / tPos / nBlocks / blockOff, blocks is the
        concatenated Block array (positions relative to qPos / tPos, as Alignment::blocks)."""
        keep = {"q": np.ascontiguousarray(batch.q, np.uint8), "qOff": np.ascontiguousarray(batch.qOff, np.uint64),
                "t": np.ascontiguousarray(batch.t, np.uint8), "tOff": np.ascontiguousarray(batch.tOff, np.uint64)}
        b = capi.Batch(batch.n, _ptr(keep["q"]), _ptr(keep["qOff"]), _ptr(keep["t"]), _ptr(keep["tOff"]))
        p = capi.SdpParams(wordSize, sdpIns, sdpDel, float(indelRate), alignType, int(detailedAlignment), int(extendFrontByLocalAlignment),
                           sdpPrefixLength, recurse, noRecurseUnder, maxMatchesPerPosition)
        f = scoreFn.c_struct()
        res = np.zeros(batch.n, dtype=capi.RESULT_DTYPE)
        arena = capi.Arena()
        rc = self._lib.bgpu_sdp_align(self._ctx, C.byref(f), C.byref(p), C.byref(b), res.ctypes.data_as(C.c_void_p), C.byref(arena))
        if rc != 0:
            self._err(rc, "bgpu_sdp_align")
        if arena.nBlocks:
            buf = (C.c_ubyte * (int(arena.nBlocks) * capi.BLOCK_DTYPE.itemsize)).from_address(arena.blocks)
            blocks = np.frombuffer(buf, dtype=capi.BLOCK_DTYPE).copy()
        else:
            blocks = np.zeros(0, dtype=capi.BLOCK_DTYPE)
        return res, blocks

    def SWAlign(self, batch: JobBatch, scoreFn, alignType: int = capi.LOCAL, computeStats: bool = False) -> BatchResult:
        """SWAlign.h:18."""
        return self._run(batch, scoreFn, SW, alignType=alignType, band=0, doStats=computeStats, statsAffine=False)
