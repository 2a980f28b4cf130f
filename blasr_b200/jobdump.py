"""Reader of refinement-job dumps: the exact (query slice, target window, guide blocks, band, parameters) tuples the
reference's call sites issue (RefineAlignment, alignment/Blasr.cpp:863-872; AlignSubstring, :1067-1076), written by the
instrumented twin of the reference program (baseline/job_dump.hpp documents the record layout).

`load(path)` returns one JobGroup per (kind, band, score parameters) combination, each holding a JobBatch that
`Aligner.submit` / `bench.py --workload dump:<file>` take as is.  Plain and gzip files are accepted.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass
from typing import List

import numpy as np

from .align import AFFINE_GUIDED, AFFINE_KBAND, GUIDED, DistanceMatrixScoreFunction, JobBatch

MAGIC = 0x314A4742
KIND_NAMES = {GUIDED: "GuidedAlign", AFFINE_GUIDED: "AffineGuidedAlign", AFFINE_KBAND: "AffineKBandAlign"}


@dataclass
class JobGroup:
    kind: int                      # capi.GUIDED / AFFINE_GUIDED / AFFINE_KBAND
    band: int
    fn: DistanceMatrixScoreFunction
    extra: tuple                   # AffineKBandAlign: (hpInsOpen, hpInsExtend, insOpen, insExtend, del)
    batch: JobBatch
    index: np.ndarray              # position of each job in the dump (dump order = call order of the reference)


def load(path: str) -> List[JobGroup]:
    raw = (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")).read()
    buf = np.frombuffer(raw, np.uint8)
    pos, n = 0, 0
    groups = {}
    while pos < len(buf):
        head = buf[pos:pos + 24].view("<u4")
        if int(head[0]) != MAGIC:
            raise ValueError(f"{path}: bad record magic at byte {pos}")
        kind, band = int(head[1].astype(np.int32)), int(head[2].astype(np.int32))
        qLen, tLen, nB = int(head[3]), int(head[4]), int(head[5])
        par = buf[pos + 24:pos + 24 + 34 * 4].view("<i4")
        pos += 24 + 34 * 4
        q = buf[pos:pos + qLen]; pos += qLen
        t = buf[pos:pos + tLen]; pos += tLen
        g = buf[pos:pos + 12 * nB].view("<u4").reshape(-1, 3); pos += 12 * nB
        key = (kind, band, par.tobytes())
        groups.setdefault(key, ([], [], [], []))
        qs, ts, gs, idx = groups[key]
        qs.append(q.tobytes()); ts.append(t.tobytes()); gs.append(g); idx.append(n)
        n += 1
    out = []
    for (kind, band, parb), (qs, ts, gs, idx) in groups.items():
        par = np.frombuffer(parb, "<i4")
        fn = DistanceMatrixScoreFunction(par[4:29].reshape(5, 5).astype(np.int32).copy(), int(par[0]), int(par[1]), int(par[2]), int(par[3]))
        b = JobBatch.from_lists(qs, ts, gs if kind != AFFINE_KBAND else None)
        out.append(JobGroup(kind, band, fn, tuple(int(x) for x in par[29:34]), b, np.asarray(idx, np.int64)))
    return out
