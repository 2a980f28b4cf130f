"""Helpers for the pipeline-derived job sets under tests/golden/jobs_*_small.bgj.gz (dumped from the unmodified
reference pipeline by baseline/make_golden_dumps.sh): run a dump group through an oracle library."""
from __future__ import annotations

import os

import numpy as np

from blasr_b200 import capi, jobdump
from . import cases, oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = {c: os.path.join(GOLDEN, f"jobs_{c}_small.bgj.gz") for c in ("c0", "c2", "c4")}


def oracle_group(which: str, g: jobdump.JobGroup, idx=None):
    fn = g.fn
    ofn = O.score_fn(fn.scoreMatrix, fn.ins, fn.del_, fn.affineOpen, fn.affineExtend)
    out = []
    for i in (range(g.batch.n) if idx is None else idx):
        q, t, gd, _ = cases.job_arrays(g.batch, i)
        if g.kind == capi.AFFINE_KBAND:
            j, keep = O.make_job(4, capi.GLOBAL, g.band, q, t, None, None, 0, g.extra[4], 1, 0, affineKBand=g.extra[:4])
        else:
            j, keep = O.make_job(g.kind, capi.GLOBAL, g.band, q, t, gd, None, 0, 0, 1, int(g.kind == capi.AFFINE_GUIDED))
        out.append(O.align(which, ofn, j))
    return out


def gpu_group(aligner, g: jobdump.JobGroup):
    if g.kind == capi.AFFINE_KBAND:
        return aligner.AffineKBandAlign(g.batch, g.fn.scoreMatrix, g.extra[0], g.extra[1], g.extra[2], g.extra[3], g.extra[4],
                                        g.band, alignType=capi.GLOBAL, computeStats=True, scoreFn=g.fn)
    if g.kind == capi.AFFINE_GUIDED:
        return aligner.AffineGuidedAlign(g.batch, g.fn, g.band)
    return aligner.GuidedAlign(g.batch, g.fn, g.band)
