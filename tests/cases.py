"""Seeded case generators shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) parity tests."""
from __future__ import annotations

import numpy as np

from blasr_b200 import JobBatch, SMRTDistanceMatrix, synth
from . import oracle as O

ACGT = np.frombuffer(b"ACGT", np.uint8)


def job_arrays(b: JobBatch, i: int):
    q = b.q[int(b.qOff[i]):int(b.qOff[i + 1])]
    t = b.t[int(b.tOff[i]):int(b.tOff[i + 1])]
    g = b.guide[int(b.guideOff[i]):int(b.guideOff[i + 1])] if b.guide is not None else None
    qv = b.qual[int(b.qOff[i]):int(b.qOff[i + 1])] if b.qual is not None else None
    return q, t, g, qv


def job_tracks(b: JobBatch, i: int):
    lo, hi = int(b.qOff[i]), int(b.qOff[i + 1])
    return {k: (getattr(b, k)[lo:hi] if getattr(b, k, None) is not None else None) for k in JobBatch.TRACKS}


def add_ids_tracks(b: JobBatch, seed: int, with_del=True) -> JobBatch:
    """bas.h5-style rich QV tracks: QVs ~ clamp(N(12,4),1,93); tags are a random base
    (substitution) / a random base or 'N' (deletion), a few of them lower-case (the compares are raw bytes)."""
    rng = np.random.default_rng(seed)
    n = len(b.q)
    qv = lambda: np.clip(np.rint(rng.normal(12, 4, n)), 1, 93).astype(np.uint8)
    b.insQV, b.subQV = qv(), qv()
    st = ACGT[rng.integers(0, 4, n)].copy(); st[rng.random(n) < 0.05] |= 0x20
    b.subTag = st
    if with_del:
        b.delQV = qv()
        dt = ACGT[rng.integers(0, 4, n)].copy(); dt[rng.random(n) < 0.3] = ord("N"); dt[rng.random(n) < 0.05] |= 0x20
        b.delTag = dt
    return b


def drop_blocks(guide: np.ndarray, rng, p_drop: float, run: int = 1) -> np.ndarray:
    """Adversarial guide: delete runs of interior blocks (keeps first and last), leaving real gaps."""
    n = len(guide)
    keep = np.ones(n, bool)
    i = 1
    while i < n - 1:
        if rng.random() < p_drop:
            L = int(rng.integers(1, run + 1)); keep[i:min(n - 1, i + L)] = False; i += L
        else:
            i += 1
    return guide[keep]


def guided_batch(seed: int, n: int, lo: int, hi: int, err=0.15, adversarial=0.0, run=1, min_block=1, n_rate=0.0,
                 with_qual=False, lower=False) -> JobBatch:
    b = synth.simulate_pairs(n, lo, hi, err=err, seed=seed, min_block=min_block, n_rate=n_rate, with_qual=with_qual)
    rng = np.random.default_rng(seed + 1000)
    if adversarial > 0:
        gs = []
        for i in range(b.n):
            g = b.guide[int(b.guideOff[i]):int(b.guideOff[i + 1])]
            gs.append(drop_blocks(g, rng, adversarial, run))
        off = np.zeros(b.n + 1, np.uint64); off[1:] = np.cumsum([len(g) for g in gs])
        b.guide = np.concatenate(gs, axis=0); b.guideOff = off
    if lower:   # soft-masked (lower-case) bases must behave like upper-case
        m = rng.random(len(b.q)) < 0.3
        b.q = b.q.copy(); b.q[m] |= 0x20
        m = rng.random(len(b.t)) < 0.3
        b.t = b.t.copy(); b.t[m] |= 0x20
    return b


def random_pair(rng, lo, hi, err=0.2, n_rate=0.0):
    L = int(rng.integers(lo, hi + 1))
    t = rng.integers(0, 4, L)
    q = []
    for c in t:
        r = rng.random()
        if r < err * 0.4:
            q.append(c); q.append(rng.integers(0, 4))
        elif r < err * 0.75:
            pass
        elif r < err:
            q.append((c + 1 + rng.integers(0, 3)) % 4)
        else:
            q.append(c)
    if not q:
        q = [0]
    qb = ACGT[np.asarray(q, dtype=np.int64)].copy(); tb = ACGT[t].copy()
    if n_rate:
        qb[rng.random(len(qb)) < n_rate] = ord("N"); tb[rng.random(len(tb)) < n_rate] = ord("N")
    return qb, tb


def compare(a: dict, r: dict, fields=None):
    """Names of the fields that differ between two oracle-style result dicts."""
    bad = []
    for k in (fields or [x for x in a if x not in ("blocks", "gaps")]):
        if k in a and k in r and a[k] != r[k]:
            bad.append((k, a[k], r[k]))
    if not np.array_equal(a["blocks"], r["blocks"]):
        bad.append(("blocks", len(a["blocks"]), len(r["blocks"])))
    if a["gaps"] != r["gaps"]:
        bad.append(("gaps", len(a["gaps"]), len(r["gaps"])))
    return bad


def gpu_to_dict(res, i: int) -> dict:
    al = res.alignment(i)
    return dict(status=al.status, score=al.score, qPos=al.qPos, tPos=al.tPos, nCells=al.nCells, nMatch=al.nMatch,
                nMismatch=al.nMismatch, nIns=al.nIns, nDel=al.nDel, pctSimilarity=np.float32(al.pctSimilarity),
                statsScore=al.statsScore, nBlocks=len(al.blocks), nGapLists=len(al.gaps),
                nGaps=sum(len(g) for g in al.gaps), blocks=al.blocks, gaps=al.gaps)


GPU_FIELDS = ["status", "score", "qPos", "tPos", "nCells", "nMatch", "nMismatch", "nIns", "nDel", "pctSimilarity",
              "statsScore", "nBlocks", "nGapLists", "nGaps"]


def oracle_batch(which: str, b: JobBatch, fn: O.OrcScoreFn, algo: int, alignType: int, band, bndIns=0, bndDel=0,
                 statsAffine=0, doStats=1):
    out = []
    for i in range(b.n):
        q, t, g, qv = job_arrays(b, i)
        bd = int(band[i]) if hasattr(band, "__len__") else int(band)
        j, keep = O.make_job(algo, alignType, bd, q, t, g, qv, bndIns, bndDel, doStats, statsAffine, tracks=job_tracks(b, i))
        out.append(O.align(which, fn, j))
    return out
