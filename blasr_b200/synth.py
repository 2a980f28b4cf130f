"""Seeded synthetic workloads for the refinement path.

The reference's own simulator needs HDF5 (simulator/Alchemy.cpp:15-16), so inputs are generated here:
uniform-ACGT target windows, PacBio-like reads (error split ins 55 % / del 35 % / sub 10 %), QV tracks
~ clamp(N(12,4),1,93), and a guide per pair.  The guide is the block list a detailed SDPAlign hands to
RefineAlignment (Blasr.cpp:1716-1722, :850-866): every diagonal run of the read-to-window alignment,
first block at (0,0), last block ending at the sequence ends.  Here it is derived from the simulated
edit script instead of running SDPAlign, so the generator has no dependency on the reference or the oracle.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from .align import JobBatch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _chunk_worker(args):
    seed, ci, lens, err, split, with_qual, anchor, min_block, n_rate = args
    return _simulate_chunk(np.random.default_rng([seed, ci]), lens, err, split, with_qual, anchor, min_block, n_rate)


def simulate_pairs(n_jobs: int, len_lo: int, len_hi: int, err: float = 0.15, split=(0.55, 0.35, 0.10), seed: int = 1,
                   bands: Optional[Sequence[int]] = None, with_qual: bool = False, anchor: int = 12,
                   min_block: int = 1, n_rate: float = 0.0, chunk_jobs: int = 512, workers: int = 0) -> JobBatch:
    """n_jobs (query, window, guide) triples; window lengths ~ U[len_lo, len_hi].

    min_block > 1 drops shorter interior guide blocks (anchor-only guides with real gaps between blocks).
    Chunks of chunk_jobs jobs are seeded by (seed, chunk index), so the result does not depend on `workers`
    (0 = all host cores for large requests, in-process for small ones).
    """
    rng = np.random.default_rng(seed)
    lens = rng.integers(len_lo, len_hi + 1, size=n_jobs, dtype=np.int64)
    lens = np.maximum(lens, 2 * anchor + 2)
    tasks = [(seed, ci, lens[i0:i0 + chunk_jobs], err, split, with_qual, anchor, min_block, n_rate)
             for ci, i0 in enumerate(range(0, n_jobs, chunk_jobs))]
    if workers == 0:
        workers = min(len(tasks), os.cpu_count() or 1) if int(lens.sum()) > 20_000_000 else 1
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_chunk_worker, tasks, chunksize=1)
    else:
        parts = [_chunk_worker(t) for t in tasks]
    if not parts:
        z = np.zeros(1, np.uint64)
        return JobBatch(np.zeros(0, np.uint8), z, np.zeros(0, np.uint8), z.copy(), np.zeros((0, 3), np.uint32), z.copy(),
                        np.zeros(0, np.uint8) if with_qual else None, None)
    q = np.concatenate([p[0] for p in parts]); t = np.concatenate([p[2] for p in parts])
    g = np.concatenate([p[4] for p in parts], axis=0)

    def cat_off(k):
        out = [np.zeros(1, np.uint64)]
        base = 0
        for p in parts:
            out.append(p[k][1:] + np.uint64(base)); base += int(p[k][-1])
        return np.concatenate(out)
    qOff, tOff, gOff = cat_off(1), cat_off(3), cat_off(5)
    qual = np.concatenate([p[6] for p in parts]) if with_qual else None
    band = None
    if bands is not None:
        band = np.asarray(bands, dtype=np.int32)[rng.integers(0, len(bands), size=n_jobs)]
    return JobBatch(q, qOff, t, tOff, g, gOff, qual, band)


def _simulate_chunk(rng, lens, err, split, with_qual, anchor, min_block, n_rate):
    n = len(lens)
    T = int(lens.sum())
    tOff = np.zeros(n + 1, np.int64); tOff[1:] = np.cumsum(lens)
    tcode = rng.integers(0, 4, size=T, dtype=np.uint8)
    r = rng.integers(0, 65536, size=T, dtype=np.uint16).astype(np.int32)
    p_ins, p_del, p_sub = (int(err * s * 65536) for s in split)
    is_sub = r < p_sub
    is_del = (r >= p_sub) & (r < p_sub + p_del)
    is_ins = rng.integers(0, 65536, size=T, dtype=np.uint16).astype(np.int32) < p_ins   # inserted base *before* position j
    # protect `anchor` positions at both ends of every window
    pos = np.arange(T, dtype=np.int64) - np.repeat(tOff[:-1], lens)
    prot = (pos < anchor) | (pos >= np.repeat(lens, lens) - anchor)
    is_sub &= ~prot; is_del &= ~prot; is_ins &= ~prot
    kept = ~is_del
    sub_to = (tcode + 1 + rng.integers(0, 3, size=T, dtype=np.uint8)) % 4
    qbase = np.where(is_sub, sub_to, tcode).astype(np.uint8)
    # interleave insertion slots and base slots
    vals = np.empty(2 * T, np.uint8); vals[0::2] = rng.integers(0, 4, size=T, dtype=np.uint8); vals[1::2] = qbase
    mask = np.empty(2 * T, bool); mask[0::2] = is_ins; mask[1::2] = kept
    qcode = vals[mask]
    emitted = is_ins.astype(np.int64) + kept.astype(np.int64)
    qcum = np.cumsum(emitted)                       # q bases emitted through position j (inclusive)
    qpos_of_base = qcum - 1                         # q index of base j when kept
    qOff = np.zeros(n + 1, np.int64); qOff[1:] = qcum[tOff[1:] - 1]
    # guide blocks: maximal runs of kept positions with no insertion in between
    prev_kept = np.empty(T, bool); prev_kept[0] = False; prev_kept[1:] = kept[:-1]
    job_start = np.zeros(T, bool); job_start[tOff[:-1]] = True
    start = kept & (job_start | ~prev_kept | is_ins)
    sidx = np.flatnonzero(start)
    # run length = distance to the next break (next start, next deleted position, or window end)
    brk = start | ~kept | job_start
    bidx = np.flatnonzero(brk)
    nxt = np.searchsorted(bidx, sidx, side="right")
    end = np.where(nxt < len(bidx), bidx[np.minimum(nxt, len(bidx) - 1)], T)
    length = end - sidx
    job_of = np.searchsorted(tOff, sidx, side="right") - 1
    g_t = sidx - tOff[job_of]
    g_q = qpos_of_base[sidx] - qOff[job_of]
    if min_block > 1:
        first = np.zeros(len(sidx), bool); last = np.zeros(len(sidx), bool)
        chg = np.flatnonzero(np.diff(job_of)) + 1
        first[0] = True; first[chg] = True; last[-1] = True; last[chg - 1] = True
        keep = (length >= min_block) | first | last
        job_of, g_t, g_q, length = job_of[keep], g_t[keep], g_q[keep], length[keep]
    guide = np.stack([g_q, g_t, length], axis=1).astype(np.uint32)
    gOff = np.zeros(n + 1, np.int64); gOff[1:] = np.cumsum(np.bincount(job_of, minlength=n))
    tb = _ACGT[tcode]; qb = _ACGT[qcode]
    if n_rate > 0:
        tb = tb.copy(); qb = qb.copy()
        tb[rng.random(T) < n_rate] = ord("N"); qb[rng.random(len(qb)) < n_rate] = ord("N")
    qual = None
    if with_qual:
        qual = np.clip(np.rint(rng.normal(12, 4, size=len(qb))), 1, 93).astype(np.uint8)
    return (qb, qOff.astype(np.uint64), tb, tOff.astype(np.uint64), guide, gOff.astype(np.uint64), qual)


def batch_cells_estimate(batch: JobBatch, band: int) -> int:
    """Rough nCells (rows x (2*band+1)), for sizing only; the exact count comes back from the library."""
    return int((batch.qOff[-1] - batch.qOff[0])) * (2 * band + 1)


_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


def simulate_genome(n: int, seed: int = 1, repeat_families: int = 8, unit_lo: int = 3000, unit_hi: int = 6000, copies: int = 3,
                    divergence: float = 0.03) -> np.ndarray:
    """Uniform ACGT with embedded repeat families (units copied `copies` times at `divergence`, SURVEY 8d C3 in miniature: a
    uniform-random genome gives every k-mer one candidate) and the trailing 'N' blasr's FASTA reader appends (FASTAReader.h:130)."""
    rng = np.random.default_rng(seed)
    g = _ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
    for _ in range(repeat_families):
        ln = int(rng.integers(unit_lo, unit_hi + 1))
        if 2 * ln >= n:
            continue
        unit = g[int(rng.integers(0, n - ln)):][:ln].copy()
        for _ in range(copies):
            u = unit.copy()
            mut = rng.random(ln) < divergence
            u[mut] = _ACGT[rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)]
            at = int(rng.integers(0, n - ln))
            g[at:at + ln] = u
    g[-1] = ord("N")
    return g


def simulate_reads(genome: np.ndarray, n_reads: int, length: int, err: float = 0.15, split=(0.55, 0.35, 0.10), seed: int = 1,
                   both_strands: bool = True):
    """Reads drawn from the genome with PacBio-like errors, odd reads reverse-complemented: (bases, readOff).
    With both_strands every read is followed by its reverse complement, the two MapReadToGenome calls of Blasr.cpp:2282-2296."""
    rng = np.random.default_rng(seed)
    n = len(genome)
    length = min(length, n - 1)
    starts = rng.integers(0, n - length, n_reads)
    base = genome[(starts[:, None] + np.arange(length)[None, :]).ravel()]
    T = len(base)
    r = rng.random(T)
    is_sub = r < err * split[2]
    is_del = (r >= err * split[2]) & (r < err * (split[2] + split[1]))
    is_ins = rng.random(T) < err * split[0]
    qb = np.where(is_sub, _ACGT[rng.integers(0, 4, T, dtype=np.uint8)], base)
    vals = np.empty(2 * T, np.uint8); vals[0::2] = _ACGT[rng.integers(0, 4, T, dtype=np.uint8)]; vals[1::2] = qb
    mask = np.empty(2 * T, bool); mask[0::2] = is_ins; mask[1::2] = ~is_del
    q = vals[mask]
    per = (is_ins.astype(np.int64) + (~is_del).astype(np.int64)).reshape(n_reads, length).sum(axis=1)
    off = np.zeros(n_reads + 1, np.int64); off[1:] = np.cumsum(per)
    reads = []
    for i in range(n_reads):
        rd = q[off[i]:off[i + 1]]
        if i & 1:
            rd = _COMP[rd[::-1]]
        reads.append(rd)
        if both_strands:
            reads.append(_COMP[rd[::-1]])
    ro = np.zeros(len(reads) + 1, np.uint64)
    ro[1:] = np.cumsum([len(x) for x in reads])
    return np.concatenate(reads), ro
