#!/usr/bin/env python3
"""baseline/make_data.py -- seeded synthetic genomes and PacBio-like reads for the pipeline-level runs.

    make_data.py c0   <dir> [--genome 4600000] [--reads 1000] [--len 10000]      configs[0]: E. coli-sized, unique genome
    make_data.py c2   <dir> [--genome 2000000] [--reads 200]  [--lo 10000 --hi 30000]   scaled configs[2]: repeat families, -bestn 10
    make_data.py c4   <dir> [--genome 3000000] [--contigs 3] [--len 300000]      scaled configs[4]: contigs for -alignContigs

Reads: windows of the genome, strand alternating, 15 % error split ins 55 % / del 35 % / sub 10 %.  Contigs (c4): windows at
0.5 % divergence (equal parts ins / del / sub).  The reference's own simulator needs HDF5 (simulator/Alchemy.cpp:15-16),
so the generator is ours; everything is determined by --seed.
"""
import argparse
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8)
COMP[list(b"ACGT")] = list(b"TGCA")


def write_fasta(path, records, width=60):
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            b = seq.tobytes()
            f.write(b"\n".join(b[i:i + width] for i in range(0, len(b), width)) + b"\n")


def mutate(rng, seq, err, p_ins, p_del, p_sub):
    """One pass over seq: each base is kept, substituted, deleted, or followed by an inserted base."""
    n = len(seq)
    r = rng.random(n)
    is_del = r < err * p_del
    is_sub = (r >= err * p_del) & (r < err * (p_del + p_sub))
    is_ins = (r >= err * (p_del + p_sub)) & (r < err)
    out = seq.copy()
    out[is_sub] = ACGT[(np.searchsorted(ACGT, seq[is_sub]) + rng.integers(1, 4, int(is_sub.sum()))) % 4]
    reps = np.ones(n, np.int64)
    reps[is_del] = 0
    reps[is_ins] = 2
    res = np.repeat(out, reps)
    # the second copy of an "insert" position becomes a random base
    idx = np.cumsum(reps)[is_ins] - 1
    res[idx] = ACGT[rng.integers(0, 4, len(idx))]
    return res


def genome_unique(rng, n):
    return ACGT[rng.integers(0, 4, n)]


def genome_with_repeats(rng, n):
    """Repeat families: units of 30-60 kb in 2..10 copies at ~3 % divergence over a random background."""
    g = genome_unique(rng, n)
    pos = 0
    fam = 0
    spans = []
    while True:
        unit = int(rng.integers(30000, 60001))
        copies = int(rng.integers(2, 11))
        copies = min(copies, int((n * 0.6 - pos) // (unit + 20000)))
        if copies < 2:
            break
        base = ACGT[rng.integers(0, 4, unit)]
        for _ in range(copies):
            pos += int(rng.integers(5000, 20000))
            c = mutate(rng, base, 0.03, 1 / 3, 1 / 3, 1 / 3)[:unit]
            g[pos:pos + len(c)] = c
            spans.append((pos, pos + len(c)))
            pos += len(c)
        fam += 1
    return g, spans


def sample_reads(rng, g, n_reads, lo, hi, err, spans=None, tag="read"):
    recs = []
    for i in range(n_reads):
        L = int(rng.integers(lo, hi + 1))
        if spans and i % 2 == 1:                       # half of the reads start inside a repeat copy
            a, b = spans[int(rng.integers(0, len(spans)))]
            s = int(rng.integers(a, max(a + 1, b - L // 2)))
        else:
            s = int(rng.integers(0, len(g) - L))
        s = min(s, len(g) - L)
        w = g[s:s + L]
        if i % 2 == 1:
            w = COMP[w[::-1]]
        r = mutate(rng, w, err, 0.55, 0.35, 0.10)
        recs.append((f"{tag}_{i}/{s}_{s + L}/{'-' if i % 2 else '+'}", r))
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c0", "c2", "c4"])
    ap.add_argument("dir")
    ap.add_argument("--genome", type=int, default=None)
    ap.add_argument("--reads", type=int, default=None)
    ap.add_argument("--len", type=int, default=None)
    ap.add_argument("--lo", type=int, default=10000)
    ap.add_argument("--hi", type=int, default=30000)
    ap.add_argument("--contigs", type=int, default=3)
    ap.add_argument("--err", type=float, default=0.15)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    rng = np.random.default_rng(a.seed)
    if a.config == "c0":
        g = genome_unique(rng, a.genome or 4600000)
        L = a.len or 10000
        reads = sample_reads(rng, g, a.reads or 1000, L, L, a.err)
    elif a.config == "c2":
        g, spans = genome_with_repeats(rng, a.genome or 2000000)
        reads = sample_reads(rng, g, a.reads or 200, a.lo, a.hi, a.err, spans)
    else:
        g = genome_unique(rng, a.genome or 3000000)
        L = a.len or 300000
        reads = []
        for i in range(a.contigs):
            s = int(rng.integers(0, len(g) - L))
            w = g[s:s + L]
            if i % 2 == 1:
                w = COMP[w[::-1]]
            reads.append((f"contig_{i}/{s}_{s + L}", mutate(rng, w, 0.005, 1 / 3, 1 / 3, 1 / 3)))
    write_fasta(os.path.join(a.dir, "genome.fa"), [("genome", g)])
    write_fasta(os.path.join(a.dir, "reads.fa"), reads)
    print(f"{a.config}: genome {len(g)} b, {len(reads)} reads -> {a.dir}")


if __name__ == "__main__":
    main()
