"""ctypes bindings of the anchoring checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

  ref_*  -> oracle/_ref/libblasr_ref_anchor.so   the unmodified MapReadToGenome / SuffixArray templates (oracle/ref_anchor.cpp)
  orc_*  -> oracle/liborc.so                      plain-C restatement (oracle/orc_anchor.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle as _o

REF_ANCHOR_PATH = os.path.join(_o.ROOT, "oracle", "_ref", "libblasr_ref_anchor.so")

# blasr's defaults: alignment/MappingParameters.h:243-244,278,309; AnchorParameters.h:26-42
DEFAULTS = dict(minPrefix=8, minMatch=12, expand=0, useLookup=1, maxAnchors=1000, advance=0, maxLCP=0, stopUnique=1, removeEncompassed=0)


def params(**kw) -> np.ndarray:
    d = dict(DEFAULTS)
    d.update(kw)
    return np.array([d["minPrefix"], d["minMatch"], d["expand"], d["useLookup"], d["maxAnchors"], d["advance"], d["maxLCP"],
                     d["stopUnique"], d["removeEncompassed"]], dtype=np.int32)


_ref = None
_orc = None


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_ANCHOR_PATH):
            _o.build()
        if not os.path.exists(REF_ANCHOR_PATH):
            return None
        L = C.CDLL(REF_ANCHOR_PATH)
        L.ref_sa_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_sa_lookup_table.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_map_read.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ref_map_read.restype = C.c_int64
        L.ref_map_reads.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_map_reads.restype = C.c_int64
        _ref = L
    return _ref


def orc():
    global _orc
    if _orc is None:
        L = _o._load("orc")
        L.orc_map_read.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_map_read.restype = C.c_int64
        _orc = L
    return _orc


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


PAD = 16   # BuildLookupTable reads up to lookupPrefixLength - 1 bytes past the text for the shortest suffixes (SuffixArray.h:227-228)


def padded(genome: np.ndarray) -> np.ndarray:
    """The genome with one readable byte behind it (the convention of bgpu_map_reads / orc_anchor.c)."""
    g = np.full(len(genome) + PAD, ord("N"), dtype=np.uint8)
    g[:len(genome)] = genome
    return g


class Index:
    """Genome + suffix array + lookup table, built by the reference's own code (sawriter's recipe)."""

    def __init__(self, genome: np.ndarray, prefixLength: int = 8, table: bool = True):
        self.gpad = padded(np.ascontiguousarray(genome, dtype=np.uint8))
        self.n = len(genome)
        self.sa = np.zeros(self.n, dtype=np.uint32)
        ref().ref_sa_build(_p(self.gpad), self.n, _p(self.sa))
        self.prefixLength = prefixLength
        self.start = self.end = None
        if table:
            self.start = np.zeros(4 ** prefixLength, dtype=np.uint32)
            self.end = np.zeros(4 ** prefixLength, dtype=np.uint32)
            ref().ref_sa_lookup_table(_p(self.gpad), self.n, _p(self.sa), prefixLength, _p(self.start), _p(self.end))

    @property
    def genome(self):
        return self.gpad[:self.n]


def map_read(which: str, ix, read: np.ndarray, prm: np.ndarray, subStart: int = 0, subEnd: int | None = None) -> np.ndarray:
    """One MapReadToGenome call through the reference ('ref') or the restatement ('orc'); (n, 3) array of (t, q, l)."""
    read = np.ascontiguousarray(read, dtype=np.uint8)
    if subEnd is None:
        subEnd = len(read)
    fn = ref().ref_map_read if which == "ref" else orc().orc_map_read
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 3), dtype=np.uint32)
        n = fn(_p(ix.gpad), ix.n, _p(ix.sa), _p(ix.start), _p(ix.end), ix.prefixLength, _p(read), len(read), subStart, subEnd,
               _p(prm), _p(out), cap)
        if n < 0:
            raise ValueError("refused")
        if n <= cap:
            return out[:n]
        cap = int(n)


def map_reads_ref(ix, reads: np.ndarray, readOff: np.ndarray, prm: np.ndarray, nThreads: int = 0, want_matches: bool = True):
    """Whole reads through the reference on nThreads threads: (matchOff[n + 1], matches (N, 3)) or the counts alone."""
    n = len(readOff) - 1
    nThreads = nThreads or (os.cpu_count() or 1)
    counts = np.zeros(n, dtype=np.uint64)
    readOff = np.ascontiguousarray(readOff, dtype=np.uint64)
    ref().ref_map_reads(_p(ix.gpad), ix.n, _p(ix.sa), _p(ix.start), _p(ix.end), ix.prefixLength, _p(reads), _p(readOff), n, _p(prm),
                        nThreads, _p(counts), None, None)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    if not want_matches:
        return off, None
    m = np.zeros((int(off[-1]), 3), dtype=np.uint32)
    ref().ref_map_reads(_p(ix.gpad), ix.n, _p(ix.sa), _p(ix.start), _p(ix.end), ix.prefixLength, _p(reads), _p(readOff), n, _p(prm),
                        nThreads, _p(counts), _p(m), _p(off))
    return off, m


def make_genome(rng: np.random.Generator, n: int, repeats: int = 4, nRuns: int = 2, lower: bool = False) -> np.ndarray:
    """Random ACGT with embedded diverged repeat copies, short low-complexity stretches and runs of N."""
    g = rng.integers(0, 4, n)
    g = np.frombuffer(b"ACGT", dtype=np.uint8)[g].copy()
    for _ in range(repeats):
        ln = int(rng.integers(n // 50 + 20, n // 10 + 40))
        if ln * 2 >= n:
            continue
        src = int(rng.integers(0, n - ln))
        for _ in range(int(rng.integers(1, 4))):
            dst = int(rng.integers(0, n - ln))
            unit = g[src:src + ln].copy()
            mut = rng.random(ln) < 0.03
            unit[mut] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(mut.sum()))]
            g[dst:dst + ln] = unit
    for _ in range(nRuns):
        ln = int(rng.integers(1, 30))
        at = int(rng.integers(0, max(1, n - ln)))
        g[at:at + ln] = ord("N")
    if n > 200:                      # a homopolymer and a dinucleotide run: deep non-unique searches
        at = int(rng.integers(0, n - 120))
        g[at:at + 50] = ord("A")
        g[at + 60:at + 110] = np.tile(np.frombuffer(b"CA", dtype=np.uint8), 25)
    if lower:
        low = rng.random(n) < 0.1
        g[low] |= 0x20
    return g


def make_read(rng: np.random.Generator, genome: np.ndarray, length: int, err: float = 0.15, rc: bool = False) -> np.ndarray:
    """A window of the genome with PacBio-like errors (ins 55 % / del 35 % / sub 10 %)."""
    n = len(genome)
    length = min(length, n)
    at = int(rng.integers(0, n - length + 1))
    w = genome[at:at + length]
    out = []
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    r = rng.random(length)
    kind = rng.random(length)
    ins = rng.integers(0, 4, length)
    for i in range(length):
        if r[i] < err:
            if kind[i] < 0.55:
                out.append(acgt[ins[i]]); out.append(w[i])
            elif kind[i] < 0.90:
                continue
            else:
                out.append(acgt[ins[i]])
        else:
            out.append(w[i])
    a = np.array(out, dtype=np.uint8)
    if rc:
        comp = np.arange(256, dtype=np.uint8)
        for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
            comp[x] = y
        a = comp[a[::-1]].copy()
    return a
