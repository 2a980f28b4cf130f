"""Index preparation for the anchoring path (what the reference's sawriter does, alignment/SAWriter.cpp:160-225): the suffix
array of a genome in the order the reference's searches assume -- suffixes compared by ThreeBit code (A < C < G < T < N,
case-insensitive; common/NucConversion.h:48-84), a suffix that is a prefix of another one first (Larsson-Sadakane with a
terminal sentinel, SuffixArray.h:256-269) -- and the k-mer look-up table over it (SuffixArray::BuildLookupTable, through
bgpu_build_lookup_table).  The array is unique, so any construction gives the reference's; this one is prefix doubling
in numpy (a few radix passes of n keys: seconds for a bacterial genome), for synthetic genomes in tests and bench.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def three_bit(genome: np.ndarray) -> np.ndarray:
    """ThreeBit[] applied to every byte (the library's own table, bgpu_base_code)."""
    L = capi.lib()
    lut = np.array([L.bgpu_base_code(c) for c in range(256)], dtype=np.uint8)
    return lut[np.ascontiguousarray(genome, np.uint8)]


def suffix_array(genome: np.ndarray) -> np.ndarray:
    """uint32 suffix array of the genome, sawriter's order."""
    n = len(genome)
    if n == 0:
        return np.zeros(0, np.uint32)
    code = three_bit(genome).astype(np.int64) + 1            # 0 = "past the end": the shorter suffix sorts first
    code[code > 6] = 7                                       # bytes outside the alphabet (ThreeBit 255) sort last
    k = min(16, n)                                           # first pass: the leading k bases packed 3 bits each
    rank = np.zeros(n, np.int64)
    for i in range(k):
        rank <<= 3
        rank[:n - i] |= code[i:]
    while True:
        sa = np.argsort(rank, kind="stable")
        sk = rank[sa]
        new = np.empty(n, np.int64)
        new[sa] = np.cumsum(np.concatenate(([1], (sk[1:] != sk[:-1]).astype(np.int64))))
        if int(new[sa[-1]]) == n or k >= n:
            return sa.astype(np.uint32)
        nxt = np.zeros(n, np.int64)
        nxt[:n - k] = new[k:]
        rank = new * (n + 1) + nxt
        k *= 2


def lookup_table(genome: np.ndarray, index: np.ndarray, prefixLength: int = 8):
    """(startPosTable, endPosTable) of SuffixArray::BuildLookupTable for this genome and suffix array."""
    g = np.ascontiguousarray(genome, np.uint8)
    ix = np.ascontiguousarray(index, np.uint32)
    start = np.zeros(4 ** prefixLength, np.uint32)
    end = np.zeros(4 ** prefixLength, np.uint32)
    rc = capi.lib().bgpu_build_lookup_table(g.ctypes.data_as(C.c_void_p), len(g), ix.ctypes.data_as(C.c_void_p), prefixLength,
                                            start.ctypes.data_as(C.c_void_p), end.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise capi.BgpuError(f"bgpu_build_lookup_table failed ({rc})")
    return start, end
