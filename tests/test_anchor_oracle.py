"""The C restatement of suffix-array anchoring (oracle/orc_anchor.c) against the reference itself
(MapReadToGenome, common/algorithms/anchoring/MapBySuffixArray.h:209-309, compiled unmodified into oracle/_ref).

The reference ships no test or golden vector for this path: parity is pinned against outputs of the
reference run here, over blasr's default AnchorParameters and every switch MapReadToGenome reads.
"""
import numpy as np
import pytest

from . import anchor_oracle as ao

pytestmark = pytest.mark.skipif(ao.ref() is None, reason="oracle/_ref not built (needs /root/reference)")

VARIANTS = [
    {},                                                     # blasr defaults
    dict(stopUnique=0),
    dict(stopUnique=0, maxLCP=20),
    dict(expand=1), dict(expand=3),
    dict(maxAnchors=2), dict(maxAnchors=1, advance=30, minMatch=30, stopUnique=1),   # MappingParameters.h:548-554
    dict(advance=2),
    # without the table the reference's match length is minPrefix + depth - 1 whatever minPrefix is (MapBySuffixArray.h:110):
    # only minPrefix = 1 is coherent (longer ones trip its own assertion at :301 near the end of the genome)
    dict(useLookup=0, minPrefix=1), dict(useLookup=0, minPrefix=1, stopUnique=0, minMatch=5),
    dict(minMatch=14), dict(minMatch=9),     # <= 8 would accept the garbage ranges BuildLookupTable leaves for N suffixes (the reference asserts, :302)
    dict(minMatch=21, maxLCP=22, maxAnchors=10),            # MappingParameters.h:457-466
]


def test_three_bit_table():
    import ctypes as C
    L = ao.orc()
    from . import oracle as o
    R = o._load("ref")
    for c in range(256):
        assert L.orc_three_bit(c) == R.ref_three_bit(c), c


@pytest.mark.parametrize("seed", range(6))
def test_restatement_matches_reference(seed):
    rng = np.random.default_rng(100 + seed)
    n = [3000, 20000, 800, 60000, 5000, 257][seed]
    g = ao.make_genome(rng, n, lower=(seed % 2 == 1))
    if seed == 4:
        g[-1] = ord("N")                                     # blasr's own genomes end with the appended 'N'
    ix = ao.Index(g)
    total = 0
    for k in range(10):
        ln = int(rng.integers(10, min(n, 1500)))
        read = ao.make_read(rng, ix.genome, ln, err=[0.0, 0.05, 0.15, 0.3][k % 4], rc=(k % 3 == 2))
        if k == 5 and len(read) > 40:
            read[20:24] = ord("N")
        for v in VARIANTS:
            prm = ao.params(**v)
            a = ao.map_read("ref", ix, read, prm)
            b = ao.map_read("orc", ix, read, prm)
            assert a.shape == b.shape and np.array_equal(a, b), (seed, k, v)
            total += len(a)
    assert total > 0


def test_subreads_and_short_reads():
    rng = np.random.default_rng(7)
    g = ao.make_genome(rng, 10000)
    ix = ao.Index(g)
    read = ao.make_read(rng, ix.genome, 900, err=0.1)
    for s, e in [(0, len(read)), (100, 600), (50, 70), (10, 22), (0, 12), (5, 13), (300, 305)]:
        for v in ({}, dict(stopUnique=0), dict(advance=3)):
            prm = ao.params(**v)
            a = ao.map_read("ref", ix, read, prm, s, e)
            b = ao.map_read("orc", ix, read, prm, s, e)
            assert np.array_equal(a, b), (s, e, v)
    for ln in (0, 1, 7, 8, 9, 12, 13, 14):
        r = read[:ln]
        assert np.array_equal(ao.map_read("ref", ix, r, ao.params()), ao.map_read("orc", ix, r, ao.params())), ln


def test_other_table_sizes():
    rng = np.random.default_rng(9)
    g = ao.make_genome(rng, 8000)
    for pl in (4, 6, 10):
        ix = ao.Index(g, prefixLength=pl)
        for k in range(4):
            read = ao.make_read(rng, ix.genome, 700, err=0.12)
            for v in ({}, dict(stopUnique=0), dict(minMatch=pl + 1)):
                prm = ao.params(minPrefix=pl, **v)
                assert np.array_equal(ao.map_read("ref", ix, read, prm), ao.map_read("orc", ix, read, prm)), (pl, k, v)


def test_bulk_entry_matches_single_calls():
    rng = np.random.default_rng(11)
    g = ao.make_genome(rng, 30000)
    ix = ao.Index(g)
    reads = [ao.make_read(rng, ix.genome, int(rng.integers(50, 2000))) for _ in range(12)]
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    flat = np.concatenate(reads)
    mo, m = ao.map_reads_ref(ix, flat, off, ao.params(), nThreads=3)
    for i, r in enumerate(reads):
        assert np.array_equal(m[int(mo[i]):int(mo[i + 1])], ao.map_read("orc", ix, r, ao.params())), i


@pytest.mark.parametrize("seed", range(5))
def test_index_preparation_matches_sawriter(seed):
    """blasr_b200/saindex.py (suffix array) and bgpu_build_lookup_table (SuffixArray::BuildLookupTable) against the
    reference's own construction: Larsson-Sadakane on the ThreeBit text, then BuildLookupTable."""
    from blasr_b200 import saindex
    rng = np.random.default_rng(900 + seed)
    n = [5000, 70000, 300, 12, 20000][seed]
    g = ao.make_genome(rng, n, lower=(seed == 1), nRuns=[2, 5, 1, 0, 0][seed])
    if seed in (0, 1):
        g[-1] = ord("N")                               # blasr's genomes end with the separator FASTAReader.h:130 appends
    if seed == 4:
        g[-9:] = np.frombuffer(b"TTTTTTTTT", np.uint8)  # the text ends in the last tuple of the table
    sa = saindex.suffix_array(g)
    ix = ao.Index(g, table=False)
    assert np.array_equal(sa, ix.sa)
    for pl in (2, 5, 8):
        want = ao.Index(g, prefixLength=pl)
        start, end = saindex.lookup_table(g, sa, pl)
        assert np.array_equal(start, want.start) and np.array_equal(end, want.end), pl
