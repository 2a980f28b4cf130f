"""GPU parity of suffix-array anchoring (bgpu_set_suffix_array / bgpu_map_reads) through the C ABI: every
read's matchPosList (t, q, l) in order against the reference's own MapReadToGenome (MapBySuffixArray.h:209-309, compiled
unmodified into oracle/_ref) over an index built by the reference's own code (sawriter's recipe), for blasr's default
AnchorParameters and every switch the function reads."""
import numpy as np
import pytest

from blasr_b200 import capi
from . import anchor_oracle as ao
from .test_anchor_oracle import VARIANTS

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(ao.ref() is None, reason="oracle/_ref not built")


def _flat(reads):
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    return (np.concatenate(reads) if reads else np.zeros(0, np.uint8)), off


def _load(aligner, ix, table=True):
    aligner.set_reference(ix.genome)
    if table:
        aligner.set_suffix_array(ix.sa, ix.start, ix.end, ix.prefixLength)
    else:
        aligner.set_suffix_array(ix.sa)


def _kw(v):
    names = dict(minPrefix="minPrefixMatchLength", minMatch="minMatchLength", expand="expand", useLookup="useLookupTable",
                 maxAnchors="maxAnchorsPerPosition", advance="advanceExactMatches", maxLCP="maxLCPLength", stopUnique="stopMappingOnceUnique")
    return {names[k]: x for k, x in v.items()}


def _compare(aligner, ix, reads, v, which="ref", subs=None):
    flat, off = _flat(reads)
    kw = _kw(v)
    if subs is not None:
        kw["subreadStart"] = np.array([s for s, _ in subs], np.uint32)
        kw["subreadEnd"] = np.array([e for _, e in subs], np.uint32)
    mo, m = aligner.MapReadToGenome(flat, off, **kw)
    prm = ao.params(**v)
    total = 0
    for i, r in enumerate(reads):
        s, e = subs[i] if subs is not None else (0, len(r))
        want = ao.map_read(which, ix, r, prm, s, e)
        got = m[int(mo[i]):int(mo[i + 1])]
        got = np.stack([got["t"], got["q"], got["l"]], axis=1) if len(got) else np.zeros((0, 3), np.uint32)
        assert got.shape == want.shape and np.array_equal(got, want), (i, v, len(got), len(want))
        total += len(want)
    return total


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_map_reads_matches_reference(aligner, seed):
    rng = np.random.default_rng(500 + seed)
    n = [20000, 300000, 3000, 60000][seed]
    ix = ao.Index(ao.make_genome(rng, n, repeats=6, lower=(seed == 1)))
    _load(aligner, ix)
    reads = []
    for k in range(24):
        r = ao.make_read(rng, ix.genome, int(rng.integers(10, min(n, 4000))), err=[0.0, 0.05, 0.15, 0.3][k % 4], rc=(k % 3 == 2))
        if k == 5 and len(r) > 40:
            r[20:24] = ord("N")
        reads.append(r)
    reads += [reads[0][:ln] for ln in (0, 1, 7, 8, 12, 13)]          # shorter than the key, than minMatchLength, just above
    total = sum(_compare(aligner, ix, reads, v) for v in VARIANTS)
    assert total > 1000


@needs_ref
def test_subreads_tables_and_no_table(aligner):
    rng = np.random.default_rng(77)
    g = ao.make_genome(rng, 40000)
    for pl in (6, 8, 10):
        ix = ao.Index(g, prefixLength=pl)
        _load(aligner, ix)
        reads = [ao.make_read(rng, ix.genome, int(rng.integers(200, 2500)), err=0.12) for _ in range(8)]
        subs = []
        for r in reads:
            s = int(rng.integers(0, len(r) // 2)); e = int(rng.integers(s, len(r) + 1))
            subs.append((s, e))
        subs[0] = (0, len(reads[0])); subs[1] = (5, 5 + pl); subs[2] = (10, 22)
        for v in ({}, dict(stopUnique=0), dict(advance=3), dict(expand=2)):
            _compare(aligner, ix, reads, dict(minPrefix=pl, **v), subs=subs)
    # an index without a lookup table (SuffixArray::startPosTable == NULL): full-range searches from depth 0
    ix = ao.Index(g, table=False)
    _load(aligner, ix, table=False)
    reads = [ao.make_read(rng, ix.genome, 600, err=0.1) for _ in range(4)]
    for v in (dict(minPrefix=1), dict(minPrefix=1, stopUnique=0, minMatch=5), dict(minPrefix=1, useLookup=0, expand=5)):
        _compare(aligner, ix, reads, v)


@needs_ref
def test_bulk_shard_matches_reference(aligner):
    """A read shard the size of a bench step in miniature (config[0] shape: 10 kb reads, 15 % error) against the reference
    on all host cores; also repeats the call on the cached buffers and re-executes the kernels (the bench's resident leg)."""
    rng = np.random.default_rng(2024)
    ix = ao.Index(ao.make_genome(rng, 2_000_000, repeats=10, nRuns=5))
    _load(aligner, ix)
    reads = [ao.make_read(rng, ix.genome, 10000, rc=bool(k & 1)) for k in range(64)]
    flat, off = _flat(reads)
    want_off, want = ao.map_reads_ref(ix, flat, off, ao.params())
    for _ in range(2):
        mo, m = aligner.MapReadToGenome(flat, off)
        assert np.array_equal(mo, want_off)
        assert np.array_equal(np.stack([m["t"], m["q"], m["l"]], axis=1), want)
    aligner.map_rerun()
    ms_locate, ms_rest, positions, h2d, d2h = aligner.map_timing()
    assert positions == sum(len(r) - 8 + 1 for r in reads) and ms_locate > 0 and h2d >= len(flat) and d2h >= 12 * len(want)


def test_refusals(aligner):
    rng = np.random.default_rng(5)
    g = ao.make_genome(rng, 5000)
    flat, off = _flat([g[100:400].copy()])
    aligner.set_reference(g)
    aligner.set_suffix_array(None)
    with pytest.raises(capi.BgpuError):
        aligner.MapReadToGenome(flat, off)                      # no index on the device
    if ao.ref() is None:
        return
    ix = ao.Index(g)
    _load(aligner, ix)
    with pytest.raises(capi.BgpuError):
        aligner.MapReadToGenome(flat, off, removeEncompassedMatches=True)
    with pytest.raises(capi.BgpuError):
        aligner.MapReadToGenome(flat, off, minPrefixMatchLength=6)     # shorter than the table's key
    with pytest.raises(capi.BgpuError):
        aligner.MapReadToGenome(flat, off, subreadStart=[10], subreadEnd=[5000])
    aligner.set_reference(g[:-1])
    with pytest.raises(capi.BgpuError):
        aligner.MapReadToGenome(flat, off)                      # genome and index disagree
    mo, m = _load(aligner, ix) or aligner.MapReadToGenome(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(m) == 0 and mo[0] == 0
