"""The chaining step of SDPAlign (SURVEY 8f N2, next scope row): the C restatement of SDPLongestCommonSubsequence
(oracle/orc_sdp.c) against the reference's own template (oracle/_ref), on fragment sets with unique (x, y)."""
import numpy as np
import pytest

from tests import oracle as O

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def kmer_fragments(rng, n, err, k, small=0):
    """Exact k-mer matches between a random target and a read of it with indels / substitutions: the (x = query
    position, y = target position) set StoreMatchingPositions would produce, de-duplicated."""
    t = rng.integers(0, 4, n)
    q = []
    for b in t:
        r = rng.random()
        if r < err * 0.5:
            q.append(int(rng.integers(0, 4))); q.append(int(b))
        elif r < err * 0.85:
            continue
        elif r < err:
            q.append(int((b + 1 + rng.integers(0, 3)) % 4))
        else:
            q.append(int(b))
    q = np.asarray(q, np.int64)

    def keys(s, kk):
        if len(s) < kk:
            return np.zeros(0, np.int64)
        v = np.zeros(len(s) - kk + 1, np.int64)
        for j in range(kk):
            v = v * 4 + s[j:len(s) - kk + 1 + j]
        return v
    out = {}
    for kk, length in ((k, k),) + (((small, small),) if small else ()):
        kt, kq = keys(t, kk), keys(q, kk)
        pos = {}
        for y, v in enumerate(kt):
            pos.setdefault(int(v), []).append(y)
        for x, v in enumerate(kq):
            for y in pos.get(int(v), ()):
                if small and kk == small and not (x < 50 or x >= len(q) - 50):
                    continue                     # short words only near the read ends, as SDPAlign's prefix / suffix sets
                out.setdefault((x, y), (length, k))
    fr = np.asarray([(x, y, l, w) for (x, y), (l, w) in out.items()], np.uint32).reshape(-1, 4)
    return fr, len(q)


@needs_ref
@pytest.mark.parametrize("align_type", [1, 0])          # Global, Local
def test_chain_matches_reference_on_kmer_matches(align_type):
    rng = np.random.default_rng(700 + align_type)
    total = 0
    for rep in range(30):
        k = int(rng.choice([5, 8, 11]))
        fr, qlen = kmer_fragments(rng, int(rng.integers(60, 900)), float(rng.choice([0.0, 0.1, 0.2, 0.3])), k,
                                  small=5 if (rep % 3 == 0 and k > 5) else 0)
        if len(fr) == 0:
            continue
        ins, dele, match = (5, 10, -5) if rep % 2 == 0 else (int(rng.integers(1, 12)), int(rng.integers(1, 12)), -int(rng.integers(1, 9)))
        want = O.sdp_chain("ref", fr, qlen, k, ins, dele, match, align_type)
        got = O.sdp_chain("orc", fr, qlen, k, ins, dele, match, align_type)
        assert np.array_equal(got, want), (rep, k, len(fr), got[:10], want[:10])
        total += len(want)
    assert total > 500


@needs_ref
def test_chain_matches_reference_on_random_fragments():
    """Dense random fragment clouds: many same-row, same-column and same-diagonal neighbours (every set operation runs)."""
    rng = np.random.default_rng(42)
    for rep in range(40):
        n = int(rng.integers(1, 400)); span = int(rng.integers(5, 120))
        xy = np.unique(rng.integers(0, span, size=(n, 2)), axis=0)
        k = int(rng.integers(2, 12))
        fr = np.concatenate([xy, np.full((len(xy), 1), k), np.full((len(xy), 1), k)], axis=1).astype(np.uint32)
        for at in (1, 0):
            want = O.sdp_chain("ref", fr, span + 3, k, 5, 10, -5, at)
            got = O.sdp_chain("orc", fr, span + 3, k, 5, 10, -5, at)
            assert np.array_equal(got, want), (rep, at, len(fr))


@needs_ref
def test_fragment_set_and_chain_match_reference_sdpalign():
    """End to end over the deterministic front half: the fragment set (prefix / middle / suffix k-mer matches, sorted with
    the restated libstdc++ std::sort, de-duplicated) and the chain over it, against what the reference's SDPAlign leaves
    in its own buffers -- including which of two equal (x, y) fragments (length 5 vs 11) survives."""
    from blasr_b200 import SMRTDistanceMatrix
    from tests import cases
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    n_dup_len5 = n_frag = 0
    for seed, lo, hi, n_rate, lower in ((1, 30, 400, 0.0, False), (2, 100, 3000, 0.01, True), (3, 2000, 6000, 0.0, False)):
        b = cases.guided_batch(seed=300 + seed, n=8, lo=lo, hi=hi, n_rate=n_rate, lower=lower)
        for i in range(b.n):
            q, t, _, _ = cases.job_arrays(b, i)
            for word, at in ((11, 0), (11, 1), (8, 1), (5, 0)):
                want, want_chain = O.ref_sdp_fragments(q, t, fn, word, 5, 10, at)
                got = O.orc_sdp_fragments(q, t, word)
                assert got is not None
                assert np.array_equal(got, want), (seed, i, word, len(got), len(want))
                if len(want):
                    match = int(fn.M[0])
                    chain = O.sdp_chain("orc", got, len(q), word, 5, 10, match, at)
                    assert np.array_equal(chain, want_chain), (seed, i, word, at)
                n_frag += len(want); n_dup_len5 += int(((want[:, 2] == 5) & (want[:, 3] == 11)).sum())
    assert n_frag > 10000 and n_dup_len5 > 100


@needs_ref
def test_whole_sdpalign_matches_reference():
    """SDPAlign as blasr calls it (Local, detailed, prefix 50, recurse 2, noRecurseUnder 1000; Blasr.cpp:1716-1722): fragment
    set, chain, chain -> blocks, SWAlign / recursive SDPAlign gap fills -- the guide the refinement receives, block for block."""
    from blasr_b200 import SMRTDistanceMatrix, synth
    from tests import cases
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    n_blocks = 0
    for seed, (lo, hi, err) in enumerate([(20, 60, 0.15), (50, 600, 0.30), (500, 3000, 0.25), (3000, 7000, 0.15), (200, 2000, 0.05)]):
        b = synth.simulate_pairs(6, lo, hi, err=err, seed=1300 + seed, n_rate=0.004 if seed == 2 else 0.0)
        for i in range(b.n):
            q, t, _, _ = cases.job_arrays(b, i)
            for word, rate in ((11, 0.30), (8, 0.9), (13, 0.30)):
                want = O.sdp_guide(q, t, fn, word, 5, 10, rate)
                got = O.orc_sdp_guide(q, t, fn, word, 5, 10, rate)
                assert np.array_equal(got, want), (seed, i, word, rate, len(got), len(want))
                n_blocks += len(want)
    assert n_blocks > 5000
