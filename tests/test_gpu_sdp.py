"""GPU parity of the device SDPAlign (bgpu_sdp_align) through the C ABI: the guide blasr hands to the
refinement (SDPAlign.h:95-637 with the argument pattern of Blasr.cpp:1716-1722), block for block against the reference
itself (oracle/_ref) and, for the other parameter patterns, against the pinned C restatement."""
import ctypes as C

import numpy as np
import pytest

from blasr_b200 import DistanceMatrixScoreFunction, JobBatch, SMRTDistanceMatrix, capi, synth
from . import cases, oracle as O

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def _absolute(res, blocks, i):
    b = blocks[int(res["blockOff"][i]):int(res["blockOff"][i]) + int(res["nBlocks"][i])]
    out = np.stack([b["qPos"] + res["qPos"][i], b["tPos"] + res["tPos"][i], b["length"]], axis=1).astype(np.uint32)
    return out.reshape(-1, 3)


def _batch_of(pairs):
    qs, ts = [p[0] for p in pairs], [p[1] for p in pairs]
    qOff = np.zeros(len(pairs) + 1, np.uint64); tOff = np.zeros(len(pairs) + 1, np.uint64)
    qOff[1:] = np.cumsum([len(x) for x in qs]); tOff[1:] = np.cumsum([len(x) for x in ts])
    return JobBatch(q=np.concatenate(qs), qOff=qOff, t=np.concatenate(ts), tOff=tOff)


@needs_ref
def test_device_sdpalign_matches_reference(aligner):
    """The 276-call set of tests/test_sdp_chain.py::test_whole_sdpalign_matches_reference, one device batch per
    (word size, indel rate): 20 b - 7 kb pairs at 5 - 30 % error, word sizes 8 / 11 / 13."""
    fn = DistanceMatrixScoreFunction(SMRTDistanceMatrix, 5, 5)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    pairs = []
    for seed, (lo, hi, err) in enumerate([(20, 60, 0.15), (50, 600, 0.30), (500, 3000, 0.25), (3000, 7000, 0.15), (200, 2000, 0.05)]):
        b = synth.simulate_pairs(6, lo, hi, err=err, seed=1300 + seed, n_rate=0.004 if seed == 2 else 0.0)
        pairs += [cases.job_arrays(b, i)[:2] for i in range(b.n)]
    batch = _batch_of(pairs)
    n_blocks = 0
    for word, rate in ((11, 0.30), (8, 0.9), (13, 0.30)):
        res, blocks = aligner.SDPAlign(batch, fn, wordSize=word, sdpIns=5, sdpDel=10, indelRate=rate)
        assert (res["status"] == 0).all(), res["status"]
        for i, (q, t) in enumerate(pairs):
            want = O.sdp_guide(q, t, ofn, word, 5, 10, rate)
            got = _absolute(res, blocks, i)
            assert np.array_equal(got, want.reshape(-1, 3)), (i, word, rate, len(got), len(want))
            n_blocks += len(want)
    assert n_blocks > 5000


def test_device_sdpalign_other_argument_patterns(aligner):
    """Global / no detail / front extension / no recursion / anchor cap (AlignSubstring's pattern, Blasr.cpp:1080-1090, and
    the -noDetailedSDP / -sdpMaxAnchorsPerPosition switches) against the C restatement, soft-masked and N-holding inputs."""
    fn = DistanceMatrixScoreFunction(SMRTDistanceMatrix, 5, 5)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    L = O._load("orc")
    b = cases.guided_batch(seed=77, n=10, lo=40, hi=2500, err=0.2, n_rate=0.01, lower=True)
    pairs = [cases.job_arrays(b, i)[:2] for i in range(b.n)]
    batch = _batch_of(pairs)
    n = 0
    for (at, detailed, front, prefix, recurse, under, maxm, word) in ((1, 1, 1, 50, 2, 1000, 0, 11), (0, 0, 0, 50, 2, 1000, 0, 11),
                                                                        (1, 1, 1, 0, 0, 1000, 0, 8), (0, 1, 0, 50, 1, 200, 20, 11)):
        res, blocks = aligner.SDPAlign(batch, fn, wordSize=word, sdpIns=5, sdpDel=10, indelRate=0.25, alignType=at,
                                       detailedAlignment=bool(detailed), extendFrontByLocalAlignment=bool(front), sdpPrefixLength=prefix,
                                       recurse=recurse, noRecurseUnder=under, maxMatchesPerPosition=maxm)
        assert (res["status"] == 0).all()
        for i, (q, t) in enumerate(pairs):
            q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
            cap = len(q) + len(t) + 8
            wb = np.zeros((cap, 3), np.uint32); qp = C.c_uint32(0); tp = C.c_uint32(0)
            # orc_sdp_align has no maxMatches argument at the top level (blasr's default 0): compare that pattern with 0 only
            if maxm:
                continue
            k = L.orc_sdp_align(C.byref(ofn), q.ctypes.data, len(q), t.ctypes.data, len(t), word, 5, 10, C.c_float(0.25), at, detailed, front,
                                prefix, recurse, under, wb.ctypes.data, cap, C.byref(qp), C.byref(tp))
            assert k >= 0
            want = wb[:k].copy(); want[:, 0] += qp.value; want[:, 1] += tp.value
            assert np.array_equal(_absolute(res, blocks, i), want), (i, at, detailed, front, recurse)
            n += k
    assert n > 500


def test_device_sdpalign_feeds_the_refinement(aligner):
    """Raw (query, target) pairs end to end on the device: SDPAlign's blocks are the guide of GuidedAlign (the pipeline of
    Blasr.cpp:1716 -> :863), and the refined alignments equal the oracle's refinement of the reference's own guide."""
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50, 0)
    b = synth.simulate_pairs(12, 300, 4000, err=0.15, seed=909)
    pairs = [cases.job_arrays(b, i)[:2] for i in range(b.n)]
    batch = _batch_of(pairs)
    res, blocks = aligner.SDPAlign(batch, fn, wordSize=11, indelRate=0.30)
    guides, off = [], [0]
    for i in range(batch.n):
        g = _absolute(res, blocks, i)
        assert len(g) > 0
        guides.append(g); off.append(off[-1] + len(g))
    batch.guide = np.concatenate(guides).astype(np.uint32); batch.guideOff = np.asarray(off, np.uint64)
    out = aligner.GuidedAlign(batch, fn, 16)
    which = "ref" if O.have_ref() else "orc"
    want = cases.oracle_batch(which, batch, ofn, 0, 1, 16)
    for i in range(batch.n):
        bad = cases.compare(cases.gpu_to_dict(out, i), want[i], cases.GPU_FIELDS)
        assert not bad, (i, bad)
