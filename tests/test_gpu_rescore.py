"""GPU parity of bgpu_rescore, the rescoring step of StoreMapQVs (alignment/Blasr.cpp:2768-2780): the reference's
ComputeAlignmentScore(alignment, qAlignedSeq, tAlignedSeq, scoreFn, useAffinePenalty) -- the Alignment overload,
AlignmentUtils.h:127-169, NOT the string form ComputeAlignmentStats uses -- under SMRTLogProbMatrix, on alignments refined with
SMRTDistanceMatrix, against the reference itself (oracle/_ref)."""
import numpy as np
import pytest

from blasr_b200 import DistanceMatrixScoreFunction, SMRTDistanceMatrix, capi
from . import cases, oracle as O

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")

# common/algorithms/alignment/ScoreMatrices.h:28-34
SMRTLogProbMatrix = [[0 if i == j else 15 for j in range(5)] for i in range(5)]
ASYM = [[(3 * i + 7 * j) % 11 - 4 for j in range(5)] for i in range(5)]


@needs_ref
@pytest.mark.parametrize("algo,adversarial", [(capi.AFFINE_GUIDED, 0.0), (capi.GUIDED, 0.0), (capi.AFFINE_GUIDED, 0.15), (capi.GUIDED, 0.2)])
def test_rescore_matches_reference(aligner, algo, adversarial):
    b = cases.guided_batch(seed=4100 + algo, n=40, lo=100, hi=4000, adversarial=adversarial, run=3, n_rate=0.004, lower=True)
    fn = DistanceMatrixScoreFunction(SMRTDistanceMatrix, 5, 5, 50, 0)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50, 0)
    tk = aligner.submit(b, fn, algo, band=16)
    res = aligner.collect(tk)
    try:
        for M, ins, dele, aopen, aext in ((SMRTLogProbMatrix, 5, 5, 50, 0), (SMRTLogProbMatrix, 4, 7, 3, 1), (ASYM, 2, 9, 6, 2)):
            fn2 = DistanceMatrixScoreFunction(M, ins, dele, aopen, aext)
            ofn2 = O.score_fn(M, ins, dele, aopen, aext)
            for aff in (False, True):
                got = aligner.rescore(tk, fn2, aff)
                for i in range(b.n):
                    q, t, g, _ = cases.job_arrays(b, i)
                    j, keep = O.make_job(algo, capi.GLOBAL, 16, q, t, g, None, 0, 0, 1, int(algo == capi.AFFINE_GUIDED))
                    want = O.rescore(ofn, j, ofn2, aff) if int(res.results["status"][i]) == 0 else 0
                    assert int(got[i]) == want, (i, aff, int(got[i]), want)
    finally:
        aligner.release(tk)


def test_rescore_refusals(aligner):
    b = cases.guided_batch(seed=5, n=3, lo=100, hi=300)
    fn = DistanceMatrixScoreFunction(SMRTDistanceMatrix, 5, 5)
    tk = aligner.submit(b, fn, capi.GUIDED, band=10)
    with pytest.raises(capi.BgpuError):
        aligner.rescore(tk, fn)                     # not collected yet
    aligner.collect(tk)
    assert len(aligner.rescore(tk, fn)) == 3
    aligner.release(tk)
