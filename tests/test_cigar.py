"""SAM CIGAR core: the reference's own printer code (oracle/_ref: CreateNoClippingCigarOps run on the result of
the reference aligner) against the C restatement (CPU) and against bgpu_cigar (GPU), op for op."""
import numpy as np
import pytest

from blasr_b200 import DistanceMatrixScoreFunction, SMRTDistanceMatrix, align as A
from tests import cases, oracle as O

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def _mixed_case_batch(seed, n, lo, hi):
    """Pairs with lower-case stretches and N's: the printer compares raw bytes, the aligner base codes."""
    b = cases.guided_batch(seed=seed, n=n, lo=lo, hi=hi, n_rate=0.01, lower=True)
    return b


@needs_ref
@pytest.mark.parametrize("algo", [0, 1])
def test_restatement_matches_reference_printer(algo):
    b = _mixed_case_batch(90 + algo, 12, 60, 900)
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0)
    n_ops = 0
    for i in range(b.n):
        q, t, g, _ = cases.job_arrays(b, i)
        j, keep = O.make_job(algo, 1, 12, q, t, g, None, 0, 0, 0, 0)
        want = O.ref_cigar(fn, j)
        aln = O.align("ref", fn, j)
        got = O.orc_cigar_from(q, t, aln)
        assert np.array_equal(got, want), (i, A.cigar_string(got)[:80], A.cigar_string(want)[:80])
        # the ops account for every aligned base of both sequences
        qlen = sum(int(o) >> 4 for o in want if int(o) & 15 in (1, 7, 8)); tlen = sum(int(o) >> 4 for o in want if int(o) & 15 in (2, 7, 8))
        if len(aln["blocks"]):
            last = aln["blocks"][-1]
            assert qlen == int(last[0] + last[2] - aln["blocks"][0][0]) and tlen == int(last[1] + last[2] - aln["blocks"][0][1])
        n_ops += len(want)
    assert n_ops > 100


@needs_ref
def test_restatement_on_anchor_only_and_local_guides():
    """Guides with real gaps between their blocks (min_block 8) and Local alignments (qPos / tPos past the guide start)."""
    b = cases.guided_batch(seed=93, n=10, lo=80, hi=700, min_block=8)
    fn = O.score_fn(SMRTDistanceMatrix, 4, 6, 0, 0)
    for at in (0, 1):
        for i in range(b.n):
            q, t, g, _ = cases.job_arrays(b, i)
            j, keep = O.make_job(0, at, 20, q, t, g, None, 0, 0, 0, 0)
            want = O.ref_cigar(fn, j)
            got = O.orc_cigar_from(q, t, O.align("ref", fn, j))
            assert np.array_equal(got, want), (at, i)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_gpu_cigar_matches_reference_printer(aligner, algo):
    b = _mixed_case_batch(190 + algo, 40, 50, 5000)
    b.band = np.random.default_rng(3).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50 if algo else 0, affineExtend=0)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0)
    n_ops = 0
    for at in (1, 0):                                   # Global, Local
        tk = aligner.submit(b, fn, algo, alignType=at, band=16)
        aligner.collect(tk)
        ops, off = aligner.cigar(tk)
        for i in range(b.n):
            q, t, g, _ = cases.job_arrays(b, i)
            j, keep = O.make_job(algo, at, int(b.band[i]), q, t, g, None, 0, 0, 0, 0)
            want = O.ref_cigar(ofn, j)
            got = ops[int(off[i]):int(off[i + 1])]
            assert np.array_equal(got, want), (at, i, A.cigar_string(got)[:80], A.cigar_string(want)[:80])
            n_ops += len(want)
        aligner.release(tk)
    assert n_ops > 1000


@pytest.mark.gpu
def test_gpu_cigar_refuses_dense_tickets(aligner):
    from blasr_b200 import BgpuError, JobBatch, capi
    b = JobBatch.from_lists([b"ACGTACGT"], [b"ACGAACGT"])
    tk = aligner.submit(b, DistanceMatrixScoreFunction(), capi.SW, alignType=capi.GLOBAL)
    aligner.collect(tk)
    with pytest.raises(BgpuError):
        aligner.cigar(tk)
    aligner.release(tk)


@needs_ref
def test_m5_strings_restatement_matches_reference():
    """CreateAlignmentStrings (the qalignedseq / matchpattern / talignedseq columns of -m 5) restated in C against the
    reference, on refined alignments (gap lists) of mixed-case / N inputs and on KBandAlign / SWAlign results."""
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50, 0)
    b = _mixed_case_batch(290, 16, 60, 1500)
    total = 0
    for i in range(b.n):
        q, t, g, _ = cases.job_arrays(b, i)
        for algo, at, band in ((1, 1, 16), (0, 0, 10), (2, 1, 12), (3, 1, 0)):
            qq, tt = (q, t) if algo < 3 else (q[:60], t[:70])
            j, keep = O.make_job(algo, at, band, qq, tt, g if algo < 2 else None, None, 5, 5, 0, 0)
            aln = O.align("ref", fn, j)
            if aln["status"] != 0:
                continue
            want = O.ref_alignment_strings(fn, j)
            got = O.orc_alignment_strings(qq, tt, aln)
            assert got == want, (i, algo, got[1][:60], want[1][:60])
            total += len(want[0])
    assert total > 20000


@needs_ref
def test_m5_strings_of_block_only_alignments():
    """The branch without gap lists (what SDPAlign returns): leading offsets and inter-block gaps laid out from the block
    coordinates alone."""
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5)
    b = cases.guided_batch(seed=295, n=10, lo=80, hi=1500)
    total = 0
    for i in range(b.n):
        q, t, _, _ = cases.job_arrays(b, i)
        blocks = O.sdp_guide(q, t, fn)
        want = O.ref_block_strings(q, t, blocks)
        got = O.orc_alignment_strings(q, t, {"blocks": blocks, "gaps": [], "qPos": 0, "tPos": 0}, with_gaps=False)
        assert got == want, (i, got[1][:60], want[1][:60])
        total += len(want[0])
    assert total > 5000


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_gpu_clipped_cigar_matches_create_cigar_string(aligner, algo):
    """bgpu_cigar_clipped against the reference's whole CreateCIGARString (SAMPrinter.h:345-400): alignments placed inside longer
    reads, -clipping hard / soft / none, both strands (the op list of a reverse-strand alignment is reversed)."""
    rng = np.random.default_rng(31 + algo)
    b = _mixed_case_batch(390 + algo, 36, 50, 3000)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50 if algo else 0, affineExtend=0)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0)
    for at in (1, 0):
        tk = aligner.submit(b, fn, algo, alignType=at, band=16)
        aligner.collect(tk)
        clips = np.zeros((b.n, 4), np.uint32); strand = np.zeros(b.n, np.uint8); want = []
        for i in range(b.n):
            q, t, g, _ = cases.job_arrays(b, i)
            j, keep = O.make_job(algo, at, 16, q, t, g, None, 0, 0, 0, 0)
            mode = int(rng.integers(0, 4)) if i % 4 else 3            # hard, soft, subread, none
            lowp, lows = (int(rng.integers(0, 30)), int(rng.integers(0, 30))) if mode in (1, 2) else (0, 0)
            qpos = lowp + int(rng.integers(0, 50)); rlen = qpos + len(q) + lows + int(rng.integers(0, 50))
            strand[i] = int(rng.integers(0, 2))
            text, cl = O.ref_cigar_string(ofn, j, mode, int(strand[i]), qpos, rlen, lowp, lows)
            want.append(text); clips[i] = cl
        ops, off = aligner.cigar_clipped(tk, clips, strand)
        for i in range(b.n):
            got = A.cigar_string(ops[int(off[i]):int(off[i + 1])])
            assert got == want[i], (at, i, got[:80], want[i][:80])
        # without clips / strands it is the core that bgpu_cigar returns
        o2, f2 = aligner.cigar_clipped(tk, None, None)
        o1, f1 = aligner.cigar(tk)
        assert np.array_equal(o1, o2) and np.array_equal(f1, f2)
        aligner.release(tk)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_gpu_alignment_strings_match_reference(aligner, algo):
    """bgpu_strings (the three columns -m 5 prints) against the reference's CreateAlignmentStrings run on the reference's own
    alignment of the same job: mixed case, N's, per-job bands, Global and Local."""
    b = _mixed_case_batch(490 + algo, 40, 50, 5000)
    b.band = np.random.default_rng(5).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50 if algo else 0, affineExtend=0)
    ofn = O.score_fn(SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0)
    total = 0
    for at in (1, 0):
        tk = aligner.submit(b, fn, algo, alignType=at, band=16)
        aligner.collect(tk)
        text, aln, query, off = aligner.strings(tk)
        for i in range(b.n):
            q, t, g, _ = cases.job_arrays(b, i)
            j, keep = O.make_job(algo, at, int(b.band[i]), q, t, g, None, 0, 0, 0, 0)
            want = O.ref_alignment_strings(ofn, j)
            lo, hi = int(off[i]), int(off[i + 1])
            got = (text[lo:hi].tobytes(), aln[lo:hi].tobytes(), query[lo:hi].tobytes())
            assert got == want, (at, i, got[1][:60], want[1][:60])
            total += hi - lo
        aligner.release(tk)
    assert total > 100000
