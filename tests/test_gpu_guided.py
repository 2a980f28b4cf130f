"""GPU parity: GuidedAlign / AffineGuidedAlign through the C ABI vs the oracle, bit-exact on every field
(score, qPos, tPos, nCells, blocks, gap lists, stats)."""
import numpy as np
import pytest

from blasr_b200 import DistanceMatrixScoreFunction, IDSScoreFunction, QualityValueScoreFunction, SMRTDistanceMatrix
from blasr_b200 import capi
from . import cases, oracle as O

pytestmark = pytest.mark.gpu
WHICH = "ref" if O.have_ref() else "orc"


def _run(aligner, b, algo, fn, band, at=1):
    if algo:
        res = aligner.AffineGuidedAlign(b, fn, band, alignType=at)
    else:
        res = aligner.GuidedAlign(b, fn, band, alignType=at)
    ofn = O.score_fn(fn.scoreMatrix, fn.ins, fn.del_, fn.affineOpen, fn.affineExtend, fn.kind, fn.substitutionPrior,
                     fn.globalDeletionPrior)
    bd = b.band if b.band is not None else band
    want = cases.oracle_batch(WHICH, b, ofn, algo, at, bd, statsAffine=algo)
    nbad = 0
    for i in range(b.n):
        got = cases.gpu_to_dict(res, i)
        # the product has no fallback: a job the library refuses (TOO_WIDE / RANGE) is a failure, never a skip
        assert got["status"] not in (capi.JOB_TOO_WIDE, capi.JOB_RANGE), f"job {i} refused with status {got['status']}"
        bad = cases.compare(got, want[i], cases.GPU_FIELDS)
        assert not bad, f"job {i}: {bad}"
    return res, want


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("band", [4, 10, 16, 32, 64])
def test_natural_guides(aligner, algo, band):
    b = cases.guided_batch(seed=200 + band, n=40, lo=100, hi=4000, n_rate=0.01, lower=True)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    res, want = _run(aligner, b, algo, fn, band)
    assert (res.results["status"] == 0).all()


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("at", [0, 1])
def test_adversarial_and_params(aligner, algo, at):
    rng = np.random.default_rng(7 + algo + 2 * at)
    ok = n = 0
    for rep in range(10):
        b = cases.guided_batch(seed=300 + rep, n=12, lo=60, hi=2500, err=float(rng.choice([0.02, 0.15, 0.3])),
                               adversarial=float(rng.choice([0.0, 0.2, 0.6])), run=int(rng.choice([1, 5, 40])), n_rate=0.01)
        M = SMRTDistanceMatrix.copy()
        if rep % 3 == 0:
            M = rng.integers(-6, 8, size=(5, 5)).astype(np.int32)
        fn = DistanceMatrixScoreFunction(M, int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.choice([0, 3, 7, 11, 50])),
                                         int(rng.choice([0, 1, 2])))
        res, want = _run(aligner, b, algo, fn, int(rng.choice([4, 10, 16, 32])), at)
        ok += int((res.results["status"] == 0).sum())
        n += b.n
    # every one of the 120 jobs was compared field by field in _run (statuses included: the few non-OK ones are inputs on
    # which the reference itself exits, "path has gone awry"); wide post-gap rows go through the looped-group kernel
    assert n == 120 and ok >= 115


def test_per_job_bands_and_lengths(aligner):
    b = cases.guided_batch(seed=41, n=64, lo=50, hi=6000)
    b.band = np.random.default_rng(1).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    _run(aligner, b, 1, fn, 16)
    _run(aligner, b, 0, fn, 16)


def test_quality_value_score_function(aligner):
    b = cases.guided_batch(seed=52, n=24, lo=100, hi=2000, with_qual=True, n_rate=0.01)
    fn = QualityValueScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    _run(aligner, b, 0, fn, 16)
    _run(aligner, b, 1, fn, 16)


def test_quality_value_all_classes(aligner):
    """The QV rides in RowInfo::cd8 through the register-ring kernels of every job class (bands 8..64, reads up to 6 kb)."""
    b = cases.guided_batch(seed=53, n=48, lo=300, hi=6000, with_qual=True)
    b.band = np.random.default_rng(2).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    b.qual[::7] = 0; b.qual[3::11] = 255            # extreme QVs: zero-cost and maximal-cost mismatches
    fn = QualityValueScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    for at in (0, 1):
        _run(aligner, b, 0, fn, 16, at)
        _run(aligner, b, 1, fn, 16, at)


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("with_del", [True, False])
def test_ids_score_function(aligner, algo, with_del):
    """BASELINE configs[3]: rich QV tracks (ins/del/sub QVs + tags) through IDSScoreFunction, function-level parity."""
    rng = np.random.default_rng(17 + algo)
    for rep in range(3):
        b = cases.guided_batch(seed=520 + rep, n=16, lo=100, hi=2500, n_rate=0.01, lower=(rep == 1),
                               adversarial=0.3 if rep == 2 else 0.0, run=10)
        cases.add_ids_tracks(b, 77 + rep, with_del)
        fn = IDSScoreFunction(ins=int(rng.integers(1, 9)), del_=int(rng.integers(1, 9)), affineOpen=int(rng.choice([0, 5, 30])),
                              affineExtend=int(rng.choice([0, 1, 3])), substitutionPrior=int(rng.choice([20, 7])),
                              globalDeletionPrior=int(rng.choice([13, 4])))
        for at in (0, 1):
            res, _ = _run(aligner, b, algo, fn, int(rng.choice([8, 16, 32])), at)
            assert (res.results["status"] == 0).sum() >= b.n - 2


def test_ids_needs_tracks(aligner):
    from blasr_b200 import BgpuError
    b = cases.guided_batch(seed=1, n=2, lo=100, hi=200)
    with pytest.raises(BgpuError):
        aligner.GuidedAlign(b, IDSScoreFunction(), 16)
    cases.add_ids_tracks(b, 1)
    with pytest.raises(BgpuError):   # SWAlign x IDS reads out of bounds in the reference (SWAlign.h:166-167)
        aligner.SWAlign(b, IDSScoreFunction())


def test_edge_cases(aligner):
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    from blasr_b200 import JobBatch
    # empty guide -> EMPTY_GUIDE, score 0, empty alignment (GuidedAlign.h:388-392); tiny single-block jobs
    qs = [b"ACGTACGT", b"A", b"ACGTTTGA", b"ACGT"]
    ts = [b"ACGTACGT", b"A", b"ACGAATGA", b"ACXT"]
    gs = [np.zeros((0, 3), np.uint32), np.array([[0, 0, 1]], np.uint32), np.array([[0, 0, 3], [5, 5, 3]], np.uint32),
          np.array([[0, 0, 4]], np.uint32)]
    b = JobBatch.from_lists(qs, ts, gs)
    res = aligner.AffineGuidedAlign(b, fn, 16)
    assert res.results["status"][0] == capi.JOB_EMPTY_GUIDE and res.results["score"][0] == 0 and res.results["nBlocks"][0] == 0
    assert res.results["status"][3] == capi.JOB_BAD_INPUT          # 'X' is outside ThreeBit 0..4
    ofn = O.score_fn(fn.scoreMatrix, 5, 5, 50, 0)
    for i in (1, 2):
        j, keep = O.make_job(1, 1, 16, b.q[int(b.qOff[i]):int(b.qOff[i + 1])], b.t[int(b.tOff[i]):int(b.tOff[i + 1])], gs[i], None, 0, 0, 1, 1)
        want = O.align(WHICH, ofn, j)
        assert not cases.compare(cases.gpu_to_dict(res, i), want, cases.GPU_FIELDS)
    # empty batch
    e = JobBatch.from_lists([], [], [])
    assert len(aligner.AffineGuidedAlign(e, fn, 16)) == 0


def test_long_reads_properties(aligner):
    """BASELINE-size jobs (10-30 kb): exact vs oracle on a few, plus size-independent properties on all:
    blocks tile the path monotonically, rescoring the returned alignment reproduces the DP score for linear gaps."""
    b = cases.guided_batch(seed=61, n=12, lo=10000, hi=30000)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    res, want = _run(aligner, b, 1, fn, 16)
    res0, want0 = _run(aligner, b, 0, fn, 16)
    for i in range(b.n):
        al = res0.alignment(i)
        blk = al.blocks.astype(np.int64)
        assert (np.diff(blk[:, 0]) >= blk[:-1, 2]).all() and (np.diff(blk[:, 1]) >= blk[:-1, 2]).all()
        # GuidedAlign (linear gaps): stats score over the whole returned alignment == DP score when the path has no
        # leading/trailing gaps (first guide block at (0,0), last at the ends)
        if al.qPos == 0 and al.tPos == 0 and blk[-1, 0] + blk[-1, 2] == b.qOff[i + 1] - b.qOff[i]:
            assert al.statsScore == al.score


def test_concurrent_contexts_match_single_context(aligner):
    """Host threads with their own contexts (blasr's MapReads pthreads) behind the per-device phase gates: every sub-batch
    comes back exactly as the single-context run returns it."""
    import threading
    from blasr_b200 import Aligner
    b = cases.guided_batch(seed=77, n=96, lo=200, hi=4000)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    want = aligner.AffineGuidedAlign(b, fn, 16)
    parts = [list(range(i, b.n, 8)) for i in range(8)]
    got, errs = {}, []

    def work(a, mine):
        try:
            for p in mine:
                got[p] = a.AffineGuidedAlign(b.slice(parts[p]), fn, 16)
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    workers = [Aligner(0) for _ in range(4)]
    th = [threading.Thread(target=work, args=(a, [w, w + 4])) for w, a in enumerate(workers)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for a in workers:
        a.close()
    assert not errs, errs
    for p, idx in enumerate(parts):
        for k, i in enumerate(idx):
            bad = cases.compare(cases.gpu_to_dict(got[p], k), cases.gpu_to_dict(want, i), cases.GPU_FIELDS)
            assert not bad, (p, i, bad)


def _embed(b, seed, pad_lo=1, pad_hi=300):
    """The same pairs with random flanks around q and t and the guide shifted accordingly: the first guide block then
    starts at (qStart, tStart) > (0, 0) and the last one ends before the ends of the sequences -- guides are used raw
    (GuidedAlign.h:115-118), the alignment runs from guide.front() to guide.back()."""
    from blasr_b200 import JobBatch
    rng = np.random.default_rng(seed)
    qs, ts, gs = [], [], []
    for i in range(b.n):
        q, t, g, _ = cases.job_arrays(b, i)
        pads = [cases.ACGT[rng.integers(0, 4, int(rng.integers(pad_lo, pad_hi)))] for _ in range(4)]
        if i % 5 == 0:
            pads[0] = pads[0][:0]            # tStart > 0 with qStart == 0 and the reverse
        if i % 5 == 1:
            pads[2] = pads[2][:0]
        g = g.copy(); g[:, 0] += len(pads[0]); g[:, 1] += len(pads[2])
        qs.append(np.concatenate([pads[0], q, pads[1]]).tobytes()); ts.append(np.concatenate([pads[2], t, pads[3]]).tobytes()); gs.append(g)
    return JobBatch.from_lists(qs, ts, gs)


@pytest.mark.parametrize("algo", [0, 1])
def test_guides_with_offsets_on_a_dirty_cache(aligner, algo):
    """Guides that start after (0,0) and end before the sequence ends, run right after a ticket that left raw ASCII in the
    cached device blocks: the boundary column t' = 0 and the columns past the guide's end are staged by the fill kernels and
    must read as valid codes whatever the allocation held before."""
    from blasr_b200 import Aligner
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    a = Aligner(0)
    try:
        for rep in range(3):
            dirty = cases.guided_batch(seed=900 + rep, n=24, lo=200, hi=3000)
            cases.add_ids_tracks(dirty, 5)
            a.GuidedAlign(dirty, IDSScoreFunction(), 16)          # keeps raw bytes ('A' = 65 ...) in its tc block
            b = _embed(cases.guided_batch(seed=910 + rep, n=24, lo=150, hi=2800, n_rate=0.01,
                                          adversarial=0.3 if rep == 2 else 0.0, run=8), seed=rep)
            for at in (0, 1):
                res, _ = _run(a, b, algo, fn, [8, 16, 32][rep], at)
                assert (res.results["status"] == 0).all()
    finally:
        a.close()


def test_traceback_pool_waves(monkeypatch):
    """A 1 MB traceback pool cuts the ticket into many waves (the pool is reused wave after wave): same results as one wave."""
    from blasr_b200 import Aligner
    b = cases.guided_batch(seed=88, n=64, lo=500, hi=5000)
    b.band = np.random.default_rng(3).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    big = Aligner(0)
    monkeypatch.setenv("BGPU_ARROW_POOL_MB", "1")
    small = Aligner(0)
    try:
        for algo in (0, 1):
            want = big.AffineGuidedAlign(b, fn, 16) if algo else big.GuidedAlign(b, fn, 16)
            got = small.AffineGuidedAlign(b, fn, 16) if algo else small.GuidedAlign(b, fn, 16)
            assert (got.results["status"] == 0).all()
            for i in range(b.n):
                bad = cases.compare(cases.gpu_to_dict(got, i), cases.gpu_to_dict(want, i), cases.GPU_FIELDS)
                assert not bad, (algo, i, bad)
    finally:
        big.close(); small.close()


@pytest.mark.parametrize("algo", [0, 1])
def test_packed_guides_and_compact_results(aligner, algo, monkeypatch):
    """The PCIe-lean forms -- guides as three bytes per block (+ a side list for blocks that do not fit a byte), results as
    run-length paths -- give exactly the alignments of the plain forms; wide guide gaps exercise the side list."""
    from blasr_b200 import capi
    from blasr_b200.align import pack_guide
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    b = cases.guided_batch(seed=640 + algo, n=48, lo=100, hi=5000, n_rate=0.01, adversarial=0.4, run=300, min_block=1)
    b = _embed(b, seed=3, pad_lo=1, pad_hi=600)                        # first blocks at offsets > 255 as well
    gp, gw = pack_guide(b.guide, b.guideOff)
    assert len(gw) > 0 and len(gp) == 3 * len(b.guide)
    a = capi.AFFINE_GUIDED if algo else capi.GUIDED
    tk0 = aligner.submit(b, fn, a, band=16)
    want = aligner.collect(tk0, copy=True)
    aligner.release(tk0)
    for compact, packed in ((True, False), (False, True), (True, True)):
        tk = aligner.submit(b, fn, a, band=16, compact=compact, packed=packed)
        got = aligner.collect(tk, copy=True)
        assert (got.runs is not None) == compact
        for i in range(b.n):
            bad = cases.compare(cases.gpu_to_dict(got, i), cases.gpu_to_dict(want, i), cases.GPU_FIELDS)
            assert not bad, (compact, packed, i, bad)
        if compact:                                                    # the formatting kernels read the device-side path either way
            ops, off = aligner.cigar(tk)
            assert int(off[-1]) == len(ops) > 0
        aligner.release(tk)
    assert got.timing.d2hBytes < want.timing.d2hBytes and got.timing.h2dBytes < want.timing.h2dBytes


def test_one_process_drives_every_device():
    """One process, threads x contexts x devices (how a pthread blasr would drive a multi-GPU box): jobs are dealt out by
    -start / -stride over 2 worker threads per visible device, every worker owns a context on its device, and the per-job
    records merged back into read order equal the single-context run and the oracle."""
    from blasr_b200 import Aligner, capi, shard
    ndev = capi.lib().bgpu_device_count()
    b = cases.guided_batch(seed=812, n=61, lo=200, hi=5000)
    b.band = np.random.default_rng(4).choice([8, 16, 32, 64], size=b.n).astype(np.int32)
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)

    def run(al, sub):
        return al.AffineGuidedAlign(sub, fn, 16).results.copy()
    merged = shard.run_on_devices(b, run, devices=list(range(ndev)), threads_per_device=2)
    al = Aligner(0)
    try:
        whole = al.AffineGuidedAlign(b, fn, 16).results
        for f in ("status", "score", "qPos", "tPos", "nCells", "nMatch", "nMismatch", "nIns", "nDel", "statsScore", "nBlocks", "nGaps"):
            assert np.array_equal(merged[f], whole[f]), f
    finally:
        al.close()
    ofn = O.score_fn(fn.scoreMatrix, 5, 5, 50, 0)
    want = cases.oracle_batch(WHICH, b, ofn, 1, 1, b.band, statsAffine=1)
    assert [int(x) for x in merged["score"]] == [w["score"] for w in want]


def test_targets_from_the_resident_reference(aligner):
    """bgpu_set_reference + bgpu_batch.tRefOff / tRefRc: targets gathered on the device from the resident genome (forward and
    reverse-complemented windows) give exactly the results of the same targets uploaded as bytes."""
    rng = np.random.default_rng(31)
    genome = cases.ACGT[rng.integers(0, 4, 24 * 4600 + 100)].copy()
    genome[rng.random(len(genome)) < 0.002] = ord("N")
    genome[rng.random(len(genome)) < 0.1] |= 0x20                 # soft-masked stretches
    comp = np.arange(256, dtype=np.uint8)
    for a, c in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[a] = c
    b = cases.guided_batch(seed=77, n=24, lo=200, hi=3000)
    # re-seat every target as a window of the genome: plant the job's target (or its reverse complement) there
    assert np.diff(b.tOff.astype(np.int64)).max() < 4000
    starts = (np.arange(b.n) * 4600 + rng.integers(0, 500, b.n)).astype(np.uint64)      # disjoint windows (targets < 4,000 b)
    rc = (rng.random(b.n) < 0.5).astype(np.uint8)
    for i in range(b.n):
        t = b.t[int(b.tOff[i]):int(b.tOff[i + 1])]
        genome[int(starts[i]):int(starts[i]) + len(t)] = comp[t][::-1] if rc[i] else t
    fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
    want = aligner.AffineGuidedAlign(b, fn, 16)
    aligner.set_reference(genome)
    try:
        b.tRefOff = starts; b.tRefRc = rc
        got = aligner.AffineGuidedAlign(b, fn, 16)
    finally:
        b.tRefOff = None
        aligner.set_reference(None)
    for i in range(b.n):
        bad = cases.compare(cases.gpu_to_dict(got, i), cases.gpu_to_dict(want, i), cases.GPU_FIELDS)
        assert not bad, (i, int(rc[i]), bad)
    assert (got.results["status"] == 0).all() and rc.sum() > 3 and (1 - rc).sum() > 3
