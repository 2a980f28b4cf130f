"""GPU: the CUDA path against the committed golden vectors (reference outputs), through the C ABI."""
import numpy as np
import pytest

from blasr_b200 import DistanceMatrixScoreFunction, JobBatch, capi
from . import cases, golden_io

pytestmark = pytest.mark.gpu


def test_cuda_matches_golden(aligner):
    recs, fields = golden_io.load()
    # group cases that share every batch-level parameter into one submission
    groups = {}
    for i, r in enumerate(recs):
        p = r["p"]
        key = (p["algo"], p["at"], p["bndIns"], p["bndDel"], p["statsAffine"], p["doStats"], p["ins"], p["del_"], p["open"],
               p["ext"], p["kind"], p["M"].tobytes())
        groups.setdefault(key, []).append(i)
    checked = 0
    for key, idx in groups.items():
        p = recs[idx[0]]["p"]
        qs = [recs[i]["q"].tobytes() for i in idx]; ts = [recs[i]["t"].tobytes() for i in idx]
        gs = [recs[i]["guide"] if recs[i]["guide"] is not None else np.zeros((0, 3), np.uint32) for i in idx] if p["algo"] < 2 else None
        qv = [recs[i]["qual"] if recs[i]["qual"] is not None else np.zeros(len(recs[i]["q"]), np.uint8) for i in idx] if p["kind"] else None
        b = JobBatch.from_lists(qs, ts, gs, qv, [recs[i]["p"]["band"] for i in idx])
        fn = DistanceMatrixScoreFunction(p["M"].reshape(5, 5), p["ins"], p["del_"], p["open"], p["ext"], p["kind"])
        tk = aligner.submit(b, fn, p["algo"], alignType=p["at"], band=0, bndIns=p["bndIns"], bndDel=p["bndDel"],
                            doStats=bool(p["doStats"]), statsAffine=bool(p["statsAffine"]))
        res = aligner.collect(tk, copy=True); aligner.release(tk)
        for j, i in enumerate(idx):
            bad = cases.compare(cases.gpu_to_dict(res, j), recs[i]["want"], fields + ["pctSimilarity"])
            assert not bad, f"golden case {i}: {bad}"
            checked += 1
    assert checked == len(recs)
