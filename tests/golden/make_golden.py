#!/usr/bin/env python
"""Generates tests/golden/refinement_golden.npz from the UNMODIFIED reference (oracle/_ref/libblasr_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container (the reference tree is absent on the GPU box):

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path; these fixtures are its own outputs on seeded
inputs, so the oracle and the CUDA path stay pinned even where the reference library cannot travel.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blasr_b200 import SMRTDistanceMatrix  # noqa: E402
from tests import cases, oracle as O  # noqa: E402

FIELDS = ["status", "score", "qPos", "tPos", "nCells", "nMatch", "nMismatch", "nIns", "nDel", "statsScore", "nBlocks",
          "nGapLists", "nGaps"]


def main():
    assert O.have_ref(), "oracle/_ref/libblasr_ref.so missing: run make -C oracle"
    rng = np.random.default_rng(20261017)
    recs = []   # (params dict, q, t, guide, qual, result)

    def add(algo, at, band, q, t, g, qv, fnargs, bndIns=0, bndDel=0, statsAffine=0, doStats=1):
        fn = O.score_fn(*fnargs)
        j, keep = O.make_job(algo, at, band, q, t, g, qv, bndIns, bndDel, doStats, statsAffine)
        if algo >= 2 and O.align("orc", fn, j)["status"] != 0:
            return   # inputs on which the reference is undefined
        r = O.align("ref", fn, j)
        recs.append((dict(algo=algo, at=at, band=band, bndIns=bndIns, bndDel=bndDel, statsAffine=statsAffine, doStats=doStats,
                          M=np.asarray(fnargs[0], np.int32).reshape(25), ins=fnargs[1], del_=fnargs[2], open=fnargs[3],
                          ext=fnargs[4], kind=fnargs[5]), q, t, g, qv, r))

    # guided: natural + adversarial guides, linear and affine, Global and Local, both score functions
    for rep in range(14):
        b = cases.guided_batch(seed=900 + rep, n=3, lo=80, hi=1500, err=float(rng.choice([0.05, 0.15, 0.3])),
                               adversarial=float(rng.choice([0.0, 0.0, 0.3, 0.6])), run=int(rng.choice([1, 8, 40])), n_rate=0.01,
                               with_qual=True, lower=rep % 2 == 0)
        for i in range(b.n):
            q, t, g, qv = cases.job_arrays(b, i)
            M = SMRTDistanceMatrix if rep % 4 else rng.integers(-6, 8, size=(5, 5)).astype(np.int32)
            kind = int(rep % 5 == 0)
            fnargs = (M, int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.choice([0, 5, 11, 50])), int(rng.choice([0, 1, 2])), kind)
            if rep % 3 == 0:
                fnargs = (SMRTDistanceMatrix, 5, 5, 50, 0, kind)      # blasr's defaults
            for algo in (0, 1):
                add(algo, int(rng.choice([0, 1, 1])), int(rng.choice([4, 10, 16, 32, 64])), q, t, g, qv if kind else None, fnargs,
                    statsAffine=algo)
    # k-band and SW
    for rep in range(60):
        q, t = cases.random_pair(rng, 3, 180, err=0.2, n_rate=0.01)
        kind = int(rep % 4 == 0)
        qv = rng.integers(1, 60, len(q)).astype(np.uint8) if kind else None
        fnargs = (SMRTDistanceMatrix, int(rng.integers(1, 8)), int(rng.integers(1, 8)), 0, 0, kind)
        at = int(rng.choice([1, 2, 3, 7]))
        k = int(rng.integers(1, 40))
        if at in (3, 7):
            k = max(1, min(k, len(t), len(q)))
        add(2, at, k, q, t, None, qv, fnargs, int(rng.integers(1, 9)), int(rng.integers(1, 9)), doStats=int(at == 1))
        q2, t2 = q[:100], t[:100]
        at = int(rng.choice([0, 1, 2, 4, 5, 6, 8, 9]))
        add(3, at, 0, q2, t2, None, qv[:100] if kind else None, fnargs)
    out = {"n": np.int64(len(recs)), "fields": np.array(FIELDS)}
    for i, (p, q, t, g, qv, r) in enumerate(recs):
        out[f"p{i}"] = np.array([p[k] for k in ("algo", "at", "band", "bndIns", "bndDel", "statsAffine", "doStats", "ins", "del_", "open", "ext", "kind")], np.int64)
        out[f"M{i}"] = p["M"]
        out[f"q{i}"] = np.asarray(q, np.uint8); out[f"t{i}"] = np.asarray(t, np.uint8)
        out[f"g{i}"] = np.asarray(g, np.uint32).reshape(-1, 3) if g is not None else np.zeros((0, 3), np.uint32)
        out[f"v{i}"] = np.asarray(qv, np.uint8) if qv is not None else np.zeros(0, np.uint8)
        out[f"r{i}"] = np.array([int(r[k]) for k in FIELDS], np.int64)
        out[f"s{i}"] = np.float32(r["pctSimilarity"])
        out[f"b{i}"] = r["blocks"].astype(np.uint32)
        out[f"c{i}"] = np.array([len(gl) for gl in r["gaps"]], np.uint32)
        out[f"a{i}"] = np.array([x for gl in r["gaps"] for x in gl], np.int32).reshape(-1, 2)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refinement_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(recs)} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
