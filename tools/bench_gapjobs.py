"""configs[4]-style microbench: the gap fills of -alignContigs (SURVEY 8a sizes: ~41 k AffineKBandAlign jobs of ~80 cells
per 1 Mb contig at 0.5 % divergence, Blasr.cpp:1064-1076 parameter pattern) through the C ABI, next to the reference
templates (oracle/_ref) replaying the same jobs on the host cores.  Diagnostic; prints one JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blasr_b200 import Aligner, JobBatch, SMRTDistanceMatrix

n = int(sys.argv[1]) if len(sys.argv) > 1 else 41000 * 8          # gap jobs of ~8 Mb of contigs
rng = np.random.default_rng(5)
ACGT = np.frombuffer(b"ACGT", np.uint8)
lens = rng.integers(2, 14, n)                                       # (|q|+1) * (2k+1) ~ 80 cells at k = 4..5
tl = np.maximum(1, lens + rng.integers(-2, 3, n))
q = ACGT[rng.integers(0, 4, int(lens.sum()))]; t = ACGT[rng.integers(0, 4, int(tl.sum()))]
qOff = np.zeros(n + 1, np.uint64); qOff[1:] = np.cumsum(lens); tOff = np.zeros(n + 1, np.uint64); tOff[1:] = np.cumsum(tl)
band = np.maximum(np.abs(lens - tl) + 3, 4).astype(np.int32)       # bandSize of AlignSubstring grows with the length difference
b = JobBatch(q, qOff, t, tOff, np.zeros((0, 3), np.uint32), np.zeros(n + 1, np.uint64), None, band)
indel = 5
pr = (indel + 2, indel - 3, indel + 2, indel - 1)                   # Blasr.cpp:1067-1076
al = Aligner(0)
res = al.AffineKBandAlign(b, SMRTDistanceMatrix, pr[0], pr[1], pr[2], pr[3], indel, 0)   # warm-up (allocations)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); res = al.AffineKBandAlign(b, SMRTDistanceMatrix, pr[0], pr[1], pr[2], pr[3], indel, 0); best = min(best, time.perf_counter() - t0)
tm = res.timing
cells = int(((lens + 1) * (2 * band + 1)).sum())
out = {"workload": f"{n} AffineKBandAlign gap jobs, |q| 2-13, k = |dq-dt|+3, mean {cells / n:.0f} cells", "jobs_ok": int((res.results['status'] == 0).sum()),
       "e2e_ms": best * 1e3, "e2e_jobs_per_s": n / best, "device_ms": {"prep": tm.msPrep, "fill": tm.msFill, "trace": tm.msTrace, "emit": tm.msEmit, "total": tm.msTotal},
       "device_jobs_per_s": n / (tm.msTotal * 1e-3), "cells": cells}
try:
    from tests import cases, oracle as O
    if O.have_ref():
        ofn = O.score_fn(SMRTDistanceMatrix, 5, 5)
        m = min(n, 200000); jobs, keep = [], []
        for i in range(m):
            qq, tt, _, _ = cases.job_arrays(b, i)
            j, k = O.make_job(4, 1, int(band[i]), qq, tt, None, None, 0, indel, 0, 0, affineKBand=pr); jobs.append(j); keep.append(k)
        t0 = time.perf_counter(); O.replay("ref", ofn, jobs, os.cpu_count() or 1); dt = time.perf_counter() - t0
        out["reference_jobs_per_s"] = m / dt; out["reference_cores"] = os.cpu_count()
except Exception as e:  # noqa: BLE001
    out["reference_error"] = str(e)
al.close()
print(json.dumps(out))
