"""Latency of small tickets (what one MapReads pthread, or RefineService, submits): N jobs of ~10 kb, AffineGuidedAlign band 16.
usage: small_ticket_probe.py [jobs per ticket ...]   (run on the GPU box)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blasr_b200 import Aligner, DistanceMatrixScoreFunction, capi, synth

sizes = [int(x) for x in sys.argv[1:]] or [1, 16, 64, 256, 1024]
fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50, affineExtend=0)
a = Aligner(0)
for n in sizes:
    b = synth.simulate_pairs(n, 10000, 10000, err=0.15, seed=3, bands=(16,))
    rows = []
    for rep in range(6):
        h0 = time.perf_counter(); tk = a.submit(b, fn, capi.AFFINE_GUIDED, band=16, doStats=True)
        h1 = time.perf_counter(); res = a.collect(tk); h2 = time.perf_counter()
        tm = res.timing; a.release(tk); h3 = time.perf_counter()
        rows.append(((h1 - h0) * 1e3, (h2 - h1) * 1e3, (h3 - h2) * 1e3, tm.msPrep, tm.msFill, tm.msTrace, tm.msEmit, tm.msTotal, tm.devAllocs, tm.pinAllocs))
    r = rows[-1]
    print("jobs %5d  submit %7.2f collect %7.2f release %6.2f | prep %6.2f fill %6.2f trace %6.2f emit %6.2f total %6.2f | allocs %d %d | first-rep submit %.1f collect %.1f"
          % ((n,) + r + (rows[0][0], rows[0][1])))
