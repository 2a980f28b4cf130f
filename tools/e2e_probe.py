"""Per-sub-batch timeline of the pipelined e2e path (diagnostic, run on the GPU box)."""
import sys, os, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from blasr_b200 import Aligner, DistanceMatrixScoreFunction, capi

jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C = int(sys.argv[3]) if len(sys.argv) > 3 else 12
STAG = float(sys.argv[4]) * 1e-3 if len(sys.argv) > 4 else 0.0
QUAL = os.environ.get("SCOREFN") == "quality"
batch = bench.make_workload(jobs, 1, with_qual=QUAL)
keep = []
for name in ("q", "qOff", "t", "tOff", "guide", "guideOff", "band") + (("qual",) if QUAL else ()):
    v, t = bench.pinned_copy(getattr(batch, name)); setattr(batch, name, v); keep.append(t)
from blasr_b200 import QualityValueScoreFunction
fn = QualityValueScoreFunction(ins=5, del_=5) if QUAL else DistanceMatrixScoreFunction(ins=5, del_=5)
bounds = np.linspace(0, batch.n, C + 1).astype(np.int64)
chunks = [bench.range_view(batch, int(bounds[i]), int(bounds[i + 1])) for i in range(C)]
workers = [Aligner(0) for _ in range(T)]
log = []
def run_pass(record):
    nxt = iter(range(C)); lock = threading.Lock(); t00 = time.perf_counter()
    def work(w, a):
        time.sleep(w * STAG)
        while True:
            with lock:
                i = next(nxt, None)
            if i is None: return
            h0 = time.perf_counter(); tk = a.submit(chunks[i], fn, capi.GUIDED, band=16, doStats=True)
            h1 = time.perf_counter(); res = a.collect(tk); h2 = time.perf_counter()
            tm = res.timing
            a.release(tk); h3 = time.perf_counter()
            if record: log.append((w, i, (h0 - t00) * 1e3, (h1 - h0) * 1e3, (h2 - h1) * 1e3, (h3 - h2) * 1e3, tm.msHostSubmit, tm.msHostCollect, tm.msPrep, tm.msFill, tm.msTrace, tm.msEmit, tm.devAllocs, tm.pinAllocs))
    th = [threading.Thread(target=work, args=(w, a)) for w, a in enumerate(workers)]
    for x in th: x.start()
    for x in th: x.join()
    return (time.perf_counter() - t00) * 1e3
for _ in range(2): run_pass(False)
ms = run_pass(True)
print("pass ms", ms)
print("w  i   start  submit collect release | c_submit c_collect | prep fill trace emit")
for r in sorted(log, key=lambda x: x[2]):
    print("%d %2d %7.1f %6.1f %6.1f %6.1f | %6.1f %6.1f | %5.1f %5.1f %5.1f %5.1f | allocs %d %d" % r)
