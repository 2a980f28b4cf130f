"""e2e pass time of the pipelined submit/collect path for several (host threads, sub-batches) settings (diagnostic)."""
import sys, os, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from blasr_b200 import Aligner, DistanceMatrixScoreFunction, capi

jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
configs = [tuple(int(x) for x in a.split("x")) for a in sys.argv[2:]] or [(3, 12), (4, 16), (6, 24)]
batch = bench.make_workload(jobs, 1)
keep = []
for name in ("q", "qOff", "t", "tOff", "guide", "guideOff", "band"):
    v, t = bench.pinned_copy(getattr(batch, name)); setattr(batch, name, v); keep.append(t)
for algo, fn in ((capi.GUIDED, DistanceMatrixScoreFunction(ins=5, del_=5)), (capi.AFFINE_GUIDED, DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50))):
    for T, Cn in configs:
        bounds = np.linspace(0, batch.n, Cn + 1).astype(np.int64)
        chunks = [bench.range_view(batch, int(bounds[i]), int(bounds[i + 1])) for i in range(Cn)]
        workers = [Aligner(0) for _ in range(T)]
        def run_pass():
            nxt = iter(range(Cn)); lock = threading.Lock(); t00 = time.perf_counter(); cells = [0]
            def work(a):
                while True:
                    with lock:
                        i = next(nxt, None)
                    if i is None: return
                    tk = a.submit(chunks[i], fn, algo, band=16, doStats=True); res = a.collect(tk)
                    with lock: cells[0] += int(res.timing.cells)
                    a.release(tk)
            th = [threading.Thread(target=work, args=(a,)) for a in workers]
            for x in th: x.start()
            for x in th: x.join()
            return (time.perf_counter() - t00) * 1e3, cells[0]
        for _ in range(2): run_pass()
        ms = [run_pass() for _ in range(3)]
        best = min(m for m, _ in ms)
        print("algo %d threads %d chunks %d: pass ms %s -> %.0f GCUPS" % (algo, T, Cn, ["%.1f" % m for m, _ in ms], ms[0][1] / best / 1e6), flush=True)
        for a in workers: a.close()
