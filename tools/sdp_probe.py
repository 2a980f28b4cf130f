"""Wall time of bgpu_sdp_align on N synthetic pairs, with and without the detailed gap fills (where does a job's time go)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from blasr_b200 import Aligner, DistanceMatrixScoreFunction
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
base = bench.make_workload(n, 5, len_lo=lo, len_hi=hi)
fn = DistanceMatrixScoreFunction(ins=5, del_=5)
al = Aligner(0)
for detailed, recurse in ((True, 2), (False, 2), (True, 0)):
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); res, blocks = al.SDPAlign(base, fn, indelRate=0.9, detailedAlignment=detailed, recurse=recurse); ts.append(time.perf_counter() - t0)
    print("pairs", n, "detailed", detailed, "recurse", recurse, "best %.3f s" % min(ts), "%.0f pairs/s" % (n / min(ts)), "ok", int((res["status"] == 0).sum()), "blocks", len(blocks))
al.close()
