"""One ticket of J pairs: submit, collect, one rerun.  Target of the ncu captures (every kernel launches twice,
with the grids of a whole-shard ticket).  Numbers printed here are never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from blasr_b200 import Aligner, DistanceMatrixScoreFunction, capi

jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
algo = capi.AFFINE_GUIDED if (len(sys.argv) > 2 and sys.argv[2] == "affine") else capi.GUIDED
batch = bench.make_workload(jobs, 1)
fn = DistanceMatrixScoreFunction(ins=5, del_=5, affineOpen=50 if algo == capi.AFFINE_GUIDED else 0, affineExtend=0)
al = Aligner(0)
tk = al.submit(batch, fn, algo, band=16, doStats=True)
res = al.collect(tk)
t = al.rerun(tk)
print("cells", int(res.timing.cells), "prep %.3f fill %.3f trace %.3f emit %.3f total %.3f ms" % (t.msPrep, t.msFill, t.msTrace, t.msEmit, t.msTotal))
al.release(tk); al.close()
