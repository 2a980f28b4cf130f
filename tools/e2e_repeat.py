"""Diagnostic (GPU box): the bench's two end-to-end legs several times in one process, pass by pass, with the contexts' allocation
counters and the host's free memory -- looking for the occasional 2.4x slower leg."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from blasr_b200 import DistanceMatrixScoreFunction, capi

def mem():
    return subprocess.run("free -m | sed -n 2p", shell=True, capture_output=True, text=True).stdout.split()[1:7]
print("host memory MB (total used free shared buff/cache available):", mem(), "cpus", len(os.sched_getaffinity(0)), flush=True)
batch = bench.make_workload(100000, 1)
keep = []
bench._pin_batch(batch, keep)
print("after pinning the shard:", mem(), flush=True)
fn = DistanceMatrixScoreFunction(ins=5, del_=5)
def barrier():
    torch.cuda.synchronize()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for rr in (False, True):
        t0 = time.perf_counter()
        e = bench.e2e_measure(0, batch, fn, capi.GUIDED, 4, int(sys.argv[2]) if len(sys.argv) > 2 else 4, 2 if not rr else 1, 3, barrier, resident_reference=rr)
        print(f"rep {rep} resident_reference={rr}: {e['sec'] * 1e3:.1f} ms/step passes {['%.1f' % x for x in e['pass_ms']]} allocs(max ctx) {e['allocs']} "
              f"leg wall {time.perf_counter() - t0:.1f} s mem {mem()[1:3]}", flush=True)
