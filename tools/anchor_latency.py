"""Diagnostic (GPU box): latency of one synchronous bgpu_map_reads call for a read and its reverse complement (what
baseline/gpu_refine.hpp::BgpuMapReadToGenome issues per read), single thread and from several threads with a context each."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from blasr_b200 import Aligner, saindex, synth

g = synth.simulate_genome(4_600_000, seed=1)
sa = saindex.suffix_array(g)
start, end = saindex.lookup_table(g, sa, 8)
reads, off = synth.simulate_reads(g, 400, 10000, seed=3)
al = Aligner(0)
al.set_reference(g); al.set_suffix_array(sa, start, end, 8)
pairs = [(reads[int(off[2 * i]):int(off[2 * i + 2])].copy(), (off[2 * i:2 * i + 3] - off[2 * i]).astype(np.uint64)) for i in range(200)]
for r, o in pairs[:20]:
    al.MapReadToGenome(r, o)
t0 = time.perf_counter()
for r, o in pairs:
    al.MapReadToGenome(r, o)
dt = time.perf_counter() - t0
print(f"1 thread: {1e3 * dt / len(pairs):.3f} ms per call (2 strands of 10 kb)", al.map_timing()[:2])
for nt in (4, 16):
    als = [Aligner(0) for _ in range(nt)]
    def work(a):
        for r, o in pairs[:20]:
            a.MapReadToGenome(r, o)
    th = [threading.Thread(target=work, args=(a,)) for a in als]
    [t.start() for t in th]; [t.join() for t in th]
    def work2(a):
        for r, o in pairs:
            a.MapReadToGenome(r, o)
    th = [threading.Thread(target=work2, args=(a,)) for a in als]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
    print(f"{nt} threads: {1e3 * dt / len(pairs):.3f} ms per call per thread, {nt * len(pairs) / dt:.0f} read pairs/s")
    [a.close() for a in als]
