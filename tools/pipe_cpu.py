"""Diagnostic (GPU box): wall / user / sys of stock and GPU-refined blasr at several -nproc."""
import os, resource, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "baseline", "_ref")
W = "/tmp/pp"
subprocess.check_call([sys.executable, os.path.join(ROOT, "baseline", "make_data.py"), "c0", W, "--reads", "2000"], stdout=subprocess.DEVNULL)
subprocess.check_call([os.path.join(B, "sawritermc"), "genome.sa", "genome.fa"], cwd=W, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
def run(exe, nproc, extra=(), env=None):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN); t0 = time.perf_counter()
    p = subprocess.run([os.path.join(B, exe), "reads.fa", "genome.fa", "-sa", "genome.sa", "-sam", "-nproc", str(nproc), "-out", "o.sam"] + list(extra),
                       cwd=W, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    dt = time.perf_counter() - t0; r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    stats = [l for l in p.stderr.splitlines() if "RefineService" in l]
    print(f"{exe:12s} nproc {nproc:4d} {' '.join(extra):22s} wall {dt:6.2f} user {r1.ru_utime - r0.ru_utime:6.2f} sys {r1.ru_stime - r0.ru_stime:6.2f} "
          f"vol-ctx {r1.ru_nvcsw - r0.ru_nvcsw} invol-ctx {r1.ru_nivcsw - r0.ru_nivcsw} {stats[0][15:] if stats else ''}", flush=True)
run("blasrmc", 16)
for n in (16, 32, 64):
    run("blasrmc", n)
    run("blasrmc", n, ["-noRefineAlignments"])
for n, e in ((16, {}), (64, {}), (64, {"BGPU_BLOCKING_SYNC": "0"}), (128, {"BGPU_SERVICE_CONTEXTS": "6"})):
    run("blasrmc_gpu", n, env=dict(e, BGPU_SERVICE_STATS="1"))
