#!/usr/bin/env python
"""tools/pipeline_bench.py -- aligned reads/s of the reference PROGRAM, stock vs GPU-refined, on the box's own host cores.

Generates configs[0] (4.6 Mb genome, 10 kb reads at 15 % error; --reads to scale), builds the suffix array with the
reference's sawriter, then times
    baseline/_ref/blasrmc      reads.fa genome.fa -sa genome.sa -sam -nproc C          (the unmodified reference)
    ... -noRefineAlignments                                                               (how much of it is refinement)
    baseline/_ref/blasrmc_gpu  ... -nproc T    for T in --gpu-threads                   (anchoring + RefineAlignments on the GPU; T MapReads
                                                                                          fibers on one pthread per core)
and prints one JSON object.  Wall times include the program's start-up (index load); `startup_s` is measured with an
empty read set so that reads/s can be quoted net of it.  Sorted SAM of every GPU run is compared with the stock run.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BL = os.path.join(ROOT, "baseline")
STOCK, GPU, SAW = (os.path.join(BL, "_ref", x) for x in ("blasrmc", "blasrmc_gpu", "sawritermc"))
RUN_TIMEOUT_S = 240      # one program run (a few seconds on configs[0]): a hung child must not hang the bench line


def run(exe, d, out, nproc, extra=(), reads="reads.fa", env=None):
    cmd = [exe, reads, "genome.fa", "-sa", "genome.sa", "-sam", "-nproc", str(nproc), "-out", out] + list(extra)
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, env=env, timeout=RUN_TIMEOUT_S)
    except subprocess.TimeoutExpired:              # the child is killed; bench.py reports the leg as unavailable and goes on
        raise SystemExit(f"{' '.join(cmd)} did not finish within {RUN_TIMEOUT_S} s")
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed: {r.stdout[-1000:]} {r.stderr[-1000:]}")
    return dt


def sam_lines(path):
    return sorted(x for x in open(path).read().splitlines() if not x.startswith("@PG"))


def measure(n_reads=1000, genome=4600000, length=10000, gpu_threads=None, workdir=None, extra=(), device=0, keep=False):
    if not all(os.path.exists(p) for p in (STOCK, GPU, SAW)):
        return {"unavailable": "baseline/_ref binaries absent"}
    cores = len(os.sched_getaffinity(0))
    gpu_threads = gpu_threads or [2 * cores, 4 * cores]   # -nproc = MapReads fibers (reads in flight) on `cores` pthreads
    tmp = workdir or tempfile.mkdtemp(prefix="bgpu_pipe_")
    subprocess.check_call([sys.executable, os.path.join(BL, "make_data.py"), "c0", tmp, "--genome", str(genome), "--reads", str(n_reads),
                           "--len", str(length)], stdout=subprocess.DEVNULL)
    subprocess.check_call([SAW, "genome.sa", "genome.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    open(os.path.join(tmp, "none.fa"), "w").write(">r\nACGT\n")
    env = dict(os.environ, BGPU_DEVICE=str(device))
    out = {"config": f"configs[0]: {genome} b genome, {n_reads} reads x {length} b at 15 % error, -sam", "host_cores": cores,
           "blasr_nproc": cores}
    run(STOCK, tmp, "warm.sam", cores, extra)                                 # page cache warm-up (index file)
    out["startup_s"] = run(STOCK, tmp, "none.sam", cores, extra, reads="none.fa")
    t_stock = run(STOCK, tmp, "stock.sam", cores, extra)
    t_norefine = run(STOCK, tmp, "norefine.sam", cores, list(extra) + ["-noRefineAlignments"])
    want = sam_lines(os.path.join(tmp, "stock.sam"))
    out["stock"] = {"wall_s": t_stock, "reads_per_s": n_reads / t_stock, "reads_per_s_net": n_reads / max(t_stock - out["startup_s"], 1e-9),
                    "no_refine_wall_s": t_norefine, "refinement_share": (t_stock - t_norefine) / max(t_stock - out["startup_s"], 1e-9)}
    run(GPU, tmp, "gpuwarm.sam", cores, extra, env=env)
    out["gpu_startup_s"] = run(GPU, tmp, "none_gpu.sam", cores, extra, reads="none.fa", env=env)
    out["gpu"] = []
    for t in gpu_threads:
        dt = run(GPU, tmp, f"gpu{t}.sam", t, extra, env=env)
        same = sam_lines(os.path.join(tmp, f"gpu{t}.sam")) == want
        out["gpu"].append({"nproc": t, "wall_s": dt, "reads_per_s": n_reads / dt,
                           "reads_per_s_net": n_reads / max(dt - out["gpu_startup_s"], 1e-9), "sam_identical_to_stock": same})
    best = max(out["gpu"], key=lambda r: r["reads_per_s_net"])
    # the same program with anchoring left to the reference's CPU code (BGPU_NO_ANCHOR): what the device anchoring adds
    dt = run(GPU, tmp, "gpu_noanchor.sam", best["nproc"], extra, env=dict(env, BGPU_NO_ANCHOR="1"))
    out["gpu_refinement_only"] = {"nproc": best["nproc"], "wall_s": dt, "reads_per_s_net": n_reads / max(dt - out["gpu_startup_s"], 1e-9),
                                  "sam_identical_to_stock": sam_lines(os.path.join(tmp, "gpu_noanchor.sam")) == want}
    out["on_device"] = "MapReadToGenome (both strands, bgpu_map_reads) and RefineAlignments (bgpu_submit / bgpu_collect); SDPAlign, clustering, mapQV, printing = the reference's CPU code"
    out["reads_per_s"] = best["reads_per_s_net"]; out["reads_per_s_stock"] = out["stock"]["reads_per_s_net"]
    out["speedup_net"] = best["reads_per_s_net"] / out["stock"]["reads_per_s_net"]
    out["sam_identical_to_stock"] = all(r["sam_identical_to_stock"] for r in out["gpu"])
    if not keep and not workdir:
        subprocess.call(["rm", "-rf", tmp])
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1000)
    ap.add_argument("--genome", type=int, default=4600000)
    ap.add_argument("--len", type=int, default=10000)
    ap.add_argument("--gpu-threads", default="")
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    th = [int(x) for x in a.gpu_threads.split(",") if x] or None
    print(json.dumps(measure(a.reads, a.genome, a.len, th, a.workdir)))
