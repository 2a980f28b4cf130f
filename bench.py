#!/usr/bin/env python
"""bench.py -- banded-DP GCUPS of the refinement hot path on N B200s (BASELINE.json metric).

Workload = BASELINE.json configs[1]: "GuidedAlign kernel microbench: 100k read/window pairs 1-20 kb,
band width 16-64, DistanceMatrixScoreFunction" (per GPU; reads shard over GPUs with no collective, so scaling
is weak: every rank aligns its own 100k-pair shard).  One step = one pass of the whole hot path
(guide construction, DP fill, traceback, block/gap/stats emission) over the shard.

  value      GCUPS with inputs resident in HBM (bgpu_rerun), cells = sum of the reference's nCells
  e2e        GCUPS through the C ABI with pinned HOST buffers: H2D + kernels + D2H every step
  roofline   the fill kernel against the HBM roofline the contract asks for (algorithmic 0.25 B/cell) plus
             int_roofline: cells/s x 8 int32 ops/cell against the int32 peak measured on this device
  cpu_baseline  the unmodified reference (oracle/_ref) or the C port replayed on all host cores, bounded sample

`--impl reference` times the reference's own CPU implementation (same metric / config) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEN_LO, LEN_HI = 1000, 20000
BANDS = (16, 32, 64)
OPS_PER_CELL = {0: 8, 1: 16}          # linear / affine int32 ops per cell
BYTES_PER_CELL = {0: 0.25, 1: 0.625}  # algorithmic traceback bytes per cell


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--jobs", type=int, default=int(os.environ.get("BGPU_BENCH_JOBS", 100000)), help="pairs per GPU")
    ap.add_argument("--algo", default="guided", choices=["guided", "affine"])
    ap.add_argument("--scorefn", default="distance", choices=["distance", "quality"],
                    help="distance = configs[1] (the headline); quality = configs[3]-style QualityValueScoreFunction over a simulated QV track")
    ap.add_argument("--len-lo", type=int, default=LEN_LO); ap.add_argument("--len-hi", type=int, default=LEN_HI)
    ap.add_argument("--bands", default=",".join(str(b) for b in BANDS), help="band sizes drawn per job (configs[1]: 16,32,64; blasr's default -bandSize: 16)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--e2e-threads", type=int, default=5, help="host threads (one context each) of the e2e measurement")
    ap.add_argument("--e2e-chunks", type=int, default=10, help="sub-batches the shard is cut into for the e2e measurement")
    return ap.parse_args()


def make_workload(n_jobs, seed, with_qual=False, len_lo=LEN_LO, len_hi=LEN_HI, bands=BANDS):
    from blasr_b200 import synth
    return synth.simulate_pairs(n_jobs, len_lo, len_hi, err=0.15, seed=seed, bands=bands, with_qual=with_qual)


def workload_args(args):
    return dict(len_lo=args.len_lo, len_hi=args.len_hi, bands=tuple(int(x) for x in args.bands.split(",")))


def config_dict(args, n_jobs):
    wl = ("configs[1]: GuidedAlign microbench, read/window pairs 1-20 kb, band 16/32/64, "
          "DistanceMatrixScoreFunction(SMRTDistanceMatrix, ins=5, del=5)")
    if (args.len_lo, args.len_hi, args.bands) != (LEN_LO, LEN_HI, ",".join(str(b) for b in BANDS)):
        wl = (f"configs[1] generator at read/window pairs {args.len_lo}-{args.len_hi} b, band {args.bands} "
              "(10 kb / band 16 = the refinement jobs of configs[0]), DistanceMatrixScoreFunction(SMRTDistanceMatrix, ins=5, del=5)")
    if args.scorefn == "quality":
        wl = ("configs[3]-style: the configs[1] pairs with a simulated QV track (clamp(N(12,4),1,93)), "
              "QualityValueScoreFunction(ins=5, del=5)")
    return {"workload": wl,
            "pairs_per_gpu": n_jobs, "algo": "AffineGuidedAlign" if args.algo == "affine" else "GuidedAlign",
            "error_rate": 0.15, "guide": "all diagonal runs of the simulated alignment (detailed-SDP-like)",
            "l2": "inputs_larger_than_L2", "parallelism": f"read-shard x{args.gpus}, no collective"}


# ---------------------------------------------------------------- CPU side (reference / port)
def cpu_replay(batch, algo, n_threads, target_seconds, est_gcups_per_core=0.05, quality=False):
    """Replays a bounded prefix of the batch through oracle/_ref (or the C port) on n_threads; returns dict."""
    from tests import cases, oracle as O
    which = "ref" if O.have_ref() else "orc"
    fn = O.score_fn(__import__("blasr_b200").SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0, kind=1 if quality else 0)
    target_cells = target_seconds * n_threads * est_gcups_per_core * 1e9 * (0.5 if algo else 1.0)
    jobs, keep, est = [], [], 0
    for i in range(batch.n):
        q, t, g, qv = cases.job_arrays(batch, i)
        j, k = O.make_job(algo, 1, int(batch.band[i]), q, t, g, qv if quality else None, 0, 0, 0, 0)
        jobs.append(j); keep.append(k)
        est += len(q) * (2 * int(batch.band[i]) + 2)
        if est >= target_cells and len(jobs) >= n_threads:
            break
    t0 = time.perf_counter()
    cells, _ = O.replay(which, fn, jobs, n_threads)
    dt = time.perf_counter() - t0
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": n_threads, "kind": "reference" if which == "ref" else "port",
            "sample": f"first {len(jobs)} pairs of the rank-0 shard ({cells} cells) replayed once in {dt:.1f} s",
            "_cells": cells, "_seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    algo = 1 if args.algo == "affine" else 0
    # a shard prefix is enough: the sample is bounded by CPU time, not by the 100k pairs
    quality = args.scorefn == "quality"
    batch = make_workload(min(args.jobs, max(256, n_threads * 32)), args.seed, with_qual=quality, **workload_args(args))
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_replay(batch, algo, n_threads, min(per_step, 2.0), quality=quality)
    vals, ms = [], []
    for _ in range(args.steps):
        r = cpu_replay(batch, algo, n_threads, per_step, quality=quality)
        vals.append(r["value"]); ms.append(r["_seconds"] * 1e3)
    v = float(np.mean(vals))
    cb = {k: r[k] for k in ("unit", "cores", "kind", "sample")}
    cb["value"] = v
    print(json.dumps({"impl": "reference", "metric": "banded_dp_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                      "config": config_dict(args, args.jobs), "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------- GPU side
def bind_to_gpu_cpus(index):
    """Binds this process (and the host threads it starts) to the CPUs NVML reports as local to GPU `index`, so the
    pinned input / result buffers are allocated on that GPU's NUMA node and the copies do not cross sockets.
    Returns the number of CPUs bound to, or None when NVML gives no answer."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


def pinned_copy(a):
    import torch
    a = np.ascontiguousarray(a)
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
    v = t.numpy()[:a.nbytes].view(a.dtype).reshape(a.shape)
    v[...] = a
    return v, t


def range_view(batch, a, b):
    """Jobs [a, b) of a JobBatch as views into the same (pinned) arrays, offsets rebased."""
    from blasr_b200 import JobBatch
    q0, q1 = int(batch.qOff[a]), int(batch.qOff[b]); t0, t1 = int(batch.tOff[a]), int(batch.tOff[b])
    g0, g1 = int(batch.guideOff[a]), int(batch.guideOff[b])
    return JobBatch(batch.q[q0:q1], batch.qOff[a:b + 1] - batch.qOff[a], batch.t[t0:t1], batch.tOff[a:b + 1] - batch.tOff[a],
                    batch.guide[g0:g1], batch.guideOff[a:b + 1] - batch.guideOff[a],
                    batch.qual[q0:q1] if batch.qual is not None else None, batch.band[a:b] if batch.band is not None else None)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from blasr_b200 import Aligner, DistanceMatrixScoreFunction, QualityValueScoreFunction, capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # the generator forks worker processes: run it before CUDA / NCCL are initialised in this process
    batch = make_workload(args.jobs, args.seed + 1000 * rank, with_qual=args.scorefn == "quality", **workload_args(args))
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local)     # before any pinned allocation: first touch then lands on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    algo = capi.AFFINE_GUIDED if args.algo == "affine" else capi.GUIDED
    # inputs in pinned host memory (the library then DMA's straight from them)
    keep = []
    for name in ("q", "qOff", "t", "tOff", "guide", "guideOff", "band") + (("qual",) if batch.qual is not None else ()):
        v, t = pinned_copy(getattr(batch, name)); setattr(batch, name, v); keep.append(t)
    fn_cls = QualityValueScoreFunction if args.scorefn == "quality" else DistanceMatrixScoreFunction
    fn = fn_cls(ins=5, del_=5, affineOpen=50 if algo == capi.AFFINE_GUIDED else 0, affineExtend=0)
    al = Aligner(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: submit (H2D + kernels) + collect (D2H) through the C ABI, host buffers in, host results out.
    # The way a multi-threaded host (blasr's MapReads pthreads) drives the library: a few host threads, each with its
    # own context, push sub-batches of the shard; copies of one sub-batch overlap the kernels of another.
    e2e_host = {}
    n_chunks = max(1, min(args.e2e_chunks, batch.n))
    bounds = np.linspace(0, batch.n, n_chunks + 1).astype(np.int64)
    chunks = [range_view(batch, int(bounds[i]), int(bounds[i + 1])) for i in range(n_chunks)]
    workers = [Aligner(local) for _ in range(max(1, min(args.e2e_threads, n_chunks)))]

    def e2e_pass():
        nxt = iter(range(n_chunks)); lock = threading.Lock()
        tot = {"cells": 0, "ok": 0, "h2d": 0, "d2h": 0}
        errs = []

        def work(a):
            try:
                while True:
                    with lock:
                        i = next(nxt, None)
                    if i is None:
                        return
                    tk = a.submit(chunks[i], fn, algo, band=16, doStats=True)
                    res = a.collect(tk)
                    with lock:
                        tot["cells"] += int(res.timing.cells); tot["ok"] += int((res.results["status"] == 0).sum())
                        tot["h2d"] += int(res.timing.h2dBytes); tot["d2h"] += int(res.timing.d2hBytes)
                    a.release(tk)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(a,)) for a in workers]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]
        return tot
    e2e_warm = max(1, min(args.warmup, 2))
    for _ in range(e2e_warm):
        e2e_pass()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        tot = e2e_pass()
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    for a in workers:
        a.close()
    e2e_cells, h2d, d2h = tot["cells"], tot["h2d"], tot["d2h"]
    e2e_host.update(threads=len(workers), sub_batches=n_chunks)

    # one ticket for the whole shard: the device-resident measurement below re-runs it; its own submit -> collect time
    # is reported as e2e.single_ticket (no copy/compute overlap)
    h0 = time.perf_counter()
    tk = al.submit(batch, fn, algo, band=16, doStats=True)
    h1 = time.perf_counter()
    res = al.collect(tk)
    h2 = time.perf_counter()
    e2e_host.update(single_submit_ms=(h1 - h0) * 1e3, single_collect_ms=(h2 - h1) * 1e3)
    tm = res.timing
    cells = int(tm.cells)
    assert cells == e2e_cells, (cells, e2e_cells)
    ok = int((res.results["status"] == 0).sum())

    # ---- device-resident: re-run every kernel of the ticket on inputs already in HBM
    for _ in range(args.warmup):
        al.rerun(tk)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ms_total, ms_fill, ms_trace, ms_prep, ms_emit, launches = [], [], [], [], [], 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        t = al.rerun(tk)
        ms_total.append(t.msTotal); ms_fill.append(t.msFill); ms_trace.append(t.msTrace); ms_prep.append(t.msPrep); ms_emit.append(t.msEmit)
        launches += int(t.kernelLaunches)
    barrier()
    wall = time.perf_counter() - w0
    sampler.stop_flag = True; sampler.join(2)
    dev_ms = float(np.sum(ms_total))

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); return float(tt.item())

    def allsum(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM); return float(tt.item())
    dev_ms_max = allmax(dev_ms); e2e_sec_max = allmax(e2e_sec); cells_all = allsum(float(cells)); jobs_all = allsum(float(ok))
    value = cells_all * args.steps / (dev_ms_max * 1e-3) / 1e9
    e2e_val = cells_all / e2e_sec_max / 1e9
    int_peak, _ = al.int_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    a = 1 if algo == capi.AFFINE_GUIDED else 0
    # DRAM bytes of the fill kernels from the committed ncu --set full capture of this very workload (same pairs, same seed)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "fill_traffic.json")))[args.algo]
        default_wl = (args.len_lo, args.len_hi, args.bands) == (LEN_LO, LEN_HI, ",".join(str(b) for b in BANDS))
        if tr["pairs"] == args.jobs and tr["seed"] == args.seed and tr["scorefn"] == args.scorefn and default_wl:
            traffic, traffic_src = tr["dram_bytes_per_step"], tr["source"]
    except Exception:  # noqa: BLE001
        pass
    fill_s = float(np.mean(ms_fill)) * 1e-3
    fill_gcups = cells / fill_s / 1e9
    out = {
        "metric": "banded_dp_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": config_dict(args, args.jobs),
        "aligned_pairs_per_s": jobs_all * args.steps / (dev_ms_max * 1e-3),
        "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "pairs_per_s": jobs_all / e2e_sec_max, "ms_per_step": e2e_sec_max * 1e3,
                "how": f"{e2e_host['threads']} host threads x own context, {e2e_host['sub_batches']} sub-batches of the shard, pinned host buffers "
                       "in, pinned result arena out, H2D + D2H inside the timed region",
                "single_ticket": {"value": cells / ((e2e_host["single_submit_ms"] + e2e_host["single_collect_ms"]) * 1e-3) / 1e9,
                                  "submit_ms": e2e_host["single_submit_ms"], "collect_ms": e2e_host["single_collect_ms"]}},
        "gpu_launches": launches,
        "stage_ms": {"prep": float(np.mean(ms_prep)), "fill": float(np.mean(ms_fill)), "trace": float(np.mean(ms_trace)),
                     "emit": float(np.mean(ms_emit)), "wall_per_step": wall / args.steps * 1e3},
        "roofline": {"bound": "hbm", "kernel": "fill_guided_kernel", "achieved": cells * BYTES_PER_CELL[a] / fill_s / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": cells * BYTES_PER_CELL[a] / fill_s / 1e9 / hbm_peak, "traffic": traffic,
                     "algorithmic_bytes": cells * BYTES_PER_CELL[a], "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                     "note": "integer DP: the HBM roofline is not the binding one, see int_roofline"},
        "int_roofline": {"bound": "int32 ALU issue", "fill_gcups": fill_gcups, "ops_per_cell": OPS_PER_CELL[a],
                         "achieved": fill_gcups * OPS_PER_CELL[a] / 1e3, "peak": int_peak / 1e12, "unit": "Tops/s",
                         "frac": fill_gcups * 1e9 * OPS_PER_CELL[a] / int_peak if int_peak else None,
                         "lane_steps_per_cell": float(tm.fillCells) / max(1, cells),
                         "peak_by_mix_tops": {k: v / 1e12 for k, v in al.int_peak_modes.items()},
                         "peak_source": "bgpu_measure_int_peak: best of add / min / mad / add+mad chains on this device"},
        "clocks": sampler.summary(), "jobs_ok": int(jobs_all), "host_cpus_bound_per_rank": numa,
    }
    if rank == 0 and world == 1:
        try:
            os.sched_setaffinity(0, all_cpus)            # the CPU baseline gets every host core back
            cb = cpu_replay(batch, a, len(all_cpus), args.cpu_seconds, quality=args.scorefn == "quality")
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    al.release(tk); al.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
