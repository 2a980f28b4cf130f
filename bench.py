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
    ap.add_argument("--ref-seconds", type=float, default=None,
                    help="--impl reference: CPU time of one step's sample (default: 2-20 s so that steps + warmup end within ~2 min)")
    ap.add_argument("--e2e-threads", type=int, default=4, help="host threads (one context each) of the e2e measurement")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="sub-batches the shard is cut into for the e2e measurement")
    ap.add_argument("--no-subrecords", dest="subrecords", action="store_false", help="skip the affine / affine_production / quality / sdp_guides sub-records")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the parity samples against oracle/_ref")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false", help="skip the reads/s leg (stock blasr vs GPU-refined blasr)")
    ap.add_argument("--prod-jobs", type=int, default=40000, help="pairs of the affine_production sub-record (10 kb, band 16)")
    ap.add_argument("--sdp-jobs", type=int, default=2048, help="pairs of the sdp_guides sub-record (0 = skip)")
    ap.add_argument("--anchor-reads", type=int, default=1000, help="reads (x 2 strands) of the anchoring sub-record (configs[0] shape; 0 = skip)")
    ap.add_argument("--gap-jobs", type=int, default=328000, help="AffineKBandAlign jobs of the gap_fills sub-record (configs[4]-style; 0 = skip)")
    ap.add_argument("--pipeline-reads", type=int, default=2000, help="reads of the configs[0] pipeline run")
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
def cpu_sample(batch, n_threads, target_seconds, algo, est_gcups_per_core=0.05, seed=7):
    """A bounded, seeded RANDOM sample of the shard (not a prefix: job sizes follow the shard's own mix) worth about
    target_seconds of CPU work on n_threads cores."""
    target_cells = target_seconds * n_threads * est_gcups_per_core * 1e9 * (0.45 if algo else 1.0)
    ql = np.diff(batch.qOff.astype(np.int64))
    band = batch.band.astype(np.int64) if batch.band is not None else np.full(batch.n, 16, np.int64)
    est = ql * (2 * band + 2)
    order = np.random.default_rng(seed).permutation(batch.n)
    cum = np.cumsum(est[order])
    k = int(np.searchsorted(cum, target_cells)) + 1
    return np.sort(order[:max(min(k, batch.n), min(n_threads, batch.n))])


def cpu_replay(batch, algo, n_threads, target_seconds, quality=False, idx=None):
    """Replays a bounded sample of the batch through oracle/_ref (the unmodified reference templates; the C port when that
    build is absent) on n_threads threads, ComputeAlignmentStats included (the GPU path computes it too); returns dict."""
    from tests import cases, oracle as O
    which = "ref" if O.have_ref() else "orc"
    fn = O.score_fn(__import__("blasr_b200").SMRTDistanceMatrix, 5, 5, 50 if algo else 0, 0, kind=1 if quality else 0)
    if idx is None:
        idx = cpu_sample(batch, n_threads, target_seconds, algo)
    jobs, keep = [], []
    for i in idx:
        q, t, g, qv = cases.job_arrays(batch, int(i))
        j, k = O.make_job(algo, 1, int(batch.band[i]) if batch.band is not None else 16, q, t, g, qv if quality else None, 0, 0, 1, algo)
        jobs.append(j); keep.append(k)
    t0 = time.perf_counter()
    cells, _ = O.replay(which, fn, jobs, n_threads)
    dt = time.perf_counter() - t0
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": n_threads, "kind": "reference" if which == "ref" else "port",
            "sample": f"{len(jobs)} pairs drawn at random (seeded) from the rank-0 shard ({cells} cells), "
                      f"{'AffineGuidedAlign' if algo else 'GuidedAlign'} + ComputeAlignmentStats, replayed once in {dt:.1f} s",
            "_cells": cells, "_seconds": dt, "_idx": idx}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    algo = 1 if args.algo == "affine" else 0
    quality = args.scorefn == "quality"
    # the sample is drawn from a shard prefix large enough to hold the shard's mix of lengths and bands; it is bounded by
    # CPU time, not by the 100k pairs
    batch = make_workload(min(args.jobs, max(4096, n_threads * 128)), args.seed, with_qual=quality, **workload_args(args))
    per_step = args.ref_seconds if args.ref_seconds else max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    idx = cpu_sample(batch, n_threads, per_step, algo)
    for _ in range(args.warmup):
        cpu_replay(batch, algo, n_threads, per_step, quality=quality, idx=idx[:max(n_threads, len(idx) // 4)])
    vals, ms = [], []
    for _ in range(args.steps):
        r = cpu_replay(batch, algo, n_threads, per_step, quality=quality, idx=idx)
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
    """Jobs [a, b) of a JobBatch as views into the same (pinned) arrays, offsets rebased; the guide also in the packed form
    the C++ adapter (blasr_gpu::RefineBatch::Add) hands to the library (host-side preparation, outside the timed region)."""
    from blasr_b200 import JobBatch
    from blasr_b200.align import pack_guide
    q0, q1 = int(batch.qOff[a]), int(batch.qOff[b]); t0, t1 = int(batch.tOff[a]), int(batch.tOff[b])
    g0, g1 = int(batch.guideOff[a]), int(batch.guideOff[b])
    v = JobBatch(batch.q[q0:q1], batch.qOff[a:b + 1] - batch.qOff[a], batch.t[t0:t1], batch.tOff[a:b + 1] - batch.tOff[a],
                 batch.guide[g0:g1], batch.guideOff[a:b + 1] - batch.guideOff[a],
                 batch.qual[q0:q1] if batch.qual is not None else None, batch.band[a:b] if batch.band is not None else None)
    gp, gw = pack_guide(v.guide, v.guideOff)
    v.guidePacked, _k1 = pinned_copy(gp); v.guideWide, _k2 = pinned_copy(gw.reshape(-1, 4) if len(gw) else np.zeros((0, 4), np.uint32))
    v._keep = (_k1, _k2)
    v.tRefOffAbs, _k3 = pinned_copy(np.ascontiguousarray(batch.tOff[a:b], np.uint64))    # where the targets sit in the whole shard's t
    v._keep = (_k1, _k2, _k3)
    return v


def _pin_batch(batch, keep):
    for name in ("q", "qOff", "t", "tOff", "guide", "guideOff", "band") + (("qual",) if batch.qual is not None else ()):
        if getattr(batch, name) is None:
            continue
        v, t = pinned_copy(getattr(batch, name)); setattr(batch, name, v); keep.append(t)
    return batch


def e2e_measure(local, batch, fn, algo, n_threads, n_chunks, warm, steps, barrier, resident_reference=False):
    """submit (H2D + kernels) + collect (D2H) through the C ABI, host buffers in, host results out.  The way a multi-threaded
    host (blasr's MapReads pthreads) drives the library: a few host threads, each with its own context, push sub-batches of
    the shard; copies of one sub-batch overlap the kernels of another."""
    from blasr_b200 import Aligner
    n_chunks = max(1, min(n_chunks, batch.n))
    bounds = np.linspace(0, batch.n, n_chunks + 1).astype(np.int64)
    chunks = [range_view(batch, int(bounds[i]), int(bounds[i + 1])) for i in range(n_chunks)]
    workers = [Aligner(local) for _ in range(max(1, min(n_threads, n_chunks)))]
    if resident_reference:
        # the targets become windows of a reference resident on the device (blasr's genome: loaded once, outside the timed
        # region, like the reference program loads it into RAM); per step only 8 bytes per job describe them
        workers[0].set_reference(batch.t)
        for c in chunks:
            c.tRefOff = c.tRefOffAbs

    def one_pass():
        lock = threading.Lock()
        tot = {"cells": 0, "ok": 0, "h2d": 0, "d2h": 0}
        errs = []

        def finish(a, tk):
            res = a.collect(tk)
            with lock:
                tot["cells"] += int(res.timing.cells); tot["ok"] += int((res.results["status"] == 0).sum())
                tot["h2d"] += int(res.timing.h2dBytes); tot["d2h"] += int(res.timing.d2hBytes)
                tot["allocs"] = max(tot.get("allocs", 0), int(res.timing.devAllocs) + int(res.timing.pinAllocs))
            a.release(tk)

        def work(w, a):
            # bgpu_submit only enqueues: a host thread hands over its next sub-batch before it collects the previous one, so
            # the device always has a ticket to copy in behind the ones that compute (two tickets in flight per thread).
            # Every thread owns its sub-batches (w, w + T, ...), the way every MapReads thread owns its reads: a context then
            # meets the same ticket sizes in every pass and its allocation cache holds after the first one (with the sub-batches
            # handed out first come first served, one pass in ~20 met a larger ticket than its context had cached slabs for and
            # paid ~400 ms of cudaMalloc / cudaHostAlloc inside the timed region)
            try:
                prev = None
                for i in range(w, n_chunks, len(workers)):
                    tk = a.submit(chunks[i], fn, algo, band=16, doStats=True, compact=True, packed=True)
                    if prev is not None:
                        finish(a, prev)
                    prev = tk
                if prev is not None:
                    finish(a, prev)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(w, a)) for w, a in enumerate(workers)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]
        return tot
    try:
        for _ in range(warm):
            one_pass()
        barrier()
        t0 = time.perf_counter()
        pass_ms = []
        for _ in range(steps):
            p0 = time.perf_counter()
            tot = one_pass()
            pass_ms.append((time.perf_counter() - p0) * 1e3)
        barrier()
        sec = (time.perf_counter() - t0) / steps
        if max(pass_ms) > 1.5 * min(pass_ms):      # an outlier pass (a cold allocation, a stalled host thread): say so
            print(f"bench: e2e passes of uneven length {['%.1f' % x for x in pass_ms]} ms (resident_reference={resident_reference})", file=sys.stderr)
    finally:                                       # also on a failed pass: the contexts (and their tickets' memory) go away
        if resident_reference:
            workers[0].set_reference(None)
        for a in workers:
            a.close()
    return {"sec": sec, "cells": tot["cells"], "ok": tot["ok"], "h2d": tot["h2d"], "d2h": tot["d2h"], "threads": len(workers), "chunks": n_chunks,
            "pass_ms": pass_ms, "allocs": tot.get("allocs")}


def resident_pipelined(local, batch, fn, algo, n_threads, n_chunks, warm, steps, barrier):
    """The shard resident in HBM as n_chunks tickets spread over n_threads host threads (a context each), every step re-executing
    all kernels of all tickets (bgpu_rerun): what the e2e leg does, minus the copies.  Tickets of different contexts run on
    different streams, so one ticket's prep / traceback / emit (latency-bound) overlap another one's fill (issue-bound) the way
    they do under a multi-threaded host.  Timed on the wall clock between device-wide synchronisations."""
    from blasr_b200 import Aligner
    n_chunks = max(1, min(n_chunks, batch.n)); n_threads = max(1, min(n_threads, n_chunks))
    bounds = np.linspace(0, batch.n, n_chunks + 1).astype(np.int64)
    chunks = [range_view(batch, int(bounds[i]), int(bounds[i + 1])) for i in range(n_chunks)]
    workers = [Aligner(local) for _ in range(n_threads)]
    tickets = [[] for _ in workers]
    cells = ok = 0
    try:
        for i, c in enumerate(chunks):
            a = workers[i % n_threads]
            tk = a.submit(c, fn, algo, band=16, doStats=True, compact=True, packed=True)
            res = a.collect(tk)
            cells += int(res.timing.cells); ok += int((res.results["status"] == 0).sum())
            tickets[i % n_threads].append(tk)
        errs = []

        def work(a, tks):
            try:
                for tk in tks:
                    a.rerun(tk)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        def one_pass():
            th = [threading.Thread(target=work, args=(a, tks)) for a, tks in zip(workers, tickets)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            if errs:
                raise errs[0]
        for _ in range(warm):
            one_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass()
        barrier()
        sec = (time.perf_counter() - t0) / steps
    finally:
        for a, tks in zip(workers, tickets):
            for tk in tks:
                a.release(tk)
            a.close()
    return {"sec": sec, "cells": cells, "ok": ok, "threads": n_threads, "tickets": n_chunks}


def torch_sync():
    import torch
    torch.cuda.synchronize()


def parity_sample(al, batch, fn, algo, quality, n=256, seed=11):
    """Outside every timed region: a seeded sample of the very shard the bench times (at least n/8 of it band-64 jobs of
    >= 15 kb when the shard has them), through the GPU and through oracle/_ref, every field compared."""
    from concurrent.futures import ThreadPoolExecutor
    from tests import cases, oracle as O
    which = "ref" if O.have_ref() else "orc"
    rng = np.random.default_rng(seed)
    ql = np.diff(batch.qOff.astype(np.int64))
    band = batch.band if batch.band is not None else np.full(batch.n, 16)
    big = np.flatnonzero((band == 64) & (ql >= 15000))
    pick = set(rng.choice(big, min(len(big), n // 8), replace=False).tolist()) if len(big) else set()
    rest = rng.permutation(batch.n)
    for i in rest:
        if len(pick) >= min(n, batch.n):
            break
        pick.add(int(i))
    idx = sorted(pick)
    sub = batch.slice(idx)
    res = al.AffineGuidedAlign(sub, fn, 16) if algo else al.GuidedAlign(sub, fn, 16)
    ofn = O.score_fn(fn.scoreMatrix, fn.ins, fn.del_, fn.affineOpen, fn.affineExtend, fn.kind)

    def one(k):
        q, t, g, qv = cases.job_arrays(sub, k)
        j, keep = O.make_job(algo, 1, int(sub.band[k]) if sub.band is not None else 16, q, t, g, qv if quality else None, 0, 0, 1, algo)
        return O.align(which, ofn, j)
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        want = list(ex.map(one, range(sub.n)))
    bad = [idx[k] for k in range(sub.n) if cases.compare(cases.gpu_to_dict(res, k), want[k], cases.GPU_FIELDS)]
    return {"n": len(idx), "mismatches": len(bad), "band64_ge15kb": int(sum(1 for i in idx if band[i] == 64 and ql[i] >= 15000)),
            "checker": "oracle/_ref (unmodified reference templates)" if which == "ref" else "oracle C port",
            "fields": "status score qPos tPos nCells blocks gaps nMatch nMismatch nIns nDel pctSimilarity statsScore",
            "first_bad": bad[:4]}


def measure(al, local, batch, fn, algo, args, steps, warmup, barrier, do_e2e=True, do_parity=True, do_cpu=True, all_cpus=None,
            quality=False, clocks=False):
    """One (workload, aligner, score function) combination: device-resident GCUPS (bgpu_rerun), per-stage times, the fill
    kernel against the int32 roofline, optionally e2e through the C ABI, the parity sample and the CPU baseline."""
    from blasr_b200 import capi
    rec = {}
    al.trim()                          # the e2e contexts below share this GPU: hand them the memory of earlier measurements
    e2e = e2e_measure(local, batch, fn, algo, args.e2e_threads, args.e2e_chunks, max(1, min(warmup, 2)), max(1, min(steps, 3)),
                      barrier) if do_e2e else None
    if e2e and clocks:       # headline record only: the same passes with the targets taken from a device-resident reference
        al.trim()
        try:
            e2e["resident_reference"] = e2e_measure(local, batch, fn, algo, args.e2e_threads, args.e2e_chunks, 1, max(1, min(steps, 3)), barrier,
                                                    resident_reference=True)
        except Exception as e:  # noqa: BLE001  (an affine 100k-pair shard: four contexts' traceback pools + the resident targets do not fit)
            print(f"bench: resident_reference leg skipped: {type(e).__name__}: {e}", file=sys.stderr)
            torch_sync()
    pipelined = None
    if e2e and clocks and int(os.environ.get("WORLD_SIZE", "1")) == 1:      # an explanatory leg of the single-GPU line only
        al.trim()
        try:
            pipelined = resident_pipelined(local, batch, fn, algo, args.e2e_threads, 2 * args.e2e_threads, max(1, min(warmup, 3)), steps, barrier)
        except Exception as e:  # noqa: BLE001  (e.g. the shard does not fit HBM twice over): the single-ticket number stands
            pipelined = {"unavailable": f"{type(e).__name__}: {e}"}
        al.trim()
    # one ticket for the whole shard: the device-resident measurement below re-runs it; its own submit -> collect time is
    # reported as e2e.single_ticket (no copy/compute overlap; second ticket of the context, i.e. with its slabs cached)
    tk = al.submit(batch, fn, algo, band=16, doStats=True)
    al.collect(tk)
    al.release(tk)
    h0 = time.perf_counter()
    tk = al.submit(batch, fn, algo, band=16, doStats=True)
    h1 = time.perf_counter()
    res = al.collect(tk)
    h2 = time.perf_counter()
    tm = res.timing
    cells = int(tm.cells)
    ok = int((res.results["status"] == 0).sum())
    if e2e:
        assert cells == e2e["cells"], (cells, e2e["cells"])
    for _ in range(warmup):
        al.rerun(tk)
    barrier()
    sampler = None
    if clocks:
        sampler = ClockSampler(local); sampler.start()
    ms = {k: [] for k in ("total", "fill", "trace", "prep", "emit")}
    launches = 0
    w0 = time.perf_counter()
    for _ in range(steps):
        t = al.rerun(tk)
        ms["total"].append(t.msTotal); ms["fill"].append(t.msFill); ms["trace"].append(t.msTrace); ms["prep"].append(t.msPrep); ms["emit"].append(t.msEmit)
        launches += int(t.kernelLaunches)
    barrier()
    wall = time.perf_counter() - w0
    if sampler:
        sampler.stop_flag = True; sampler.join(2)
    lane_steps = float(al.timing(tk).fillCells)
    al.release(tk)
    a = 1 if algo == capi.AFFINE_GUIDED else 0
    fill_s = float(np.mean(ms["fill"])) * 1e-3
    fill_gcups = cells / fill_s / 1e9
    rec.update(cells=cells, jobs_ok=ok, dev_ms=float(np.sum(ms["total"])), launches=launches, wall=wall, sampler=sampler,
               stage_ms={"prep": float(np.mean(ms["prep"])), "fill": float(np.mean(ms["fill"])), "trace": float(np.mean(ms["trace"])),
                         "emit": float(np.mean(ms["emit"])), "wall_per_step": wall / steps * 1e3},
               fill_gcups=fill_gcups, fill_s=fill_s, lane_steps_per_cell=lane_steps / max(1, cells), algo=a,
               single={"submit_ms": (h1 - h0) * 1e3, "collect_ms": (h2 - h1) * 1e3}, e2e=e2e, pipelined=pipelined)
    if do_parity:
        rec["parity_sample"] = parity_sample(al, batch, fn, a, quality)
    if do_cpu:
        try:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)            # the CPU baseline gets every host core back
            cb = cpu_replay(batch, a, len(all_cpus or os.sched_getaffinity(0)), args.cpu_seconds, quality=quality)
            rec["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # noqa: BLE001
            rec["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    return rec


def sub_record(rec, steps, int_peak, world=1):
    """A sub-record of the JSON line (single rank's view)."""
    v = rec["cells"] * steps / (rec["dev_ms"] * 1e-3) / 1e9
    out = {"value": v, "unit": "GCUPS", "ms_per_step": rec["dev_ms"] / steps, "cells_per_step": rec["cells"], "jobs_ok": rec["jobs_ok"],
           "aligned_pairs_per_s": rec["jobs_ok"] * steps / (rec["dev_ms"] * 1e-3), "stage_ms": rec["stage_ms"], "gpu_launches": rec["launches"],
           "int_roofline": {"fill_gcups": rec["fill_gcups"], "ops_per_cell": OPS_PER_CELL[rec["algo"]],
                            "achieved": rec["fill_gcups"] * OPS_PER_CELL[rec["algo"]] / 1e3, "peak": int_peak / 1e12, "unit": "Tops/s",
                            "frac": rec["fill_gcups"] * 1e9 * OPS_PER_CELL[rec["algo"]] / int_peak if int_peak else None,
                            "lane_steps_per_cell": rec["lane_steps_per_cell"]}}
    if rec.get("e2e"):
        e = rec["e2e"]
        out["e2e"] = {"value": e["cells"] / e["sec"] / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"],
                      "pairs_per_s": e["ok"] / e["sec"], "ms_per_step": e["sec"] * 1e3}
    for k in ("parity_sample", "cpu_baseline"):
        if k in rec:
            out[k] = rec[k]
    return out


def sdp_guided_batch(n, seed, len_lo, len_hi, bands):
    """configs[1] as SURVEY 8(d) words it: guides = the reference's own SDPAlign(k=11, sdpIns 5, sdpDel 10, indelRate 0.3 x 3) output
    (oracle/_ref ref_sdp_guide, the argument pattern of Blasr.cpp:1716-1722), sliced the way RefineAlignment slices them
    (Blasr.cpp:850-859).  Workload preparation, outside every timed region."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from blasr_b200 import JobBatch, SMRTDistanceMatrix
    from tests import cases, oracle as O
    if not O.have_ref():
        return None
    L = O._load("ref")
    base = make_workload(n, seed, len_lo=len_lo, len_hi=len_hi, bands=bands)
    fn = O.score_fn(SMRTDistanceMatrix, 5, 5, 0, 0)

    def one(i):
        q, t, _, _ = cases.job_arrays(base, i)
        q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
        cap = len(q) + 16
        blocks = np.zeros((cap, 3), np.uint32)
        nb = L.ref_sdp_guide(q.ctypes.data, len(q), t.ctypes.data, len(t), C.byref(fn), 11, 5, 10, C.c_float(0.9), blocks.ctypes.data, cap)
        if nb <= 0:
            return None
        g = blocks[:nb].astype(np.int64)
        q0, t0 = int(g[0, 0]), int(g[0, 1]); q1, t1 = int(g[-1, 0] + g[-1, 2]), int(g[-1, 1] + g[-1, 2])
        g[:, 0] -= q0; g[:, 1] -= t0
        return q[q0:q1].tobytes(), t[t0:t1].tobytes(), g.astype(np.uint32), int(base.band[i]), blocks[:nb].copy()
    nthr = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nthr) as ex:
        raw = list(ex.map(one, range(base.n)))
    cpu_s = time.perf_counter() - t0
    got = [x for x in raw if x is not None]
    out = JobBatch.from_lists([x[0] for x in got], [x[1] for x in got], [x[2] for x in got], None, [x[3] for x in got])
    # what the device SDPAlign is checked against / timed beside: the raw pairs, the reference's absolute blocks per pair
    # and the wall time the reference's SDPAlign took on the host cores (ctypes releases the GIL)
    out.sdp_pairs = base
    out.sdp_ref = raw
    out.sdp_cpu = dict(seconds=cpu_s, threads=nthr)
    return out


def sdp_device_record(al, sdp, fn):
    """SDPAlign itself on the device (bgpu_sdp_align) on the pairs of the sdp_guides workload: wall time of the
    synchronous call from host buffers (H2D + kernel + D2H), every pair's blocks compared with the reference's own SDPAlign,
    and the reference's time for the same pairs on the host cores beside it."""
    base, ref = sdp.sdp_pairs, sdp.sdp_ref
    res = blocks = None
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        res, blocks = al.SDPAlign(base, fn, wordSize=11, sdpIns=5, sdpDel=10, indelRate=0.9)
        times.append(time.perf_counter() - t0)
    bad = refused = 0
    for i in range(base.n):
        if int(res["status"][i]) != 0:
            refused += 1
            continue
        b = blocks[int(res["blockOff"][i]):int(res["blockOff"][i]) + int(res["nBlocks"][i])]
        got = np.stack([b["qPos"] + res["qPos"][i], b["tPos"] + res["tPos"][i], b["length"]], axis=1).astype(np.uint32).reshape(-1, 3)
        want = ref[i][4] if ref[i] is not None else np.zeros((0, 3), np.uint32)
        bad += int(not np.array_equal(got, want.reshape(-1, 3)))
    best = min(times)
    return dict(metric="sdp_pairs_per_s", value=base.n / best, unit="pairs/s", ms_per_call=1e3 * best, pairs=base.n,
                bases=int(len(base.q) + len(base.t)), parity=dict(n=base.n, mismatches=bad, refused=refused, checker="oracle/_ref SDPAlign (unmodified reference)"),
                cpu_baseline=dict(value=base.n / sdp.sdp_cpu["seconds"], unit="pairs/s", cores=sdp.sdp_cpu["threads"], kind="reference",
                                  sample=f"the same {base.n} pairs through the reference's SDPAlign, {sdp.sdp_cpu['seconds']:.2f} s"),
                how="bgpu_sdp_align: one warp per pair, walked by lane 0 (first device version: every phase after the k-mer matching is an order-dependent sequential algorithm), synchronous call from host buffers, best of 3",
                workload="SDPAlign(k=11, sdpIns 5, sdpDel 10, indelRate 0.9, Local, detailed, sdpPrefix 50, recurse 2, recurseOver 1000) -- Blasr.cpp:1716-1722")


def gap_fill_record(al, n):
    """configs[4]-style dense aligner leg: the gap fills of -alignContigs (~41 k AffineKBandAlign jobs of ~80 cells per 1 Mb
    contig, the parameter pattern of Blasr.cpp:1064-1076) through bgpu_submit / bgpu_collect from host buffers, a seeded
    sample compared field by field with the reference's own AffineKBandAlign, and the reference replaying the same jobs on
    the host cores beside it."""
    from blasr_b200 import JobBatch, SMRTDistanceMatrix
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    lens = rng.integers(2, 14, n)                                       # (|q|+1) * (2k+1) ~ 80 cells at k = 4..5
    tl = np.maximum(1, lens + rng.integers(-2, 3, n))
    q = acgt[rng.integers(0, 4, int(lens.sum()))]; t = acgt[rng.integers(0, 4, int(tl.sum()))]
    qOff = np.zeros(n + 1, np.uint64); qOff[1:] = np.cumsum(lens); tOff = np.zeros(n + 1, np.uint64); tOff[1:] = np.cumsum(tl)
    band = np.maximum(np.abs(lens - tl) + 3, 4).astype(np.int32)       # bandSize of AlignSubstring grows with the length difference
    b = JobBatch(q, qOff, t, tOff, np.zeros((0, 3), np.uint32), np.zeros(n + 1, np.uint64), None, band)
    indel = 5
    pr = (indel + 2, indel - 3, indel + 2, indel - 1)                   # Blasr.cpp:1067-1076
    res = al.AffineKBandAlign(b, SMRTDistanceMatrix, pr[0], pr[1], pr[2], pr[3], indel, 0, computeStats=True)   # warm-up (allocations)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); res = al.AffineKBandAlign(b, SMRTDistanceMatrix, pr[0], pr[1], pr[2], pr[3], indel, 0, computeStats=True)
        best = min(best, time.perf_counter() - t0)
    tm = res.timing
    cells = int(((lens + 1) * (2 * band + 1)).sum())
    out = dict(metric="gap_fill_jobs_per_s", value=n / best, unit="jobs/s", ms_per_call=1e3 * best, jobs=n, jobs_ok=int((res.results["status"] == 0).sum()),
               cells=cells, gcups=cells / best * 1e-9,
               device_ms=dict(prep=tm.msPrep, fill=tm.msFill, trace=tm.msTrace, emit=tm.msEmit, total=tm.msTotal),
               host_ms=dict(bgpu_submit=tm.msHostSubmit, bgpu_collect=tm.msHostCollect),
               how="bgpu_submit + bgpu_collect from host buffers, best of 5 (batches of small matrices take the planner-free path: "
                   "submission order, traceback offsets laid out by bgpu_submit, nothing read back before the kernels run)",
               workload=f"{n} AffineKBandAlign jobs, |q| 2-13, k = |dq-dt|+3, mean {cells / n:.0f} cells, hpInsOpen/hpInsExtend/insOpen/insExtend/del = "
                        f"{pr[0]}/{pr[1]}/{pr[2]}/{pr[3]}/{indel} (Blasr.cpp:1067-1076)")
    try:
        from tests import cases, oracle as O
        if O.have_ref():
            ofn = O.score_fn(SMRTDistanceMatrix, 5, 5)
            pick = np.random.default_rng(9).choice(n, size=min(n, 512), replace=False)
            bad = 0
            for i in pick:
                qq, tt, _, _ = cases.job_arrays(b, int(i))
                j, keep = O.make_job(4, 1, int(band[i]), qq, tt, None, None, 0, indel, 1, 0, affineKBand=pr)
                bad += bool(cases.compare(cases.gpu_to_dict(res, int(i)), O.align("ref", ofn, j), cases.GPU_FIELDS))
            out["parity_sample"] = dict(n=len(pick), mismatches=int(bad), checker="oracle/_ref AffineKBandAlign (unmodified reference template)")
            m = min(n, 200000); jobs, keeps = [], []
            for i in range(m):
                qq, tt, _, _ = cases.job_arrays(b, i)
                j, k = O.make_job(4, 1, int(band[i]), qq, tt, None, None, 0, indel, 0, 0, affineKBand=pr); jobs.append(j); keeps.append(k)
            nthr = len(os.sched_getaffinity(0))
            t0 = time.perf_counter(); O.replay("ref", ofn, jobs, nthr); dt = time.perf_counter() - t0
            out["cpu_baseline"] = dict(value=m / dt, unit="jobs/s", cores=nthr, kind="reference", sample=f"the first {m} jobs replayed once in {dt:.2f} s")
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def anchoring_record(al, n_reads, genome_len=4_600_000):
    """Suffix-array anchoring (bgpu_map_reads) on configs[0]'s shape: a 4.6 Mb genome with a few repeat families,
    n_reads reads of 10 kb at 15 % error, each mapped in both strands as MapRead does (Blasr.cpp:2282-2296) with blasr's default
    AnchorParameters.  Wall time of the synchronous call from host buffers, device time of the kernels on resident reads, every
    read's match list compared with the reference's own MapReadToGenome, whose time on the host cores is the CPU baseline."""
    from blasr_b200 import saindex, synth
    g = synth.simulate_genome(genome_len, seed=1)
    sa = saindex.suffix_array(g)
    start, end = saindex.lookup_table(g, sa, 8)
    reads, off = synth.simulate_reads(g, n_reads, 10000, seed=3)
    reads, _keep = pinned_copy(reads)                            # the caller's read buffer is pinned, like the refinement's inputs
    n = len(off) - 1
    al.set_reference(g)
    al.set_suffix_array(sa, start, end, 8)
    mo, m = al.MapReadToGenome(reads, off)                       # warm-up: allocations
    wall = []
    for _ in range(3):
        t0 = time.perf_counter(); mo, m = al.MapReadToGenome(reads, off, copy=False); wall.append(time.perf_counter() - t0)
    dev = []
    for _ in range(3):
        al.map_rerun(); dev.append(al.map_timing())
    ms_locate, ms_rest, positions, h2d, d2h = min(dev, key=lambda x: x[0] + x[1])
    m = m.copy()
    best = min(wall)
    out = dict(metric="anchored_reads_per_s", value=n / (1e-3 * (ms_locate + ms_rest)), unit="read strands/s", reads=n, positions=int(positions),
               matches=int(mo[-1]), device_ms=dict(locate=ms_locate, count_scan_emit=ms_rest),
               positions_per_s=positions / (1e-3 * (ms_locate + ms_rest)),
               e2e=dict(value=n / best, unit="read strands/s", ms_per_call=1e3 * best, h2d_bytes_per_call=int(h2d), d2h_bytes_per_call=int(d2h),
                        how="synchronous bgpu_map_reads from pinned host reads (H2D reads, kernels, D2H offsets + matches into the library's pinned buffer), best of 3"),
               index=dict(genome=int(len(g)), suffix_array_bytes=int(sa.nbytes), lookup_table_bytes=int(start.nbytes + end.nbytes), resident=True),
               how="bgpu_map_rerun on the resident reads, CUDA events, best of 3; one thread per read position, the reference's probe sequence",
               workload=f"{n_reads} reads x 2 strands of 10 kb, 15 % error, on a {genome_len / 1e6:.1f} Mb genome; MapReadToGenome with "
                        "minPrefixMatchLength 8 (lookup table), minMatchLength 12, maxAnchorsPerPosition 1000, stopMappingOnceUnique (blasr's defaults)")
    try:
        from tests import anchor_oracle as ao
        if ao.ref() is not None:
            class _Ix:
                pass
            ix = _Ix(); ix.gpad = ao.padded(g); ix.n = len(g); ix.sa = sa; ix.start = start; ix.end = end; ix.prefixLength = 8
            nthr = len(os.sched_getaffinity(0))
            t0 = time.perf_counter(); ro, rm = ao.map_reads_ref(ix, reads, off, ao.params(), nThreads=nthr, want_matches=False); dt = time.perf_counter() - t0
            ro, rm = ao.map_reads_ref(ix, reads, off, ao.params(), nThreads=nthr)
            same = bool(np.array_equal(ro, mo)) and bool(np.array_equal(rm, np.stack([m["t"], m["q"], m["l"]], axis=1)))
            bad = 0 if same else int(sum(not np.array_equal(rm[int(ro[i]):int(ro[i + 1])],
                                                                np.stack([m["t"], m["q"], m["l"]], axis=1)[int(mo[i]):int(mo[i + 1])]) for i in range(n)))
            out["parity"] = dict(n=n, mismatches=bad, matches_compared=int(ro[-1]), checker="oracle/_ref MapReadToGenome (unmodified reference templates)")
            out["cpu_baseline"] = dict(value=n / dt, unit="read strands/s", cores=nthr, kind="reference",
                                       sample=f"the same {n} read strands through the reference's MapReadToGenome in {dt:.2f} s")
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"] = {"unavailable": str(e)}
    al.set_suffix_array(None)
    al.set_reference(None)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from blasr_b200 import Aligner, DistanceMatrixScoreFunction, QualityValueScoreFunction, capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # the generator forks worker processes: run it before CUDA / NCCL are initialised in this process
    headline_quality = args.scorefn == "quality"
    # the sub-records are single-rank views (one of them, sdp_guides, exists on rank 0 only): they belong to the N = 1 run; with
    # more ranks every rank runs exactly the headline measurement, so that the ranks meet at the same barriers
    want_subs = args.subrecords and world == 1 and (args.algo, args.scorefn) == ("guided", "distance")
    batch = make_workload(args.jobs, args.seed + 1000 * rank, with_qual=headline_quality or want_subs, **workload_args(args))
    prod = make_workload(args.prod_jobs, args.seed + 1000 * rank + 500, len_lo=10000, len_hi=10000, bands=(16,)) if want_subs else None
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    sdp = sdp_guided_batch(args.sdp_jobs, args.seed + 77, args.len_lo, args.len_hi, tuple(int(x) for x in args.bands.split(","))) \
        if (want_subs and rank == 0 and args.sdp_jobs > 0) else None
    numa = bind_to_gpu_cpus(local)     # before any pinned allocation: first touch then lands on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    algo = capi.AFFINE_GUIDED if args.algo == "affine" else capi.GUIDED
    keep = []
    _pin_batch(batch, keep)            # inputs in pinned host memory (the library then DMA's straight from them)
    if prod is not None:
        _pin_batch(prod, keep)

    def mkfn(a, quality):
        return (QualityValueScoreFunction if quality else DistanceMatrixScoreFunction)(ins=5, del_=5, affineOpen=50 if a == capi.AFFINE_GUIDED else 0, affineExtend=0)
    al = Aligner(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solo = rank == 0 and world == 1
    head = measure(al, local, batch, mkfn(algo, headline_quality), algo, args, args.steps, args.warmup, barrier, do_e2e=True,
                   do_parity=solo and args.parity, do_cpu=solo, all_cpus=all_cpus, quality=headline_quality, clocks=True)
    os.sched_setaffinity(0, all_cpus) if numa is None else bind_to_gpu_cpus(local)

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); return float(tt.item())

    def allsum(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM); return float(tt.item())
    cells = head["cells"]
    dev_ms_max = allmax(head["dev_ms"]); e2e_sec_max = allmax(head["e2e"]["sec"]); cells_all = allsum(float(cells)); jobs_all = allsum(float(head["jobs_ok"]))
    value = cells_all * args.steps / (dev_ms_max * 1e-3) / 1e9
    ms_per_step = dev_ms_max / args.steps
    resident = {"single_ticket": {"value": value, "ms_per_step": ms_per_step,
                                  "how": "one ticket for the whole shard, bgpu_rerun, CUDA events first kernel start -> last kernel end (stage_ms is this run)"}}
    pl = head.get("pipelined")
    if pl and "sec" in pl:
        pl_sec = allmax(pl["sec"])
        resident["pipelined"] = {"value": cells_all / pl_sec / 1e9, "ms_per_step": pl_sec * 1e3, "threads": pl["threads"], "tickets": pl["tickets"],
                                 "how": "the shard resident as several tickets on several contexts (host threads), all re-executed concurrently per "
                                        "step: one ticket's prep / traceback / emit overlap another one's fill; wall clock between device-wide synchronisations"}
        if resident["pipelined"]["value"] > value:
            value = resident["pipelined"]["value"]; ms_per_step = pl_sec * 1e3
    elif pl:
        resident["pipelined"] = pl
    e2e_val = cells_all / e2e_sec_max / 1e9
    resident_ref = None
    # the leg may have been skipped on ONE rank (out of memory): every rank takes the same branch, or the all-reduce below hangs
    have_rr = allsum(1.0 if head["e2e"].get("resident_reference") else 0.0) == float(world)
    if have_rr:
        rr = head["e2e"]["resident_reference"]
        rr_sec = allmax(rr["sec"])
        resident_ref = {"value": cells_all / rr_sec / 1e9, "unit": "GCUPS", "ms_per_step": rr_sec * 1e3, "h2d_bytes_per_step": rr["h2d"],
                        "d2h_bytes_per_step": rr["d2h"], "jobs_ok": rr["ok"], "pass_ms": rr.get("pass_ms"),
                        "how": "the same passes with the targets given as 8-byte offsets into a reference resident on the device "
                               "(bgpu_set_reference, uploaded once outside the timed region the way blasr loads its genome; here the shard's "
                               "own target array) and gathered there: reads, guides and results still cross PCIe every step"}
    single = head["single"]
    uploaded = {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": head["e2e"]["h2d"], "d2h_bytes_per_step": head["e2e"]["d2h"],
                "pairs_per_s": jobs_all / e2e_sec_max, "ms_per_step": e2e_sec_max * 1e3, "pass_ms": head["e2e"].get("pass_ms"),
                "how": "the same passes with every job's target window uploaded from host memory as well (the round-1 form of this number)"}
    how = (f"{head['e2e']['threads']} host threads x own context, {head['e2e']['chunks']} sub-batches of the shard, pinned host buffers in (reads; guides "
           "packed 3 B / block, the form the C++ adapter writes), pinned result arena out (run-length paths, expanded by the adapter's Store), "
           "H2D + D2H inside the timed region")
    if resident_ref:
        # the configuration a blasr host runs: the genome is loaded to the device once (bgpu_set_reference; the anchoring needs it
        # there anyway), a job names its target window by offset -- like the reference, which keeps the genome in RAM for the whole
        # run and hands the aligners pointers into it.  Reads, guides and results cross PCIe every step.
        e2e_rec = {"value": resident_ref["value"], "unit": "GCUPS", "h2d_bytes_per_step": resident_ref["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": resident_ref["d2h_bytes_per_step"], "pairs_per_s": jobs_all / (resident_ref["ms_per_step"] * 1e-3),
                   "ms_per_step": resident_ref["ms_per_step"], "pass_ms": resident_ref.get("pass_ms"),
                   "how": how + "; targets = 8-byte offsets into the genome resident on the device (bgpu_set_reference, loaded once outside the "
                                "timed region the way blasr loads its genome into RAM; here the shard's own target array), gathered there",
                   "targets_uploaded": uploaded}
    else:
        e2e_rec = dict(uploaded, how=how + "; every job's target window uploaded from host memory")
    e2e_rec["single_ticket"] = {"value": cells / ((single["submit_ms"] + single["collect_ms"]) * 1e-3) / 1e9,
                                "submit_ms": single["submit_ms"], "collect_ms": single["collect_ms"],
                                "how": "one bgpu_submit + bgpu_collect for the whole shard, targets uploaded (no copy / compute overlap)"}
    int_peak, _ = al.int_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    a = head["algo"]
    # DRAM bytes of the fill kernels from the committed ncu --set full capture of this very workload (same pairs, same seed)
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "fill_traffic.json")))[args.algo]
        default_wl = (args.len_lo, args.len_hi, args.bands) == (LEN_LO, LEN_HI, ",".join(str(b) for b in BANDS))
        if tr["pairs"] == args.jobs and tr["seed"] == args.seed and tr["scorefn"] == args.scorefn and default_wl:
            traffic, traffic_src = tr["dram_bytes_per_step"], tr["source"]
    except Exception:  # noqa: BLE001
        pass
    fill_s, fill_gcups = head["fill_s"], head["fill_gcups"]
    out = {
        "metric": "banded_dp_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "resident": resident,
        "dtype": "int32", "data": "synthetic", "config": config_dict(args, args.jobs),
        "aligned_pairs_per_s": jobs_all / (ms_per_step * 1e-3),
        "e2e": e2e_rec,
        "gpu_launches": head["launches"],
        "stage_ms": head["stage_ms"],
        "roofline": {"bound": "hbm", "kernel": "fill_guided_kernel", "achieved": cells * BYTES_PER_CELL[a] / fill_s / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": cells * BYTES_PER_CELL[a] / fill_s / 1e9 / hbm_peak, "traffic": traffic,
                     "algorithmic_bytes": cells * BYTES_PER_CELL[a], "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                     "note": "integer DP: the HBM roofline is not the binding one, see int_roofline"},
        "int_roofline": {"bound": "int32 ALU issue", "fill_gcups": fill_gcups, "ops_per_cell": OPS_PER_CELL[a],
                         "achieved": fill_gcups * OPS_PER_CELL[a] / 1e3, "peak": int_peak / 1e12, "unit": "Tops/s",
                         "frac": fill_gcups * 1e9 * OPS_PER_CELL[a] / int_peak if int_peak else None,
                         "lane_steps_per_cell": head["lane_steps_per_cell"],
                         "peak_by_mix_tops": {k: v / 1e12 for k, v in al.int_peak_modes.items()},
                         "peak_source": "bgpu_measure_int_peak: best of add / min / mad / add+mad chains on this device"},
        "clocks": head["sampler"].summary(), "jobs_ok": int(jobs_all), "host_cpus_bound_per_rank": numa,
    }
    for k in ("parity_sample", "cpu_baseline"):
        if k in head:
            out[k] = head[k]
    if want_subs:
        # the other aligners of the hot path, on this rank's GPU, so that the driver's record pins them too (single-rank views).
        # A leg that fails (out of memory on a smaller device, a missing prebuilt checker) is reported as unavailable: the
        # headline above was measured already and must reach the driver whatever happens here.
        sub_steps, sub_warm = max(2, min(args.steps, 3)), 3

        def leg(name, f):
            try:
                out[name] = f()
            except Exception as e:  # noqa: BLE001
                out[name] = {"unavailable": f"{type(e).__name__}: {e}"}
                print(f"bench: sub-record {name} failed: {type(e).__name__}: {e}", file=sys.stderr)
                try:
                    torch_sync(); al.trim()
                except Exception:  # noqa: BLE001
                    pass

        def affine_leg():
            r = measure(al, local, batch, mkfn(capi.AFFINE_GUIDED, False), capi.AFFINE_GUIDED, args, sub_steps, sub_warm, barrier,
                        do_e2e=solo, do_parity=solo and args.parity, do_cpu=solo, all_cpus=all_cpus)
            return dict(sub_record(r, sub_steps, int_peak), workload="the configs[1] pairs above through AffineGuidedAlign (affineOpen 50, affineExtend 0)")

        def production_leg():
            r = measure(al, local, prod, mkfn(capi.AFFINE_GUIDED, False), capi.AFFINE_GUIDED, args, sub_steps, sub_warm, barrier,
                        do_e2e=solo, do_parity=solo and args.parity, do_cpu=solo, all_cpus=all_cpus)
            return dict(sub_record(r, sub_steps, int_peak),
                        workload=f"blasr's refinement jobs (configs[0] shape): {args.prod_jobs} pairs of 10 kb, band 16, AffineGuidedAlign, "
                                 "ins 5 / del 5 / affineOpen 50 / affineExtend 0 (MappingParameters.h:338-342,395-397)")

        def quality_leg():
            r = measure(al, local, batch, mkfn(capi.GUIDED, True), capi.GUIDED, args, sub_steps, sub_warm, barrier,
                        do_e2e=False, do_parity=solo and args.parity, do_cpu=solo, all_cpus=all_cpus, quality=True)
            return dict(sub_record(r, sub_steps, int_peak),
                        workload="configs[3]-style: the configs[1] pairs with a simulated QV track, GuidedAlign x QualityValueScoreFunction")

        def sdp_guides_leg():
            _pin_batch(sdp, keep)
            r = measure(al, local, sdp, mkfn(capi.GUIDED, False), capi.GUIDED, args, sub_steps, sub_warm, barrier, do_e2e=False,
                        do_parity=args.parity, do_cpu=False)
            return dict(sub_record(r, sub_steps, int_peak),
                        workload=f"{sdp.n} pairs of the configs[1] generator with guides = the reference's own SDPAlign(k=11, sdpIns 5, "
                                 "sdpDel 10, indelRate 0.3 x 3) output sliced as RefineAlignment does (SURVEY 8d C2); a small ticket: "
                                 "the fill kernel's tail is a visible share of it")
        leg("affine", affine_leg)
        leg("affine_production", production_leg)
        leg("quality", quality_leg)
        if sdp is not None and sdp.n:
            leg("sdp_guides", sdp_guides_leg)
            leg("sdp_device", lambda: sdp_device_record(al, sdp, mkfn(capi.GUIDED, False)))
        if solo and args.gap_jobs > 0:
            leg("gap_fills", lambda: gap_fill_record(al, args.gap_jobs))
        if solo and args.anchor_reads > 0:
            leg("anchoring", lambda: anchoring_record(al, args.anchor_reads))
    try:
        al.close()
    except Exception as e:  # noqa: BLE001
        print(f"bench: closing the context failed: {type(e).__name__}: {e}", file=sys.stderr)
    if solo and args.pipeline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import pipeline_bench
            os.sched_setaffinity(0, all_cpus)
            out["pipeline"] = pipeline_bench.measure(n_reads=args.pipeline_reads, device=local)
        except BaseException as e:  # noqa: BLE001
            out["pipeline"] = {"unavailable": f"{type(e).__name__}: {e}"}
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
