"""CPU: the assembly of bench.py's JSON line with the device stubbed out (host logic only -- no aligner runs here, the
numbers are canned): every key of the bench contract is present, a sub-record that throws is reported as unavailable without
losing the line, and the headline end-to-end number is the resident-genome leg when every rank has one."""
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Sampler:
    def summary(self):
        return {"sm_mhz": 1965.0, "sm_max_mhz": 1965, "reasons": []}


def _canned(algo, with_e2e, with_rr):
    e2e = None
    if with_e2e:
        e2e = {"sec": 0.1, "cells": 8_000_000_000, "ok": 1000, "h2d": 10, "d2h": 5, "threads": 4, "chunks": 4, "pass_ms": [100.0], "allocs": 0}
        if with_rr:
            e2e["resident_reference"] = dict(e2e, sec=0.08, h2d=6)
    return {"cells": 8_000_000_000, "jobs_ok": 1000, "dev_ms": 40.0, "launches": 20, "wall": 0.05, "sampler": _Sampler(),
            "stage_ms": {"prep": 1.0, "fill": 5.0, "trace": 1.0, "emit": 1.0, "wall_per_step": 8.1}, "fill_gcups": 1600.0, "fill_s": 0.005,
            "lane_steps_per_cell": 1.25, "algo": algo, "single": {"submit_ms": 1.0, "collect_ms": 19.0}, "e2e": e2e, "pipelined": None,
            "parity_sample": {"n": 256, "mismatches": 0}, "cpu_baseline": {"value": 1.6, "unit": "GCUPS", "cores": 16, "kind": "reference", "sample": "canned"}}


def _run(monkeypatch, argv, with_rr=True, fail=()):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    import blasr_b200
    from blasr_b200 import capi

    class FakeAligner:
        int_peak_modes = {"add": 36e12}

        def __init__(self, *_):
            pass

        def int_peak(self):
            return 36e12, 1965.0

        def trim(self):
            pass

        def close(self):
            pass

    def fake_measure(al, local, batch, fn, algo, args, steps, warmup, barrier, do_e2e=True, clocks=False, **kw):
        name = "affine" if algo == capi.AFFINE_GUIDED else "guided"
        if name in fail and not clocks:
            raise RuntimeError("canned failure")
        return _canned(1 if algo == capi.AFFINE_GUIDED else 0, do_e2e, with_rr and clocks)

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_: None)
    monkeypatch.setattr(blasr_b200, "Aligner", FakeAligner)
    monkeypatch.setattr(bench, "measure", fake_measure)
    monkeypatch.setattr(bench, "make_workload", lambda n, seed, **kw: types.SimpleNamespace(n=n))
    monkeypatch.setattr(bench, "sdp_guided_batch", lambda n, *a: types.SimpleNamespace(n=n))
    monkeypatch.setattr(bench, "_pin_batch", lambda *a: None)
    monkeypatch.setattr(bench, "bind_to_gpu_cpus", lambda *_: None)
    monkeypatch.setattr(bench, "sdp_device_record", lambda *a: {"metric": "sdp_pairs_per_s", "value": 7000.0})
    monkeypatch.setattr(bench, "gap_fill_record", lambda *a: (_ for _ in ()).throw(MemoryError("canned")) if "gap_fills" in fail else {"value": 1.5e7})
    monkeypatch.setattr(bench, "anchoring_record", lambda *a: {"value": 8e5})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--no-pipeline"] + argv)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_ours(bench.parse())
    lines = [x for x in buf.getvalue().splitlines() if x.strip()]
    assert len(lines) == 1, lines                       # ONE JSON line
    return json.loads(lines[0])


CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks")


def test_bench_line_has_the_contract_keys(monkeypatch):
    d = _run(monkeypatch, ["--steps", "5", "--warmup", "3"])
    for k in CONTRACT:
        assert k in d, k
    assert d["metric"] == "banded_dp_gcups" and d["unit"] == "GCUPS" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["vs_baseline"] is None and d["dtype"] == "int32"
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    assert np.isclose(d["value"], 8e9 * 5 / 0.040 / 1e9) and np.isclose(d["ms_per_step"], 8.0)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert np.isclose(d["roofline"]["frac"], d["roofline"]["achieved"] / d["roofline"]["peak"])
    assert np.isclose(d["int_roofline"]["frac"], 1600e9 * 8 / 36e12)
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    # the headline end-to-end number is the resident-genome leg; the targets-uploaded form stays in the line
    assert np.isclose(d["e2e"]["value"], 8e9 / 0.08 / 1e9) and d["e2e"]["h2d_bytes_per_step"] == 6
    assert np.isclose(d["e2e"]["targets_uploaded"]["value"], 8e9 / 0.1 / 1e9)
    assert np.isclose(d["e2e"]["single_ticket"]["value"], 8e9 / 0.020 / 1e9)
    for k in ("affine", "affine_production", "quality", "sdp_guides", "sdp_device", "gap_fills", "anchoring"):
        assert k in d and "unavailable" not in d[k], k
    assert d["affine"]["int_roofline"]["ops_per_cell"] == 16 and d["quality"]["int_roofline"]["ops_per_cell"] == 8


def test_bench_line_survives_failing_sub_records(monkeypatch):
    d = _run(monkeypatch, [], with_rr=False, fail=("affine", "gap_fills"))
    assert d["affine"]["unavailable"].startswith("RuntimeError") and d["affine_production"]["unavailable"].startswith("RuntimeError")
    assert d["gap_fills"]["unavailable"].startswith("MemoryError")
    assert "unavailable" not in d["quality"] and "unavailable" not in d["anchoring"]
    # no rank had a resident-genome leg: the end-to-end number is the targets-uploaded one
    assert np.isclose(d["e2e"]["value"], 80.0) and "targets_uploaded" not in d["e2e"] and d["value"] > 0


def test_bench_line_without_sub_records(monkeypatch):
    d = _run(monkeypatch, ["--no-subrecords"])
    assert "affine" not in d and "anchoring" not in d and d["value"] > 0
