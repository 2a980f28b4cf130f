"""GPU, pipeline level (VERDICT row g).

1. Job sets dumped from the UNMODIFIED reference pipeline (tests/golden/jobs_*_small.bgj.gz) replayed through the C ABI and
   through oracle/_ref: every field equal, no job refused.
2. The reference PROGRAM with RefineAlignments routed through the library (baseline/_ref/blasrmc_gpu = INTEGRATION.md
   section 2 compiled for real) against the stock program (baseline/_ref/blasrmc): sorted -sam, -m 4 and -m 5 output must be
   identical line for line, at -nproc 1 and with several MapReads pthreads sharing the GPU through RefineService.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from blasr_b200 import capi, jobdump
from . import cases, dumps, oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BL = os.path.join(ROOT, "baseline")
STOCK, GPU = os.path.join(BL, "_ref", "blasrmc"), os.path.join(BL, "_ref", "blasrmc_gpu")


@pytest.mark.parametrize("cfg", ["c0", "c2", "c4"])
def test_dump_replay_matches_reference(aligner, cfg):
    which = "ref" if O.have_ref() else "orc"
    n = 0
    for g in jobdump.load(dumps.DUMPS[cfg]):
        res = dumps.gpu_group(aligner, g)
        assert (res.results["status"] == 0).all(), (cfg, jobdump.KIND_NAMES[g.kind], np.unique(res.results["status"]))
        want = dumps.oracle_group(which, g)
        for i in range(g.batch.n):
            bad = cases.compare(cases.gpu_to_dict(res, i), want[i], cases.GPU_FIELDS)
            assert not bad, (cfg, jobdump.KIND_NAMES[g.kind], i, bad)
        n += g.batch.n
    assert n > 0


def _data(tmp, cfg):
    d = str(tmp / cfg)
    args = {"c0": ["c0", d, "--genome", "400000", "--reads", "60", "--len", "4000", "--seed", "11"],
            "c2": ["c2", d, "--genome", "900000", "--reads", "16", "--lo", "4000", "--hi", "9000", "--seed", "12"],
            "c4": ["c4", d, "--genome", "1000000", "--contigs", "2", "--len", "80000", "--seed", "13"]}[cfg]
    subprocess.check_call([sys.executable, os.path.join(BL, "make_data.py")] + args, stdout=subprocess.DEVNULL)
    return d


def _run(exe, d, out, flags, nproc):
    cmd = [exe, "reads.fa", "genome.fa", "-nproc", str(nproc), "-out", out] + flags
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (cmd, r.stdout[-2000:], r.stderr[-2000:])
    lines = [x for x in open(os.path.join(d, out)).read().splitlines() if not x.startswith("@PG")]
    return sorted(lines)


@pytest.mark.parametrize("cfg,extra", [("c0", []), ("c2", ["-bestn", "10"]), ("c4", ["-alignContigs"])])
def test_cli_output_identical_to_stock_blasr(tmp_path, cfg, extra):
    if not (os.path.exists(STOCK) and os.path.exists(GPU)):
        pytest.skip("baseline/_ref binaries absent (run `make -C baseline all` where /root/reference is mounted)")
    d = _data(tmp_path, cfg)
    for fmt, flags in (("sam", ["-sam"]), ("m4", ["-m", "4"]), ("m5", ["-m", "5"])):
        want = _run(STOCK, d, f"stock.{fmt}", flags + extra, 1)
        assert len(want) > 0
        for nproc in (1, 4):
            got = _run(GPU, d, f"gpu{nproc}.{fmt}", flags + extra, nproc)
            assert len(got) == len(want), (cfg, fmt, nproc, len(got), len(want))
            diff = [(a, b) for a, b in zip(got, want) if a != b]
            assert not diff, (cfg, fmt, nproc, len(diff), diff[0][0][:300], diff[0][1][:300])
