"""CPU: the C-ABI library loads and exports every symbol include/blasr_gpu.h declares (no compute without a GPU),
host-side containers, workload generator, and the N>1 sharding path under gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header():
    from blasr_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "blasr_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(bgpu_[a-z_]+)\s*\(", hdr)))
    assert set(declared) == set(capi.EXPORTS), (declared, capi.EXPORTS)
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bgpu_version() == 107


def test_struct_layouts_match_header(tmp_path):
    """sizeof of every POD of include/blasr_gpu.h, taken from the header by the C compiler, against the ctypes / numpy mirrors."""
    from blasr_b200 import capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "blasr_gpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(bgpu_scorefn),sizeof(bgpu_params),sizeof(bgpu_batch),sizeof(bgpu_job),sizeof(bgpu_result),"
                   "sizeof(bgpu_block),sizeof(bgpu_gap),sizeof(bgpu_arena),sizeof(bgpu_timing));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    want = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    got = [C.sizeof(capi.ScoreFn), C.sizeof(capi.Params), C.sizeof(capi.Batch), C.sizeof(capi.Job), capi.RESULT_DTYPE.itemsize,
           capi.BLOCK_DTYPE.itemsize, capi.GAP_DTYPE.itemsize, C.sizeof(capi.Arena), C.sizeof(capi.Timing)]
    assert got == want, (got, want)
    assert capi.RESULT_DTYPE.itemsize == 88 and capi.BLOCK_DTYPE.itemsize == 12 and capi.GAP_DTYPE.itemsize == 8


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path must refuse to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from blasr_b200 import Aligner, BgpuError
    with pytest.raises(BgpuError):
        Aligner(0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "blasr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no dependency on the reference or the oracle", ""), f


def test_synth_is_deterministic_and_well_formed():
    from blasr_b200 import synth
    a = synth.simulate_pairs(40, 200, 900, seed=7, bands=[16, 32], with_qual=True, workers=1)
    b = synth.simulate_pairs(40, 200, 900, seed=7, bands=[16, 32], with_qual=True, workers=2)
    for k in ("q", "t", "qOff", "tOff", "guide", "guideOff", "qual", "band"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for i in range(a.n):
        g = a.guide[int(a.guideOff[i]):int(a.guideOff[i + 1])].astype(np.int64)
        ql = int(a.qOff[i + 1] - a.qOff[i]); tl = int(a.tOff[i + 1] - a.tOff[i])
        assert g[0, 0] == 0 and g[0, 1] == 0 and g[-1, 0] + g[-1, 2] == ql and g[-1, 1] + g[-1, 2] == tl
        assert (g[1:, 0] >= g[:-1, 0] + g[:-1, 2]).all() and (g[1:, 1] >= g[:-1, 1] + g[:-1, 2]).all()


def test_jobbatch_roundtrip():
    from blasr_b200 import JobBatch
    b = JobBatch.from_lists([b"ACGT", b"", b"GG"], [b"AC", b"T", b""], [np.array([[0, 0, 2]]), np.zeros((0, 3)), np.zeros((0, 3))])
    assert b.n == 3 and list(b.qOff) == [0, 4, 4, 6] and list(b.tOff) == [0, 2, 3, 3] and list(b.guideOff) == [0, 1, 1, 1]
    s = b.slice([2, 0])
    assert s.q.tobytes() == b"GGACGT" and list(s.guideOff) == [0, 0, 1]


def test_shard_and_merge():
    from blasr_b200 import shard
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            parts = [shard.shard_indices(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            merged = shard.merge_in_read_order(n, [p * 10 for p in parts])
            assert merged.tolist() == [10 * i for i in range(n)]


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["BGPU_ROOT"])
from blasr_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
mine = shard.shard_indices(n, rank, world)
local = np.stack([mine * 3 + 1, mine * mine], axis=1)          # stand-in per-job records (score, qPos)
merged = shard.gather_records(local, n)
assert merged[:, 0].tolist() == [3 * i + 1 for i in range(n)] and merged[:, 1].tolist() == [i * i for i in range(n)]
assert shard.all_reduce_scalar(10.0 + rank, "max") == 10.0 + world - 1
assert shard.all_reduce_scalar(float(len(mine)), "sum") == float(n)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, BGPU_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29677", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


# ThreeBit[] of the reference (common/NucConversion.h:48-84) as (code -> bytes); generated here from
# oracle/_ref (ref_three_bit) and re-checked against it whenever the reference build is present.
_THREE_BIT = {0: b"\x00Aa", 1: b"\x01Cc", 2: b"\x02Gg", 3: b"\x03Tt", 4: b"\x04BDHKMNRSUVWY_bdhkmnrsuvwxy", 5: b"$"}


def test_base_code_table():
    """All 256 entries of the kernels' base table (bgpu_base_code) equal the reference's ThreeBit[]."""
    from blasr_b200 import capi
    from . import oracle as O
    lib = capi.lib()
    want = np.full(256, 255, np.int64)
    for code, chars in _THREE_BIT.items():
        want[list(chars)] = code
    got = np.array([lib.bgpu_base_code(c) for c in range(256)])
    assert np.array_equal(got, want), np.nonzero(got != want)
    if O.have_ref():
        L = O._load("ref")
        assert [L.ref_three_bit(c) for c in range(256)] == want.tolist()


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the reference's own templates on the host cores, no GPU): rank 0 prints one JSON line with
    the contract's keys; the other ranks of a torchrun launch exit 0 without work."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-seconds", "0.5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "banded_dp_gcups" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "pairs" in cb["sample"]
    assert line["config"]["workload"].startswith("configs[1]") and line["config"]["pairs_per_gpu"] == 100000
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""
